// Sample-rate conversion in front of the path (reference: SttEngine::resample_audio,
// /root/reference/src/stt_engine.cpp:87-115 -> libsamplerate src_simple(SRC_SINC_FASTEST), called at :138-145
// when the input is not 16 kHz). libsamplerate is third-party code that is not in the
// tree: parity with it is NOT claimed. This is the same published method (band-limited interpolation with a
// windowed-sinc table, linear interpolation between table points, cut-off scaled by the ratio when
// downsampling) with its own 16-zero-crossing Kaiser(9) window: > 89 dB tone SNR for 8 / 22.05 / 44.1 / 48 kHz
// -> 16 kHz. One thread per output sample, taps walked left to right in single precision without FMA
// contraction, so that the CPU statement of this converter kept with the tests defines the result bit for bit.
#include "common.cuh"
#include "kernels.cuh"

namespace sw {
namespace {

__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ in, int64_t n_in, int sr_in, int sr_out, const float* __restrict__ table,
                float scale, float gscale, int half, int64_t n_out, float* __restrict__ out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_out) return;
  const int64_t num = n * sr_in;
  const int64_t xi = num / sr_out;
  const float frac = (float)((double)(num % sr_out) / (double)sr_out);
  float acc = 0.0f;
  for (int k = -half + 1; k <= half; ++k) {
    const int64_t idx = xi + k;
    if (idx < 0 || idx >= n_in) continue;
    const float dist = __fmul_rn(fabsf(__fsub_rn((float)k, frac)), gscale);
    const int ti = (int)dist;
    if (ti >= RS_ZEROS * RS_GRID) continue;
    const float tf = __fsub_rn(dist, (float)ti);
    const float t0 = __ldg(table + ti), t1 = __ldg(table + ti + 1);
    const float w = __fadd_rn(t0, __fmul_rn(tf, __fsub_rn(t1, t0)));
    acc = __fadd_rn(acc, __fmul_rn(__ldg(in + idx), w));
  }
  out[n] = __fmul_rn(acc, scale);
}

}  // namespace

int resample_f32(const float* d_in, int64_t n_in, int sr_in, int sr_out, const float* d_table, float scale,
                 float gscale, int half, int64_t n_out, float* d_out, cudaStream_t stream) {
  if (n_out <= 0) return 0;
  resample_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, stream>>>(d_in, n_in, sr_in, sr_out, d_table, scale,
                                                                        gscale, half, n_out, d_out);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sw
