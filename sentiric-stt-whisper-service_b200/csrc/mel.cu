// PCM -> log-mel front end (whisper.cpp log_mel_spectrogram; reached from
// whisper_full_with_state, reference call site stt_engine.cpp:245). Upstream runs this on CPU
// threads even in CUDA builds; here it is two kernels:
//   1. mel_log_power_kernel : framing (reflect + zero pad), Hann, 400-point DFT, |.|^2,
//      mel filterbank (f64 accumulate like upstream), log10, per-utterance max (atomic);
//      the int16 -> f32 /32768 of transcribe_pcm16 (stt_engine.cpp:117-125) is folded into the load.
//   2. mel_finalize_*       : clamp to max-8, (x+4)/4, and the layout change to what the conv stem
//      consumes (time-major bf16 with zero pad rows).
// HBM-bound by bytes (1.92 / 2.50 MB per 30 s window). Generation 1 evaluated the 400-point DFT
// directly (201 bins x 400 samples per frame: 482 MFMA per window, 66 us per window, 0.6 % of HBM).
// Generation 2 factors it as 400 = 16 x 25 (Cooley-Tukey, n = 25 n1 + n2, k = k1 + 16 k2):
//   stage A: 25 real 16-point DFTs over n1 (only k1 = 0..8 computed, the rest by conjugate symmetry),
//            times the twiddle W400^(n2 k1);
//   stage B: for each of the 201 wanted bins a 25-point DFT over n2;
// 5.9 x fewer multiply-adds, all of it out of shared memory, f32 like upstream's own FFT. The mel
// filterbank walks only the groups of four bins where the triangle of that filter has weight
// (Model::filter_span): the partial sums are formed in upstream's order (four f32 products, then a
// double add), and a group of zero weights adds exactly 0.0, so this is bit-identical to the dense loop.
#include "common.cuh"
#include "kernels.cuh"

#include <math.h>

#include <mutex>
#include <vector>

namespace sw {
namespace {

constexpr int FR = 8;        // frames per CTA
constexpr int PW_LD = 209;   // power row stride (bank-conflict free for 8 frames)
constexpr int N1 = 16, N2 = 25;
constexpr int ZS_K1 = N2 * FR * 2 + 4;  // floats per k1 row of the stage-A output (+4: conflict-free LDS.128)

struct MelTables {
  float2* tw400 = nullptr;  // (cos, sin)(2*pi*i/400)
  float2* tw16 = nullptr;   // (cos, sin)(2*pi*i/16)
  float2* tw25 = nullptr;   // (cos, sin)(2*pi*i/25)
  float* hann = nullptr;    // periodic Hann
};
MelTables g_tables[16];
std::mutex g_tables_mu;

int get_tables(MelTables* out) {
  int dev = 0;
  SW_CUDA_CHECK(cudaGetDevice(&dev));
  SW_CHECK(dev < 16, "device ordinal %d too large", dev);
  std::lock_guard<std::mutex> lk(g_tables_mu);
  if (!g_tables[dev].tw400) {
    auto upload_tw = [](int n, float2** dst) -> int {
      std::vector<float2> tw(n);
      for (int i = 0; i < n; ++i) {
        const double th = (2.0 * M_PI * i) / n;
        tw[i] = make_float2((float)cos(th), (float)sin(th));
      }
      SW_CUDA_CHECK(cudaMalloc(dst, sizeof(float2) * n));
      SW_CUDA_CHECK(cudaMemcpy(*dst, tw.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
      return 0;
    };
    if (upload_tw(MEL_N_FFT, &g_tables[dev].tw400) || upload_tw(N1, &g_tables[dev].tw16) ||
        upload_tw(N2, &g_tables[dev].tw25))
      return -1;
    std::vector<float> hann(MEL_N_FFT);
    for (int i = 0; i < MEL_N_FFT; ++i) hann[i] = 0.5 * (1.0 - cosf((2.0 * M_PI * i) / MEL_N_FFT));
    SW_CUDA_CHECK(cudaMalloc(&g_tables[dev].hann, sizeof(float) * MEL_N_FFT));
    SW_CUDA_CHECK(cudaMemcpy(g_tables[dev].hann, hann.data(), sizeof(float) * MEL_N_FFT, cudaMemcpyHostToDevice));
  }
  *out = g_tables[dev];
  return 0;
}

__device__ __forceinline__ unsigned enc_ordered(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(unsigned e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

template <bool F32>
__device__ __forceinline__ float load_sample(const void* pcm, int64_t off, int i) {
  if (F32) return static_cast<const float*>(pcm)[off + i];
  return static_cast<float>(static_cast<const int16_t*>(pcm)[off + i]) / 32768.0f;
}

template <bool F32>
__global__ void __launch_bounds__(256)
mel_log_power_kernel(const void* __restrict__ pcm, const MelUtt* __restrict__ utts, MelTables tabs,
                     const float* __restrict__ filters, const int2* __restrict__ filter_span, int n_mel,
                     float* __restrict__ log_out, unsigned* __restrict__ max_enc) {
  __shared__ __align__(16) float xs[MEL_N_FFT * FR];  // [n][frame], windowed
  __shared__ __align__(16) float zs[N1 * ZS_K1];      // stage A: [k1][n2][frame](re, im), twiddled
  __shared__ float2 tw16[N1], tw25[N2];
  __shared__ float pw[FR * PW_LD];
  __shared__ float red[8];

  const MelUtt u = utts[blockIdx.y];
  const int f0 = blockIdx.x * FR;
  if (f0 >= u.n_active) return;
  const int tid = threadIdx.x;

  if (tid < N1) tw16[tid] = tabs.tw16[tid];
  if (tid >= 32 && tid < 32 + N2) tw25[tid - 32] = tabs.tw25[tid - 32];
  // framing: padded index p = f*160 + n; p < 200 reflects (pcm[200 - p]); else pcm[p - 200]; 0 past the end.
  // The 8 frames of the CTA cover 1520 consecutive padded samples: they are staged in shared memory once,
  // with 16-byte loads (8 int16 or 2 x float4; the span starts at a multiple of 8 samples), the int16 -> f32
  // / 32768 of transcribe_pcm16 applied in registers, and every frame is windowed out of that copy.
  {
    constexpr int SPAN = (FR - 1) * MEL_HOP + MEL_N_FFT;  // 1520
    float* raw = zs;                                      // stage A's output buffer is free until then
    const int p0 = f0 * MEL_HOP;
    if (p0 >= 200) {
      const int s0 = p0 - 200;
      for (int v = tid; v < SPAN / 8; v += 256) {
        const int sidx = s0 + v * 8;
        float x[8];
        if (sidx + 8 <= u.n_samples) {
          if (F32) {
            const float4* gp = reinterpret_cast<const float4*>(static_cast<const float*>(pcm) + u.pcm_off + sidx);
            const float4 a = __ldg(gp), b = __ldg(gp + 1);
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
          } else {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(static_cast<const int16_t*>(pcm) + u.pcm_off + sidx));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              x[2 * e] = static_cast<float>(static_cast<int16_t>(w[e] & 0xffffu)) / 32768.0f;
              x[2 * e + 1] = static_cast<float>(static_cast<int16_t>(w[e] >> 16)) / 32768.0f;
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = sidx + e < u.n_samples ? load_sample<F32>(pcm, u.pcm_off, sidx + e) : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) raw[v * 8 + e] = x[e];
      }
    } else {  // the first CTA of an utterance: reflected head
      for (int q = tid; q < SPAN; q += 256) {
        const int p = p0 + q;
        const int sidx = p < 200 ? 200 - p : p - 200;
        raw[q] = sidx < u.n_samples ? load_sample<F32>(pcm, u.pcm_off, sidx) : 0.f;
      }
    }
    __syncthreads();
    for (int i = tid; i < MEL_N_FFT * FR; i += 256) {
      const int n = i / FR, fr = i % FR;
      xs[i] = f0 + fr < u.n_active ? tabs.hann[n] * raw[fr * MEL_HOP + n] : 0.f;
    }
  }
  __syncthreads();

  // ---- stage A: thread (n2, k1 <= 8): Y[k1] = sum_n1 x[25 n1 + n2] W16^(n1 k1) for the 8 frames, then
  // Z[k1] = Y[k1] W400^(n2 k1) and, for k1 = 1..7, Z[16 - k1] = conj(Y[k1]) W400^(n2 (16 - k1))
  if (tid < N2 * 9) {
    const int n2 = tid / 9, k1 = tid % 9;
    float re[FR], im[FR];
#pragma unroll
    for (int f = 0; f < FR; ++f) re[f] = im[f] = 0.f;
    const float4* x4 = reinterpret_cast<const float4*>(xs);
#pragma unroll
    for (int n1 = 0; n1 < N1; ++n1) {
      const float2 w = tw16[(n1 * k1) & (N1 - 1)];
      const int n = N2 * n1 + n2;
      const float4 a = x4[2 * n], b = x4[2 * n + 1];
      re[0] += a.x * w.x; im[0] -= a.x * w.y;
      re[1] += a.y * w.x; im[1] -= a.y * w.y;
      re[2] += a.z * w.x; im[2] -= a.z * w.y;
      re[3] += a.w * w.x; im[3] -= a.w * w.y;
      re[4] += b.x * w.x; im[4] -= b.x * w.y;
      re[5] += b.y * w.x; im[5] -= b.y * w.y;
      re[6] += b.z * w.x; im[6] -= b.z * w.y;
      re[7] += b.w * w.x; im[7] -= b.w * w.y;
    }
    {
      const float2 t = __ldg(tabs.tw400 + n2 * k1);  // W400^(n2 k1) = cos - i sin
      float* z = zs + k1 * ZS_K1 + n2 * (2 * FR);
#pragma unroll
      for (int f = 0; f < FR; f += 2)
        *reinterpret_cast<float4*>(z + 2 * f) =
            make_float4(re[f] * t.x + im[f] * t.y, im[f] * t.x - re[f] * t.y,
                        re[f + 1] * t.x + im[f + 1] * t.y, im[f + 1] * t.x - re[f + 1] * t.y);
    }
    if (k1 >= 1 && k1 <= 7) {
      const int kk = N1 - k1;
      const float2 t = __ldg(tabs.tw400 + n2 * kk);
      float* z = zs + kk * ZS_K1 + n2 * (2 * FR);
#pragma unroll
      for (int f = 0; f < FR; f += 2)  // conj(Y) = (re, -im)
        *reinterpret_cast<float4*>(z + 2 * f) =
            make_float4(re[f] * t.x - im[f] * t.y, -im[f] * t.x - re[f] * t.y,
                        re[f + 1] * t.x - im[f + 1] * t.y, -im[f + 1] * t.x - re[f + 1] * t.y);
    }
  }
  __syncthreads();

  // ---- stage B: thread k = k1 + 16 k2: X[k] = sum_n2 Z[k1][n2] W25^(n2 k2); power spectrum
  if (tid < MEL_N_BINS) {
    const int k1 = tid & (N1 - 1), k2 = tid >> 4;
    float re[FR], im[FR];
#pragma unroll
    for (int f = 0; f < FR; ++f) re[f] = im[f] = 0.f;
    const float4* z4 = reinterpret_cast<const float4*>(zs + k1 * ZS_K1);
    int idx = 0;
#pragma unroll 5
    for (int n2 = 0; n2 < N2; ++n2) {
      const float2 w = tw25[idx];  // multiply by cos - i sin
#pragma unroll
      for (int q = 0; q < FR / 2; ++q) {
        const float4 z = z4[n2 * (FR / 2) + q];  // (re, im) of frames 2q, 2q+1
        re[2 * q] += z.x * w.x + z.y * w.y;
        im[2 * q] += z.y * w.x - z.x * w.y;
        re[2 * q + 1] += z.z * w.x + z.w * w.y;
        im[2 * q + 1] += z.w * w.x - z.z * w.y;
      }
      idx += k2;
      if (idx >= N2) idx -= N2;
    }
#pragma unroll
    for (int f = 0; f < FR; ++f) pw[f * PW_LD + tid] = re[f] * re[f] + im[f] * im[f];
  }
  __syncthreads();

  float lmax = -1e30f;
  for (int o = tid; o < n_mel * FR; o += 256) {
    const int fr = o % FR, m = o / FR;
    if (f0 + fr >= u.n_active) continue;
    const float* fl = filters + (size_t)m * MEL_N_BINS;
    const float* p = pw + fr * PW_LD;
    const int2 span = __ldg(filter_span + m);  // groups of four bins outside it hold zero weights: they add 0.0
    double sum = 0.0;
    int k = span.x;
    const int k_end = span.y < MEL_N_BINS - 3 ? span.y : MEL_N_BINS - 3;
    for (; k < k_end; k += 4) {
      // upstream sums four float products in float, then adds to the double accumulator
      float g = __fmul_rn(p[k], __ldg(fl + k));
      g = __fadd_rn(g, __fmul_rn(p[k + 1], __ldg(fl + k + 1)));
      g = __fadd_rn(g, __fmul_rn(p[k + 2], __ldg(fl + k + 2)));
      g = __fadd_rn(g, __fmul_rn(p[k + 3], __ldg(fl + k + 3)));
      sum += (double)g;
    }
    if (span.y > MEL_N_BINS - 1)  // the scalar tail of upstream's loop: bin 200
      for (k = MEL_N_BINS - 1; k < MEL_N_BINS; ++k) sum += (double)__fmul_rn(p[k], __ldg(fl + k));
    const float v = (float)log10(fmax(sum, 1e-10));
    log_out[u.log_off + (int64_t)m * u.n_active + f0 + fr] = v;
    lmax = fmaxf(lmax, v);
  }
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) red[tid >> 5] = lmax;
  __syncthreads();
  if (tid == 0) {
    float m = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    if (m > -1e29f) atomicMax(max_enc + blockIdx.y, enc_ordered(m));
  }
}

__device__ __forceinline__ float mel_norm(float x, float mx) {
  const double mm = (double)mx - 8.0;
  if ((double)x < mm) x = (float)mm;
  return (float)(((double)x + 4.0) / 4.0);
}

// value of absolute frame f, mel bin m of utterance u after normalisation
__device__ __forceinline__ float mel_value(const float* log_in, const MelUtt& u, float mx, int m, int f) {
  if (f >= u.n_len) return 0.f;  // beyond the spectrogram: the encoder input is zero-filled
  const float raw = f < u.n_active ? log_in[u.log_off + (int64_t)m * u.n_active + f] : -10.0f;
  return mel_norm(raw, mx);
}

__global__ void __launch_bounds__(256)
mel_finalize_windows_kernel(const float* __restrict__ log_in, const MelUtt* __restrict__ utts,
                            const unsigned* __restrict__ max_enc, const int* __restrict__ win_utt,
                            const int* __restrict__ win_seek, int n_mel, bf16* __restrict__ out_bf16,
                            float* __restrict__ out_f32) {
  __shared__ float tile[128][33];
  const int w = blockIdx.y;
  const int ui = win_utt[w];
  const MelUtt u = utts[ui];
  const float mx = dec_ordered(max_enc[ui]);
  const int seek = win_seek[w];
  const int t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int m = ty; m < n_mel; m += 8) {
    const int t = t0 + tx;
    float v = 0.f;
    if (t < MEL_WIN_FRAMES) {
      v = mel_value(log_in, u, mx, m, seek + t);
      if (out_f32) out_f32[((int64_t)w * n_mel + m) * MEL_WIN_FRAMES + t] = v;
    }
    tile[m][tx] = v;
  }
  __syncthreads();
  bf16* ob = out_bf16 + (int64_t)w * (MEL_WIN_FRAMES + 2) * n_mel;
  for (int t = ty; t < 32; t += 8) {
    if (t0 + t >= MEL_WIN_FRAMES) break;
    for (int m = tx; m < n_mel; m += 32)
      ob[(int64_t)(1 + t0 + t) * n_mel + m] = __float2bfloat16_rn(tile[m][t]);
  }
  if (blockIdx.x == 0)
    for (int m = threadIdx.x; m < n_mel; m += 256) {
      ob[m] = __float2bfloat16_rn(0.f);
      ob[(int64_t)(MEL_WIN_FRAMES + 1) * n_mel + m] = __float2bfloat16_rn(0.f);
    }
}

__global__ void mel_finalize_full_kernel(const float* __restrict__ log_in, const MelUtt* __restrict__ utts,
                                         const unsigned* __restrict__ max_enc, int utt, int n_mel,
                                         int n_len, float* __restrict__ out) {
  const MelUtt u = utts[utt];
  const float mx = dec_ordered(max_enc[utt]);
  const int64_t total = (int64_t)n_mel * n_len;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / n_len), f = (int)(i % n_len);
    out[i] = mel_value(log_in, u, mx, m, f);
  }
}

__global__ void __launch_bounds__(256)
mel_f32_to_conv_input_kernel(const float* __restrict__ mel, int n_mel, bf16* __restrict__ out) {
  __shared__ float tile[128][33];
  const int w = blockIdx.y;
  const int t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int m = ty; m < n_mel; m += 8) {
    const int t = t0 + tx;
    tile[m][tx] = t < MEL_WIN_FRAMES ? mel[((int64_t)w * n_mel + m) * MEL_WIN_FRAMES + t] : 0.f;
  }
  __syncthreads();
  bf16* ob = out + (int64_t)w * (MEL_WIN_FRAMES + 2) * n_mel;
  for (int t = ty; t < 32; t += 8) {
    if (t0 + t >= MEL_WIN_FRAMES) break;
    for (int m = tx; m < n_mel; m += 32)
      ob[(int64_t)(1 + t0 + t) * n_mel + m] = __float2bfloat16_rn(tile[m][t]);
  }
  if (blockIdx.x == 0)
    for (int m = threadIdx.x; m < n_mel; m += 256) {
      ob[m] = __float2bfloat16_rn(0.f);
      ob[(int64_t)(MEL_WIN_FRAMES + 1) * n_mel + m] = __float2bfloat16_rn(0.f);
    }
}

template <bool F32>
__global__ void __launch_bounds__(256)
signal_energy_kernel(const void* __restrict__ pcm, const MelUtt* __restrict__ utts, int hw,
                     float* __restrict__ out, float* __restrict__ blk_min, float* __restrict__ blk_max) {
  extern __shared__ float sh[];  // 256 + 2*hw, |x|
  __shared__ float rmin[8], rmax[8];
  const MelUtt u = utts[blockIdx.y];
  const int n = u.n_samples;
  if ((int)blockIdx.x * 256 >= n) return;
  const int base = blockIdx.x * 256 - hw;
  for (int i = threadIdx.x; i < 256 + 2 * hw; i += 256) {
    const int s = base + i;
    sh[i] = (s >= 0 && s < n) ? fabsf(load_sample<F32>(pcm, u.pcm_off, s)) : 0.f;
  }
  __syncthreads();
  const int i = blockIdx.x * 256 + threadIdx.x;
  float v = 0.f;
  if (i < n) {
    float sum = 0.f;
    for (int j = -hw; j <= hw; ++j)
      if (i + j >= 0 && i + j < n) sum = __fadd_rn(sum, sh[threadIdx.x + hw + j]);
    v = sum / (float)(2 * hw + 1);
    out[u.pcm_off + i] = v;
  }
  // min / max of this block's valid samples
  float lo = i < n ? v : INFINITY, hi = i < n ? v : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    rmin[threadIdx.x >> 5] = lo;
    rmax[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      lo = fminf(lo, rmin[w]);
      hi = fmaxf(hi, rmax[w]);
    }
    blk_min[u.blk_off + blockIdx.x] = lo;
    blk_max[u.blk_off + blockIdx.x] = hi;
  }
}

}  // namespace

int mel_log_power(const void* pcm, int is_f32, const MelUtt* d_utts, int n_utts, int max_active,
                  const float* d_filters, const int2* d_filter_span, int n_mel, float* d_log,
                  unsigned* d_max_enc, cudaStream_t stream) {
  if (n_utts <= 0 || max_active <= 0) return 0;
  SW_CHECK(n_mel <= 128, "n_mel %d > 128", n_mel);
  MelTables t;
  if (get_tables(&t)) return -1;
  dim3 grid((max_active + FR - 1) / FR, n_utts);
  if (is_f32)
    mel_log_power_kernel<true><<<grid, 256, 0, stream>>>(pcm, d_utts, t, d_filters, d_filter_span, n_mel, d_log, d_max_enc);
  else
    mel_log_power_kernel<false><<<grid, 256, 0, stream>>>(pcm, d_utts, t, d_filters, d_filter_span, n_mel, d_log, d_max_enc);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int mel_finalize_windows(const float* d_log, const MelUtt* d_utts, const unsigned* d_max_enc,
                         const int* d_win_utt, const int* d_win_seek, int n_win, int n_mel,
                         bf16* out_bf16, float* out_f32, cudaStream_t stream) {
  if (n_win <= 0) return 0;
  dim3 grid((MEL_WIN_FRAMES + 31) / 32, n_win);
  mel_finalize_windows_kernel<<<grid, 256, 0, stream>>>(d_log, d_utts, d_max_enc, d_win_utt, d_win_seek,
                                                        n_mel, out_bf16, out_f32);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int mel_finalize_full(const float* d_log, const MelUtt* d_utts, const unsigned* d_max_enc, int utt,
                      int n_mel, int n_len, int n_active, float* out, cudaStream_t stream) {
  (void)n_active;
  mel_finalize_full_kernel<<<296, 256, 0, stream>>>(d_log, d_utts, d_max_enc, utt, n_mel, n_len, out);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int mel_f32_to_conv_input(const float* d_mel, int n_win, int n_mel, bf16* out, cudaStream_t stream) {
  dim3 grid((MEL_WIN_FRAMES + 31) / 32, n_win);
  mel_f32_to_conv_input_kernel<<<grid, 256, 0, stream>>>(d_mel, n_mel, out);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int signal_energy(const void* pcm, int is_f32, const MelUtt* d_utts, int n_utts, int max_n, int hw,
                  float* out, float* blk_min, float* blk_max, cudaStream_t stream) {
  if (n_utts <= 0 || max_n <= 0) return 0;
  dim3 grid((max_n + 255) / 256, n_utts);
  const size_t sh = (256 + 2 * hw) * sizeof(float);
  if (is_f32)
    signal_energy_kernel<true><<<grid, 256, sh, stream>>>(pcm, d_utts, hw, out, blk_min, blk_max);
  else
    signal_energy_kernel<false><<<grid, 256, sh, stream>>>(pcm, d_utts, hw, out, blk_min, blk_max);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sw
