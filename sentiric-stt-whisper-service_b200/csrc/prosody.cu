// Per-segment prosody DSP of the service (extract_prosody, /root/reference/src/prosody_extractor.cpp:31-224,
// called once per segment at stt_engine.cpp:313-334) for all segments of an utterance in two launches.
// SURVEY.md §8(f) rank 3. The reference walks every sample of every segment on the host, carrying a
// one-pole low-pass state through the whole segment; here
//   1. prosody_frames_kernel: one thread per 10 ms frame. The low-pass state at the start of a frame is
//      rebuilt by running the recurrence over the `warm` samples before it (state error decays as
//      (1 - alpha)^warm: 48 / alpha samples leave < 1e-20 of it, i.e. the same float), then the frame is
//      processed exactly as the reference does: energy, zero crossings and hysteresis cycle count of the
//      filtered frame, first-difference centroid. 16 bytes out per frame.
//   2. prosody_reduce_kernel: one warp per segment; five lanes each walk the frames IN ORDER for one
//      statistic (the reference's means and deviations are sequential float sums, so the order is kept and
//      the results are bit-identical); the median pitch comes from a histogram of cycle counts.
// All arithmetic uses __fmul_rn / __fadd_rn so that nvcc cannot contract into FMAs the host code does not
// have. HBM traffic: the PCM of the utterance once (L1/L2 serve the re-reads of the warm-up).
#include "common.cuh"
#include "kernels.cuh"

namespace sw {
namespace {

template <bool F32>
__device__ __forceinline__ float ps_load(const void* pcm, int64_t i) {
  if (F32) return __ldg(static_cast<const float*>(pcm) + i);
  return static_cast<float>(__ldg(static_cast<const int16_t*>(pcm) + i)) / 32768.0f;  // stt_engine.cpp:117-125
}

template <bool F32>
__global__ void __launch_bounds__(128)
prosody_frames_kernel(const void* __restrict__ pcm, const ProsodySeg* __restrict__ segs, int shift, int warm,
                      float alpha, ProsodyFrame* __restrict__ frames) {
  const ProsodySeg sg = segs[blockIdx.y];
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= sg.n_frames) return;
  const int fs = shift < 1600 ? shift : 1600;
  const int64_t base = sg.begin + (int64_t)f * shift;
  // low-pass state at the start of the frame (prosody_extractor.cpp:58-59, 72-74)
  float lpf = 0.0f;
  {
    int64_t ws = base - warm;
    if (ws < sg.begin) ws = sg.begin;
    for (int64_t i = ws; i < base; ++i) {
      const float x = ps_load<F32>(pcm, i);
      lpf = __fadd_rn(lpf, __fmul_rn(alpha, __fsub_rn(x, lpf)));
    }
  }
  float r0 = 0.0f, power = 0.0f, weighted = 0.0f, prev = 0.0f;
  for (int k = 0; k < fs; ++k) {  // :68-76 energy, :119-125 first-difference centroid
    const float x = ps_load<F32>(pcm, base + k);
    r0 = __fadd_rn(r0, __fmul_rn(x, x));
    if (k > 0) {
      const float dd = fabsf(__fsub_rn(x, prev));
      weighted = __fadd_rn(weighted, __fmul_rn(dd, (float)k));
      power = __fadd_rn(power, dd);
    }
    prev = x;
  }
  const float rms = sqrtf(__fdiv_rn(r0, (float)fs));
  const float clip = fmaxf(0.002f, __fmul_rn(rms, 0.15f));  // :84
  int cycles = 0, zc = 0;
  bool positive = false, started = false;
  float vprev = 0.0f;
  for (int k = 0; k < fs; ++k) {  // :90-108 on the filtered frame
    const float x = ps_load<F32>(pcm, base + k);
    lpf = __fadd_rn(lpf, __fmul_rn(alpha, __fsub_rn(x, lpf)));
    const float v = lpf;
    if (k > 0) {
      if ((v >= 0) != (vprev >= 0)) ++zc;
      if (!started) {
        if (v > clip) positive = true, started = true;
        else if (v < -clip) positive = false, started = true;
      } else if (positive && v < -clip) {
        positive = false;
        ++cycles;
      } else if (!positive && v > clip) {
        positive = true;
      }
    }
    vprev = v;
  }
  ProsodyFrame o;
  o.rms = rms;
  o.zc = zc;
  o.cycles = cycles;
  o.sc = power > 0 ? __fdiv_rn(weighted, power) : 0.0f;
  frames[sg.frame_off + f] = o;
}

// One warp per segment. The frame records are staged through shared memory in chunks (coalesced loads by
// all 32 lanes); five lanes then walk a chunk in frame order, one statistic each - the sums stay sequential
// single-precision sums like the reference's, but the walk reads shared memory instead of paying an L2
// round trip per frame (generation 1: 1.3 ms for a 30 s segment, profiles/r1_ncu_prosody_mel_v3.txt).
constexpr int PR_CHUNK = 512;

__global__ void __launch_bounds__(32)
prosody_reduce_kernel(const ProsodySeg* __restrict__ segs, const ProsodyFrame* __restrict__ frames, int shift,
                      int sample_rate, float min_pitch, float max_pitch, ProsodyRaw* __restrict__ out) {
  __shared__ ProsodyFrame buf[PR_CHUNK];
  __shared__ int hist[1601];
  __shared__ float means[2];  // energy mean, pitch mean (pass 2 needs them)
  const ProsodySeg sg = segs[blockIdx.x];
  const ProsodyFrame* fr = frames + sg.frame_off;
  const int n = sg.n_frames, lane = threadIdx.x;
  const int fs = shift < 1600 ? shift : 1600;
  const float dur = __fdiv_rn((float)shift, (float)sample_rate);
  ProsodyRaw& o = out[blockIdx.x];
  for (int c = lane; c <= fs; c += 32) hist[c] = 0;
  __syncwarp();
  // ---- pass 1: sums
  float s = 0.0f, last = 0.0f;  // lanes 0, 1, 2, 4: their running sum; lane 3: previous rms
  int cnt = 0;                  // lane 3: onsets; lane 4: pitch candidates
  for (int base = 0; base < n; base += PR_CHUNK) {
    const int m = n - base < PR_CHUNK ? n - base : PR_CHUNK;
    for (int i = lane; i < m; i += 32) buf[i] = fr[base + i];
    __syncwarp();
    if (lane == 0) {  // :130 energy mean
      for (int i = 0; i < m; ++i) s = __fadd_rn(s, buf[i].rms);
    } else if (lane == 1) {  // :133 zero-crossing rate
      for (int i = 0; i < m; ++i) s = __fadd_rn(s, __fdiv_rn((float)buf[i].zc, (float)fs));
    } else if (lane == 2) {  // :132 centroid
      for (int i = 0; i < m; ++i) s = __fadd_rn(s, buf[i].sc);
    } else if (lane == 3) {  // :79-82 onsets
      for (int i = 0; i < m; ++i) {
        const float r = buf[i].rms;
        if (r > 0.05f && last <= 0.05f) ++cnt;
        last = r;
      }
    } else if (lane == 4) {  // :112-117 pitch candidates
      for (int i = 0; i < m; ++i) {
        const int c = buf[i].cycles;
        if (buf[i].rms > 0.015f && c > 0) {
          const float f0 = __fdiv_rn((float)c, dur);
          if (f0 >= min_pitch && f0 <= max_pitch) {
            ++cnt;
            s = __fadd_rn(s, f0);
            ++hist[c];
          }
        }
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    const float mean = n ? __fdiv_rn(s, (float)n) : 0.01f;
    means[0] = mean;
    o.energy_mean = mean;
    o.n_frames = n;
  } else if (lane == 1) {
    o.zcr_mean = n ? __fdiv_rn(s, (float)n) : 0.1f;
  } else if (lane == 2) {
    o.sc_mean = n ? __fdiv_rn(s, (float)n) : 50.0f;
  } else if (lane == 3) {
    o.peaks = cnt;
  } else if (lane == 4) {
    means[1] = cnt ? __fdiv_rn(s, (float)cnt) : 0.0f;
    o.n_f0 = cnt;
  }
  __syncwarp();
  // ---- pass 2: deviations (:129, :131)
  float acc = 0.0f;
  for (int base = 0; base < n; base += PR_CHUNK) {
    const int m = n - base < PR_CHUNK ? n - base : PR_CHUNK;
    for (int i = lane; i < m; i += 32) buf[i] = fr[base + i];
    __syncwarp();
    if (lane == 0) {
      const float mean = means[0];
      for (int i = 0; i < m; ++i) {
        const float d = __fsub_rn(buf[i].rms, mean);
        acc = __fadd_rn(acc, __fmul_rn(d, d));
      }
    } else if (lane == 4 && cnt) {
      const float mean = means[1];
      for (int i = 0; i < m; ++i) {
        const int c = buf[i].cycles;
        if (buf[i].rms > 0.015f && c > 0) {
          const float f0 = __fdiv_rn((float)c, dur);
          if (f0 >= min_pitch && f0 <= max_pitch) {
            const float d = __fsub_rn(f0, mean);
            acc = __fadd_rn(acc, __fmul_rn(d, d));
          }
        }
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    o.energy_std = n ? sqrtf(__fdiv_rn(acc, (float)n)) : 0.0f;
  } else if (lane == 4) {
    float med = 0.0f, sd = 0.0f;
    if (cnt) {
      sd = sqrtf(__fdiv_rn(acc, (float)cnt));
      int run = 0;  // element cnt/2 of the sorted candidates (std::nth_element in the reference)
      for (int c = 0; c <= fs; ++c) {
        run += hist[c];
        if (run > cnt / 2) {
          med = __fdiv_rn((float)c, dur);
          break;
        }
      }
    }
    o.pitch_median = med;
    o.pitch_std = sd;
  }
}

}  // namespace

int prosody_frames(const void* d_pcm, int is_f32, const ProsodySeg* d_segs, int n_segs, int max_frames, int shift,
                   int warm, float alpha, ProsodyFrame* d_frames, cudaStream_t stream) {
  if (n_segs <= 0 || max_frames <= 0) return 0;
  dim3 grid((max_frames + 127) / 128, n_segs);
  if (is_f32)
    prosody_frames_kernel<true><<<grid, 128, 0, stream>>>(d_pcm, d_segs, shift, warm, alpha, d_frames);
  else
    prosody_frames_kernel<false><<<grid, 128, 0, stream>>>(d_pcm, d_segs, shift, warm, alpha, d_frames);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int prosody_reduce(const ProsodySeg* d_segs, const ProsodyFrame* d_frames, int n_segs, int shift, int sample_rate,
                   float min_pitch, float max_pitch, ProsodyRaw* d_out, cudaStream_t stream) {
  if (n_segs <= 0) return 0;
  SW_CHECK(shift >= 1, "prosody: sample rate %d too low", sample_rate);
  prosody_reduce_kernel<<<n_segs, 32, 0, stream>>>(d_segs, d_frames, shift, sample_rate, min_pitch, max_pitch, d_out);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sw
