// Token-level timestamps, second half: the expand / contract of every text token's [t0, t1] on the smoothed
// signal energy (whisper_exp_compute_token_level_timestamps upstream; the first half - timestamp
// probabilities and the voice-length split - is host work in sequencer.cpp).
//
// The energy of a batch (4 bytes per PCM sample: 245 MB for 128 windows of 30 s) used to travel back to the host
// for this; it now stays in HBM and only the tokens' times cross the bus (16 bytes per token each way).
// One warp per segment. Results are bit-identical to the host loops they replace:
//   * the threshold of a token is 0.5 * (sum of the energy over [s0 - hw, s1 + hw)) / count with the sum taken
//     in index order in f32 by ONE lane (a different order would round differently); lanes take different tokens;
//   * the four threshold scans are "first sample in a direction that fails a comparison, or the bound": the warp
//     tests 32 samples per step and steps over whole 256-sample blocks whose min / max (signal_energy_kernel)
//     proves that every sample passes - any search order finds the same index;
//   * the clamps against the neighbouring tokens run in token order, as upstream (token j reads the final t1 of
//     token j - 1 and the not yet refined t0 of token j + 1).
#include "common.cuh"
#include "kernels.cuh"

namespace sw {
namespace {

constexpr int TT_SR = 16000;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int tt_ts_to_sample(long long t, int n) {
  const int s = (int)((t * TT_SR) / 100);
  return max(0, min(n - 1, s));
}
__device__ __forceinline__ long long tt_sample_to_ts(int i) { return (100ll * i) / TT_SR; }

// k moves UP (+1) or down (-1) from k0 while it has not reached `bound` (k < bound going up, k > bound going
// down) and en[k] > th (GT) or en[k] < th (!GT); returns the k the loop ends on. All 32 lanes call it with
// identical arguments and get the same result.
template <bool UP, bool GT>
__device__ int tt_scan(const float* __restrict__ en, const float* __restrict__ bmin, const float* __restrict__ bmax,
                       int k0, int bound, float th) {
  const int lane = threadIdx.x & 31;
  auto cont = [&](int i) { return GT ? en[i] > th : en[i] < th; };
  auto block_all = [&](int b) { return GT ? bmin[b] > th : bmax[b] < th; };  // every sample of block b passes
  auto at_bound = [&](int i) { return UP ? i >= bound : i <= bound; };
  int k = k0;
  while (true) {
    if (at_bound(k)) return k;
    const int b = k >> 8;
    if (block_all(b)) {
      // run over consecutive blocks that pass as a whole, 32 per step, up to the block of the bound
      const int bb = bound >> 8;
      int nb = b;
      while (true) {
        const int bi = UP ? nb + lane : nb - lane;
        const bool in_range = UP ? bi <= bb : bi >= bb;
        const bool skip = in_range && block_all(bi);
        const unsigned m = __ballot_sync(FULL, !skip);
        if (m) {
          const int l = __ffs(m) - 1;
          nb = UP ? nb + l : nb - l;
          break;
        }
        nb = UP ? nb + 32 : nb - 32;
      }
      k = UP ? min(bound, nb << 8) : max(bound, ((nb + 1) << 8) - 1);
      continue;
    }
    const int edge = UP ? min(bound, (b << 8) + 255) : max(bound, b << 8);  // last index of this block to look at
    const int idx = UP ? k + lane : k - lane;
    const bool valid = UP ? idx <= edge : idx >= edge;
    const bool stop = valid && (at_bound(idx) || !cont(idx));
    const unsigned m = __ballot_sync(FULL, stop);
    if (m) {
      const int l = __ffs(m) - 1;
      return UP ? k + l : k - l;
    }
    k = UP ? k + 32 : k - 32;
    if (UP ? k > edge : k < edge) k = UP ? edge + 1 : edge - 1;
  }
}

__global__ void __launch_bounds__(32)
token_time_refine_kernel(const float* __restrict__ energy, const float* __restrict__ blk_min,
                         const float* __restrict__ blk_max, const TtSeg* __restrict__ segs, long long* __restrict__ t0_all,
                         long long* __restrict__ t1_all, const uint8_t* __restrict__ is_text_all,
                         float* __restrict__ thold_all) {
  const TtSeg sg = segs[blockIdx.x];
  const float* en = energy + sg.en_off;
  const float* bmin = blk_min + sg.blk_off;
  const float* bmax = blk_max + sg.blk_off;
  long long* T0 = t0_all + sg.tok_off;
  long long* T1 = t1_all + sg.tok_off;
  const uint8_t* is_text = is_text_all + sg.tok_off;
  float* thold = thold_all + sg.tok_off;
  const int n = sg.n_tok, n_samples = sg.n_samples, lane = threadIdx.x;
  const int hw = TT_SR / 8;
  // ---- thresholds: one lane per token, sequential f32 sum in index order
  for (int j = lane; j < n; j += 32) {
    if (!is_text[j]) continue;
    const int s0 = tt_ts_to_sample(T0[j], n_samples), s1 = tt_ts_to_sample(T1[j], n_samples);
    const int ss0 = max(s0 - hw, 0), ss1 = min(s1 + hw, n_samples);
    float sum = 0.f;
    for (int k = ss0; k < ss1; ++k) sum = __fadd_rn(sum, en[k]);
    thold[j] = __fdiv_rn(__fmul_rn(0.5f, sum), (float)(ss1 - ss0));
  }
  __syncwarp();
  // ---- scans and clamps in token order (every lane follows the same control flow; lane 0 stores)
  for (int j = 0; j < n; ++j) {
    if (!is_text[j]) continue;
    int s0 = tt_ts_to_sample(T0[j], n_samples), s1 = tt_ts_to_sample(T1[j], n_samples);
    const float th = thold[j];
    {
      int k = s0;
      if (en[k] > th && j > 0) {
        k = tt_scan<false, true>(en, bmin, bmax, k, 0, th);
        long long t = tt_sample_to_ts(k);
        if (t < T1[j - 1]) t = T1[j - 1];
        else s0 = k;
        if (lane == 0) T0[j] = t;
      } else {
        k = tt_scan<true, false>(en, bmin, bmax, k, s1, th);
        s0 = k;
        if (lane == 0) T0[j] = tt_sample_to_ts(k);
      }
    }
    {
      int k = s1;
      if (en[k] > th) {
        k = tt_scan<true, true>(en, bmin, bmax, k, n_samples - 1, th);
        long long t = tt_sample_to_ts(k);
        if (j < n - 1 && t > T0[j + 1]) t = T0[j + 1];
        else s1 = k;
        if (lane == 0) T1[j] = t;
      } else {
        k = tt_scan<false, false>(en, bmin, bmax, k, s0, th);
        s1 = k;
        if (lane == 0) T1[j] = tt_sample_to_ts(k);
      }
    }
    __syncwarp();  // the next token reads T1[j]
  }
}

}  // namespace

int token_time_refine(const float* d_energy, const float* d_blk_min, const float* d_blk_max, const TtSeg* d_segs,
                      int n_segs, long long* d_t0, long long* d_t1, const uint8_t* d_is_text, float* d_thold,
                      cudaStream_t stream) {
  if (n_segs <= 0) return 0;
  token_time_refine_kernel<<<n_segs, 32, 0, stream>>>(d_energy, d_blk_min, d_blk_max, d_segs, d_t0, d_t1, d_is_text,
                                                      d_thold);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sw
