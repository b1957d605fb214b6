// Weight-streaming GEMM for the decoder step: out[R][N] = X[R][K] . W[N][K]^T with R = the few
// decoder rows of a step (<= 64 per row block) and W a bf16 weight matrix that is read from HBM
// exactly once. This is what ggml's mul_mat_vec / small-batch mul_mat does per token in
// whisper_decode_internal; the 128-row tcgen05 tiles of gemm_tcgen05.cu leave
// most SMs idle at these shapes (N/64 CTAs), so the step is bound by how many SMs pull weights.
//
// CTA tile 64 (rows) x 32 (weight rows) x 64 (k) per stage, 8-stage smem ring, two CTAs per SM.
// One producer thread issues two TMA tile loads per stage (128B-swizzled, rows past R / N are
// zero-filled by the TMA unit) that complete on an mbarrier, so the 8 consumer warps execute
// nothing but ldmatrix + mma.sync m16n8k16 (the math is ~1 % of the tensor peak: HBM/L2-bound by
// construction). Generation 1 issued per-thread cp.async and was instruction-latency bound at
// ~1 TB/s (profiles/r1_ncu_skinny_v1.txt); per-row 128-byte bulk copies were slower still.
// N-small matrices are split along K so that >= 148 CTAs stream weights; split partials are f32
// and are reduced by the consumer (fused LayerNorm / reduce_partials).
//
// Tried and rejected, measured with tools/dev_decode_kernels.py (profiles/r1_skinny_ab.log): (a) one
// persistent CTA per SM with a 16-stage ring: ~2x slower (half the consumer warps per SM); (b) a
// software-pipelined consumer (fragments of k-block i+1 loaded before the MMAs of k-block i, four
// accumulator chains): 8.1 vs 6.4 us on QKV, 13.0 vs 9.5 us on FC1. The launch is bound by ring bytes
// in flight x latency (96 KB per CTA, two thirds of it re-read activations), not by the consumers.
// The weight tiles of the first ring fill do not depend on the previous kernel, so the producer
// issues them before griddepcontrol.wait (only matters when SW_PDL=1).
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.cuh"

namespace sw {
namespace {

constexpr int SK_BM = 64, SK_BN = 32, SK_BK = 64, SK_STAGES = 8;
constexpr int SK_X_BYTES = SK_BM * SK_BK * 2, SK_W_BYTES = SK_BN * SK_BK * 2;
constexpr int SK_STAGE_BYTES = SK_X_BYTES + SK_W_BYTES;
constexpr int SK_SMEM = SK_STAGES * SK_STAGE_BYTES + 1024 + 2 * SK_STAGES * 8;
constexpr int SK_CONSUMERS = 8;
constexpr int SK_THREADS = (SK_CONSUMERS + 1) * 32;

__global__ void __launch_bounds__(SK_THREADS, 2)
skinny_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, int R,
                   int N, int k_slice, const float* __restrict__ bias, int gelu, bf16* __restrict__ out, int ldo,
                   float* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t sbase = smem_u32(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SK_STAGES * SK_STAGE_BYTES);
  uint64_t* empty = full + SK_STAGES;
  const int n0 = blockIdx.x * SK_BN;
  const int k_begin = blockIdx.y * k_slice;
  const int r0 = blockIdx.z * SK_BM;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_kb = k_slice / SK_BK;

  pdl_launch_dependents();
  if (tid == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < SK_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], SK_CONSUMERS);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == SK_CONSUMERS) {
    if (lane == 0) {
      // weights first (independent of the predecessor), activations once it has finished
      const int pre = n_kb < SK_STAGES ? n_kb : SK_STAGES;
      for (int kb = 0; kb < pre; ++kb) {
        mbar_arrive_expect_tx(&full[kb], SK_STAGE_BYTES);
        tma_load_2d(smem + kb * SK_STAGE_BYTES + SK_X_BYTES, &map_w, &full[kb], k_begin + kb * SK_BK, n0);
      }
      pdl_wait();
      for (int kb = 0; kb < pre; ++kb)
        tma_load_2d(smem + kb * SK_STAGE_BYTES, &map_x, &full[kb], k_begin + kb * SK_BK, r0);
      for (int kb = pre; kb < n_kb; ++kb) {
        const int s = kb % SK_STAGES;
        const uint32_t ph = (kb / SK_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], SK_STAGE_BYTES);
        uint8_t* xs = smem + s * SK_STAGE_BYTES;
        tma_load_2d(xs, &map_x, &full[s], k_begin + kb * SK_BK, r0);
        tma_load_2d(xs + SK_X_BYTES, &map_w, &full[s], k_begin + kb * SK_BK, n0);
      }
    }
    return;
  }

  // ---- consumers: warp (rw, nh) owns rows 16*rw.. of the row block and weight rows 16*nh..
  const int rw = warp & 3, nh = warp >> 2;
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  // 128B-swizzled tiles: 16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) << 4)
  const int a_row = rw * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, a_ch = lane >> 4;
  const int b_row = nh * 16 + (lane & 7) + (lane >> 4) * 8, b_ch = (lane >> 3) & 1;
  for (int kb = 0; kb < n_kb; ++kb) {
    const int s = kb % SK_STAGES;
    const uint32_t ph = (kb / SK_STAGES) & 1;
    mbar_wait(&full[s], ph);
    const uint32_t xs = sbase + s * SK_STAGE_BYTES, ws = xs + SK_X_BYTES;
    uint32_t a[SK_BK / 16][4], b[SK_BK / 16][4];
#pragma unroll
    for (int ks = 0; ks < SK_BK / 16; ++ks) {
      ldmatrix_x4(a[ks], xs + a_row * 128 + (((ks * 2 + a_ch) ^ (a_row & 7)) << 4));
      ldmatrix_x4(b[ks], ws + b_row * 128 + (((ks * 2 + b_ch) ^ (b_row & 7)) << 4));
    }
#pragma unroll
    for (int ks = 0; ks < SK_BK / 16; ++ks) {
      const uint32_t b01[2] = {b[ks][0], b[ks][1]}, b23[2] = {b[ks][2], b[ks][3]};
      mma_m16n8k16_bf16(acc[0], a[ks], b01);
      mma_m16n8k16_bf16(acc[1], a[ks], b23);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);  // every fragment has been consumed: the slot can be refilled
  }

  pdl_wait();  // (already satisfied: the activations we consumed were loaded after the producer's wait)
  const int g = lane >> 2, t4 = lane & 3;
  const int row0 = r0 + rw * 16 + g, row1 = row0 + 8;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const int col = n0 + nh * 16 + nt * 8 + 2 * t4;
    if (col >= N) continue;
    float v0 = acc[nt][0], v1 = acc[nt][1], v2 = acc[nt][2], v3 = acc[nt][3];
    if (partial) {
      float* p = partial + ((int64_t)blockIdx.y * R) * N;
      if (row0 < R) *reinterpret_cast<float2*>(p + (int64_t)row0 * N + col) = make_float2(v0, v1);
      if (row1 < R) *reinterpret_cast<float2*>(p + (int64_t)row1 * N + col) = make_float2(v2, v3);
    } else {
      if (bias) {
        const float b0 = __ldg(bias + col), b1 = __ldg(bias + col + 1);
        v0 += b0; v1 += b1; v2 += b0; v3 += b1;
      }
      if (gelu) {
        v0 = gelu_tanh(v0); v1 = gelu_tanh(v1); v2 = gelu_tanh(v2); v3 = gelu_tanh(v3);
      }
      if (row0 < R) *reinterpret_cast<uint32_t*>(out + (int64_t)row0 * ldo + col) = pack_bf16x2(v0, v1);
      if (row1 < R) *reinterpret_cast<uint32_t*>(out + (int64_t)row1 * ldo + col) = pack_bf16x2(v2, v3);
    }
  }
}

}  // namespace

int skinny_split_for(int N, int K) {
  // The most CTAs that are still ONE wave at two per SM (a second wave costs a whole CTA lifetime):
  // the largest divisor of the k-block count with n_blocks * split <= 2 * 148, slices >= 2 k-blocks.
  const int n_blocks = (N + SK_BN - 1) / SK_BN;
  const int n_kb = K / SK_BK;
  int split = 1;
  for (int s = 2; s <= 32 && s <= n_kb / 2; ++s)
    if (n_kb % s == 0 && n_blocks * s <= 2 * 148) split = s;
  return split;
}

int skinny_gemm(const bf16* X, int ldx, const bf16* W, int R, int N, int K, const float* bias, int gelu,
                bf16* out, int ldo, float* partial, int split, cudaStream_t stream) {
  if (R <= 0) return 0;
  SW_CHECK(K % SK_BK == 0 && ldx % 8 == 0 && N % 2 == 0, "skinny_gemm: unsupported shape N=%d K=%d ldx=%d", N, K, ldx);
  SW_CHECK(split >= 1 && split <= 32 && (split == 1 || partial), "skinny_gemm: split-K needs a partial buffer");
  SW_CHECK(K % (split * SK_BK) == 0, "skinny_gemm: K=%d not divisible into %d slices of 64-element blocks", K, split);
  SW_CHECK(partial || out, "skinny_gemm: null output");
  SW_CHECK((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
           "skinny_gemm: operands must be 16-byte aligned");
  static bool attr = false;
  if (!attr) {
    SW_CUDA_CHECK(cudaFuncSetAttribute(skinny_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM));
    attr = true;
  }
  const int k_slice = K / split;
  CUtensorMap map_x, map_w;
  if (make_tma_map_2d_bf16(&map_x, X, K, R, ldx, SK_BK, SK_BM)) return -1;
  if (make_tma_map_2d_bf16(&map_w, W, K, N, K, SK_BK, SK_BN)) return -1;
  dim3 grid((N + SK_BN - 1) / SK_BN, split, (R + SK_BM - 1) / SK_BM);
  SW_CUDA_CHECK(launch_pdl(skinny_gemm_kernel, grid, dim3(SK_THREADS), SK_SMEM, stream, map_x, map_w, R, N, k_slice,
                           bias, gelu, out, ldo, partial));
  return 0;
}

}  // namespace sw
