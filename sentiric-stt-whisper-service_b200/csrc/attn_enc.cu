// Encoder self-attention: non-causal softmax(Q K^T / sqrt(64)) V over the 1500 positions of a
// window, one (window, head, 128-query tile) per CTA, flash-style (S and P never leave the SM).
// Replaces ggml's flash_attn_ext as whisper_encode_internal uses it.
// Generation 1: legacy mma.sync m16n8k16 bf16 tensor path with cp.async double buffering;
// the tcgen05/TMEM version is the planned replacement.
#include "common.cuh"
#include "kernels.cuh"

namespace sw {
namespace {

constexpr int BQ = 128, BKV = 64, DH = 64;
constexpr int ATT_THREADS = 256;
constexpr int ATT_SMEM = BQ * DH * 2 + 2 * 2 * BKV * DH * 2;  // 48 KB

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// smem tile of rows x 64 bf16, 16-byte chunks XOR-swizzled by (row & 7)
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
  return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
encoder_attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int T, int d) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK0 = sQ + BQ * DH * 2;
  const uint32_t sV0 = sK0 + 2 * BKV * DH * 2;

  const int qt = blockIdx.x, h = blockIdx.y, w = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int64_t ld = 3 * (int64_t)d;
  const bf16* base = qkv + (int64_t)w * T * ld + h * DH;
  const int q0 = qt * BQ;
  const int n_tiles = (T + BKV - 1) / BKV;

  // ---- async loads
  {
#pragma unroll
    for (int i = 0; i < (BQ * 8) / ATT_THREADS; ++i) {
      const int c = tid + i * ATT_THREADS;
      const int row = c >> 3, ch = c & 7;
      const bool ok = q0 + row < T;
      const bf16* src = base + (int64_t)(ok ? q0 + row : 0) * ld + ch * 8;
      cp_async_16(tile_addr(sQ, row, ch), src, ok);
    }
  }
  auto load_kv = [&](int j, int buf) {
    const int k0 = j * BKV;
#pragma unroll
    for (int i = 0; i < (BKV * 8) / ATT_THREADS; ++i) {
      const int c = tid + i * ATT_THREADS;
      const int row = c >> 3, ch = c & 7;
      const bool ok = k0 + row < T;
      const bf16* src = base + (int64_t)(ok ? k0 + row : 0) * ld + ch * 8;
      cp_async_16(tile_addr(sK0 + buf * BKV * DH * 2, row, ch), src + d, ok);
      cp_async_16(tile_addr(sV0 + buf * BKV * DH * 2, row, ch), src + 2 * d, ok);
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float sc = 0.125f * 1.4426950408889634f;

  for (int j = 0; j < n_tiles; ++j) {
    const int buf = j & 1;
    if (j + 1 < n_tiles) {
      load_kv(j + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = ks * 2 + (lane >> 4);
        ldmatrix_x4(qf[ks], tile_addr(sQ, row, ch));
      }
    }
    const uint32_t sK = sK0 + buf * BKV * DH * 2, sV = sV0 + buf * BKV * DH * 2;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        const int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int ch = ks * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(b, tile_addr(sK, key, ch));
        const uint32_t b01[2] = {b[0], b[1]}, b23[2] = {b[2], b[3]};
        mma_m16n8k16_bf16(s[2 * np], qf[ks], b01);
        mma_m16n8k16_bf16(s[2 * np + 1], qf[ks], b23);
      }
    }
    if (j == n_tiles - 1) {
      const int k0 = j * BKV;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = k0 + nt * 8 + 2 * t4;
        if (key >= T) s[nt][0] = s[nt][2] = -INFINITY;
        if (key + 1 >= T) s[nt][1] = s[nt][3] = -INFINITY;
      }
    }
    float r0 = -INFINITY, r1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      r0 = fmaxf(r0, fmaxf(s[nt][0], s[nt][1]));
      r1 = fmaxf(r1, fmaxf(s[nt][2], s[nt][3]));
    }
    r0 = fmaxf(r0, __shfl_xor_sync(0xffffffffu, r0, 1));
    r0 = fmaxf(r0, __shfl_xor_sync(0xffffffffu, r0, 2));
    r1 = fmaxf(r1, __shfl_xor_sync(0xffffffffu, r1, 1));
    r1 = fmaxf(r1, __shfl_xor_sync(0xffffffffu, r1, 2));
    const float mn0 = fmaxf(m0, r0), mn1 = fmaxf(m1, r1);
    const float a0 = fast_exp2((m0 - mn0) * sc), a1 = fast_exp2((m1 - mn1) * sc);
    m0 = mn0;
    m1 = mn1;
    l0 *= a0;
    l1 *= a1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i][0] *= a0;
      o[i][1] *= a0;
      o[i][2] *= a1;
      o[i][3] *= a1;
    }
    const float ms0 = m0 * sc, ms1 = m1 * sc;
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = fast_exp2(s[nt][0] * sc - ms0), p1 = fast_exp2(s[nt][1] * sc - ms0);
      const float p2 = fast_exp2(s[nt][2] * sc - ms1), p3 = fast_exp2(s[nt][3] * sc - ms1);
      l0 += p0 + p1;
      l1 += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t b[4];
        const int key = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = dp * 2 + (lane >> 4);
        ldmatrix_x4_trans(b, tile_addr(sV, key, ch));
        const uint32_t b01[2] = {b[0], b[1]}, b23[2] = {b[2], b[3]};
        mma_m16n8k16_bf16(o[2 * dp], pf[kk], b01);
        mma_m16n8k16_bf16(o[2 * dp + 1], pf[kk], b23);
      }
    }
    __syncthreads();
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;
  bf16* ob = out + (int64_t)w * T * d + h * DH;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = nt * 8 + 2 * t4;
    if (row0 < T)
      *reinterpret_cast<uint32_t*>(ob + (int64_t)row0 * d + col) = pack_bf16x2(o[nt][0] * i0, o[nt][1] * i0);
    if (row1 < T)
      *reinterpret_cast<uint32_t*>(ob + (int64_t)row1 * d + col) = pack_bf16x2(o[nt][2] * i1, o[nt][3] * i1);
  }
}

}  // namespace

int encoder_attention(const bf16* qkv, bf16* out, int n_win, int T, int d, int n_head,
                      cudaStream_t stream) {
  if (n_win <= 0) return 0;
  SW_CHECK(d == n_head * DH, "encoder_attention: head dim must be 64 (d=%d heads=%d)", d, n_head);
  static SmemOptIn opt_in;  // per device (host_common.h)
  SW_CUDA_CHECK(opt_in.ensure(encoder_attention_kernel, ATT_SMEM));
  dim3 grid((T + BQ - 1) / BQ, n_head, n_win);
  encoder_attention_kernel<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(qkv, out, T, d);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sw
