// Fused LayerNorm (+ split-K reduce + residual), embedding lookup and dtype conversion kernels.
// Restates ggml_norm (eps 1e-5, f32 statistics) + gamma/beta as whisper.cpp applies it in
// whisper_encode_internal / whisper_decode_internal; upstream launches
// norm, mul, add as three kernels, here it is one pass with the bf16 cast for the next GEMM.
#include "common.cuh"
#include "kernels.cuh"

#include <cuda_fp16.h>

namespace sw {
namespace {

constexpr int LN_MAX_V4 = 12;  // d <= 1536

__global__ void __launch_bounds__(128)
layer_norm_kernel(float* __restrict__ x, int rows, int d, const float* __restrict__ g,
                  const float* __restrict__ b, bf16* __restrict__ out_bf16, float* __restrict__ out_f32,
                  const float* __restrict__ partial, int n_split, int64_t split_stride,
                  const float* __restrict__ add_bias) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nv = d >> 7;  // float4 per lane
  float4 v[LN_MAX_V4];
  float4* xr = reinterpret_cast<float4*>(x + (int64_t)row * d);
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) v[i] = xr[i * 32 + lane];
  if (partial) {
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i)
      if (i < nv) {
        if (add_bias) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(add_bias) + i * 32 + lane);
          v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
        }
        for (int s = 0; s < n_split; ++s) {
          const float4 t = reinterpret_cast<const float4*>(partial + s * split_stride + (int64_t)row * d)[i * 32 + lane];
          v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
        }
        xr[i * 32 + lane] = v[i];
      }
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  sum = warp_sum(sum);
  const float mean = sum / (float)d;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      sq += (a * a + bb * bb) + (c * c + e * e);
    }
  sq = warp_sum(sq);
  const float scale = rsqrtf(sq / (float)d + 1e-5f);
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i)
    if (i < nv) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i * 32 + lane);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + i * 32 + lane);
      float4 o;
      o.x = (v[i].x - mean) * scale * gg.x + bb.x;
      o.y = (v[i].y - mean) * scale * gg.y + bb.y;
      o.z = (v[i].z - mean) * scale * gg.z + bb.z;
      o.w = (v[i].w - mean) * scale * gg.w + bb.w;
      if (out_bf16) {
        uint2 p;
        p.x = pack_bf16x2(o.x, o.y);
        p.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(out_bf16 + (int64_t)row * d)[i * 32 + lane] = p;
      }
      if (out_f32) reinterpret_cast<float4*>(out_f32 + (int64_t)row * d)[i * 32 + lane] = o;
    }
}

__global__ void zero_conv_pad_rows_kernel(bf16* buf, int d) {
  bf16* w = buf + (int64_t)blockIdx.x * (MEL_WIN_FRAMES + 2) * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    w[i] = __float2bfloat16_rn(0.f);
    w[(int64_t)(MEL_WIN_FRAMES + 1) * d + i] = __float2bfloat16_rn(0.f);
  }
}

__global__ void embed_tokens_kernel(const bf16* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                                    const int* __restrict__ tok, const int* __restrict__ pos, int d,
                                    float* __restrict__ x) {
  const int r = blockIdx.x;
  const bf16* e = tok_emb + (int64_t)tok[r] * d;
  const float* p = pos_emb + (int64_t)pos[r] * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) x[(int64_t)r * d + i] = __bfloat162float(e[i]) + p[i];
}

__global__ void f16_to_bf16_kernel(const __half* __restrict__ src, bf16* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(__half2float(src[i]));
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}

}  // namespace

int layer_norm(float* x, int rows, int d, const float* g, const float* b, bf16* out_bf16,
               float* out_f32, const float* partial, int n_split, int64_t split_stride,
               const float* add_bias, cudaStream_t stream) {
  if (rows <= 0) return 0;
  SW_CHECK(d % 128 == 0 && d <= 128 * LN_MAX_V4, "layer_norm: unsupported width %d", d);
  layer_norm_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(x, rows, d, g, b, out_bf16, out_f32, partial,
                                                        n_split, split_stride, add_bias);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int zero_conv_pad_rows(bf16* buf, int n_win, int d, cudaStream_t stream) {
  if (n_win <= 0) return 0;
  zero_conv_pad_rows_kernel<<<n_win, 256, 0, stream>>>(buf, d);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int embed_tokens(const bf16* tok_emb, const float* pos_emb, const int* d_tok, const int* d_pos,
                 int rows, int d, float* x, cudaStream_t stream) {
  if (rows <= 0) return 0;
  embed_tokens_kernel<<<rows, 128, 0, stream>>>(tok_emb, pos_emb, d_tok, d_pos, d, x);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int convert_f16_to_bf16(const uint16_t* src, bf16* dst, int64_t n, cudaStream_t stream) {
  if (n <= 0) return 0;
  f16_to_bf16_kernel<<<1184, 256, 0, stream>>>(reinterpret_cast<const __half*>(src), dst, n);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}
int convert_f32_to_bf16(const float* src, bf16* dst, int64_t n, cudaStream_t stream) {
  if (n <= 0) return 0;
  f32_to_bf16_kernel<<<1184, 256, 0, stream>>>(src, dst, n);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sw
