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

// Decoder-step variant: one CTA per row so that a 64-row step still spreads over 64 SMs
// (the warp-per-row kernel above would use 16 CTAs). Thread t owns 8 consecutive elements.
__global__ void __launch_bounds__(192)
layer_norm_row_kernel(float* __restrict__ x, int d, const float* __restrict__ g, const float* __restrict__ b,
                      bf16* __restrict__ out_bf16, float* __restrict__ out_f32, const float* __restrict__ partial,
                      int n_split, int64_t split_stride, const float* __restrict__ add_bias) {
  __shared__ float red[8];
  const int row = blockIdx.x, t = threadIdx.x;
  const int nch = d >> 3;
  const bool act = t < nch;
  pdl_launch_dependents();
  // model constants first: they do not depend on the previous kernel of the step
  float gg[8], bb[8], ab[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) gg[i] = bb[i] = ab[i] = 0.f;
  if (act) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + t * 8)), g1 = __ldg(reinterpret_cast<const float4*>(g + t * 8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + t * 8)), b1 = __ldg(reinterpret_cast<const float4*>(b + t * 8) + 1);
    gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w; gg[4] = g1.x; gg[5] = g1.y; gg[6] = g1.z; gg[7] = g1.w;
    bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
    if (partial && add_bias) {
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(add_bias + t * 8));
      const float4 c1 = __ldg(reinterpret_cast<const float4*>(add_bias + t * 8) + 1);
      ab[0] = c0.x; ab[1] = c0.y; ab[2] = c0.z; ab[3] = c0.w; ab[4] = c1.x; ab[5] = c1.y; ab[6] = c1.z; ab[7] = c1.w;
    }
  }
  pdl_wait();
  float v[8];
  float* xr = x + (int64_t)row * d + t * 8;
  if (act) {
    const float4 a0 = reinterpret_cast<const float4*>(xr)[0], a1 = reinterpret_cast<const float4*>(xr)[1];
    v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
    if (partial) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += ab[i];
      float4 p0[8], p1[8];
#pragma unroll
      for (int s = 0; s < 8; ++s)
        if (s < n_split) {
          const float4* p = reinterpret_cast<const float4*>(partial + s * split_stride + (int64_t)row * d + t * 8);
          p0[s] = p[0];
          p1[s] = p[1];
        }
#pragma unroll
      for (int s = 0; s < 8; ++s)
        if (s < n_split) {
          v[0] += p0[s].x; v[1] += p0[s].y; v[2] += p0[s].z; v[3] += p0[s].w;
          v[4] += p1[s].x; v[5] += p1[s].y; v[6] += p1[s].z; v[7] += p1[s].w;
        }
      for (int s = 8; s < n_split; ++s) {
        const float4* p = reinterpret_cast<const float4*>(partial + s * split_stride + (int64_t)row * d + t * 8);
        const float4 q0 = p[0], q1 = p[1];
        v[0] += q0.x; v[1] += q0.y; v[2] += q0.z; v[3] += q0.w; v[4] += q1.x; v[5] += q1.y; v[6] += q1.z; v[7] += q1.w;
      }
      reinterpret_cast<float4*>(xr)[0] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(xr)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  }
  float sum = act ? ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7])) : 0.f;
  sum = warp_sum(sum);
  if ((t & 31) == 0) red[t >> 5] = sum;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
  const float mean = tot / (float)d;
  __syncthreads();
  float sq = 0.f;
  if (act) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float c = v[i] - mean;
      sq += c * c;
    }
  }
  sq = warp_sum(sq);
  if ((t & 31) == 0) red[t >> 5] = sq;
  __syncthreads();
  float tsq = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tsq += red[i];
  const float scale = rsqrtf(tsq / (float)d + 1e-5f);
  if (!act) return;
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = (v[i] - mean) * scale * gg[i] + bb[i];
  if (out_bf16) {
    uint4 p;
    p.x = pack_bf16x2(o[0], o[1]); p.y = pack_bf16x2(o[2], o[3]);
    p.z = pack_bf16x2(o[4], o[5]); p.w = pack_bf16x2(o[6], o[7]);
    reinterpret_cast<uint4*>(out_bf16 + (int64_t)row * d)[t] = p;
  }
  if (out_f32) {
    reinterpret_cast<float4*>(out_f32 + (int64_t)row * d + t * 8)[0] = make_float4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<float4*>(out_f32 + (int64_t)row * d + t * 8)[1] = make_float4(o[4], o[5], o[6], o[7]);
  }
}

// out[r][:] = bf16(bias + sum_s partial[s][r][:])  (split-K reduce of a projection that is consumed as bf16)
__global__ void __launch_bounds__(192)
reduce_partials_kernel(const float* __restrict__ partial, int n_split, int64_t split_stride, int d,
                       const float* __restrict__ bias, bf16* __restrict__ out) {
  const int row = blockIdx.x, t = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (t >= (d >> 3)) return;
  float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (bias) {
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(bias + t * 8));
    const float4 c1 = __ldg(reinterpret_cast<const float4*>(bias + t * 8) + 1);
    v[0] = c0.x; v[1] = c0.y; v[2] = c0.z; v[3] = c0.w; v[4] = c1.x; v[5] = c1.y; v[6] = c1.z; v[7] = c1.w;
  }
  for (int s = 0; s < n_split; ++s) {
    const float4* p = reinterpret_cast<const float4*>(partial + s * split_stride + (int64_t)row * d + t * 8);
    const float4 q0 = p[0], q1 = p[1];
    v[0] += q0.x; v[1] += q0.y; v[2] += q0.z; v[3] += q0.w; v[4] += q1.x; v[5] += q1.y; v[6] += q1.z; v[7] += q1.w;
  }
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  reinterpret_cast<uint4*>(out + (int64_t)row * d)[t] = o;
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
  pdl_launch_dependents();
  pdl_wait();  // x may still be read by the previous step's last kernels
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
  if (rows <= 2048) {  // decoder step: spread the few rows over as many SMs
    const int threads = ((d >> 3) + 31) / 32 * 32;
    SW_CUDA_CHECK(launch_pdl(layer_norm_row_kernel, dim3(rows), dim3(threads), 0, stream, x, d, g, b, out_bf16,
                             out_f32, partial, n_split, split_stride, add_bias));
    return 0;
  }
  layer_norm_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(x, rows, d, g, b, out_bf16, out_f32, partial,
                                                        n_split, split_stride, add_bias);
  SW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int reduce_partials(const float* partial, int n_split, int64_t split_stride, int rows, int d,
                    const float* bias, bf16* out, cudaStream_t stream) {
  if (rows <= 0) return 0;
  SW_CHECK(d % 8 == 0 && d <= 1536, "reduce_partials: unsupported width %d", d);
  SW_CUDA_CHECK(launch_pdl(reduce_partials_kernel, dim3(rows), dim3(((d >> 3) + 31) / 32 * 32), 0, stream, partial,
                           n_split, split_stride, d, bias, out));
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
