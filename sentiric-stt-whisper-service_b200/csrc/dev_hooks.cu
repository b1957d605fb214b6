// Device-pointer kernel hooks of the C ABI (tests / bench only).
#include "../../include/sw_whisper.h"
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.cuh"

extern "C" {

int sw_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i);
    if (major == 10) ++ok;
  }
  return ok;
}

int sw_dev_gemm_bf16(const void* dA, const void* dB, void* dC, const float* d_bias,
                     const float* d_residual, int M, int N, int K, int lda, int ldb, int ldc,
                     int flags, int block_n, void* stream) {
  sw::GemmArgs a;
  a.A = static_cast<const __nv_bfloat16*>(dA);
  a.B = static_cast<const __nv_bfloat16*>(dB);
  a.C = dC;
  a.lda = lda;
  a.ldb = ldb;
  a.ldc = ldc;
  a.bias = d_bias;
  a.residual = d_residual;
  a.ldr = ldc;
  a.M = M;
  a.N = N;
  a.K = K;
  a.flags = flags;
  a.block_n = block_n;
  return sw::gemm_bf16_tn(a, static_cast<cudaStream_t>(stream));
}

int sw_dev_skinny_gemm(const void* dX, const void* dW, int R, int N, int K, const float* d_bias, int gelu,
                       void* d_out, float* d_partial, int split, void* stream) {
  if (split <= 0) split = sw::skinny_split_for(N, K);
  return sw::skinny_gemm(static_cast<const sw::bf16*>(dX), K, static_cast<const sw::bf16*>(dW), R, N, K, d_bias,
                         gelu, static_cast<sw::bf16*>(d_out), N, d_partial, split, static_cast<cudaStream_t>(stream));
}
int sw_dev_skinny_split(int N, int K) { return sw::skinny_split_for(N, K); }
int sw_dev_layer_norm(float* d_x, int rows, int d, const float* g, const float* b, void* d_out_bf16,
                      const float* d_partial, int n_split, const float* d_bias, void* stream) {
  return sw::layer_norm(d_x, rows, d, g, b, static_cast<sw::bf16*>(d_out_bf16), nullptr, d_partial, n_split,
                        (int64_t)rows * d, d_bias, static_cast<cudaStream_t>(stream));
}
}
