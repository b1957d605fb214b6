// Device-pointer kernel hooks of the C ABI (tests / bench only).
#include "../../include/sw_whisper.h"
#include "common.cuh"
#include "gemm.cuh"

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
}
