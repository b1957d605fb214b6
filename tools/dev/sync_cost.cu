// Development micro-benchmark: what does a dependency between two tiny phases cost on B200?
//   (a) kernel -> kernel inside a CUDA graph, (b) the same with programmatic dependent launch,
//   (c) a grid-wide barrier inside one persistent kernel (148 CTAs, one per SM).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/dev/sync_cost.cu -o gpurun_tmp/sync_cost
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void phase_kernel(float* buf, int n, int pdl) {
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = buf[(i + 4097) % n] * 0.5f + 1.0f;
}

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(352, 1) persistent_kernel(float* buf, int n, int phases, unsigned* counter, unsigned base) {
  extern __shared__ char smem[];
  for (int p = 0; p < phases; ++p) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
      buf[i] = __ldcg(buf + (i + 4097) % n) * 0.5f + 1.0f;
    grid_barrier(counter, base + (unsigned)(p + 1) * gridDim.x);
  }
}

int main() {
  const int n = 64 * 1280, phases = 400;
  float* buf; unsigned* counter;
  CK(cudaMalloc(&buf, n * sizeof(float))); CK(cudaMemset(buf, 0, n * sizeof(float)));
  CK(cudaMalloc(&counter, 4)); CK(cudaMemset(counter, 0, 4));
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int pdl = 0; pdl < 2; ++pdl) {
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int p = 0; p < phases; ++p) {
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(n / 128); cfg.blockDim = dim3(128); cfg.stream = st;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1; cfg.attrs = at; cfg.numAttrs = pdl;
      CK(cudaLaunchKernelEx(&cfg, phase_kernel, buf, n, pdl));
    }
    CK(cudaStreamEndCapture(st, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int r = 0; r < 5; ++r) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("graph of %d tiny kernels, pdl=%d: %.3f us per kernel\n", phases, pdl, ms * 1e3 / (5 * phases));
  }
  for (int smem = 0; smem <= 200 * 1024; smem += 200 * 1024) {
    CK(cudaFuncSetAttribute(persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    unsigned base = 0;
    persistent_kernel<<<148, 352, smem, st>>>(buf, n, phases, counter, base); base += phases * 148;
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int r = 0; r < 5; ++r) { persistent_kernel<<<148, 352, smem, st>>>(buf, n, phases, counter, base); base += phases * 148; }
    CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("persistent kernel, 148 CTAs x 352 thr, smem %d: %.3f us per phase+grid barrier\n", smem, ms * 1e3 / (5 * phases));
  }
  return 0;
}
