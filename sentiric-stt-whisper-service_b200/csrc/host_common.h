// Host-side error/log helpers shared by .cpp and .cu translation units.
#pragma once
#include <cuda_runtime.h>

namespace sw {

void set_last_error(const char* fmt, ...);
void log_msg(int level, const char* fmt, ...);  // level: 2 info, 3 warn, 4 error

#define SW_CUDA_CHECK(expr)                                                       \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      ::sw::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,          \
                           cudaGetErrorString(_e));                               \
      return -1;                                                                  \
    }                                                                             \
  } while (0)

#define SW_CHECK(cond, ...)                                                       \
  do {                                                                            \
    if (!(cond)) {                                                                \
      ::sw::set_last_error(__VA_ARGS__);                                          \
      return -1;                                                                  \
    }                                                                             \
  } while (0)

// ---- per-device launch state -------------------------------------------------------------------
// cudaFuncSetAttribute and the SM count belong to a DEVICE, not to the process: a context on GPU 1 created
// after one on GPU 0 needs its own opt-in to > 48 KB of dynamic shared memory. One SmemOptIn per kernel
// (a function-local static at the launch site) remembers, per ordinal and under a mutex, the largest size
// already granted; lanes and contexts of several devices may launch from different host threads.
constexpr int SW_MAX_DEVICES = 64;

struct SmemOptIn {
  int granted[SW_MAX_DEVICES] = {0};
  unsigned lock = 0;  // spin flag (no <mutex> in device translation units' hot includes); held for nanoseconds
  template <typename Kernel>
  cudaError_t ensure(Kernel kernel, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= SW_MAX_DEVICES) return cudaErrorInvalidDevice;
    if (__atomic_load_n(&granted[dev], __ATOMIC_ACQUIRE) >= bytes) return cudaSuccess;
    while (__atomic_exchange_n(&lock, 1u, __ATOMIC_ACQUIRE)) {
    }
    if (granted[dev] < bytes) {
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (e == cudaSuccess) __atomic_store_n(&granted[dev], bytes, __ATOMIC_RELEASE);
    }
    __atomic_store_n(&lock, 0u, __ATOMIC_RELEASE);
    return e;
  }
};

// SM count of the calling thread's current device (cached per ordinal)
inline int device_sm_count() {
  static int cache[SW_MAX_DEVICES] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= SW_MAX_DEVICES) return 148;
  int n = __atomic_load_n(&cache[dev], __ATOMIC_RELAXED);
  if (n > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  __atomic_store_n(&cache[dev], n, __ATOMIC_RELAXED);
  return n;
}

}  // namespace sw
