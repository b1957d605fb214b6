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

}  // namespace sw
