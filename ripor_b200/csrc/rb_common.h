// Shared helpers for the riporb200 C-ABI library: error reporting and CUDA call checking.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/riporb200.h"

namespace rb {

inline std::string& last_error_slot() {
  static thread_local std::string s;
  return s;
}

inline int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_slot() = buf;
  return code;
}

#define RB_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return rb::fail(RB200_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,         \
                      cudaGetErrorString(_e));                                                     \
  } while (0)

#define RB_TRY(expr)             \
  do {                           \
    int _s = (expr);             \
    if (_s != 0) return _s;      \
  } while (0)

#define RB_REQUIRE(cond, ...)                                         \
  do {                                                                \
    if (!(cond)) return rb::fail(RB200_ERR_INVALID, __VA_ARGS__);     \
  } while (0)

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Number of kernels this library has launched from the calling thread (bench.py reports the delta over
// the timed region as gpu_launches).
inline int64_t& launch_count() {
  static thread_local int64_t n = 0;
  return n;
}

}  // namespace rb
