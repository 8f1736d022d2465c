// Shared helpers for the riporb200 C-ABI library: error reporting and CUDA call checking.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
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

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// Every hot kernel of the engine calls pdl_wait() before it touches global memory and pdl_trigger() once its
// main work is done: the next kernel of the stream may then be scheduled (and run its prologue: barrier init,
// TMEM allocation, descriptor prefetch) while this one drains, and nobody reads global memory before the previous
// kernel has fully completed and flushed. That hides the ~3 us launch/scheduling gap paid ~4400 times per search.
// The trigger is deliberately LATE: triggered at kernel entry, the dependent grid's CTAs (a GEMM CTA owns
// ~200 KB of shared memory) become resident while a multi-wave grid still has CTAs to schedule and the step
// gets slower (measured on B200: 129 ms vs 124 ms without PDL).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = []() {
    const char* e = getenv("RB200_PDL");     // default on (+4 % on the bench step); RB200_PDL=0 disables
    return !(e && e[0] == '0');
  }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace rb
