// Shared host-side helpers for the C-ABI translation units: status codes and the thread-local
// error string behind sod_last_error().
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>

#include "../../include/sod.h"

// Device-side bounds checks of every data-dependent index (compute-sanitizer is closed on the B200 pool):
// compiled in with -DSOD_DEVICE_CHECKS (python -m sod_b200.build --variant checked -DSOD_DEVICE_CHECKS), a
// violated check traps, i.e. the launch fails and the test that ran it reports a CUDA error.
#ifdef SOD_DEVICE_CHECKS
#define SOD_DCHECK(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define SOD_DCHECK(cond) do { } while (0)
#endif

namespace sod {

void set_error(const char* fmt, ...);
int device_sm_count();

// Measurement hook behind sod_timing_enable / sod_timing_read: when enabled on the calling thread, the
// entry points bracket the named stage's launches with CUDA events on the caller's stream.
void stage_begin(int stage, cudaStream_t st);
void stage_end(int stage, cudaStream_t st);
struct StageScope {
  int stage;
  cudaStream_t st;
  StageScope(int s, cudaStream_t stream) : stage(s), st(stream) { stage_begin(stage, st); }
  ~StageScope() { stage_end(stage, st); }
};

#define SOD_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      ::sod::set_error(__VA_ARGS__);    \
      return SOD_ERR_INVALID_ARGUMENT;  \
    }                                   \
  } while (0)

#define SOD_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      ::sod::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                       __LINE__);                                                         \
      return SOD_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define SOD_CHECK_LAUNCH(name)                                                       \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      ::sod::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));    \
      return SOD_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

}  // namespace sod
