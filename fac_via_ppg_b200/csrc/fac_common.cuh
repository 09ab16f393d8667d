// Shared helpers for the fac_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include "../../include/fac_b200.h"

namespace fac {

// Error text returned by fac_last_error(); set by every failing entry point.
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

#define FAC_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      fac::set_error(__VA_ARGS__);        \
      return 1;                           \
    }                                     \
  } while (0)

__device__ __forceinline__ float sigmoidf_exact(float v) { return 1.0f / (1.0f + expf(-v)); }
// MUFU-based forms for the recurrent cells (ex2.approx + rcp.approx, ~2^-21 relative error): the cell
// non-linearities are 40 % of the instructions of the LSTM step when evaluated with the libm versions.
__device__ __forceinline__ float sigmoidf_fast(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float tanhf_fast(float v) {
  v = fminf(fmaxf(v, -15.f), 15.f);
  return 1.0f - __fdividef(2.0f, __expf(2.0f * v) + 1.0f);
}

// Per-process caches of device properties / function attributes are keyed by the CURRENT device: a process
// that drives several GPUs must not reuse what it learnt (or set) on another one.
constexpr int FAC_MAX_DEVICES = 64;
inline int current_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < FAC_MAX_DEVICES ? dev : 0;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace fac
