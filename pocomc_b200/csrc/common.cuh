// Shared helpers for libpmc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/pmc_b200.h"

namespace pmc {

void set_error(const char* fmt, ...);
int sm_count();

#define PMC_TRY(expr)                                                                     \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      pmc::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                 \
                     cudaGetErrorString(_e));                                             \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

#define PMC_REQUIRE(cond, msg)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      pmc::set_error("%s (%s) at %s:%d", msg, #cond, __FILE__, __LINE__);                 \
      return 2;                                                                           \
    }                                                                                     \
  } while (0)

#define PMC_LAUNCH_CHECK() PMC_TRY(cudaGetLastError())

static inline cudaStream_t as_stream(pmc_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// grid for a grid-stride kernel: ceil(work / per_block) blocks, capped at per_sm resident blocks per SM
static inline int grid_for(long long work, int per_block, int per_sm) {
  long long b = (work + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ int warp_and(int v) { return __all_sync(FULL, v); }

// np.logaddexp semantics (numpy npy_logaddexp): equal args -> x + log 2; else max + log1p(exp(-|d|))
__device__ __forceinline__ double logaddexp(double x, double y) {
  if (x == y) return x + 0.69314718055994530942;
  double d = x - y;
  if (d > 0) return x + log1p(exp(-d));
  if (d <= 0) return y + log1p(exp(d));
  return x + y;  // NaN
}

}  // namespace pmc
