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

// ---- programmatic dependent launch (the kernels of one MCMC step form a dependent chain of short launches) ----------
// Every chain kernel starts with pdl_enter(): it lets the NEXT launch of the stream start being scheduled right away
// (griddepcontrol.launch_dependents) and then waits until the PREVIOUS launch has completed and its writes are visible
// (griddepcontrol.wait) -- nothing before the wait may touch memory a previous kernel writes.  launch_chain() marks a launch
// as allowed to start early; without the attribute both instructions are no-ops.  PMC_B200_PDL=0 turns the attribute off.
bool pdl_enabled();
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_launch_dependents(); pdl_wait(); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
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
