// Shared device/host helpers for libdcg_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <stdint.h>
#include <stddef.h>
#include "../../include/dcg.h"

#define DCG_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return -(int)_e;                  \
  } while (0)

#define DCG_LAUNCH_CHECK()                                   \
  do {                                                       \
    cudaError_t _e = cudaPeekAtLastError();                  \
    if (_e != cudaSuccess) return -(int)_e;                  \
  } while (0)

namespace dcg {

constexpr int kNumSMs = 148;  // B200

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) costs ~100 us of host time per call (measured:
// it was the whole cost of the small kernels that called it on every launch), so it is issued only
// when the requested size exceeds what this kernel already has on this device.  The table is the
// library's only mutable global state (a lazily filled per-device, per-kernel cache).
cudaError_t ensure_dynamic_smem(const void* func, size_t bytes);
// cached cudaOccupancyMaxActiveBlocksPerMultiprocessor (same reason)
cudaError_t cached_occupancy(int* per_sm, const void* func, int threads, size_t smem);

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// Streaming (read-once) global loads: keep them out of L1.
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ldg_stream2(const float* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];"
               : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// Vector width usable for rows of `ld` floats starting at `p` (4, 2 or 1).
static inline int row_vec_width(const void* p, int64_t ld) {
  uintptr_t a = (uintptr_t)p;
  if ((a % 16 == 0) && (ld % 4 == 0)) return 4;
  if ((a % 8 == 0) && (ld % 2 == 0)) return 2;
  return 1;
}

// Correctly rounded (x - m) / r for normal-range operands in 4 instructions instead of the
// ~10 of div.rn.f32: q0 = d * RN(1/r); rem = fma(-q0, r, d) (exact); q = fma(rem, rinv, q0).
// (Markstein's theorem; rinv must be the IEEE-rounded reciprocal.)  Falls back to nothing
// special for inf/nan/denormal inputs -- results there follow fma semantics.
__device__ __forceinline__ float standardize1(float x, float m, float r, float rinv) {
  const float d = x - m;
  const float q0 = d * rinv;
  const float rem = fmaf(-q0, r, d);
  return fmaf(rem, rinv, q0);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace dcg
