// Device helpers shared by the CUDA-core and the tcgen05 KMeans E-step kernels.
#pragma once
#include "dcg_common.cuh"

namespace dcg {

// dcg_kmeans_iterate_n: the run's control words [stop, iterations done, tolerance] sit this many doubles
// after `stats` in the work buffer ([sums | counts | stats 3 | info 2 | ctl 3]); a kernel launched with bit 1
// of `update_sums` set returns at once when stop != 0.
constexpr int kKmCtlOffset = 5;

// FP64 re-evaluation of one frame over all centres (rare path), by the whole warp: the frame is
// broadcast through shared memory, lane l scores centres l, l + 32, ..., and the partial
// (best, second, label) triples are merged by shuffles with the lowest-index tie-break.  All lanes
// return the same result.  (A per-thread loop over k centres kept 31 lanes idle for ~25 k
// instructions per refined frame: 12 % of all stall samples at k = 1000.)
__device__ __forceinline__ void km_refine_warp(const double* __restrict__ yy_s, int d,
                                               const double* __restrict__ centers, int k,
                                               int& lab, double& best, double& second) {
  const int lane = threadIdx.x & 31;
  best = INFINITY; second = INFINITY; lab = 0x7fffffff;
  // four centres per round: their loads are independent, so a round costs one L2 latency (the
  // centres do not stay in L1 next to 160+ KB of shared memory), not four
  for (int j0 = lane; j0 < k; j0 += 128) {
    double dot[4] = {0.0, 0.0, 0.0, 0.0}, csq[4] = {0.0, 0.0, 0.0, 0.0};
    for (int q = 0; q < d; ++q) {
      const double yq = yy_s[q];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = min(j0 + 32 * u, k - 1);
        const double cq = __ldg(centers + (size_t)j * d + q);
        dot[u] = fma(yq, cq, dot[u]);
        csq[u] = fma(cq, cq, csq[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 32 * u;
      if (j < k) {
        const double s = csq[u] - 2.0 * dot[u];
        if (s < best) { second = best; best = s; lab = j; }    // j ascending per lane: strict < keeps the lowest
        else if (s < second) second = s;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const double os = __shfl_xor_sync(0xffffffffu, second, o);
    const int ol = __shfl_xor_sync(0xffffffffu, lab, o);
    const bool take = ob < best || (ob == best && ol < lab);
    // second best of the union: the larger of the two bests, or either side's second
    const double hi = take ? best : ob;
    second = fmin(fmin(second, os), hi);
    if (take) { best = ob; lab = ol; }
  }
}

// slot (8 bytes of shared memory, two's complement in two 32-bit words) += round(v)
__device__ __forceinline__ void km_add_fixed(double* slot, double v) {
  const long long iv = __double2ll_rn(v);
  const unsigned int lo = (unsigned int)iv;
  unsigned int* w = reinterpret_cast<unsigned int*>(slot);
  const unsigned int old = atomicAdd(w, lo);
  const int hi = (int)(iv >> 32) + (((old + lo) < lo) ? 1 : 0);
  if (hi != 0) atomicAdd(reinterpret_cast<int*>(w + 1), hi);
}


// Tensor-core E-step for many centres: register-resident scan on mma.sync (kmeans_mma.cu).  Returns
// DCG_E_MODE when the shape is outside its range (the caller then takes the CUDA-core kernel).
int kmeans_mma_launch(const void* Y, int dtype_bytes, int64_t n, int d, int64_t ld, const double* centers, int k,
                      int32_t* labels, double* sums, double* counts, double* stats, void* gap,
                      int update_sums, const double* y_absmax, cudaStream_t st);

}  // namespace dcg
