// N1: per-cluster dispersion sums for the model-selection scores of optimize_clustering
// (reference modules/statistics/statistics.py:73-74: sklearn calinski_harabasz_score and
// davies_bouldin_score on the labels of every k of the search interval).
//
// Given labels and the cluster means, ONE pass over the frames yields, per cluster,
//     ssq[c]   = sum_{t in c} ||y_t - m_c||^2      (Calinski-Harabasz: within-cluster dispersion)
//     sdist[c] = sum_{t in c} ||y_t - m_c||        (Davies-Bouldin: mean intra-cluster distance)
// in FP64 (4 d bytes per frame + the label; HBM-bound).  Frames of a warp that share a label (the usual
// case on trajectories) are reduced by shuffles before one atomic per warp.
#include "dcg_common.cuh"

namespace dcg {
namespace {

template <typename T>
__global__ void __launch_bounds__(256)
cluster_dispersion_kernel(const T* __restrict__ Y, int64_t n, int d, int64_t ld, const int* __restrict__ labels,
                          const double* __restrict__ means, int k, double* __restrict__ ssq, double* __restrict__ sdist) {
  const int lane = threadIdx.x & 31;
  for (int64_t t0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) - lane; t0 < n; t0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = t0 + lane;
    int c = -1;
    double q = 0.0;
    if (t < n) {
      c = labels[t];
      if (c >= 0 && c < k) {
        const T* row = Y + (size_t)t * ld;
        const double* m = means + (size_t)c * d;
        for (int j = 0; j < d; ++j) { const double dl = (double)row[j] - m[j]; q += dl * dl; }
      } else {
        c = -1;
      }
    }
    const double s = sqrt(q);
    const int c0 = __shfl_sync(0xffffffffu, c, 0);
    if (__all_sync(0xffffffffu, c == c0)) {
      const double qs = warp_sum(q), ss = warp_sum(s);
      if (lane == 0 && c0 >= 0) { atomicAdd(ssq + c0, qs); atomicAdd(sdist + c0, ss); }
    } else if (c >= 0) {
      atomicAdd(ssq + c, q);
      atomicAdd(sdist + c, s);
    }
  }
}

}  // namespace
}  // namespace dcg

using namespace dcg;

extern "C" int dcg_cluster_dispersion(const void* Y, int64_t n, int d, int64_t ld, int dtype_bytes, const int* labels,
                                      const double* means, int k, double* ssq, double* sdist, void* stream) {
  if (!Y || !labels || !means || !ssq || !sdist) return DCG_E_NULL;
  if (n <= 0 || d <= 0 || ld < d || k <= 0) return DCG_E_SHAPE;
  if (dtype_bytes != 4 && dtype_bytes != 8) return DCG_E_MODE;
  cudaStream_t st = (cudaStream_t)stream;
  DCG_CUDA_TRY(cudaMemsetAsync(ssq, 0, (size_t)k * sizeof(double), st));
  DCG_CUDA_TRY(cudaMemsetAsync(sdist, 0, (size_t)k * sizeof(double), st));
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 8);
  if (dtype_bytes == 4)
    cluster_dispersion_kernel<float><<<grid, 256, 0, st>>>((const float*)Y, n, d, ld, labels, means, k, ssq, sdist);
  else
    cluster_dispersion_kernel<double><<<grid, 256, 0, st>>>((const double*)Y, n, d, ld, labels, means, k, ssq, sdist);
  DCG_LAUNCH_CHECK();
  return 0;
}
