// A4: in-place IEEE float32 standardisation x = (x - mean[j]) / range[j].
// Replaces LinearCalculator.normalize_data (reference cv_calculator.py:806-837) and the CV
// normalisation of projected data (cv_calculator.py:966-970).  HBM-bound: 8 bytes per element.
#include "dcg_common.cuh"

namespace dcg {

template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<1> { using type = float; };

template <int VEC>
__device__ __forceinline__ void std_vec(typename VecT<VEC>::type& x, const float* m, const float* r,
                                        const float* ri) {
  float* e = reinterpret_cast<float*>(&x);
#pragma unroll
  for (int v = 0; v < VEC; ++v) e[v] = standardize1(e[v], m[v], r[v], ri[v]);
}

// Wide matrices: thread -> fixed column group (mean / range / 1/range in registers), loops rows.
template <int VEC>
__global__ void __launch_bounds__(256)
standardize_wide_kernel(float* __restrict__ X, int64_t n, int f, int64_t ld,
                        const float* __restrict__ mean, const float* __restrict__ range,
                        int rows_per_cta) {
  using V = typename VecT<VEC>::type;
  const int col0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (col0 >= f) return;
  float m[VEC], r[VEC], ri[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    m[v] = mean[col0 + v];
    r[v] = range[col0 + v];
    ri[v] = 1.0f / r[v];
  }
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = min(n, r0 + rows_per_cta);
  int64_t row = r0;
  for (; row + 3 < r1; row += 4) {
    V x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) x[u] = *reinterpret_cast<V*>(X + (row + u) * ld + col0);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      std_vec<VEC>(x[u], m, r, ri);
      *reinterpret_cast<V*>(X + (row + u) * ld + col0) = x[u];
    }
  }
  for (; row < r1; ++row) {
    V x = *reinterpret_cast<V*>(X + row * ld + col0);
    std_vec<VEC>(x, m, r, ri);
    *reinterpret_cast<V*>(X + row * ld + col0) = x;
  }
}

// Narrow matrices (projected CVs, f = d small): grid-stride over rows, thread per row.
__global__ void __launch_bounds__(256)
standardize_narrow_kernel(float* __restrict__ X, int64_t n, int f, int64_t ld,
                          const float* __restrict__ mean, const float* __restrict__ range) {
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n;
       row += (int64_t)gridDim.x * blockDim.x) {
    float* p = X + row * ld;
    for (int j = 0; j < f; ++j) {
      const float r = __ldg(range + j);
      p[j] = standardize1(p[j], __ldg(mean + j), r, 1.0f / r);
    }
  }
}

// DeepTICA minibatches (A11 / A12 feeder): Z[i, :] = (X[idx[i] + offset, :] - mean) / range in ONE
// pass -- the reference gathers the batch, then `Normalization` (norm_in) subtracts and divides in
// two more read+write passes over B x F (mlcolvar Normalization.forward, model built at
// cv_calculator.py:2569-2590).  A warp owns a row at a time; the column parameters of a lane's
// vectors are re-read from L1 (rows are random, so there is nothing to keep in registers across
// rows beyond one vector).
template <int VEC>
__global__ void __launch_bounds__(256)
gather_standardize_kernel(const float* __restrict__ X, int f, int64_t ld, const int64_t* __restrict__ idx,
                          int64_t nb, int64_t offset, const float* __restrict__ mean,
                          const float* __restrict__ range, float* __restrict__ Z) {
  using V = typename VecT<VEC>::type;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp0; i < nb; i += nwarps) {
    const float* src = X + (idx[i] + offset) * ld;
    float* dst = Z + i * (int64_t)f;
    int col = lane * VEC;
    for (; col + VEC <= f; col += 32 * VEC) {
      V x = *reinterpret_cast<const V*>(src + col);
      float* e = reinterpret_cast<float*>(&x);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float r = __ldg(range + col + v);
        e[v] = standardize1(e[v], __ldg(mean + col + v), r, 1.0f / r);
      }
      *reinterpret_cast<V*>(dst + col) = x;
    }
    for (int c = col; c < f && c < col + VEC; ++c) {          // partial last vector
      const float r = __ldg(range + c);
      dst[c] = standardize1(src[c], __ldg(mean + c), r, 1.0f / r);
    }
  }
}

}  // namespace dcg

using namespace dcg;

extern "C" int dcg_gather_standardize_f32(const float* X, int64_t n, int f, int64_t ld,
                                          const int64_t* idx, int64_t nb, int64_t offset,
                                          const float* mean, const float* range, float* Z, void* stream) {
  if (!X || !idx || !mean || !range || !Z) return DCG_E_NULL;
  if (n <= 0 || f <= 0 || ld < f || nb <= 0 || offset < 0 || offset >= n) return DCG_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  // source rows: ld-strided (vector width from X / ld); destination rows are f floats apart
  int vec = row_vec_width(X, ld);
  while (vec > 1 && (f % vec || (uintptr_t)Z % (4 * vec))) vec >>= 1;
  const int64_t blocks = std::min<int64_t>(ceil_div(nb, 8), (int64_t)kNumSMs * 8);
  if (vec == 4) gather_standardize_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(X, f, ld, idx, nb, offset, mean, range, Z);
  else if (vec == 2) gather_standardize_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(X, f, ld, idx, nb, offset, mean, range, Z);
  else gather_standardize_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(X, f, ld, idx, nb, offset, mean, range, Z);
  DCG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dcg_standardize_f32(float* X, int64_t n, int f, int64_t ld,
                                   const float* mean, const float* range, void* stream) {
  if (!X || !mean || !range) return DCG_E_NULL;
  if (n <= 0 || f <= 0 || ld < f) return DCG_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  if (f <= 64) {
    const int64_t blocks = std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 16);
    standardize_narrow_kernel<<<(unsigned)blocks, 256, 0, st>>>(X, n, f, ld, mean, range);
    DCG_LAUNCH_CHECK();
    return 0;
  }
  int vec = row_vec_width(X, ld);
  while (f % vec) vec >>= 1;
  const int threads = 256;
  const int gx = (int)ceil_div(ceil_div(f, vec), threads);
  // enough row blocks to fill the machine several times over, at least 16 rows each
  int64_t want = (int64_t)kNumSMs * 16 / gx + 1;
  int rows_per_cta = (int)max((int64_t)16, ceil_div(n, want));
  int64_t gy = ceil_div(n, rows_per_cta);
  if (gy > 65535) { rows_per_cta = (int)ceil_div(n, 65535); gy = ceil_div(n, rows_per_cta); }
  dim3 grid(gx, (unsigned)gy);
  if (vec == 4) standardize_wide_kernel<4><<<grid, threads, 0, st>>>(X, n, f, ld, mean, range, rows_per_cta);
  else if (vec == 2) standardize_wide_kernel<2><<<grid, threads, 0, st>>>(X, n, f, ld, mean, range, rows_per_cta);
  else standardize_wide_kernel<1><<<grid, threads, 0, st>>>(X, n, f, ld, mean, range, rows_per_cta);
  DCG_LAUNCH_CHECK();
  return 0;
}
