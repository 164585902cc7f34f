// A9/A10: fused standardise + projection P = ((X - mean)/range) @ W with per-column min/max.
// Replaces LinearCalculator.normalize_cv / project_data (reference cv_calculator.py:918-991).
//
// HBM-bound for the small CV dimensions the reference uses (d <= ~12): 4*f bytes read and
// 4*d bytes written per frame, 2*f*d FLOPs.  A warp owns kProjRows rows at a time; lanes stride
// the feature axis with 16-byte loads, keep kProjRows x D accumulators in registers, and read
// the (padded) weight rows through L1 (W is small and hot).  No shared-memory staging: every
// X byte is used exactly once.
#include "dcg_common.cuh"

namespace dcg {

constexpr int kProjWarps = 8;
constexpr int kProjThreads = kProjWarps * 32;

// Per-column operands laid out for coalesced loads in the main kernel (a lane owns VEC consecutive
// columns): Wt[dp][f] (weights transposed, zero rows up to dp) and Cp[3][f] = mean | range |
// RN(1/range), so the inner loop does 16-byte loads that are contiguous across the warp instead of
// strided scalar loads and a division per column.
// Rows of both arrays are fp = f rounded up to 4 floats long; the padding columns hold weight 0
// (and mean 0, range 1), so a partial last vector of an X row contributes nothing.
__global__ void pad_weights_kernel(const float* __restrict__ W, int f, int fp, int d, int d0, int dc,
                                   int dp, float* __restrict__ Wt, const float* __restrict__ mean,
                                   const float* __restrict__ range, float* __restrict__ Cp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= fp * dp) return;
  const int j = i / fp, col = i % fp;
  Wt[i] = (j < dc && col < f) ? W[(size_t)col * d + d0 + j] : 0.f;
  if (j == 0 && mean) {
    const float rg = col < f ? range[col] : 1.f;
    Cp[col] = col < f ? mean[col] : 0.f;
    Cp[fp + col] = rg;
    Cp[2 * fp + col] = 1.0f / rg;
  }
}

// small, hot operands (weights, column parameters): cached loads
template <int VEC>
__device__ __forceinline__ void load_par_vec(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) { float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  else if constexpr (VEC == 2) { float2 t = __ldg(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y; }
  else { v[0] = __ldg(p); }
}

template <int VEC>
__device__ __forceinline__ void load_row_vec(const float* p, float (&v)[VEC]);

// the last vector of a row may be partial (f not a multiple of VEC): missing columns read as 0
template <int VEC>
__device__ __forceinline__ void load_row_tail(const float* p, int valid, float (&v)[VEC]) {
  if (valid >= VEC) { load_row_vec<VEC>(p, v); return; }
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = i < valid ? ldg_stream1(p + i) : 0.f;
}

template <int VEC>
__device__ __forceinline__ void load_row_vec(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) { float4 t = ldg_stream4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  else if constexpr (VEC == 2) { float2 t = ldg_stream2(p); v[0] = t.x; v[1] = t.y; }
  else { v[0] = ldg_stream1(p); }
}

// DP4 = padded output width / 4 (1..4).  STD: apply (x-mean)/range.
// Rows per warp step (amortises the W loads): 4 for narrow outputs, 2 for wide (register budget).
__host__ __device__ constexpr int proj_rows(int dp4) { return dp4 <= 3 ? 4 : 2; }

// resident CTAs per SM the register budget is compiled for: narrow outputs need few registers, and
// more resident warps mean more 16-byte loads in flight (the kernel is HBM-latency bound)
__host__ __device__ constexpr int proj_min_ctas(int dp4) { return dp4 == 1 ? 3 : 2; }

template <int VEC, int DP4, bool STD>
__global__ void __launch_bounds__(kProjThreads, proj_min_ctas(DP4))
project_kernel(const float* __restrict__ X, int64_t n, int f, int fp, int64_t ld,
               const float* __restrict__ Cp,
               const float* __restrict__ Wt, int d_total, int d0, int dc,
               float* __restrict__ P, float* __restrict__ part_min, float* __restrict__ part_max) {
  constexpr int DP = DP4 * 4;
  constexpr int kProjRows = proj_rows(DP4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t gwarp = (int64_t)blockIdx.x * kProjWarps + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kProjWarps;
  const int64_t ngroups = (n + kProjRows - 1) / kProjRows;

  float lo[DP], hi[DP];
#pragma unroll
  for (int j = 0; j < DP; ++j) { lo[j] = INFINITY; hi[j] = -INFINITY; }

  for (int64_t g = gwarp; g < ngroups; g += nwarps) {
    const int64_t row0 = g * kProjRows;
    float acc[kProjRows][DP];
#pragma unroll
    for (int r = 0; r < kProjRows; ++r)
#pragma unroll
      for (int j = 0; j < DP; ++j) acc[r][j] = 0.f;

    for (int col = lane * VEC; col < f; col += 32 * VEC) {
      float x[kProjRows][VEC];
#pragma unroll
      for (int r = 0; r < kProjRows; ++r) {
        // rows past the end re-read the last valid row (results discarded)
        const int64_t row = min(row0 + r, n - 1);
        load_row_tail<VEC>(X + row * ld + col, f - col, x[r]);
      }
      float z[kProjRows][VEC];
      if constexpr (STD) {
        float m[VEC], rg[VEC], ri[VEC];
        load_par_vec<VEC>(Cp + col, m);
        load_par_vec<VEC>(Cp + fp + col, rg);
        load_par_vec<VEC>(Cp + 2 * (size_t)fp + col, ri);
#pragma unroll
        for (int r = 0; r < kProjRows; ++r)
#pragma unroll
          for (int v = 0; v < VEC; ++v) z[r][v] = standardize1(x[r][v], m[v], rg[v], ri[v]);
      } else {
#pragma unroll
        for (int r = 0; r < kProjRows; ++r)
#pragma unroll
          for (int v = 0; v < VEC; ++v) z[r][v] = x[r][v];
      }
#pragma unroll
      for (int j = 0; j < DP; ++j) {
        float w[VEC];
        load_par_vec<VEC>(Wt + (size_t)j * fp + col, w);
#pragma unroll
        for (int r = 0; r < kProjRows; ++r)
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[r][j] = fmaf(z[r][v], w[v], acc[r][j]);
      }
    }
    // warp all-reduce of the kProjRows x DP partial dot products
#pragma unroll
    for (int r = 0; r < kProjRows; ++r)
#pragma unroll
      for (int j = 0; j < DP; ++j) acc[r][j] = warp_sum(acc[r][j]);
#pragma unroll
    for (int r = 0; r < kProjRows; ++r) {
      if (row0 + r < n) {
#pragma unroll
        for (int j = 0; j < DP; ++j) { lo[j] = fminf(lo[j], acc[r][j]); hi[j] = fmaxf(hi[j], acc[r][j]); }
        if (lane == r) {
          float* out = P + (row0 + r) * (int64_t)d_total + d0;
#pragma unroll
          for (int j = 0; j < DP; ++j) if (j < dc) out[j] = acc[r][j];
        }
      }
    }
  }

  // per-CTA min / max partials (all lanes of a warp hold identical lo/hi)
  __shared__ float s_lo[kProjWarps][DP], s_hi[kProjWarps][DP];
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < DP; ++j) { s_lo[warp][j] = lo[j]; s_hi[warp][j] = hi[j]; }
  }
  __syncthreads();
  if (threadIdx.x < dc) {
    float a = INFINITY, b = -INFINITY;
#pragma unroll
    for (int w = 0; w < kProjWarps; ++w) { a = fminf(a, s_lo[w][threadIdx.x]); b = fmaxf(b, s_hi[w][threadIdx.x]); }
    part_min[(size_t)blockIdx.x * d_total + d0 + threadIdx.x] = a;
    part_max[(size_t)blockIdx.x * d_total + d0 + threadIdx.x] = b;
  }
}

__global__ void project_minmax_kernel(const float* __restrict__ part_min, const float* __restrict__ part_max,
                                      int nparts, int d, float* __restrict__ pmin, float* __restrict__ pmax) {
  const int j = blockIdx.x;
  float a = INFINITY, b = -INFINITY;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
    a = fminf(a, part_min[(size_t)i * d + j]);
    b = fmaxf(b, part_max[(size_t)i * d + j]);
  }
  __shared__ float sa[32], sb[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
    b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
  }
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { a = fminf(a, sa[w]); b = fmaxf(b, sb[w]); }
    if (pmin) pmin[j] = a;
    if (pmax) pmax[j] = b;
  }
}

// persistent grid: every resident CTA slot of the device, at most (upper bound used for workspace)
static int project_grid(int64_t n, int ctas_per_sm = 4) {
  const int64_t groups = ceil_div(n, 2);
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(groups, kProjWarps), (int64_t)kNumSMs * ctas_per_sm));
}

template <int VEC, bool STD>
static void launch_project(int dp4, dim3 grid, cudaStream_t st, const float* X, int64_t n, int f, int fp,
                           int64_t ld, const float* Cp, const float* Wp,
                           int d, int d0, int dc, float* P, float* pmn, float* pmx) {
  switch (dp4) {
    case 1: project_kernel<VEC, 1, STD><<<grid, kProjThreads, 0, st>>>(X, n, f, fp, ld, Cp, Wp, d, d0, dc, P, pmn, pmx); break;
    case 2: project_kernel<VEC, 2, STD><<<grid, kProjThreads, 0, st>>>(X, n, f, fp, ld, Cp, Wp, d, d0, dc, P, pmn, pmx); break;
    case 3: project_kernel<VEC, 3, STD><<<grid, kProjThreads, 0, st>>>(X, n, f, fp, ld, Cp, Wp, d, d0, dc, P, pmn, pmx); break;
    default: project_kernel<VEC, 4, STD><<<grid, kProjThreads, 0, st>>>(X, n, f, fp, ld, Cp, Wp, d, d0, dc, P, pmn, pmx); break;
  }
}

}  // namespace dcg

using namespace dcg;

extern "C" size_t dcg_project_workspace_bytes(int64_t n, int f, int d) {
  if (n <= 0 || f <= 0 || d <= 0) return 0;
  return align_up((size_t)(f + 3) * 16 * sizeof(float), 256) + align_up((size_t)(f + 3) * 3 * sizeof(float), 256) +
         2 * align_up((size_t)project_grid(n) * d * sizeof(float), 256);
}

extern "C" int dcg_project_f32(const float* X, int64_t n, int f, int64_t ld,
                               const float* mean, const float* range,
                               const float* W, int d, float* P, float* pmin, float* pmax,
                               void* ws, size_t ws_bytes, void* stream) {
  if (!X || !W || !P) return DCG_E_NULL;
  if ((mean == nullptr) != (range == nullptr)) return DCG_E_NULL;
  if (n <= 0 || f <= 0 || ld < f || d < 1 || d > 64) return DCG_E_SHAPE;
  if (!ws || ws_bytes < dcg_project_workspace_bytes(n, f, d)) return DCG_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)ws;
  float* Wp = (float*)w;
  w += align_up((size_t)(f + 3) * 16 * sizeof(float), 256);
  float* Cp = (float*)w;
  w += align_up((size_t)(f + 3) * 3 * sizeof(float), 256);
  const int fp = (f + 3) / 4 * 4;
  const int grid_max = project_grid(n);
  float* pmn = (float*)w;
  w += align_up((size_t)grid_max * d * sizeof(float), 256);
  float* pmx = (float*)w;
  int grid = grid_max;
  const int vec = row_vec_width(X, ld);     // f need not be a multiple: the last vector of a row is guarded
  const bool stdz = mean != nullptr;
  for (int d0 = 0; d0 < d; d0 += 16) {
    const int dc = min(16, d - d0);
    const int dp4 = (dc + 3) / 4, dp = dp4 * 4;
    pad_weights_kernel<<<(unsigned)ceil_div((int64_t)fp * dp, 256), 256, 0, st>>>(W, f, fp, d, d0, dc, dp, Wp, mean, range, Cp);
    DCG_LAUNCH_CHECK();
    // every chunk of 16 output columns uses the same grid (the min/max partials are indexed by CTA)
    if (d0 == 0) grid = project_grid(n, proj_min_ctas(d > 16 ? 4 : dp4));
#define DCG_PROJ(V, S) launch_project<V, S>(dp4, dim3(grid), st, X, n, f, fp, ld, Cp, Wp, d, d0, dc, P, pmn, pmx)
    if (vec == 4) { if (stdz) DCG_PROJ(4, true); else DCG_PROJ(4, false); }
    else if (vec == 2) { if (stdz) DCG_PROJ(2, true); else DCG_PROJ(2, false); }
    else { if (stdz) DCG_PROJ(1, true); else DCG_PROJ(1, false); }
#undef DCG_PROJ
    DCG_LAUNCH_CHECK();
  }
  if (pmin || pmax) {
    project_minmax_kernel<<<d, 256, 0, st>>>(pmn, pmx, grid, d, pmin, pmax);
    DCG_LAUNCH_CHECK();
  }
  return 0;
}
