// A9/A10: fused standardise + projection P = ((X - mean)/range) @ W with per-column min/max.
// Replaces LinearCalculator.normalize_cv / project_data (reference cv_calculator.py:918-991).
//
// HBM-bound for the small CV dimensions the reference uses (d <= ~12): 4*f bytes read and
// 4*d bytes written per frame, 2*f*d FLOPs.  A warp owns kProjRows rows at a time; lanes stride
// the feature axis with 16-byte loads, keep kProjRows x D accumulators in registers, and read
// the (padded) weight rows through L1 (W is small and hot).  No shared-memory staging: every
// X byte is used exactly once.
#include "dcg_common.cuh"
#include "project_mma.cuh"

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

// ---- streamed tensor-core path (project_mma.cuh) ---------------------------------------------
namespace {

constexpr int kCombineGrid = 2 * kNumSMs;
constexpr size_t kPartBudget = (size_t)256 << 20;      // bytes of partial projections per row batch

struct PmPlan {
  int nr, rw;                 // ranges of the feature axis
  int64_t batch_rows;         // rows per launch (several ranges only: bounds the partial buffer)
  int64_t nbatches;
  size_t bf_bytes, mf_bytes, part_bytes, mm_slots;
};

static PmPlan pm_plan_dense(int64_t n, int f) {
  PmPlan p;
  p.nr = (int)ceil_div(f, pm::kRangeMax);
  p.rw = (int)(ceil_div(ceil_div(f, p.nr), 16) * 16);
  p.bf_bytes = align_up((size_t)p.nr * pm::kWarps * 8 * 32 * 16 * sizeof(float), 256);
  p.mf_bytes = align_up((size_t)p.nr * pm::kWarps * 8 * 32 * 4 * sizeof(float), 256);
  if (p.nr == 1) {
    p.batch_rows = n; p.nbatches = 1; p.part_bytes = 0; p.mm_slots = kNumSMs;
  } else {
    const int64_t cap = std::max<int64_t>(16, (int64_t)(kPartBudget / ((size_t)p.nr * 16 * sizeof(float))) / 16 * 16);
    p.batch_rows = std::min<int64_t>(n, cap);
    p.nbatches = ceil_div(n, p.batch_rows);
    p.part_bytes = align_up((size_t)p.nr * p.batch_rows * 16 * sizeof(float), 256);
    p.mm_slots = (size_t)p.nbatches * kCombineGrid;
  }
  return p;
}

static size_t pm_dense_bytes(int64_t n, int f, int d) {
  const PmPlan p = pm_plan_dense(n, f);
  return p.bf_bytes + p.mf_bytes + p.part_bytes + 2 * align_up(p.mm_slots * (size_t)d * sizeof(float), 256);
}

static bool pm_usable(const void* X, int64_t ld) { return ((uintptr_t)X % 16 == 0) && (ld % 4 == 0); }

// rows per work item: single range -> fine-grained (no operand reload between items); several
// ranges -> coarse enough that reloading the B fragments (<= 150 KB per CTA from L2) is noise
static int pm_rows_per_item(int nr) { return nr == 1 ? 64 : 512; }

template <int NT>
static cudaError_t pm_launch(int grid, cudaStream_t st, const float* X, int64_t n, int64_t ld, const pm::Geom& g,
                             const float* Bf, const float* Mf, float* out, int64_t out_rs, int64_t out_cs,
                             float* pmn, float* pmx, int mm_ld) {
  cudaError_t e = ensure_dynamic_smem((const void*)pm::project_mma_kernel<NT>, pm::kSmemBytes);
  if (e != cudaSuccess) return e;
  pm::project_mma_kernel<NT><<<grid, pm::kThreads, pm::kSmemBytes, st>>>(
      X, n, ld, g, pm_rows_per_item(g.nr), reinterpret_cast<const float4*>(Bf), reinterpret_cast<const float4*>(Mf),
      out, out_rs, out_cs, pmn, pmx, mm_ld);
  return cudaPeekAtLastError();
}

static cudaError_t pm_launch_any(int nt, int grid, cudaStream_t st, const float* X, int64_t n, int64_t ld,
                                 const pm::Geom& g, const float* Bf, const float* Mf, float* out, int64_t out_rs,
                                 int64_t out_cs, float* pmn, float* pmx, int mm_ld) {
  return nt == 1 ? pm_launch<1>(grid, st, X, n, ld, g, Bf, Mf, out, out_rs, out_cs, pmn, pmx, mm_ld)
                 : pm_launch<2>(grid, st, X, n, ld, g, Bf, Mf, out, out_rs, out_cs, pmn, pmx, mm_ld);
}

static int pm_grid(int64_t n, const pm::Geom& g) {
  const int64_t items = ceil_div(n, pm_rows_per_item(g.nr)) * g.nr;
  return (int)std::max<int64_t>(1, std::min<int64_t>(items, kNumSMs));
}

// dense projection through the streamed path; `w` points at pm_dense_bytes() of workspace
static int pm_project_dense(const float* X, int64_t n, int f, int64_t ld, const float* mean, const float* range,
                            const float* W, int d, float* P, float* pmin, float* pmax, char* w, cudaStream_t st) {
  const PmPlan p = pm_plan_dense(n, f);
  float* Bf = (float*)w; w += p.bf_bytes;
  float* Mf = (float*)w; w += p.mf_bytes;
  float* part = (float*)w; w += p.part_bytes;
  float* pmn = (float*)w; w += align_up(p.mm_slots * (size_t)d * sizeof(float), 256);
  float* pmx = (float*)w;
  const bool want_mm = pmin || pmax;
  int slots = 0;
  for (int d0 = 0; d0 < d; d0 += 16) {
    const int dc = std::min(16, d - d0);
    pm::Geom g{f, p.rw, p.nr, 0, d, d0, dc};
    const int nthreads = p.nr * pm::kWarps * 8 * 32;
    pm::prepare_kernel<<<(unsigned)ceil_div(nthreads, 256), 256, 0, st>>>(g, W, mean, range, Bf, Mf);
    DCG_LAUNCH_CHECK();
    const int nt = dc <= 8 ? 1 : 2;
    if (p.nr == 1) {
      const int grid = pm_grid(n, g);
      DCG_CUDA_TRY(pm_launch_any(nt, grid, st, X, n, ld, g, Bf, Mf, P + d0, d, 0,
                                 want_mm ? pmn + d0 : nullptr, want_mm ? pmx + d0 : nullptr, want_mm ? d : 0));
      slots = grid;
    } else {
      for (int64_t b = 0; b < p.nbatches; ++b) {
        const int64_t r0 = b * p.batch_rows, nb = std::min<int64_t>(p.batch_rows, n - r0);
        DCG_CUDA_TRY(pm_launch_any(nt, pm_grid(nb, g), st, X + r0 * ld, nb, ld, g, Bf, Mf, part, dc, nb * dc,
                                   nullptr, nullptr, 0));
        // unused slots of a short batch must not be read: every batch publishes kCombineGrid slots
        pm::combine_kernel<<<kCombineGrid, pm::kCombineThreads, 0, st>>>(
            part, p.nr, nb, dc, nb * dc, P + r0 * d + d0, d,
            want_mm ? pmn + (size_t)b * kCombineGrid * d + d0 : nullptr,
            want_mm ? pmx + (size_t)b * kCombineGrid * d + d0 : nullptr, want_mm ? d : 0);
        DCG_LAUNCH_CHECK();
      }
      slots = (int)(p.nbatches * kCombineGrid);
    }
  }
  if (want_mm) {
    project_minmax_kernel<<<d, 256, 0, st>>>(pmn, pmx, slots, d, pmin, pmax);
    DCG_LAUNCH_CHECK();
  }
  return 0;
}

static size_t legacy_bytes(int64_t n, int f, int d) {
  return align_up((size_t)(f + 3) * 16 * sizeof(float), 256) + align_up((size_t)(f + 3) * 3 * sizeof(float), 256) +
         2 * align_up((size_t)project_grid(n) * d * sizeof(float), 256);
}

}  // namespace

extern "C" size_t dcg_project_workspace_bytes(int64_t n, int f, int d) {
  if (n <= 0 || f <= 0 || d <= 0) return 0;
  return std::max(legacy_bytes(n, f, d), pm_dense_bytes(n, f, d));
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
  // 16-byte aligned rows (the layout load_training_tensor produces): streamed tensor-core path
  if (pm_usable(X, ld)) return pm_project_dense(X, n, f, ld, mean, range, W, d, P, pmin, pmax, w, st);
  // otherwise: register-tile kernel with 8- or 4-byte loads
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

// hTICA level 1: all diagonal blocks projected in one pass over X.
extern "C" size_t dcg_project_blocks_workspace_bytes(int64_t n, int f, int block) {
  if (n <= 0 || f <= 0 || block <= 0) return 0;
  const size_t nr = (size_t)ceil_div(f, block);
  return align_up(nr * pm::kWarps * 8 * 32 * 16 * sizeof(float), 256) + align_up(nr * pm::kWarps * 8 * 32 * 4 * sizeof(float), 256);
}

extern "C" int dcg_project_blocks_f32(const float* X, int64_t n, int f, int64_t ld,
                                      const float* mean, const float* range,
                                      const float* W, int block, int s, float* P, int64_t p_ld,
                                      void* ws, size_t ws_bytes, void* stream) {
  if (!X || !W || !P) return DCG_E_NULL;
  if ((mean == nullptr) != (range == nullptr)) return DCG_E_NULL;
  if (n <= 0 || f <= 0 || ld < f || block < 1 || block > pm::kRangeMax - 6 || s < 1 || s > 16) return DCG_E_SHAPE;
  const int nr = (int)ceil_div(f, block);
  if (p_ld < (int64_t)(nr - 1) * s + std::min(s, f - (nr - 1) * block)) return DCG_E_SHAPE;
  if (!pm_usable(X, ld)) return DCG_E_ALIGN;
  if (!ws || ws_bytes < dcg_project_blocks_workspace_bytes(n, f, block)) return DCG_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)ws;
  float* Bf = (float*)w;
  w += align_up((size_t)nr * pm::kWarps * 8 * 32 * 16 * sizeof(float), 256);
  float* Mf = (float*)w;
  pm::Geom g{f, block, nr, s, s, 0, s};
  pm::prepare_kernel<<<(unsigned)ceil_div((int64_t)nr * pm::kWarps * 8 * 32, 256), 256, 0, st>>>(g, W, mean, range, Bf, Mf);
  DCG_LAUNCH_CHECK();
  DCG_CUDA_TRY(pm_launch_any(s <= 8 ? 1 : 2, pm_grid(n, g), st, X, n, ld, g, Bf, Mf, P, p_ld, s, nullptr, nullptr, 0));
  return 0;
}
