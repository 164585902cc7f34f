// K1 / K3: KMeans E-step + M-step sums, and nearest-sample-to-centre, in the low-dimensional
// CV space.  Replaces one sklearn lloyd_iter as triggered by statistics.kmeans_clustering
// (reference modules/statistics/statistics.py:159-197; sklearn _k_means_lloyd.pyx:193-214) and
// statistics.find_centroids (statistics.py:370-377).
//
// E-step arithmetic: every (frame, centre) score ||c||^2 - 2 y.c is SCREENED in FP32 on CUDA
// cores (centres broadcast from shared memory, frames in registers).  A frame whose best and
// second-best FP32 scores are closer than a rigorous rounding bound is RE-EVALUATED over all
// centres in FP64 from the original data, so the label returned is always the FP64 argmin with
// lowest-index tie-break -- identical to a float64 evaluation -- at FP32 speed.
// M-step: FP64 sums / counts, privatised per CTA in shared memory when k*(d+1) fits, flushed
// with FP64 atomics.
//
// Roofline: 4*d (or 8*d) bytes read + 4 bytes written per frame, 2*k*d FLOPs: HBM-bound for
// k <~ 24, FP32-FMA-bound beyond (k=1000, d=10 is 500 FLOP/B).
#include <cstdlib>
#include "dcg_common.cuh"
#include "kmeans_common.cuh"

namespace dcg {

constexpr int kKmThreads = 512;
// frames per thread: more frames amortise the shared-memory centre loads and the loop overhead of
// the screening pass; bounded by the 128-register budget of a 512-thread CTA
__host__ __device__ constexpr int km_frames_per_thread(int dp) { return dp <= 12 ? 4 : 2; }
constexpr size_t kKmSmemBudget = 200 * 1024;

struct KmSmemPlan {
  size_t centers_off, csq_off, acc_off, total;
  int smem_acc;   // number of privatised FP64 accumulator copies in shared memory (0 = global atomics)
};

static KmSmemPlan km_plan(int d, int dp, int k) {
  KmSmemPlan p;
  p.centers_off = 0;
  p.csq_off = (size_t)k * dp * sizeof(float);
  size_t o = align_up(p.csq_off + (size_t)k * sizeof(float), 16);
  p.acc_off = o;
  // FP64 shared-memory atomics are CAS loops: with few clusters every thread of the CTA hits the
  // same k*(d+1) words, so warps get private copies (warp w uses copy w % copies) while they fit
  const size_t acc_bytes = (size_t)k * (d + 1) * sizeof(double);
  const size_t room = kKmSmemBudget > o ? kKmSmemBudget - o : 0;
  p.smem_acc = (int)std::min<size_t>(room / acc_bytes, kKmThreads / 32);
  p.total = o + (size_t)p.smem_acc * acc_bytes;
  return p;
}

// DP = padded centre row in shared memory (multiple of 4 floats); DU <= DP = dimensions that enter
// the FMA chain (DU == d when d has its own instantiation, e.g. C5's d = 10 with DP = 12: 10 FMAs
// per score instead of 12; otherwise DU == DP and the padding multiplies zeros).
template <typename T, int DP, int DU>
__global__ void __launch_bounds__(kKmThreads, 1)
kmeans_step_kernel(const T* __restrict__ Y, int64_t n, int d, int64_t ld,
                   const double* __restrict__ centers, int k, int32_t* __restrict__ labels,
                   double* __restrict__ sums, double* __restrict__ counts, double* __restrict__ stats,
                   T* __restrict__ gap, int update_sums, const double* __restrict__ y_absmax, KmSmemPlan plan) {
  if ((update_sums & 2) && stats[kKmCtlOffset] != 0.0) return;       // dcg_kmeans_iterate_n: the run has stopped
  constexpr int kKmR = km_frames_per_thread(DP);
  constexpr int kKmTile = kKmThreads * kKmR;
  extern __shared__ __align__(16) unsigned char smem[];
  float* c_s = reinterpret_cast<float*>(smem + plan.centers_off);
  float* csq_s = reinterpret_cast<float*>(smem + plan.csq_off);
  double* acc_s = reinterpret_cast<double*>(smem + plan.acc_off);
  __shared__ float s_cmax2;
  __shared__ double s_yy[kKmThreads / 32][32];      // frame broadcast for the cooperative FP64 refine
  __shared__ double s_scale[2];                     // fixed-point scale 2^s and 2^-s (0: FP64 atomics)

  const int tid = threadIdx.x;
  if (tid == 0) {
    // Fixed-point M-step sums.  Shared memory has no native 64-bit atomic add (FP64 and u64 both
    // compile to a CAS loop, which at small k retries ~7 times per add: ncu, 46 % of all stall
    // samples), but 32-bit integer adds are native.  With a bound |y| <= y_absmax from the caller
    // every value is added as round(y * 2^s) in two 32-bit words (low word first, its carry goes
    // to the high word): exact integer arithmetic, independent of the order of the adds.
    // s is the largest exponent for which the sum over all frames of this CTA fits in 62 bits;
    // a value is rounded at 2^-s ~ |y|max * frames_per_cta * 2^-61 (<= 2^-44 |y|max for 128 k frames).
    double sc = 0.0, isc = 0.0;
    if (y_absmax && update_sums && plan.smem_acc) {
      const double b = *y_absmax;
      if (b >= 0.0 && b < 1e300) {
        const int64_t tiles_cta = (((n + kKmTile - 1) / kKmTile) + gridDim.x - 1) / gridDim.x;
        const int64_t frames_cta = min(n, tiles_cta * (int64_t)kKmTile);
        const int e_f = 64 - __clzll((long long)frames_cta);
        const int e_y = b > 0.0 ? ilogb(b) + 1 : 0;
        const int e = 62 - e_f - e_y;
        sc = ldexp(1.0, e);
        isc = ldexp(1.0, -e);
      }
    }
    s_scale[0] = sc;
    s_scale[1] = isc;
  }
  for (int i = tid; i < k * DP; i += kKmThreads) {
    const int j = i / DP, q = i - j * DP;
    c_s[i] = (q < d) ? (float)centers[(size_t)j * d + q] : 0.f;
  }
  if (tid == 0) s_cmax2 = 0.f;
  __syncthreads();
  for (int j = tid; j < k; j += kKmThreads) {
    double s = 0.0;
    for (int q = 0; q < d; ++q) { const double c = centers[(size_t)j * d + q]; s = fma(c, c, s); }
    csq_s[j] = (float)s;
    atomicMax(reinterpret_cast<int*>(&s_cmax2), __float_as_int((float)s));  // s >= 0: int order == float order
  }
  if (plan.smem_acc && update_sums)
    for (int i = tid; i < plan.smem_acc * k * (d + 1); i += kKmThreads) acc_s[i] = 0.0;
  double* acc_w = acc_s + (size_t)((tid >> 5) % (plan.smem_acc > 0 ? plan.smem_acc : 1)) * k * (d + 1);
  __syncthreads();
  const double fx_scale = s_scale[0];
  const bool fixed = fx_scale != 0.0;
  const float cmax2 = s_cmax2 * 1.0001f + 1e-30f;
  const float cmaxn = sqrtf(cmax2);
  // |err(score_a) - err(score_b)| <= 2 u (2d+6) (cmax2 + |y| |c|max), u = 2^-24; x1.5 margin
  const float eps_k = 1.5f * 1.1920929e-7f * (float)(2 * d + 6);

  double t_inertia = 0.0;
  unsigned int t_changed = 0, t_ties = 0;

  const int64_t ntiles = (n + kKmTile - 1) / kKmTile;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    float x[kKmR][DP];
    int64_t idx[kKmR];
    float xsq[kKmR];
    int lab_old[kKmR];          // previous labels: loaded with the frames, needed only after the scan
#pragma unroll
    for (int r = 0; r < kKmR; ++r) {
      idx[r] = tile * kKmTile + (int64_t)r * kKmThreads + tid;
      lab_old[r] = labels[min(idx[r], n - 1)];
      const T* yrow = Y + min(idx[r], n - 1) * ld;
      xsq[r] = 0.f;
#pragma unroll
      for (int q = 0; q < DP; ++q) {
        x[r][q] = (q < d) ? (float)yrow[q] : 0.f;
        xsq[r] = fmaf(x[r][q], x[r][q], xsq[r]);
      }
    }
    float best[kKmR], second[kKmR];
    int lab[kKmR];
#pragma unroll
    for (int r = 0; r < kKmR; ++r) { best[r] = INFINITY; second[r] = INFINITY; lab[r] = 0; }

#pragma unroll 2
    for (int j = 0; j < k; ++j) {
      float c[DP];
#pragma unroll
      for (int q4 = 0; q4 < DP / 4; ++q4) {
        const float4 t = reinterpret_cast<const float4*>(c_s + (size_t)j * DP)[q4];
        c[4 * q4] = t.x; c[4 * q4 + 1] = t.y; c[4 * q4 + 2] = t.z; c[4 * q4 + 3] = t.w;
      }
      const float csq = csq_s[j];
#pragma unroll
      for (int r = 0; r < kKmR; ++r) {
        float dot = 0.f;
#pragma unroll
        for (int q = 0; q < DU; ++q) dot = fmaf(x[r][q], c[q], dot);
        const float s = fmaf(-2.f, dot, csq);
        second[r] = fminf(second[r], fmaxf(s, best[r]));
        lab[r] = (s < best[r]) ? j : lab[r];
        best[r] = fminf(best[r], s);
      }
    }

#pragma unroll
    for (int r = 0; r < kKmR; ++r) {
      const bool live = idx[r] < n;
      const T* yrow = Y + min(idx[r], n - 1) * ld;
      double b = (double)best[r], s2 = (double)second[r];
      int l = lab[r];
      const float eps = eps_k * (cmax2 + sqrtf(xsq[r]) * cmaxn);
      // frames whose two best FP32 scores are within the rounding bound: FP64 over all centres,
      // one frame at a time by the whole warp (warp-uniform loop over the ballot)
      unsigned pending = __ballot_sync(0xffffffffu, live && k > 1 && !(second[r] - best[r] > eps));
      while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const int64_t row = __shfl_sync(0xffffffffu, idx[r], src);
        const int lane = tid & 31;
        __syncwarp();
        if (lane < d) s_yy[tid >> 5][lane] = (double)Y[row * ld + lane];
        __syncwarp();
        int rl; double rb, rs;
        km_refine_warp(s_yy[tid >> 5], d, centers, k, rl, rb, rs);
        if (lane == src) { l = rl; b = rb; s2 = rs; }
      }
      if (live) {
        const double g = s2 - b;
        if (k > 1 && g <= 0.0) ++t_ties;
        if (gap) gap[idx[r]] = (T)g;
        if (lab_old[r] != l) { ++t_changed; labels[idx[r]] = l; }
        t_inertia += fmax(b + (double)xsq[r], 0.0);
      }
      if (update_sums) {
        // Frames of a trajectory are time-ordered, so the 32 frames of a warp usually share one
        // label: then the warp reduces each coordinate by shuffles (FP64) and one lane adds once --
        // per-lane FP64 shared-memory atomics are CAS loops that would retry up to 32 times on the
        // same word.  Mixed warps fall back to per-lane adds.
        int same = 0;
        __match_all_sync(0xffffffffu, live ? l : (-1 - (tid & 31)), &same);
        double* a = plan.smem_acc ? (acc_w + (size_t)l * (d + 1)) : nullptr;
        // float32 frames are still in registers (x[r][.]); float64 frames are re-read (L1)
        auto yval = [&](int q) -> double {
          if constexpr (sizeof(T) == 4) return (double)x[r][q]; else return (double)yrow[q];
        };
        if (same) {
#pragma unroll
          for (int q = 0; q < DP; ++q) {
            if (q >= d) break;
            const double v = warp_sum(yval(q));
            if ((tid & 31) == 0) {
              if (fixed) km_add_fixed(a + q, v * fx_scale);
              else if (a) atomicAdd(a + q, v); else atomicAdd(sums + (size_t)l * d + q, v);
            }
          }
          if ((tid & 31) == 0) {
            if (fixed) atomicAdd(reinterpret_cast<unsigned int*>(a + d), 32u);
            else if (a) atomicAdd(a + d, 32.0); else atomicAdd(counts + l, 32.0);
          }
        } else if (live) {
          if (fixed) {
#pragma unroll
            for (int q = 0; q < DP; ++q)
              if (q < d) km_add_fixed(a + q, yval(q) * fx_scale);
            atomicAdd(reinterpret_cast<unsigned int*>(a + d), 1u);
          } else if (a) {
            for (int q = 0; q < d; ++q) atomicAdd(a + q, (double)yrow[q]);
            atomicAdd(a + d, 1.0);
          } else {
            for (int q = 0; q < d; ++q) atomicAdd(sums + (size_t)l * d + q, (double)yrow[q]);
            atomicAdd(counts + l, 1.0);
          }
        }
      }
    }
  }

  // CTA-level stats
  t_inertia = warp_sum(t_inertia);
  t_changed = __reduce_add_sync(0xffffffffu, t_changed);
  t_ties = __reduce_add_sync(0xffffffffu, t_ties);
  if ((tid & 31) == 0) {
    if (t_changed) atomicAdd(stats + 0, (double)t_changed);
    atomicAdd(stats + 1, t_inertia);
    if (t_ties) atomicAdd(stats + 2, (double)t_ties);
  }
  if (plan.smem_acc && update_sums) {
    __syncthreads();
    for (int i = tid; i < k * (d + 1); i += kKmThreads) {
      double v = 0.0;
      const int j = i / (d + 1), q = i - j * (d + 1);
      if (fixed) {
        long long tot = 0;
        for (int c = 0; c < plan.smem_acc; ++c) {
          const uint2 w = *reinterpret_cast<const uint2*>(&acc_s[(size_t)c * k * (d + 1) + i]);
          tot += (long long)(((unsigned long long)w.y << 32) | w.x);
        }
        v = q < d ? (double)tot * s_scale[1] : (double)tot;
      } else {
        for (int c = 0; c < plan.smem_acc; ++c) v += acc_s[(size_t)c * k * (d + 1) + i];
      }
      if (v != 0.0) {
        if (q < d) atomicAdd(sums + (size_t)j * d + q, v);
        else atomicAdd(counts + j, v);
      }
    }
  }
}

// ---- M-step finish: centres = sums / counts (sklearn _average_centers: sums * (1 / count)) --------
// One CTA.  info[0] = number of empty clusters.  When there is none, the centres are updated in
// place and info[1] = sum ||c_new - c_old||^2 (the Lloyd driver's convergence test); with empty
// clusters nothing is touched: the driver relocates them first (rare path).
// `ctl` (dcg_kmeans_iterate_n; null otherwise) = [stop, iterations done, tolerance]: a stopped run is left
// untouched; otherwise the iteration is counted and the run stops when a cluster is empty, no label changed
// (stats[0], three doubles before info) or the centre shift is within the tolerance -- the Lloyd driver's own
// tests (statistics.py:159-197 of the reference, via scikit-learn), taken on the device.
__global__ void kmeans_update_kernel(const double* __restrict__ sums, const double* __restrict__ counts,
                                     int k, int d, double* __restrict__ centers, double* __restrict__ info,
                                     double* __restrict__ ctl) {
  if (ctl && ctl[0] != 0.0) return;
  __shared__ int s_empty;
  __shared__ double s_shift[32];
  if (threadIdx.x == 0) s_empty = 0;
  __syncthreads();
  int empty = 0;
  for (int j = threadIdx.x; j < k; j += blockDim.x) empty += counts[j] <= 0.0;
  if (empty) atomicAdd(&s_empty, empty);
  __syncthreads();
  const int n_empty = s_empty;
  double shift = 0.0;
  if (n_empty == 0) {
    for (int i = threadIdx.x; i < k * d; i += blockDim.x) {
      const int j = i / d;
      const double c_new = sums[i] * (1.0 / counts[j]);
      const double df = c_new - centers[i];
      shift = fma(df, df, shift);
      centers[i] = c_new;
    }
  }
  shift = warp_sum(shift);
  if ((threadIdx.x & 31) == 0) s_shift[threadIdx.x >> 5] = shift;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_shift[w];
    info[0] = (double)n_empty;
    info[1] = t;
    if (ctl) {
      ctl[1] += 1.0;
      if (n_empty > 0 || info[-3] == 0.0 || t <= ctl[2]) ctl[0] = 1.0;
    }
  }
}

// zero `n` doubles unless the run has stopped; a fresh run (init) also resets ctl = [0, 0, tol]
__global__ void kmeans_zero_kernel(double* __restrict__ p, int n, double* __restrict__ ctl, int init, double tol) {
  if (init && blockIdx.x == 0 && threadIdx.x == 0) { ctl[0] = 0.0; ctl[1] = 0.0; ctl[2] = tol; }
  if (!init && ctl[0] != 0.0) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0.0;
}

// ---- K3 ----------------------------------------------------------------------------------------
// thread <-> centre, frames of the CTA's tile broadcast from shared memory; FP64 throughout.
constexpr int kNcThreads = 128;
constexpr int kNcMaxTile = 1024;   // frames per CTA tile (fewer when d is large)

static int nearest_tile_rows(int d) {
  int rows = (int)((160 * 1024) / ((size_t)d * sizeof(double)));
  rows = rows / 32 * 32;
  return rows > kNcMaxTile ? kNcMaxTile : rows;
}

template <typename T>
__global__ void __launch_bounds__(kNcThreads)
nearest_partial_kernel(const T* __restrict__ Y, int64_t n, int d, int64_t ld, int tile_rows,
                       const double* __restrict__ centers, int k,
                       double* __restrict__ pdist, int64_t* __restrict__ pidx) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* y_s = reinterpret_cast<double*>(smem);   // [tile_rows][d]
  const int j = blockIdx.x * kNcThreads + threadIdx.x;
  const int64_t t0 = (int64_t)blockIdx.y * tile_rows;
  const int rows = (int)min((int64_t)tile_rows, n - t0);
  for (int i = threadIdx.x; i < rows * d; i += kNcThreads) {
    const int r = i / d, q = i - r * d;
    y_s[i] = (double)Y[(t0 + r) * ld + q];
  }
  __syncthreads();
  if (j >= k) return;
  double c[32];
  for (int q = 0; q < d; ++q) c[q] = centers[(size_t)j * d + q];
  // np.linalg.norm semantics: compare sqrt(sum sq); the sqrt is only taken when the squared
  // distance is not clearly larger than the incumbent's (monotone => cannot win otherwise).
  double best = INFINITY, guard = INFINITY;
  int besti = 0;
  for (int r = 0; r < rows; ++r) {
    double s = 0.0;
    for (int q = 0; q < d; ++q) { const double df = y_s[r * d + q] - c[q]; s = fma(df, df, s); }
    if (s <= guard) {
      const double rt = sqrt(s);
      if (rt < best) { best = rt; besti = r; guard = s * (1.0 + 1e-15); }
    }
  }
  pdist[(size_t)blockIdx.y * k + j] = best;
  pidx[(size_t)blockIdx.y * k + j] = t0 + besti;
}

__global__ void nearest_merge_kernel(const double* __restrict__ pdist, const int64_t* __restrict__ pidx,
                                     int ntiles, int k, int64_t* __restrict__ argmin) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  double best = INFINITY;
  int64_t bi = 0;
  for (int t = 0; t < ntiles; ++t) {          // tiles in frame order: strict < keeps the first index
    const double s = pdist[(size_t)t * k + j];
    if (s < best) { best = s; bi = pidx[(size_t)t * k + j]; }
  }
  argmin[j] = bi;
}

template <typename T, int DP, int DU = DP>
static int launch_kmeans(const T* Y, int64_t n, int d, int64_t ld, const double* centers, int k,
                         int32_t* labels, double* sums, double* counts, double* stats, T* gap,
                         int update_sums, const double* y_absmax, cudaStream_t st) {
  const KmSmemPlan plan = km_plan(d, DP, k);
  if (plan.total > kKmSmemBudget) return DCG_E_SHAPE;
  auto kern = kmeans_step_kernel<T, DP, DU>;
  DCG_CUDA_TRY(ensure_dynamic_smem((const void*)kern, (size_t)(plan.total)));
  int per_sm = 1;
  DCG_CUDA_TRY(cached_occupancy(&per_sm, (const void*)kern, kKmThreads, plan.total));
  if (per_sm < 1) per_sm = 1;
  const int64_t ntiles = ceil_div(n, kKmThreads * km_frames_per_thread(DP));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)kNumSMs * per_sm));
  kern<<<grid, kKmThreads, plan.total, st>>>(Y, n, d, ld, centers, k, labels, sums, counts, stats,
                                             gap, update_sums, y_absmax, plan);
  DCG_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int dispatch_kmeans(const T* Y, int64_t n, int d, int64_t ld, const double* centers, int k,
                           int32_t* labels, double* sums, double* counts, double* stats, T* gap,
                           int update_sums, const double* y_absmax, cudaStream_t st) {
#define DCG_KM(DPV) return launch_kmeans<T, DPV>(Y, n, d, ld, centers, k, labels, sums, counts, stats, gap, update_sums, y_absmax, st)
  if (d <= 4) DCG_KM(4);
  if (d <= 8) DCG_KM(8);
  if (d == 10) return launch_kmeans<T, 12, 10>(Y, n, d, ld, centers, k, labels, sums, counts, stats, gap, update_sums, y_absmax, st);
  if (d <= 12) DCG_KM(12);
  if (d <= 16) DCG_KM(16);
  if (d <= 24) DCG_KM(24);
  DCG_KM(32);
#undef DCG_KM
}

}  // namespace dcg

using namespace dcg;

extern "C" size_t dcg_kmeans_workspace_bytes(int64_t, int, int, int) { return 256; }

extern "C" int dcg_kmeans_step(const void* Y, int64_t n, int d, int64_t ld, int dtype_bytes,
                               const double* centers, int k, int32_t* labels,
                               double* sums, double* counts, double* stats, void* gap,
                               int update_sums, const double* y_absmax, void* ws, size_t ws_bytes, void* stream) {
  (void)ws; (void)ws_bytes;
  if (!Y || !centers || !labels || !stats) return DCG_E_NULL;
  if (update_sums && (!sums || !counts)) return DCG_E_NULL;
  if (n <= 0 || d < 1 || d > 32 || ld < d || k < 1) return DCG_E_SHAPE;
  if (dtype_bytes != 4 && dtype_bytes != 8) return DCG_E_MODE;
  cudaStream_t st = (cudaStream_t)stream;
  if (update_sums & 2) {
    // dcg_kmeans_iterate_n zeroes the buffer itself (conditionally: a stopped run keeps its last result)
  } else if (update_sums && counts == sums + (size_t)k * d && stats == counts + k) {
    // [sums | counts | stats] in one buffer (dcg_kmeans_iterate): one memset
    DCG_CUDA_TRY(cudaMemsetAsync(sums, 0, ((size_t)k * d + k + 3) * sizeof(double), st));
  } else {
    DCG_CUDA_TRY(cudaMemsetAsync(stats, 0, 3 * sizeof(double), st));
    if (update_sums) {
      DCG_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)k * d * sizeof(double), st));
      DCG_CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)k * sizeof(double), st));
    }
  }
  // many centres: the scan runs on the tensor cores (kmeans_mma.cu) when the data bound is known
  // (DCG_KMEANS_TC=0 forces the CUDA-core kernel)
  {
    const char* sel = getenv("DCG_KMEANS_TC");
    int rc = DCG_E_MODE;
    if (!(sel && sel[0] == '0'))
      rc = kmeans_mma_launch(Y, dtype_bytes, n, d, ld, centers, k, labels, sums, counts, stats, gap, update_sums, y_absmax, st);
    if (rc != DCG_E_MODE) return rc;
  }
  if (dtype_bytes == 4)
    return dispatch_kmeans<float>((const float*)Y, n, d, ld, centers, k, labels, sums, counts, stats,
                                  (float*)gap, update_sums, y_absmax, st);
  return dispatch_kmeans<double>((const double*)Y, n, d, ld, centers, k, labels, sums, counts, stats,
                                 (double*)gap, update_sums, y_absmax, st);
}

extern "C" int dcg_kmeans_update(const double* sums, const double* counts, int k, int d,
                                 double* centers, double* info, void* stream) {
  if (!sums || !counts || !centers || !info) return DCG_E_NULL;
  if (k < 1 || d < 1 || d > 32) return DCG_E_SHAPE;
  kmeans_update_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(sums, counts, k, d, centers, info, nullptr);
  DCG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dcg_kmeans_iterate(const void* Y, int64_t n, int d, int64_t ld, int dtype_bytes,
                                  double* centers, int k, int32_t* labels, double* work,
                                  const double* y_absmax, void* ws, size_t ws_bytes, void* stream) {
  if (!work) return DCG_E_NULL;
  double* sums = work;
  double* counts = work + (size_t)k * d;
  double* stats = counts + k;
  double* info = stats + 3;
  const int rc = dcg_kmeans_step(Y, n, d, ld, dtype_bytes, centers, k, labels, sums, counts, stats, nullptr,
                                 1, y_absmax, ws, ws_bytes, stream);
  if (rc) return rc;
  return dcg_kmeans_update(sums, counts, k, d, centers, info, stream);
}

extern "C" int dcg_kmeans_iterate_n(const void* Y, int64_t n, int d, int64_t ld, int dtype_bytes,
                                    double* centers, int k, int32_t* labels, double* work,
                                    const double* y_absmax, int iters, double tol,
                                    void* ws, size_t ws_bytes, void* stream) {
  if (!work) return DCG_E_NULL;
  if (iters < 1 || k < 1 || d < 1) return DCG_E_SHAPE;
  double* sums = work;
  double* counts = work + (size_t)k * d;
  double* stats = counts + k;
  double* info = stats + 3;
  double* ctl = stats + kKmCtlOffset;
  cudaStream_t st = (cudaStream_t)stream;
  const int nz = k * d + k + 3;
  const int zb = std::min(64, (nz + 255) / 256);
  for (int it = 0; it < iters; ++it) {
    kmeans_zero_kernel<<<zb, 256, 0, st>>>(sums, nz, ctl, it == 0, tol);
    DCG_LAUNCH_CHECK();
    const int rc = dcg_kmeans_step(Y, n, d, ld, dtype_bytes, centers, k, labels, sums, counts, stats, nullptr,
                                   3, y_absmax, ws, ws_bytes, stream);
    if (rc) return rc;
    kmeans_update_kernel<<<1, 512, 0, st>>>(sums, counts, k, d, centers, info, ctl);
    DCG_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" size_t dcg_nearest_workspace_bytes(int64_t n, int d, int k) {
  if (n <= 0 || k <= 0 || d < 1 || d > 32) return 0;
  const size_t tiles = (size_t)ceil_div(n, nearest_tile_rows(d));
  return align_up(tiles * k * sizeof(double), 256) + align_up(tiles * k * sizeof(int64_t), 256);
}

extern "C" int dcg_nearest_to_centers(const void* Y, int64_t n, int d, int64_t ld, int dtype_bytes,
                                      const double* centers, int k, int64_t* argmin,
                                      void* ws, size_t ws_bytes, void* stream) {
  if (!Y || !centers || !argmin) return DCG_E_NULL;
  if (n <= 0 || d < 1 || d > 32 || ld < d || k < 1) return DCG_E_SHAPE;
  if (dtype_bytes != 4 && dtype_bytes != 8) return DCG_E_MODE;
  if (!ws || ws_bytes < dcg_nearest_workspace_bytes(n, d, k)) return DCG_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int tr = nearest_tile_rows(d);
  const int64_t tiles = ceil_div(n, tr);
  if (tiles > 65535) return DCG_E_SHAPE;   // ~67M frames per call; callers chunk beyond that
  double* pdist = (double*)ws;
  int64_t* pidx = (int64_t*)((char*)ws + align_up((size_t)tiles * k * sizeof(double), 256));
  dim3 grid((unsigned)ceil_div(k, kNcThreads), (unsigned)tiles);
  const size_t smem = (size_t)tr * d * sizeof(double);
  if (dtype_bytes == 4) {
    DCG_CUDA_TRY(ensure_dynamic_smem((const void*)nearest_partial_kernel<float>, (size_t)(smem)));
    nearest_partial_kernel<float><<<grid, kNcThreads, smem, st>>>((const float*)Y, n, d, ld, tr, centers, k, pdist, pidx);
  } else {
    DCG_CUDA_TRY(ensure_dynamic_smem((const void*)nearest_partial_kernel<double>, (size_t)(smem)));
    nearest_partial_kernel<double><<<grid, kNcThreads, smem, st>>>((const double*)Y, n, d, ld, tr, centers, k, pdist, pidx);
  }
  DCG_LAUNCH_CHECK();
  nearest_merge_kernel<<<(unsigned)ceil_div(k, 128), 128, 0, st>>>(pdist, pidx, (int)tiles, k, argmin);
  DCG_LAUNCH_CHECK();
  return 0;
}
