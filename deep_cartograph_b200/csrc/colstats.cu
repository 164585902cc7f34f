// A2: single-pass per-feature statistics (mean, M2, min, max) of a row-major float32 matrix.
// Replaces training_df.agg(['mean','std','min','max']) (reference cv_calculator.py:295-297).
//
// Layout: X[n][ld]; a warp reads 32*VEC consecutive floats of one row (fully coalesced), the
// CTA's 8 warps take 8 different rows per step; a CTA owns a (row block) x (column strip) tile.
// Per thread: shifted FP32 sums over 4 rows at a time (shift = first row of the tile, so the sums
// are conditioned by the column's spread even when |mean| >> std), promoted to FP64, then merged
// across the CTA and across row blocks in FP64 with Chan's formula.
// HBM-bound: 4*f bytes per frame, read once.
#include "dcg_common.cuh"

namespace dcg {

constexpr int kStatWarps = 8;
constexpr int kStatThreads = kStatWarps * 32;
constexpr int kStatUnroll = 4;

struct StatPartial {  // one per (row block, column)
  double mean;
  double m2;
};

template <int VEC>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    float4 t = ldg_stream4(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = ldg_stream2(p);
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = ldg_stream1(p);
  }
}

// VEC floats of one row; the last vector of a row may be partial (f not a multiple of VEC)
template <int VEC>
__device__ __forceinline__ void load_vec_tail(const float* p, int valid, float (&v)[VEC]) {
  if (valid >= VEC) { load_vec<VEC>(p, v); return; }
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = i < valid ? ldg_stream1(p + i) : 0.f;
}

template <int VEC>
__global__ void __launch_bounds__(kStatThreads)
colstats_partial_kernel(const float* __restrict__ X, int64_t n, int f, int64_t ld, int64_t rows_per_cta,
                        StatPartial* __restrict__ part, float* __restrict__ pmin,
                        float* __restrict__ pmax) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = (blockIdx.x * 32 + lane) * VEC;            // first column of this thread
  const int64_t row_begin = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t row_end = min(n, row_begin + rows_per_cta);
  const bool active = col0 < f;
  const int valid = f - col0;                                  // columns of this thread's vector inside the matrix

  // Shifted sums: d = x - K with K = the tile's first row of this column, so every FP32 quantity
  // is of the size of the column's spread (not its mean).  FP32 partials are promoted to FP64
  // every kStatUnroll rows; the per-thread result is (mean, M2) in FP64.
  float shift[VEC], mn[VEC], mx[VEC];
  double s64[VEC], q64[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { shift[v] = 0.f; s64[v] = 0.0; q64[v] = 0.0; mn[v] = INFINITY; mx[v] = -INFINITY; }
  int cnt = 0;

  if (active) {
    const float* base = X + col0;
    load_vec_tail<VEC>(base + row_begin * ld, valid, shift);
    int64_t r = row_begin + warp;
    // unrolled: kStatUnroll independent loads in flight per thread
    for (; r + (int64_t)(kStatUnroll - 1) * kStatWarps < row_end; r += (int64_t)kStatUnroll * kStatWarps) {
      float x[kStatUnroll][VEC];
#pragma unroll
      for (int u = 0; u < kStatUnroll; ++u) load_vec_tail<VEC>(base + (r + (int64_t)u * kStatWarps) * ld, valid, x[u]);
      float s[VEC], q[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) { s[v] = 0.f; q[v] = 0.f; }
#pragma unroll
      for (int u = 0; u < kStatUnroll; ++u) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const float d = x[u][v] - shift[v];
          s[v] += d;
          q[v] = fmaf(d, d, q[v]);
          mn[v] = fminf(mn[v], x[u][v]);
          mx[v] = fmaxf(mx[v], x[u][v]);
        }
      }
      cnt += kStatUnroll;
#pragma unroll
      for (int v = 0; v < VEC; ++v) { s64[v] += (double)s[v]; q64[v] += (double)q[v]; }
    }
    for (; r < row_end; r += kStatWarps) {
      float x[VEC];
      load_vec_tail<VEC>(base + r * ld, valid, x);
      cnt += 1;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float d = x[v] - shift[v];
        s64[v] += (double)d;
        q64[v] += (double)d * (double)d;
        mn[v] = fminf(mn[v], x[v]);
        mx[v] = fmaxf(mx[v], x[v]);
      }
    }
  }
  double mean[VEC], m2[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const double c = cnt > 0 ? (double)cnt : 1.0;
    const double ms = s64[v] / c;
    mean[v] = (double)shift[v] + ms;
    m2[v] = fmax(q64[v] - s64[v] * ms, 0.0);
  }

  // CTA merge (FP64 Chan) of the 8 warps' partials, per column.
  __shared__ double s_mean[kStatWarps][32 * VEC];
  __shared__ double s_m2[kStatWarps][32 * VEC];
  __shared__ float s_cnt[kStatWarps];
  __shared__ float s_mn[kStatWarps][32 * VEC];
  __shared__ float s_mx[kStatWarps][32 * VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    s_mean[warp][lane * VEC + v] = mean[v];
    s_m2[warp][lane * VEC + v] = m2[v];
    s_mn[warp][lane * VEC + v] = mn[v];
    s_mx[warp][lane * VEC + v] = mx[v];
  }
  if (lane == 0) {
    // rows handled by this warp (same for all its lanes)
    int64_t rows = row_end - row_begin;
    int64_t c = rows > warp ? (rows - warp + kStatWarps - 1) / kStatWarps : 0;
    s_cnt[warp] = (float)c;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VEC; c += kStatThreads) {
    const int col = blockIdx.x * 32 * VEC + c;
    if (col >= f) continue;
    double na = 0.0, ma = 0.0, qa = 0.0;
    float lo = INFINITY, hi = -INFINITY;
#pragma unroll
    for (int w = 0; w < kStatWarps; ++w) {
      const double nb = (double)s_cnt[w];
      if (nb > 0.0) {
        const double mb = s_mean[w][c], qb = s_m2[w][c];
        const double nt = na + nb, dl = mb - ma;
        ma += dl * (nb / nt);
        qa += qb + dl * dl * (na * nb / nt);
        na = nt;
        lo = fminf(lo, s_mn[w][c]);
        hi = fmaxf(hi, s_mx[w][c]);
      }
    }
    const size_t o = (size_t)blockIdx.y * f + col;
    part[o].mean = ma;
    part[o].m2 = qa;
    pmin[o] = lo;
    pmax[o] = hi;
  }
}

// Chan merge of the row-block partials.  A CTA owns 32 columns; its 8 warps each merge every 8th
// row block sequentially (loads coalesced across the 32 columns), then warp 0 merges the 8 partial
// triples.  (One thread per column walking all ~150 row blocks took 0.11 ms at F = 1000: a chain of
// 150 dependent FP64 divisions on 1000 threads -- 14 % of the statistics pass.)
constexpr int kMergeWarps = 8;
__global__ void __launch_bounds__(kMergeWarps * 32)
colstats_merge_kernel(const StatPartial* __restrict__ part,
                      const float* __restrict__ pmin, const float* __restrict__ pmax,
                      int64_t n, int f, int row_blocks, int64_t rows_per_cta,
                      double* __restrict__ mean, double* __restrict__ m2,
                      float* __restrict__ minv, float* __restrict__ maxv) {
  __shared__ double s_n[kMergeWarps][32], s_m[kMergeWarps][32], s_q[kMergeWarps][32];
  __shared__ float s_lo[kMergeWarps][32], s_hi[kMergeWarps][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  double na = 0.0, ma = 0.0, qa = 0.0;
  float lo = INFINITY, hi = -INFINITY;
  if (col < f) {
    for (int b = w; b < row_blocks; b += kMergeWarps) {
      const int64_t r0 = (int64_t)b * rows_per_cta;
      const double nb = (double)min(rows_per_cta, n - r0);
      const StatPartial p = part[(size_t)b * f + col];
      const double nt = na + nb, dl = p.mean - ma;
      ma += dl * (nb / nt);
      qa += p.m2 + dl * dl * (na * nb / nt);
      na = nt;
      lo = fminf(lo, pmin[(size_t)b * f + col]);
      hi = fmaxf(hi, pmax[(size_t)b * f + col]);
    }
  }
  s_n[w][lane] = na; s_m[w][lane] = ma; s_q[w][lane] = qa; s_lo[w][lane] = lo; s_hi[w][lane] = hi;
  __syncthreads();
  if (w != 0 || col >= f) return;
  for (int k = 1; k < kMergeWarps; ++k) {
    const double nb = s_n[k][lane];
    if (nb > 0.0) {
      const double nt = na + nb, dl = s_m[k][lane] - ma;
      ma += dl * (nb / nt);
      qa += s_q[k][lane] + dl * dl * (na * nb / nt);
      na = nt;
    }
    lo = fminf(lo, s_lo[k][lane]);
    hi = fmaxf(hi, s_hi[k][lane]);
  }
  mean[col] = ma;
  m2[col] = qa;
  minv[col] = lo;
  maxv[col] = hi;
}

}  // namespace dcg

using namespace dcg;

// Rows per CTA: ~8 CTAs per SM over the (column strip) x (row block) grid, so that the per-block
// partials stay tiny (the merge walks them sequentially per column) and HBM sees long row streams.
static int64_t stat_rows_per_cta(int64_t n, int f) {
  const int64_t strips = ceil_div(f, 128);
  const int64_t want_blocks = std::max<int64_t>(1, ceil_div((int64_t)kNumSMs * 8, strips));
  const int64_t step = (int64_t)kStatWarps * kStatUnroll;
  return std::max<int64_t>(8 * step, ceil_div(ceil_div(n, want_blocks), step) * step);
}

extern "C" size_t dcg_colstats_workspace_bytes(int64_t n, int f) {
  if (n <= 0 || f <= 0) return 0;
  const size_t rb = (size_t)ceil_div(n, stat_rows_per_cta(n, f));
  return align_up(rb * f * sizeof(StatPartial), 256) + 2 * align_up(rb * f * sizeof(float), 256);
}

extern "C" int dcg_colstats_f32(const float* X, int64_t n, int f, int64_t ld,
                                double* mean, double* m2, float* minv, float* maxv,
                                void* ws, size_t ws_bytes, void* stream) {
  if (!X || !mean || !m2 || !minv || !maxv) return DCG_E_NULL;
  if (n <= 0 || f <= 0 || ld < f) return DCG_E_SHAPE;
  if (!ws || ws_bytes < dcg_colstats_workspace_bytes(n, f)) return DCG_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rpc = stat_rows_per_cta(n, f);
  const int rb = (int)ceil_div(n, rpc);
  char* w = (char*)ws;
  StatPartial* part = (StatPartial*)w;
  w += align_up((size_t)rb * f * sizeof(StatPartial), 256);
  float* pmin = (float*)w;
  w += align_up((size_t)rb * f * sizeof(float), 256);
  float* pmax = (float*)w;
  const int vec = row_vec_width(X, ld);      // f need not be a multiple: the last vector of a row is guarded
  dim3 grid((unsigned)ceil_div(f, 32 * vec), (unsigned)rb);
  if (vec == 4)
    colstats_partial_kernel<4><<<grid, kStatThreads, 0, st>>>(X, n, f, ld, rpc, part, pmin, pmax);
  else if (vec == 2)
    colstats_partial_kernel<2><<<grid, kStatThreads, 0, st>>>(X, n, f, ld, rpc, part, pmin, pmax);
  else
    colstats_partial_kernel<1><<<grid, kStatThreads, 0, st>>>(X, n, f, ld, rpc, part, pmin, pmax);
  DCG_LAUNCH_CHECK();
  colstats_merge_kernel<<<(unsigned)ceil_div(f, 32), kMergeWarps * 32, 0, st>>>(part, pmin, pmax, n, f, rb, rpc,
                                                                    mean, m2, minv, maxv);
  DCG_LAUNCH_CHECK();
  return 0;
}


// ---- statistics of frame shards ------------------------------------------------------------
// pack: [n | mean f | m2 f | min f | max f] as doubles, the record every rank contributes to the all-gather
__global__ void stats_pack_kernel(double n, const double* __restrict__ mean, const double* __restrict__ m2,
                                  const float* __restrict__ mn, const float* __restrict__ mx, int f,
                                  double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) out[0] = n;
  if (j >= f) return;
  out[1 + j] = mean[j];
  out[1 + f + j] = m2[j];
  out[1 + 2 * f + j] = (double)mn[j];
  out[1 + 3 * f + j] = (double)mx[j];
}

// merge of `world` packed records (Chan et al., FP64): N = sum n_r, mean = sum n_r mean_r / N,
// M2 = sum M2_r + sum n_r (mean_r - mean)^2, min / max over the ranks that hold frames.
__global__ void stats_merge_kernel(const double* __restrict__ all, int world, int f,
                                   double* __restrict__ mean, double* __restrict__ m2,
                                   float* __restrict__ mn, float* __restrict__ mx, double* __restrict__ n_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= f) return;
  const size_t rec = 1 + 4 * (size_t)f;
  double N = 0.0, s = 0.0;
  for (int r = 0; r < world; ++r) {
    const double n = all[r * rec];
    if (n > 0.0) { N += n; s += n * all[r * rec + 1 + j]; }
  }
  const double mu = s / N;
  double q = 0.0, lo = INFINITY, hi = -INFINITY;
  for (int r = 0; r < world; ++r) {
    const double n = all[r * rec];
    if (!(n > 0.0)) continue;
    const double d = all[r * rec + 1 + j] - mu;
    q += all[r * rec + 1 + f + j] + n * d * d;
    lo = fmin(lo, all[r * rec + 1 + 2 * f + j]);
    hi = fmax(hi, all[r * rec + 1 + 3 * f + j]);
  }
  mean[j] = mu;
  m2[j] = q;
  mn[j] = (float)lo;
  mx[j] = (float)hi;
  if (j == 0 && n_out) *n_out = N;
}

extern "C" int dcg_stats_pack(double n, const double* mean, const double* m2, const float* minv, const float* maxv,
                              int f, double* packed, void* stream) {
  if (!mean || !m2 || !minv || !maxv || !packed) return DCG_E_NULL;
  if (f <= 0 || n < 0) return DCG_E_SHAPE;
  stats_pack_kernel<<<(unsigned)dcg::ceil_div(f, 256), 256, 0, (cudaStream_t)stream>>>(n, mean, m2, minv, maxv, f, packed);
  DCG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dcg_stats_merge(const double* all_packed, int world, int f, double* mean, double* m2,
                               float* minv, float* maxv, double* n_total, void* stream) {
  if (!all_packed || !mean || !m2 || !minv || !maxv) return DCG_E_NULL;
  if (f <= 0 || world <= 0) return DCG_E_SHAPE;
  stats_merge_kernel<<<(unsigned)dcg::ceil_div(f, 128), 128, 0, (cudaStream_t)stream>>>(all_packed, world, f, mean, m2, minv,
                                                                                        maxv, n_total);
  DCG_LAUNCH_CHECK();
  return 0;
}
