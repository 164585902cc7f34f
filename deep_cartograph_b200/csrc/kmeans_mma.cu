// K1 with many centres, register-resident scan: the scores of 32 frames against 8 centres are
// warp-level tensor-core MMAs (mma.sync.m16n8k16, FP16 hi/lo split, K = d + 1 <= 16) whose
// accumulators ARE the scores, already in registers, so the arg-min scan costs 5 ALU operations
// per score and nothing else.  (A tcgen05 / TMEM variant -- M = 128 x N = 256 x K = 16 kind::f16 MMAs,
// scan warps on tcgen05.ld -- was measured in round 1 and dropped: every score has to come back through
// the 64 B/clk TMEM read port, 512 KB per 128-frame tile at k = 1000 = 8.2 k cycles, 7.4 ms per 12.5 M
// frames against 4.4 ms here, where the bound is the scan itself, 5 k / 128 cycles per frame and SM,
// next to 3 HMMAs per 128 scores.)
//
// Precision.  score[t, j] = ||c_j||^2 - 2 y_t . c_j is the GEMM [y_t, 1] x [-2 c_j ; ||c_j||^2] with
// K = d + 1 <= 16.  Both operands are split into FP16 hi + lo pieces (11 bits each, three MMAs:
// hi*hi + hi*lo + lo*hi) after scaling the data and the centres by a power of two so that everything
// lies in [-1, 1]; the result carries FP32-like error.  As in the CUDA-core kernel this is only a
// SCREEN: a frame whose two best screened scores are closer than a rigorous bound on that error is
// re-evaluated over all centres in FP64 from the original data by its whole warp, so the label is
// always the FP64 arg-min with the lowest-index tie-break.  The bound, in scaled units, with
// B = max_j ||c_j||^2 + 2 |y| max_j |c_j| >= |score|:  dropped lo*lo and split residuals
// 3 * 2^-24 * 2|y||c|, float32 rounding of centres / float64 frames 2 * 2^-24 * 2|y||c|, ||c||^2
// pieces 2 * 2^-24 * ||c||^2, FP16 subnormal floor 3 * 2^-25 * d, accumulator adds plus the in-MMA
// product sums 6 * 2^-23 * B: <= 9.5 * 2^-23 * B + 0.75 * 2^-23 * d per score, twice that for a
// difference of two, times 1.5:  eps = 3.4e-6 * B + 3e-7 * d.
//
// Per warp: 32 frames (two m16 tiles).  Lane (g = lane / 4, t = lane % 4) holds, of rows g, g + 8,
// g + 16, g + 24, the coordinates q in {2t, 2t+1, 2t+8, 2t+9} -- raw (for the M-step sums) and as
// split A fragments.  The split centres sit in shared memory in fragment order (64 B per centre,
// one conflict-free 16-byte load per lane and 8 centres).  After the scan the four lanes of a
// group merge their partial (best, second, label) triples by shuffles and lane t finishes row
// slot t: screening test, warp-cooperative FP64 refine, label, statistics; every lane adds its own
// coordinates of its four rows to the fixed-point partial sums.
#include <cstdlib>
#include <cuda_fp16.h>
#include "dcg_common.cuh"
#include "kmeans_common.cuh"

namespace dcg {

namespace {

constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr int kWarpFrames = 32;
constexpr int kTile = kWarps * kWarpFrames;     // frames per CTA tile
constexpr int kMaxK = 2048;
constexpr size_t kSmemBudget = 200 * 1024;
constexpr float kDummyScore = 1000.f;           // padding centres (scaled scores are <= 3 d <= 45)

struct Plan {
  int kpad;                                     // centres rounded up to 8
  size_t yy_off, acc_off, total;
  int copies;                                   // privatised fixed-point accumulator copies (0: global FP64 atomics)
};

Plan make_plan(int d, int k) {
  Plan p;
  p.kpad = (k + 7) / 8 * 8;
  p.yy_off = (size_t)p.kpad * 64;
  p.acc_off = p.yy_off + kWarps * 32 * sizeof(double);
  const size_t acc_bytes = (size_t)k * (d + 1) * sizeof(double);
  const size_t room = kSmemBudget > p.acc_off ? kSmemBudget - p.acc_off : 0;
  p.copies = (int)std::min<size_t>(room / acc_bytes, kWarps);
  p.total = p.acc_off + (size_t)p.copies * acc_bytes + 128;
  return p;
}

__device__ __forceinline__ uint32_t pack_split(float v0, float v1, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
  lo = *reinterpret_cast<const uint32_t*>(&l);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void hmma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// merge a (best, second, label) triple into another one; screening only (near-ties are refined)
__device__ __forceinline__ void merge3(float& b, float& s, int& l, float ob, float os, int ol) {
  if (ob < b) { s = fminf(b, os); b = ob; l = ol; }
  else s = fminf(s, ob);
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
kmeans_mma_kernel(const T* __restrict__ Y, int64_t n, int d, int64_t ld,
                  const double* __restrict__ centers, int k, int32_t* __restrict__ labels,
                  double* __restrict__ sums, double* __restrict__ counts, double* __restrict__ stats,
                  T* __restrict__ gap, int update_sums, const double* __restrict__ y_absmax, Plan plan) {
  if ((update_sums & 2) && stats[kKmCtlOffset] != 0.0) return;       // dcg_kmeans_iterate_n: the run has stopped
  extern __shared__ __align__(128) unsigned char smem[];
  uint4* b_s = reinterpret_cast<uint4*>(smem);                         // [n-tile of 8 centres][lane] -> bhi0 bhi1 blo0 blo1
  double* yy_s = reinterpret_cast<double*>(smem + plan.yy_off);
  double* acc_s = reinterpret_cast<double*>(smem + plan.acc_off);
  __shared__ float s_cmax, s_cmax2;
  __shared__ double s_par[4];          // data scale 2^-e, its inverse square, fixed-point scale, its inverse

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  if (tid == 0) { s_cmax = 0.f; s_cmax2 = 0.f; }
  __syncthreads();
  {
    float m = 0.f;
    for (int i = tid; i < k * d; i += kThreads) m = fmaxf(m, fabsf((float)centers[i]));
    float m2 = 0.f;
    for (int j = tid; j < k; j += kThreads) {
      double sq = 0.0;
      for (int q = 0; q < d; ++q) { const double c = centers[(size_t)j * d + q]; sq = fma(c, c, sq); }
      m2 = fmaxf(m2, (float)sq * 1.0000002f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    }
    if (lane == 0) {
      atomicMax(reinterpret_cast<int*>(&s_cmax), __float_as_int(m));
      atomicMax(reinterpret_cast<int*>(&s_cmax2), __float_as_int(m2));
    }
  }
  __syncthreads();
  const int64_t ntiles = (n + kTile - 1) / kTile;
  if (tid == 0) {
    const double yb = *y_absmax;
    const double bound = fmax(yb, (double)s_cmax * 1.0000002);
    const int e = bound > 0.0 ? ilogb(bound) + 1 : 0;          // |y|, |c| <= 2^e
    s_par[0] = ldexp(1.0, -e);
    s_par[1] = ldexp(1.0, 2 * e);
    // fixed-point M-step sums (see kmeans.cu): 2^s such that this CTA's total fits 62 bits
    const int64_t frames_cta = min(n, ((ntiles + gridDim.x - 1) / gridDim.x) * (int64_t)kTile);
    const int e_f = 64 - __clzll((long long)frames_cta);
    const int e_y = yb > 0.0 ? ilogb(yb) + 1 : 0;
    s_par[2] = ldexp(1.0, 62 - e_f - e_y);
    s_par[3] = ldexp(1.0, -(62 - e_f - e_y));
  }
  __syncthreads();
  const double inv_s2 = s_par[1], fx_scale = s_par[2];
  const float scale_f = (float)s_par[0];

  // ---- B fragments: for n-tile j8 and lane (g, t): centre 8 j8 + g, rows k = 2t, 2t+1 (b0) and
  //      2t+8, 2t+9 (b1) of  [-2 c ; ||c||^2 ; 0]  (scaled), hi and lo pieces --------------------
  for (int i = tid; i < plan.kpad * 4; i += kThreads) {
    const int j = i >> 2, tt = i & 3;             // centre, t
    double csq = 0.0;
    if (j < k)
      for (int q = 0; q < d; ++q) { const double c = centers[(size_t)j * d + q] * s_par[0]; csq = fma(c, c, csq); }
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int q = 2 * tt + (e & 1) + 8 * (e >> 1);
      float x = 0.f;
      if (j < k) {
        if (q < d) x = -2.f * ((float)centers[(size_t)j * d + q] * scale_f);
        else if (q == d) x = (float)csq;
      } else if (q == d) {
        x = kDummyScore;
      }
      v[e] = x;
    }
    uint32_t lo0, lo1;
    const uint32_t hi0 = pack_split(v[0], v[1], lo0);
    const uint32_t hi1 = pack_split(v[2], v[3], lo1);
    b_s[(size_t)(j >> 3) * 32 + (j & 7) * 4 + tt] = make_uint4(hi0, hi1, lo0, lo1);
  }
  const bool fixed = plan.copies > 0 && update_sums;
  if (fixed)
    for (int i = tid; i < plan.copies * k * (d + 1); i += kThreads) acc_s[i] = 0.0;
  __syncthreads();

  const float cmax2_s = s_cmax2 * scale_f * scale_f, cmaxn_s = sqrtf(s_cmax2) * scale_f;
  const float eps_a = 3.4e-6f, eps_b = 3e-7f * (float)d;
  double* acc_w = acc_s + (size_t)(warp % (plan.copies > 0 ? plan.copies : 1)) * k * (d + 1);
  double* yy_w = yy_s + warp * 32;
  const int n8 = plan.kpad >> 3;
  // this lane's coordinates
  int qv[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) qv[e] = 2 * t + (e & 1) + 8 * (e >> 1);

  double t_inertia = 0.0;
  unsigned int t_changed = 0, t_ties = 0;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * kTile + (int64_t)warp * kWarpFrames;
    if (base >= n) continue;
    // ---- frames: rows g, g+8, g+16, g+24 of the warp's 32, this lane's 4 coordinates ----------
    float x[4][4];
    float xsq[4];
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int64_t row = base + g + 8 * r;
      const T* yrow = Y + min(row, n - 1) * ld;
      xsq[r] = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        x[r][e] = (qv[e] < d && row < n) ? (float)yrow[qv[e]] : 0.f;
        xsq[r] = fmaf(x[r][e], x[r][e], xsq[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float s[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) s[e] = qv[e] < d ? x[r][e] * scale_f : (qv[e] == d ? 1.f : 0.f);
      // a0: (row g, k 2t..), a1: (row g+8, k 2t..), a2: (row g, k 2t+8..), a3: (row g+8, k 2t+8..)
      const int m = r >> 1, up = r & 1;
      ah[m][up] = pack_split(s[0], s[1], al[m][up]);
      ah[m][2 + up] = pack_split(s[2], s[3], al[m][2 + up]);
    }
    // full |y|^2 of the four rows (sum over the four lanes of the group)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      xsq[r] += __shfl_xor_sync(0xffffffffu, xsq[r], 1);
      xsq[r] += __shfl_xor_sync(0xffffffffu, xsq[r], 2);
    }

    // pull the next tile's frames and labels into L2 while this one is scanned (a register prefetch
    // costs the scan loop its registers: measured 4.4 -> 5.0 ms at k = 1000)
    {
      const int64_t nbase = (tile + gridDim.x) * kTile + (int64_t)warp * kWarpFrames;
      if (nbase < n) {
        const int64_t nrow = min(nbase + lane, n - 1);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(Y + nrow * ld));
        if (lane == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(labels + nrow));
      }
    }
    // ---- scan: 8 centres per step, scores straight from the MMA accumulators -------------------
    float bb[4], ss[4];
    int ll[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { bb[r] = INFINITY; ss[r] = INFINITY; ll[r] = 0; }
    int jl = 2 * t;
    const uint4* bp = b_s + lane;
#pragma unroll 2
    for (int j8 = 0; j8 < n8; ++j8, jl += 8, bp += 32) {
      const uint4 b = *bp;
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        hmma16816(c, ah[m], b.x, b.y);
        hmma16816(c, al[m], b.x, b.y);
        hmma16816(c, ah[m], b.z, b.w);
        // c0, c1: row g + 16 m, centres jl, jl + 1;  c2, c3: row g + 8 + 16 m
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = 2 * m + (u >> 1);
          const float s = c[u];
          const bool p = s < bb[r];
          const float tmx = fmaxf(s, bb[r]);
          bb[r] = fminf(s, bb[r]);
          ss[r] = fminf(ss[r], tmx);
          ll[r] = p ? (jl + (u & 1)) : ll[r];
        }
      }
    }
    // the four lanes of a group saw different centres: merge (all four end up with the result)
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float ob = __shfl_xor_sync(0xffffffffu, bb[r], o);
        const float os = __shfl_xor_sync(0xffffffffu, ss[r], o);
        const int ol = __shfl_xor_sync(0xffffffffu, ll[r], o);
        merge3(bb[r], ss[r], ll[r], ob, os, ol);
      }
    }

    // ---- lane t finishes row slot t (row g + 8 t of the warp's 32) ------------------------------
    const float gb = t == 0 ? bb[0] : t == 1 ? bb[1] : t == 2 ? bb[2] : bb[3];
    const float gs = t == 0 ? ss[0] : t == 1 ? ss[1] : t == 2 ? ss[2] : ss[3];
    int l = t == 0 ? ll[0] : t == 1 ? ll[1] : t == 2 ? ll[2] : ll[3];
    const float myxsq = t == 0 ? xsq[0] : t == 1 ? xsq[1] : t == 2 ? xsq[2] : xsq[3];
    const int64_t row = base + g + 8 * t;
    const bool live = row < n;
    double b = (double)gb * inv_s2, s2 = (double)gs * inv_s2;
    const float eps = fmaf(eps_a, fmaf(2.f * sqrtf(myxsq) * scale_f, cmaxn_s, cmax2_s), eps_b);
    unsigned pending = __ballot_sync(0xffffffffu, live && k > 1 && !(gs - gb > eps));
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      const int64_t rrow = __shfl_sync(0xffffffffu, row, src);
      __syncwarp();
      if (lane < d) yy_w[lane] = (double)Y[rrow * ld + lane];
      __syncwarp();
      int rl; double rb, rs;
      km_refine_warp(yy_w, d, centers, k, rl, rb, rs);
      if (lane == src) { l = rl; b = rb; s2 = rs; }
    }
    if (live) {
      const double gg = s2 - b;
      if (k > 1 && gg <= 0.0) ++t_ties;
      if (gap) gap[row] = (T)gg;
      if (labels[row] != l) { ++t_changed; labels[row] = l; }
      t_inertia += fmax(b + (double)myxsq, 0.0);
    }
    if (update_sums) {
      // labels of the group's four rows (row slot r is owned by lane t = r)
      int lr[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) lr[r] = __shfl_sync(0xffffffffu, l, (lane & ~3) | r);
      int same = 0;
      __match_all_sync(0xffffffffu, live ? l : (-1 - lane), &same);
      if (same) {
        // one label for the warp's 32 frames: reduce over rows, lanes of group 0 add once
        double* a = plan.copies ? (acc_w + (size_t)l * (d + 1)) : nullptr;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          double v;
          if constexpr (sizeof(T) == 4) {
            v = (double)x[0][e] + (double)x[1][e] + (double)x[2][e] + (double)x[3][e];
          } else {
            v = 0.0;
#pragma unroll
            for (int r = 0; r < 4; ++r) v += qv[e] < d ? (double)Y[(base + g + 8 * r) * ld + qv[e]] : 0.0;
          }
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (g == 0 && qv[e] < d) {
            if (fixed) km_add_fixed(a + qv[e], v * fx_scale);
            else atomicAdd(sums + (size_t)l * d + qv[e], v);
          }
        }
        if (lane == 0) {
          if (fixed) atomicAdd(reinterpret_cast<unsigned int*>(a + d), 32u);
          else atomicAdd(counts + l, 32.0);
        }
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int64_t rr = base + g + 8 * r;
          if (rr >= n) continue;
          double* a = plan.copies ? (acc_w + (size_t)lr[r] * (d + 1)) : nullptr;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (qv[e] >= d) continue;
            double v;
            if constexpr (sizeof(T) == 4) v = (double)x[r][e]; else v = (double)Y[rr * ld + qv[e]];
            if (fixed) km_add_fixed(a + qv[e], v * fx_scale);
            else atomicAdd(sums + (size_t)lr[r] * d + qv[e], v);
          }
          if (t == r) {
            if (fixed) atomicAdd(reinterpret_cast<unsigned int*>(a + d), 1u);
            else atomicAdd(counts + lr[r], 1.0);
          }
        }
      }
    }
  }

  t_inertia = warp_sum(t_inertia);
  t_changed = __reduce_add_sync(0xffffffffu, t_changed);
  t_ties = __reduce_add_sync(0xffffffffu, t_ties);
  if (lane == 0) {
    if (t_changed) atomicAdd(stats + 0, (double)t_changed);
    atomicAdd(stats + 1, t_inertia);
    if (t_ties) atomicAdd(stats + 2, (double)t_ties);
  }
  __syncthreads();
  if (fixed) {
    for (int i = tid; i < k * (d + 1); i += kThreads) {
      const int j = i / (d + 1), c = i - j * (d + 1);
      long long tot = 0;
      for (int cp = 0; cp < plan.copies; ++cp) {
        const uint2 w = *reinterpret_cast<const uint2*>(&acc_s[(size_t)cp * k * (d + 1) + i]);
        tot += (long long)(((unsigned long long)w.y << 32) | w.x);
      }
      if (tot != 0) {
        if (c < d) atomicAdd(sums + (size_t)j * d + c, (double)tot * s_par[3]);
        else atomicAdd(counts + j, (double)tot);
      }
    }
  }
}

template <typename T>
int launch(const T* Y, int64_t n, int d, int64_t ld, const double* centers, int k, int32_t* labels,
           double* sums, double* counts, double* stats, T* gap, int update_sums, const double* y_absmax,
           cudaStream_t st) {
  const Plan plan = make_plan(d, k);
  auto kern = kmeans_mma_kernel<T>;
  DCG_CUDA_TRY(ensure_dynamic_smem((const void*)kern, plan.total));
  const int64_t ntiles = ceil_div(n, kTile);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, kNumSMs));
  kern<<<grid, kThreads, plan.total, st>>>(Y, n, d, ld, centers, k, labels, sums, counts, stats, gap, update_sums,
                                           y_absmax, plan);
  DCG_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int kmeans_mma_launch(const void* Y, int dtype_bytes, int64_t n, int d, int64_t ld, const double* centers, int k,
                      int32_t* labels, double* sums, double* counts, double* stats, void* gap,
                      int update_sums, const double* y_absmax, cudaStream_t st) {
  // needs the data bound (operand scaling), K = d + 1 <= 16, and enough centres to pay for it
  // (measured: k = 100, d = 4 is 10 % faster on the CUDA cores, k = 100, d = 10 10 % faster here)
  const char* mk = getenv("DCG_KMEANS_MMA_MINK");            // experiments: lowest k that takes this engine
  const int min_k = mk ? atoi(mk) : 0;
  if (!y_absmax || d > 15 || k > kMaxK) return DCG_E_MODE;
  if (min_k > 0 ? k < min_k : !(k >= 256 || (k >= 64 && d >= 8))) return DCG_E_MODE;
  if (dtype_bytes == 4)
    return launch<float>((const float*)Y, n, d, ld, centers, k, labels, sums, counts, stats, (float*)gap, update_sums, y_absmax, st);
  return launch<double>((const double*)Y, n, d, ld, centers, k, labels, sums, counts, stats, (double*)gap, update_sums, y_absmax, st);
}

}  // namespace dcg
