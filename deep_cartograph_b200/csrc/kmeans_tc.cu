// K1 on the tensor cores: the KMeans E-step for many centres (C5: k = 1000, d = 10), where the
// CUDA-core kernel is FP32-FMA-bound (2*k*d FLOP against 4*d + 4 bytes per frame = 500 FLOP/B).
// Replaces the same sklearn lloyd_iter as kmeans.cu (statistics.py:159-197,
// sklearn _k_means_lloyd.pyx:193-214), same outputs, same exact-label guarantee.
//
// The scores of 128 frames against 256 centres,  score[t, j] = ||c_j||^2 - 2 y_t . c_j,  are ONE
// GEMM tile  [y_t, 1] x [-2 c_j ; ||c_j||^2]  with K = d + 1 <= 16: a single tcgen05.mma K-step
// (kind::f16, M = 128, N = 256, K = 16, accumulator in TMEM).  Precision: both operands are
// split into FP16 hi + lo pieces (11 bits each, three MMAs: hi*hi + hi*lo + lo*hi) after scaling
// the data and the centres by a power of two so that everything lies in [-1, 1]; the result
// carries FP32-like error.  As in the CUDA-core kernel this is only a SCREEN: a frame whose two
// best screened scores are closer than a rigorous bound on that error is re-evaluated over all
// centres in FP64 from the original data by its whole warp, so the label is always the FP64
// arg-min with the lowest-index tie-break.  The bound, in scaled units, with
// B = max_j ||c_j||^2 + 2 |y| max_j |c_j| >= |score|:  dropped lo*lo and split residuals
// 3 * 2^-24 * 2|y||c|, float32 rounding of centres / float64 frames 2 * 2^-24 * 2|y||c|, ||c||^2
// pieces 2 * 2^-24 * ||c||^2, FP16 subnormal floor 3 * 2^-25 * d, three truncating accumulator adds
// plus the in-MMA product sums 6 * 2^-23 * B: <= 9.5 * 2^-23 * B + 0.75 * 2^-23 * d per score, twice
// that for a difference of two, times 1.5:  eps = 3.4e-6 * B + 3e-7 * d.
//
// Per CTA (persistent, one per SM): the split centres live in shared memory as the B operand for
// the whole kernel (16 KB per 256 centres); per 128-frame tile 4 warps load their frames, stage
// the split A operand (8 KB) and later finish the frames (refine, labels, statistics, fixed-point
// M-step sums); one warp issues 3 MMAs per 256-centre block into a double-buffered TMEM
// accumulator (2 x 256 columns), so the MMAs of block j + 1 run under the scan of block j; 8 warps
// scan the accumulator (tcgen05.ld, 32 columns at a time; two warps per TMEM lane quarter, each
// takes half of the columns) keeping (best, second, label) per frame: 5 ALU ops per score instead
// of the 17 (10 FMA + 7) of the CUDA-core kernel.  The scan is what bounds the kernel
// (5 * k / 128 cycles per frame and SM); the tensor work is 3 * 128 cycles per 128 x 256 block.
#include <cuda_fp16.h>
#include "dcg_common.cuh"
#include "kmeans_common.cuh"
#include "tc_common.cuh"

namespace dcg {

using namespace tc;

namespace {

constexpr int kTile = 128;                    // frames per tile = UMMA M
constexpr int kNt = 256;                      // centres per accumulator block = UMMA N
constexpr int kMmaWarp = 8;
constexpr int kThreads = 9 * 32;              // 8 scan warps (the first 4 also stage / finish frames) + 1 MMA warp
constexpr int kAPiece = 2 * kTile * 16;       // one piece (hi or lo): 2 K-groups x 128 rows x 16 B
constexpr int kABytes = 2 * kAPiece;
constexpr int kBPiece = 2 * kNt * 16;
constexpr int kBTile = 2 * kBPiece;           // 16 KB per 256 centres
constexpr int kMaxNt = 8;                     // k <= 2048
constexpr size_t kSmemBudget = 200 * 1024;
constexpr float kDummyScore = 1000.f;         // padding centres (scaled scores are <= 3 d <= 45)

struct Plan {
  int nt, kpad;
  size_t a_off, merge_off, yy_off, acc_off, total;
  int copies;                                 // privatised fixed-point accumulator copies (0: global FP64 atomics)
};

Plan make_plan(int d, int k) {
  Plan p;
  p.kpad = (k + 31) / 32 * 32;
  p.nt = (p.kpad + kNt - 1) / kNt;
  p.a_off = (size_t)p.nt * kBTile;
  p.merge_off = p.a_off + kABytes;
  p.yy_off = p.merge_off + 2 * 4 * 32 * 12;
  p.acc_off = p.yy_off + 4 * 32 * sizeof(double);
  const size_t acc_bytes = (size_t)k * (d + 1) * sizeof(double);
  const size_t room = kSmemBudget > p.acc_off ? kSmemBudget - p.acc_off : 0;
  p.copies = (int)std::min<size_t>(room / acc_bytes, 4);
  p.total = p.acc_off + (size_t)p.copies * acc_bytes + 128;
  return p;
}

__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16 instruction descriptor: FP16 operands, FP32 accumulate, both operands K-major, dense
__device__ __forceinline__ uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void pair_sync(int q) { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); }

// v = hi + lo (+ <= 2^-24 |v|, or 2^-25 absolute below the FP16 normal range)
__device__ __forceinline__ void split_f16(float v, unsigned short& hi, unsigned short& lo) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(l);
}
__device__ __forceinline__ void st_shared_v4u(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  st_shared_v4(smem_u32(p), a, b, c, d);
}

// merge a (best, second, label) triple into another one; screening only (ties are refined)
__device__ __forceinline__ void merge3(float& b, float& s, int& l, float ob, float os, int ol) {
  if (ob < b) { s = fminf(b, os); b = ob; l = ol; }
  else s = fminf(s, ob);
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
kmeans_tc_kernel(const T* __restrict__ Y, int64_t n, int d, int64_t ld,
                 const double* __restrict__ centers, int k, int32_t* __restrict__ labels,
                 double* __restrict__ sums, double* __restrict__ counts, double* __restrict__ stats,
                 T* __restrict__ gap, int update_sums, const double* __restrict__ y_absmax, Plan plan) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* b_s = smem;
  unsigned char* a_s = smem + plan.a_off;
  float* merge_s = reinterpret_cast<float*>(smem + plan.merge_off);
  double* yy_s = reinterpret_cast<double*>(smem + plan.yy_off);
  double* acc_s = reinterpret_cast<double*>(smem + plan.acc_off);
  __shared__ uint64_t a_full, acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_cmax, s_cmax2;
  __shared__ double s_par[4];          // data scale 2^-e, its inverse square, fixed-point scale, its inverse

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&a_full, 4);
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 8); mbar_init(&acc_empty[1], 8);
    fence_barrier_init();
    s_cmax = 0.f;
    s_cmax2 = 0.f;
  }
  if (warp == kMmaWarp) tmem_alloc(&tmem_base_s, 512);
  __syncthreads();
  {
    float m = 0.f;
    for (int i = tid; i < k * d; i += kThreads) m = fmaxf(m, fabsf((float)centers[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) atomicMax(reinterpret_cast<int*>(&s_cmax), __float_as_int(m));
    float m2 = 0.f;
    for (int j = tid; j < k; j += kThreads) {
      double sq = 0.0;
      for (int q = 0; q < d; ++q) { const double c = centers[(size_t)j * d + q]; sq = fma(c, c, sq); }
      m2 = fmaxf(m2, (float)sq * 1.0000002f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    if (lane == 0) atomicMax(reinterpret_cast<int*>(&s_cmax2), __float_as_int(m2));
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
  const double scale = s_par[0], inv_s2 = s_par[1], fx_scale = s_par[2];
  const float scale_f = (float)scale;

  // ---- B operand: split centres, K-major [block][piece][K-group][256 rows][8 halves] ----------
  for (int i = tid; i < plan.kpad * 2; i += kThreads) {
    const int j = i >> 1, kg = i & 1;
    unsigned short hi[8], lo[8];
    double csq = 0.0;
    if (j < k)
      for (int q = 0; q < d; ++q) { const double c = centers[(size_t)j * d + q] * scale; csq = fma(c, c, csq); }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int q = 8 * kg + e;
      float v = 0.f;
      if (j < k) {
        if (q < d) v = -2.f * ((float)centers[(size_t)j * d + q] * scale_f);
        else if (q == d) v = (float)csq;
      } else if (q == d) {
        v = kDummyScore;
      }
      split_f16(v, hi[e], lo[e]);
    }
    unsigned char* dst = b_s + (size_t)(j >> 8) * kBTile + (size_t)kg * (kNt * 16) + (size_t)(j & 255) * 16;
    st_shared_v4u(dst, hi[0] | (uint32_t)hi[1] << 16, hi[2] | (uint32_t)hi[3] << 16,
                  hi[4] | (uint32_t)hi[5] << 16, hi[6] | (uint32_t)hi[7] << 16);
    st_shared_v4u(dst + kBPiece, lo[0] | (uint32_t)lo[1] << 16, lo[2] | (uint32_t)lo[3] << 16,
                  lo[4] | (uint32_t)lo[5] << 16, lo[6] | (uint32_t)lo[7] << 16);
  }
  const bool fixed = plan.copies > 0 && update_sums;
  if (fixed)
    for (int i = tid; i < plan.copies * k * (d + 1); i += kThreads) acc_s[i] = 0.0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  uint32_t it = 0, gc = 0;      // tiles / accumulator blocks processed so far by this CTA (all roles agree)

  if (warp == kMmaWarp) {
    // =========================== MMA issue ===========================
    const uint32_t a_base = smem_u32(a_s), b_base = smem_u32(b_s);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      mbar_wait(&a_full, it & 1);
      tc_fence_after();
      for (int nb = 0; nb < plan.nt; ++nb, ++gc) {
        const uint32_t buf = gc & 1;
        mbar_wait(&acc_empty[buf], ((gc >> 1) & 1) ^ 1);
        tc_fence_after();
        if (elect_one_sync()) {
          const int ncols = min(kNt, plan.kpad - nb * kNt);
          const uint32_t idesc = idesc_f16(kTile, ncols);
          const uint64_t a_hi = make_smem_desc(a_base, kTile * 16, 128);
          const uint64_t a_lo = make_smem_desc(a_base + kAPiece, kTile * 16, 128);
          const uint64_t b_hi = make_smem_desc(b_base + nb * kBTile, kNt * 16, 128);
          const uint64_t b_lo = make_smem_desc(b_base + nb * kBTile + kBPiece, kNt * 16, 128);
          const uint32_t dcol = tmem + buf * kNt;
          mma_f16_ss(dcol, a_hi, b_hi, idesc, 0);
          mma_f16_ss(dcol, a_hi, b_lo, idesc, 1);
          mma_f16_ss(dcol, a_lo, b_hi, idesc, 1);
          mma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== scan warps ===========================
    const int q = warp & 3, half = warp >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    // screening bound of a frame (scaled units): eps_a * (cmax2 + 2 |y| cmax) + eps_b
    const float cmax2_s = s_cmax2 * scale_f * scale_f, cmaxn_s = sqrtf(s_cmax2) * scale_f;
    const float eps_a = 3.4e-6f, eps_b = 3e-7f * (float)d;
    double* acc_w = acc_s + (size_t)(q % (plan.copies > 0 ? plan.copies : 1)) * k * (d + 1);
    double t_inertia = 0.0;
    unsigned int t_changed = 0, t_ties = 0;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int64_t row = tile * kTile + q * 32 + lane;
      const bool live = row < n;
      const T* yrow = Y + min(row, n - 1) * ld;
      float x[16];
      float xsq = 0.f;
      int lab_old = 0;
      if (half == 0) {
        // frames of this lane quarter: load, keep for the M-step, stage the split A operand
        lab_old = labels[min(row, n - 1)];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          x[c] = (c < d && live) ? (float)yrow[c] : 0.f;
          xsq = fmaf(x[c], x[c], xsq);
        }
        unsigned short hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) split_f16(c < d ? x[c] * scale_f : (c == d ? 1.f : 0.f), hi[c], lo[c]);
        unsigned char* dst = a_s + (size_t)(q * 32 + lane) * 16;
#pragma unroll
        for (int kg = 0; kg < 2; ++kg) {
          st_shared_v4u(dst + kg * (kTile * 16), hi[8 * kg] | (uint32_t)hi[8 * kg + 1] << 16, hi[8 * kg + 2] | (uint32_t)hi[8 * kg + 3] << 16,
                        hi[8 * kg + 4] | (uint32_t)hi[8 * kg + 5] << 16, hi[8 * kg + 6] | (uint32_t)hi[8 * kg + 7] << 16);
          st_shared_v4u(dst + kAPiece + kg * (kTile * 16), lo[8 * kg] | (uint32_t)lo[8 * kg + 1] << 16, lo[8 * kg + 2] | (uint32_t)lo[8 * kg + 3] << 16,
                        lo[8 * kg + 4] | (uint32_t)lo[8 * kg + 5] << 16, lo[8 * kg + 6] | (uint32_t)lo[8 * kg + 7] << 16);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full);
      }

      float gb = INFINITY, gs = INFINITY;
      int gl = 0;
      for (int nb = 0; nb < plan.nt; ++nb, ++gc) {
        const uint32_t buf = gc & 1;
        const int ncols = min(kNt, plan.kpad - nb * kNt);
        mbar_wait(&acc_full[buf], (gc >> 1) & 1);
        tc_fence_after();
        float lb[4], ls[4];
        int ll[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { lb[c] = INFINITY; ls[c] = INFINITY; ll[c] = 0; }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c0 = half * 128 + 32 * i;
          if (c0 < ncols) {
            uint32_t v[32];
            tmem_ld_x32(tmem + buf * kNt + lane_base + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const int c = jj & 3;
              const float s = __uint_as_float(v[jj]);
              const bool p = s < lb[c];
              const float t = fmaxf(s, lb[c]);
              lb[c] = fminf(s, lb[c]);
              ls[c] = fminf(ls[c], t);
              ll[c] = p ? (32 * i + jj) : ll[c];
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        float b = lb[0], s2 = ls[0];
        int l = ll[0];
#pragma unroll
        for (int c = 1; c < 4; ++c) merge3(b, s2, l, lb[c], ls[c], ll[c]);
        merge3(gb, gs, gl, b, s2, l + half * 128 + nb * kNt);
      }

      // the two warps of a lane quarter scanned different columns: combine
      float* mbuf = merge_s + ((it & 1) * 4 + q) * 32 * 3;
      if (half == 1) {
        mbuf[lane] = gb; mbuf[32 + lane] = gs; mbuf[64 + lane] = __int_as_float(gl);
        pair_sync(q);
        continue;
      }
      pair_sync(q);
      merge3(gb, gs, gl, mbuf[lane], mbuf[32 + lane], __float_as_int(mbuf[64 + lane]));

      // ---- finish the frames (same protocol as kmeans_step_kernel) ------------------------------
      double b = (double)gb * inv_s2 , s2 = (double)gs * inv_s2;
      int l = gl;
      const float eps = fmaf(eps_a, fmaf(2.f * sqrtf(xsq) * scale_f, cmaxn_s, cmax2_s), eps_b);
      unsigned pending = __ballot_sync(0xffffffffu, live && k > 1 && !(gs - gb > eps));
      while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const int64_t rrow = __shfl_sync(0xffffffffu, row, src);
        __syncwarp();
        if (lane < d) yy_s[q * 32 + lane] = (double)Y[rrow * ld + lane];
        __syncwarp();
        int rl; double rb, rs;
        km_refine_warp(yy_s + q * 32, d, centers, k, rl, rb, rs);
        if (lane == src) { l = rl; b = rb; s2 = rs; }
      }
      if (live) {
        const double g = s2 - b;
        if (k > 1 && g <= 0.0) ++t_ties;
        if (gap) gap[row] = (T)g;
        if (lab_old != l) { ++t_changed; labels[row] = l; }
        t_inertia += fmax(b + (double)xsq, 0.0);
      }
      if (update_sums) {
        auto yval = [&](int c) -> double {
          if constexpr (sizeof(T) == 4) return (double)x[c]; else return (double)yrow[c];
        };
        int same = 0;
        __match_all_sync(0xffffffffu, live ? l : (-1 - lane), &same);
        double* a = plan.copies ? (acc_w + (size_t)l * (d + 1)) : nullptr;
        if (same) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            if (c >= d) break;
            const double v = warp_sum(yval(c));
            if (lane == 0) {
              if (fixed) km_add_fixed(a + c, v * fx_scale);
              else atomicAdd(sums + (size_t)l * d + c, v);
            }
          }
          if (lane == 0) {
            if (fixed) atomicAdd(reinterpret_cast<unsigned int*>(a + d), 32u);
            else atomicAdd(counts + l, 32.0);
          }
        } else if (live) {
          if (fixed) {
#pragma unroll
            for (int c = 0; c < 16; ++c)
              if (c < d) km_add_fixed(a + c, yval(c) * fx_scale);
            atomicAdd(reinterpret_cast<unsigned int*>(a + d), 1u);
          } else {
#pragma unroll
            for (int c = 0; c < 16; ++c)
              if (c < d) atomicAdd(sums + (size_t)l * d + c, yval(c));
            atomicAdd(counts + l, 1.0);
          }
        }
      }
    }

    if (half == 0) {
      t_inertia = warp_sum(t_inertia);
      t_changed = __reduce_add_sync(0xffffffffu, t_changed);
      t_ties = __reduce_add_sync(0xffffffffu, t_ties);
      if (lane == 0) {
        if (t_changed) atomicAdd(stats + 0, (double)t_changed);
        atomicAdd(stats + 1, t_inertia);
        if (t_ties) atomicAdd(stats + 2, (double)t_ties);
      }
    }
  }

  tc_fence_before();
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
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <typename T>
int launch(const T* Y, int64_t n, int d, int64_t ld, const double* centers, int k, int32_t* labels,
           double* sums, double* counts, double* stats, T* gap, int update_sums, const double* y_absmax,
           cudaStream_t st) {
  const Plan plan = make_plan(d, k);
  auto kern = kmeans_tc_kernel<T>;
  DCG_CUDA_TRY(ensure_dynamic_smem((const void*)kern, plan.total));
  const int64_t ntiles = ceil_div(n, kTile);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, kNumSMs));
  kern<<<grid, kThreads, plan.total, st>>>(Y, n, d, ld, centers, k, labels, sums, counts, stats, gap, update_sums,
                                           y_absmax, plan);
  DCG_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int kmeans_tc_launch(const void* Y, int dtype_bytes, int64_t n, int d, int64_t ld, const double* centers, int k,
                     int32_t* labels, double* sums, double* counts, double* stats, void* gap,
                     int update_sums, const double* y_absmax, cudaStream_t st) {
  // needs the data bound (operand scaling), K = d + 1 <= 16, and enough centres to pay for the tile
  if (!y_absmax || d > 15 || k < 64 || k > kMaxNt * kNt) return DCG_E_MODE;
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return DCG_E_MODE;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) return DCG_E_MODE;
  if (dtype_bytes == 4)
    return launch<float>((const float*)Y, n, d, ld, centers, k, labels, sums, counts, stats, (float*)gap, update_sums, y_absmax, st);
  return launch<double>((const double*)Y, n, d, ld, centers, k, labels, sums, counts, stats, (double*)gap, update_sums, y_absmax, st);
}

}  // namespace dcg
