// tcgen05 covariance engine, CTA-pair edition (DCG_COV_TC_3XTF32 / DCG_COV_TC_1XTF32 / DCG_COV_TC_3XF16).
//
// Computes, for 256 x 256 super-tiles (I, J) of the feature axis and a range of frames,
//     S0[I,J] += sum_t z_t[I] (x) z_t[J]        or        St[I,J] += sum_t z_t[I] (x) z_{t+lag}[J]
// (one matrix per work item) as a dense contraction over the FRAME axis on the 5th-generation
// tensor cores: tcgen05.mma.cta_group::2.kind::tf32, M = N = 256, K = 8 frames per instruction.
// Two CTAs of a cluster (one TPC) share every MMA: CTA r stages the 128 A rows I0+128r.. and the
// 128 B rows J0+128r.. and owns accumulator rows 128r.. in its own TMEM.  Per staged operand
// element this does twice the tensor work of a 128 x 128 single-CTA tile, and the MMA reads only
// 64 B/clk of shared memory per SM (the single-CTA 128 x 128 x 8 form needs all 128 B/clk, which
// the operand stores then compete with).
//
// Operands.  X is row-major (frames x features), so the contraction axis (frames) is the slow one
// in memory, and kind::tf32 wants K-major operands.  Producer threads transpose in registers: a
// thread loads a 4-frame x 4-feature block (one 16-byte load per frame row; a warp reads 512
// contiguous bytes per row), standardises, splits into TF32 hi + lo, and writes four 16-byte
// chunks (4 frames of one feature) into the K-major SWIZZLE_NONE canonical layout
//     [frame group (4 frames)][operand row 0..127][4 frames]       (core matrix = 8 rows x 16 B)
// Operand row r holds feature 4*(r % 32) + r / 32 of the CTA's 128 (a fixed permutation that makes
// the 16-byte stores of a warp contiguous, i.e. bank-conflict free); the epilogue undoes it.
// Row strides that are not a multiple of 4 floats fall back to 8- or 4-byte loads (VEC = 2, 1).
//
// Precision.  3xF16 (DCG_COV_TC_3XF16): the same split with FP16 pieces (hi = RN_f16(z),
// lo = RN_f16(z - hi); FP16 and TF32 both carry 11 significant bits, so the split is as accurate
// as long as |z| < 65504 and 6e-8 absolute is negligible -- true for standardised features) on
// kind::f16, whose K is 16: half the MMA instructions per frame.  A producer thread then converts
// 8 frames x 4 features per visit and a pipeline stage holds 32 frames in the same 32 KB.
// 3xTF32: D += Ahi*Bhi + Ahi*Blo + Alo*Bhi.  The tensor core adds into its FP32
// accumulator with truncation (measured: -5e-8 relative per MMA on same-sign sums), so
// accumulation is two-level: the level-1 accumulator (TMEM columns [0,256)) takes `kc` frames
// (default 256: 2.4e-6 relative to float64, measured; DCG_TC_KC overrides), then the epilogue warps
// add it with round-to-nearest into the level-2 FP32 accumulator (TMEM columns [256,512)); at the
// end of a work item (<= 16384 frames) level 2 is added to the FP64 result with red.global.add.f64.
// The drain is bound by the TMEM read port (64 B/clk/SM: 2 x 128 KB per chunk = ~4.1k cycles
// against 12.3k cycles of MMAs per 256-frame chunk) and is not overlapped (level 1 is single-
// buffered: TMEM is full).
//
// Persistent clusters (one CTA per SM, static strided work-item schedule), warp-specialised roles
// joined by mbarrier pipelines:  warps 0-15 producers (two sets of 4 A + 4 B warps, alternating
// stages) | 16-19 epilogue, warp 16 of the leader CTA also issuing the MMAs.  "full" / "accumulator drained" barriers live in the leader CTA and are arrived on
// remotely by the peer; "stage free" / "accumulator ready" are tcgen05.commit multicasts.
//
// Roofline: tensor pipe.  Algorithmic work 3*F^2 FLOP per frame pair (2F^2 for St + F^2 for the
// upper triangle of S0); issued MMA FLOPs = 3x that for 3xTF32.  X is re-read once per super-tile
// row and column from L2 (work items of the same frame range run concurrently).
#include <cstdlib>
#include <cuda_fp16.h>
#include "dcg_common.cuh"
#include "cov_engines.cuh"
#include "tc_common.cuh"

namespace dcg {

using namespace tc;

namespace {

constexpr int kSup = 256;            // super-tile edge = UMMA M = N (features)
constexpr int kHalf = 128;           // operand rows staged per CTA
// precision modes of the contraction
constexpr int kPrecTf32x1 = 0, kPrecTf32x3 = 1, kPrecF16x3 = 2;
// frames per pipeline stage = two MMA K-steps: kind::tf32 has K = 8 (16 frames), kind::f16 K = 16
// (32 frames).  The shared-memory geometry is the same in both: 16-byte K-groups (4 x tf32 or
// 8 x f16), four groups per operand plane.
__host__ __device__ constexpr int stage_frames(int prec) { return prec == kPrecF16x3 ? 32 : 16; }
constexpr int kStageMin = 16;
constexpr int kNS = 6;               // pipeline stages
constexpr int kSets = 2;              // producer warp sets; set k stages the item's stages k, k + kSets, ...
constexpr int kSetWarps = 8;         // warps per set: 4 per operand (one per 4-frame group of the stage)
constexpr int kProdWarps = kSets * kSetWarps;
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = kProdWarps;
constexpr int kThreads = (kProdWarps + kEpiWarps) * 32;       // 640: 5 warps per scheduler, 96 registers
constexpr int kGroupBytes = kHalf * 16;                       // one K-group (16 bytes) of 128 rows
constexpr int kPlaneBytes = 4 * kGroupBytes;                  // 8 KB: one operand plane (hi or lo)
constexpr int kStageBytes = 4 * kPlaneBytes;                  // A_hi A_lo B_hi B_lo
constexpr size_t kSmemBytes = (size_t)kNS * kStageBytes + 1024;
constexpr int kMaxItemFrames = 16384;                         // level-2 FP32 accumulation span
constexpr int kDefaultKc = 256;                               // level-1 chunk (frames); error vs float64 ~ 1e-8 * kc

struct TileDesc {
  int i0, j0;          // first feature of the super-tile rows / columns (multiples of 4)
  int i_lo, i_hi;      // valid features [lo, hi) along I (others are staged as zeros)
  int j_lo, j_hi;
  int kind;            // 0 = S0, 1 = St
  int diag;            // S0 tile with I == J: the B operand is the A operand
};

// Super-tiles of the dense matrix (block == 0) or of each diagonal block (hTICA level 1); tile
// origins are aligned to the block start (rounded down to a multiple of 4 features).
__host__ __device__ inline int enum_tiles(int f, int block, bool want_s0, bool want_st, TileDesc* out) {
  int n = 0;
  const int w = block > 0 ? block : f;
  for (int b0 = 0; b0 < f; b0 += w) {
    const int b1 = b0 + w < f ? b0 + w : f;
    const int org = b0 & ~3;
    const int nt = (b1 - org + kSup - 1) / kSup;
    for (int ti = 0; ti < nt; ++ti)
      for (int tj = 0; tj < nt; ++tj)
        for (int kind = 0; kind < 2; ++kind) {
          if (kind == 0 ? !(want_s0 && tj >= ti) : !want_st) continue;
          if (out) {
            TileDesc t;
            t.i0 = org + ti * kSup; t.j0 = org + tj * kSup;
            t.i_lo = t.i0 > b0 ? t.i0 : b0; t.i_hi = t.i0 + kSup < b1 ? t.i0 + kSup : b1;
            t.j_lo = t.j0 > b0 ? t.j0 : b0; t.j_hi = t.j0 + kSup < b1 ? t.j0 + kSup : b1;
            t.kind = kind; t.diag = (kind == 0 && ti == tj);
            out[n] = t;
          }
          ++n;
        }
  }
  return n;
}

struct Params {
  const float* X;
  int64_t n_rows, ld;
  int f, lag;
  const float* mean;
  const float* range;
  double* S0;
  double* St;
  double* colsum;       // [f] sum_{t<M} z_t accumulated by the A producers of diagonal S0 tiles (or null)
  const TileDesc* tiles;
  int n_tiles;
  int64_t granule;      // frames per work item (multiple of kc)
  int64_t n_items;      // n_tiles * n_ranges, range-major
  int kc;               // frames per level-1 chunk (multiple of kStage)
};

// sum_{t>=lag} z_t = sum_{t<M} z_t - sum_{t<lag} z_t + sum_{t>=M} z_t: 2*lag rows instead of a pass
// over X.  `sum_t` may be null: then `sum_lag` holds sum_{t<M} z_t on entry.
__global__ void tc_colsum_lag_kernel(const float* __restrict__ X, int64_t n_rows, int f, int64_t ld, int lag,
                                     const float* __restrict__ mean, const float* __restrict__ range,
                                     const double* __restrict__ sum_t, double* __restrict__ sum_lag) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= f) return;
  float mu = 0.f, ri = 1.f;
  if (mean) { mu = mean[col]; ri = 1.0f / range[col]; }
  const int64_t M = n_rows - lag;
  double head = 0.0, tail = 0.0;
  for (int64_t t = 0; t < lag; ++t) {
    head += (double)((X[t * ld + col] - mu) * ri);
    tail += (double)((X[(M + t) * ld + col] - mu) * ri);
  }
  sum_lag[col] = (sum_t ? sum_t[col] : sum_lag[col]) - head + tail;
}

__global__ void plan_kernel(TileDesc* tiles, int f, int block, bool want_s0, bool want_st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) enum_tiles(f, block, want_s0, want_st, tiles);
}

// ---- cluster / CTA-pair primitives ---------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (count 1) on the barrier at the same offset in CTA `rank`.  .relaxed on purpose: a
// .release arrive on a shared::cluster address costs MEMBAR.ALL.CTA (+ MEMBAR.ALL.GPU + ERRBAR at
// cluster scope), which also waits for the producer's prefetched global loads (ncu: 20 % of all
// stall samples).  The data handed over is shared memory read by the tensor core -- already made
// visible by fence.proxy.async, which completes before the arrive issues -- or TMEM
// (tcgen05.wait + tcgen05.fence), never generic-proxy global memory.
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
               ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem desc, per CTA] * B[smem desc, per CTA]; issued by the leader CTA
__device__ __forceinline__ void mma2_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma2_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
template <bool F16>
__device__ __forceinline__ void mma2_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  if constexpr (F16) mma2_f16_ss(d_tmem, a_desc, b_desc, idesc, accumulate);
  else mma2_tf32_ss(d_tmem, a_desc, b_desc, idesc, accumulate);
}
// kind::f16 instruction descriptor: FP16 operands (format 0), FP32 accumulate, both K-major, dense
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// the barrier at this offset in BOTH CTAs gets one arrival when all MMAs issued so far are done
__device__ __forceinline__ void mma2_commit_both(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// 4 consecutive floats of one row with the widest load the row alignment allows
template <int VEC>
__device__ __forceinline__ void load_row4(const float* p, float* x) {
  if (VEC == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
  } else if (VEC == 2) {
    const float2 a = __ldg(reinterpret_cast<const float2*>(p));
    const float2 b = __ldg(reinterpret_cast<const float2*>(p + 2));
    x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
  } else {
    x[0] = __ldg(p); x[1] = __ldg(p + 1); x[2] = __ldg(p + 2); x[3] = __ldg(p + 3);
  }
}

template <int PREC, int VEC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) cov_tc_kernel(const Params p) {
  constexpr bool X3 = PREC != kPrecTf32x1;
  constexpr bool F16 = PREC == kPrecF16x3;
  constexpr int kStage = stage_frames(PREC);
  constexpr int kRows = kStage / 4;                // frames per producer thread block (4 or 8)
  extern __shared__ unsigned char smem_raw[];
  // identical offsets in both CTAs (the dynamic window starts at the same shared address)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  __shared__ uint64_t full_bar[kNS], empty_bar[kNS], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();

  if (tid == 0) {
    for (int i = 0; i < kNS; ++i) { mbar_init(&full_bar[i], 2 * kSetWarps); mbar_init(&empty_bar[i], 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, 2 * kEpiWarps);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc2(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync();                 // peer barriers initialised, both allocations done
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t l1 = tmem, l2 = tmem + kSup;

  const int64_t M = p.n_rows - p.lag;
  const uint32_t chunk_stages = (uint32_t)(p.kc / kStage);
  uint32_t gs = 0, gc = 0;      // pipeline stages / accumulator chunks consumed so far (all roles agree)
  const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  for (int64_t item = cluster_id; item < p.n_items; item += n_clusters) {
    // ---- decode the work item: (frame range, super-tile) ---------------------------------------
    const int64_t range = item / p.n_tiles;
    const TileDesc td = p.tiles[item - range * p.n_tiles];
    const int64_t f0 = range * p.granule;
    const int64_t f1 = f0 + p.granule < M ? f0 + p.granule : M;
    const uint32_t nS = (uint32_t)((f1 - f0 + kStage - 1) / kStage);
    const uint32_t nC = (nS + chunk_stages - 1) / chunk_stages;

    if (warp < kProdWarps) {
      // =============================== producers ==============================================
      const int set = warp / kSetWarps;                // this warp stages the item's stages set, set + kSets, ...
      const int op = (warp >> 2) & 1;                  // 0 = A (z_t[I]), 1 = B (z_t[J] or z_{t+lag}[J])
      const int g = warp & 3;                          // K-group of the stage (kRows frames)
      const bool needed = op == 0 || !td.diag;
      const int c = (op == 0 ? td.i0 : td.j0) + (int)rank * kHalf + 4 * lane;   // first of 4 features
      const int c_lo = op == 0 ? td.i_lo : td.j_lo, c_hi = op == 0 ? td.i_hi : td.j_hi;
      const bool all_ok = c >= c_lo && c + 3 < c_hi;
      bool ok[4];
      float mu[4], ri[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        ok[v] = c + v >= c_lo && c + v < c_hi;
        mu[v] = 0.f; ri[v] = 1.f;
        if (ok[v] && p.mean) { mu[v] = p.mean[c + v]; ri[v] = 1.0f / p.range[c + v]; }
      }
      const int shift = (op == 1 && td.kind == 1) ? p.lag : 0;
      // frames are counted from the item start f0.  A rows beyond the item's range must be exactly
      // zero (also what the column sums need); B rows only need to be in bounds.
      const int t_lim = (int)((op == 0 || td.diag ? f1 : p.n_rows - shift) - f0);
      const size_t ld = (size_t)p.ld;
      const float* pst = p.X + (size_t)(f0 + shift + set * kStage + kRows * g) * ld + c;   // this thread's first block
      // x[4 * r + v] = frame (t0 + r), feature (c + v)
      auto load_block = [&](float (&x)[4 * kRows], int t0, const float* pt) {
        if (all_ok && t0 + kRows - 1 < t_lim) {
#pragma unroll
          for (int r = 0; r < kRows; ++r) load_row4<VEC>(pt + r * ld, &x[4 * r]);
        } else {
#pragma unroll
          for (int r = 0; r < kRows; ++r)
#pragma unroll
            for (int v = 0; v < 4; ++v)
              x[4 * r + v] = (ok[v] && t0 + r < t_lim) ? __ldg(pt + r * ld + v) : mu[v];   // mu -> z == 0 exactly
        }
      };
      constexpr int kStep = kSets * kStage;                 // frames between this warp's stages
      const size_t step_stride = (size_t)kStep * ld;
      int t0 = set * kStage + kRows * g;
      uint32_t s = (uint32_t)set;
      if (needed) {
        const uint32_t dst0 = smem_base + (uint32_t)(2 * op) * kPlaneBytes + (uint32_t)g * kGroupBytes + (uint32_t)lane * 16;
        // standardise, split, store K-major, publish the stage
        auto emit_block = [&](const float (&x)[4 * kRows], uint32_t g_stage) {
          const uint32_t slot = g_stage % kNS;
          mbar_wait(&empty_bar[slot], ((g_stage / kNS) & 1) ^ 1);
          const uint32_t dst = dst0 + slot * kStageBytes;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint32_t hi[4], lo[4];
            if constexpr (F16) {
              // 8 frames of one feature -> 8 x f16 hi (RN) and 8 x f16 lo = RN(z - hi), packed in pairs
#pragma unroll
              for (int r2 = 0; r2 < 4; ++r2) {
                const float z0 = (x[4 * (2 * r2) + v] - mu[v]) * ri[v];
                const float z1 = (x[4 * (2 * r2 + 1) + v] - mu[v]) * ri[v];
                const __half2 h = __floats2half2_rn(z0, z1);
                const float2 hf = __half22float2(h);
                const __half2 l = __floats2half2_rn(z0 - hf.x, z1 - hf.y);
                hi[r2] = *reinterpret_cast<const uint32_t*>(&h);
                lo[r2] = *reinterpret_cast<const uint32_t*>(&l);
              }
            } else {
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                // (x - mean) * RN(1/range): within 1 ulp of the reference's IEEE division (the
                // difference is a per-feature scale factor of at most 1 + 2^-24)
                const float z = (x[4 * r + v] - mu[v]) * ri[v];
                split_tf32_fast(z, hi[r], lo[r]);
              }
            }
            st_shared_v4(dst + v * 512, hi[0], hi[1], hi[2], hi[3]);                 // operand row v * 32 + lane
            if (X3) st_shared_v4(dst + kPlaneBytes + v * 512, lo[0], lo[1], lo[2], lo[3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive_cta(&full_bar[slot], 0);
        };
        if constexpr (F16) {
          // 8 frames x 4 features per thread in ONE buffer: the loads of the warp's next stage are
          // issued right after the current stage is published, so they fly while the warp waits for
          // its next slot (the other warp set works in between) and nothing is in flight at the
          // proxy fence (its MEMBAR would wait for it)
          float xa[4 * kRows];
          if (s < nS) load_block(xa, t0, pst);
          for (; s < nS; s += kSets) {
            emit_block(xa, gs + s);
            t0 += kStep;
            pst += step_stride;
            if (s + kSets < nS) load_block(xa, t0, pst);
          }
        } else {
          // software pipeline: the loads of this warp's next stage are in flight while the current
          // one is converted (and the other warp set works on the stages in between)
          float xa[4 * kRows], xb[4 * kRows];
          if (s < nS) load_block(xa, t0, pst);
          for (; s + kSets < nS; s += 2 * kSets) {
            load_block(xb, t0 + kStep, pst + step_stride);
            emit_block(xa, gs + s);
            if (s + 2 * kSets < nS) load_block(xa, t0 + 2 * kStep, pst + 2 * step_stride);
            emit_block(xb, gs + s + kSets);
            t0 += 2 * kStep;
            pst += 2 * step_stride;
          }
          if (s < nS) emit_block(xa, gs + s);
        }
      } else {
        // Diagonal S0 tile: B is A, so the B warps have nothing to stage.  They keep the pipeline
        // protocol and, if asked, accumulate the column sums sum_t z_t of the tile's features (same
        // rows and bounds as the A operand) -- FP32 within a block of frames, FP64 across blocks.
        const bool want_sum = p.colsum != nullptr;
        double zsum[4] = {0.0, 0.0, 0.0, 0.0};
        for (; s < nS; s += kSets, t0 += kStep, pst += step_stride) {
          const uint32_t slot = (gs + s) % kNS;
          mbar_wait(&empty_bar[slot], (((gs + s) / kNS) & 1) ^ 1);
          __syncwarp();
          if (lane == 0) mbar_arrive_cta(&full_bar[slot], 0);
          if (want_sum) {
            float x[4 * kRows];
            load_block(x, t0, pst);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              float sv = 0.f;
#pragma unroll
              for (int r = 0; r < kRows; ++r) sv += (x[4 * r + v] - mu[v]) * ri[v];
              zsum[v] += (double)sv;
            }
          }
        }
        if (want_sum) {
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (ok[v]) atomicAdd(p.colsum + c + v, zsum[v]);
        }
      }
      gs += nS;
      gc += nC;
    } else {
      // ============ epilogue warps; the first one of the leader CTA also issues the MMAs ========
      // Level 1 is single-buffered, so MMA issue and accumulator drain of one CTA pair strictly
      // alternate: one warp can do both (issue the chunk's stages, wait for them to complete, drain
      // its lane quarter), which keeps the CTA at 20 warps = 96 registers per thread.
      const int q = warp & 3;                          // TMEM lane quarter this warp may access
      const uint32_t lane_base = (uint32_t)(q * 32) << 16;
      const bool issuer = warp == kMmaWarp && rank == 0;
      constexpr uint32_t idesc = F16 ? make_idesc_f16(kSup, kSup) : make_idesc_tf32(kSup, kSup, 0, 0);   // both operands K-major
      uint32_t s = 0;
      for (uint32_t c = 0; c < nC; ++c, ++gc) {
        if (issuer) {
          mbar_wait_cluster(&acc_empty, (gc & 1) ^ 1);                   // level 1 drained in both CTAs
          const uint32_t s_end = s + chunk_stages < nS ? s + chunk_stages : nS;
          for (bool first = true; s < s_end; ++s, first = false) {
            const uint32_t slot = (gs + s) % kNS;
            mbar_wait_cluster(&full_bar[slot], ((gs + s) / kNS) & 1);
            tc_fence_after();
            if (elect_one_sync()) {
              const uint32_t st_base = smem_base + slot * kStageBytes;
              const uint64_t a_h0 = make_smem_desc(st_base + 0 * kPlaneBytes, kGroupBytes, 128);
              const uint64_t a_l0 = make_smem_desc(st_base + 1 * kPlaneBytes, kGroupBytes, 128);
              const uint64_t b_h0 = td.diag ? a_h0 : make_smem_desc(st_base + 2 * kPlaneBytes, kGroupBytes, 128);
              const uint64_t b_l0 = td.diag ? a_l0 : make_smem_desc(st_base + 3 * kPlaneBytes, kGroupBytes, 128);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint64_t off = (uint64_t)((h * 2 * kGroupBytes) >> 4);     // one K-step = two 16-byte K-groups
                mma2_ss<F16>(l1, a_h0 + off, b_h0 + off, idesc, (first && h == 0) ? 0u : 1u);
                if (X3) {
                  mma2_ss<F16>(l1, a_h0 + off, b_l0 + off, idesc, 1);
                  mma2_ss<F16>(l1, a_l0 + off, b_h0 + off, idesc, 1);
                }
              }
              mma2_commit_both(&empty_bar[slot]);                              // stage consumed
              if (s + 1 == s_end) mma2_commit_both(&acc_full);                 // chunk ready to drain
            }
            __syncwarp();
          }
        }
        mbar_wait(&acc_full, gc & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < kSup; c0 += 32) {
          uint32_t v[32];
          tmem_ld_x32(l1 + lane_base + c0, v);
          if (c != 0) {
            uint32_t u[32];
            tmem_ld_x32(l2 + lane_base + c0, u);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
          } else {
            tmem_ld_wait();
          }
          tmem_st_x32(l2 + lane_base + c0, v);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(&acc_empty, 0);
      }
      // item done: level 2 -> FP64 result (red.global.add.f64, fire and forget).  TMEM lane
      // 32 q + l holds feature 4 l + q of this CTA's rows; column 128 b + 32 v + l' holds feature
      // 128 b + 4 l' + v of the super-tile's columns.
      const int gi = td.i0 + (int)rank * kHalf + 4 * lane + q;
      const bool row_ok = gi >= td.i_lo && gi < td.i_hi;
      double* out = (td.kind == 0 ? p.S0 : p.St) + (size_t)gi * p.f;
#pragma unroll 1
      for (int k = 0; k < kSup / 32; ++k) {
        uint32_t v[32];
        tmem_ld_x32(l2 + lane_base + 32 * k, v);
        tmem_ld_wait();
        const int gj0 = td.j0 + kHalf * (k >> 2) + (k & 3);
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int gj = gj0 + 4 * j;
            if (gj >= td.j_lo && gj < td.j_hi) atomicAdd(out + gj, (double)__uint_as_float(v[j]));
          }
        }
      }
      gs += nS;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();                 // the peer may still arrive on / multicast to this CTA's barriers
  if (warp == kMmaWarp) tmem_dealloc2(tmem, 512);
}

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

template <int PREC, int VEC>
int launch_variant(const Params& p, int64_t n_items, cudaStream_t st) {
  auto kern = cov_tc_kernel<PREC, VEC>;
  DCG_CUDA_TRY(ensure_dynamic_smem((const void*)kern, (size_t)(kSmemBytes)));
  const int n_clusters = (int)std::min<int64_t>(kNumSMs / 2, n_items);
  kern<<<2 * n_clusters, kThreads, kSmemBytes, st>>>(p);
  DCG_LAUNCH_CHECK();
  return 0;
}

}  // namespace

size_t cov_tc_workspace_bytes(int64_t n_rows, int f, int lag, int block, int engine) {
  (void)n_rows; (void)lag; (void)engine;
  const size_t n_tiles = (size_t)enum_tiles(f, block, true, true, nullptr);
  return 256 + align_up(n_tiles * sizeof(TileDesc), 256);
}

int cov_tc_launch(const CovArgs& a, cudaStream_t st) {
  int dev = 0, major = 0;
  DCG_CUDA_TRY(cudaGetDevice(&dev));
  DCG_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return DCG_E_ARCH;
  const int prec = a.engine == DCG_COV_TC_3XF16 ? kPrecF16x3 : a.engine == DCG_COV_TC_3XTF32 ? kPrecTf32x3 : kPrecTf32x1;
  const int stage = stage_frames(prec);
  // kind::f16 adds into the accumulator once per 16 frames (kind::tf32: per 8): at the same chunk
  // length its truncation error is smaller (1.6e-6 against 2.4e-6 at kc = 256, measured); chunks
  // twice as long (3.2e-6) gain only 5 % and cost eigenvector accuracy, so both use the same kc
  int kc = env_int("DCG_TC_KC", kDefaultKc);
  kc = std::max(stage, kc / stage * stage);

  const int64_t M = a.n_rows - a.lag;
  TileDesc* d_tiles = (TileDesc*)((char*)a.ws + 256);
  const int n_tiles = enum_tiles(a.f, a.block, a.S0 != nullptr, a.St != nullptr, nullptr);
  if (n_tiles == 0 || M <= 0) return 0;
  // the list itself is built on the device (same enumeration) so the launch stays asynchronous
  plan_kernel<<<1, 32, 0, st>>>(d_tiles, a.f, a.block, a.S0 != nullptr, a.St != nullptr);
  DCG_LAUNCH_CHECK();

  // frames per work item: ~8 items per cluster, a multiple of kc, at most kMaxItemFrames
  const int n_clusters = kNumSMs / 2;
  int64_t g = ceil_div(M * n_tiles, (int64_t)n_clusters * 8);
  g = std::min<int64_t>(std::max<int64_t>(ceil_div(g, kc) * kc, kc), std::max(kc, kMaxItemFrames / kc * kc));
  const int64_t n_ranges = ceil_div(M, g);

  double* colsum = a.colsum_t ? a.colsum_t : a.colsum_lag;      // sum_{t<M} z_t lands here first
  if (!cov_tc_fuses_colsums(a)) colsum = nullptr;
  Params p{a.X, a.n_rows, a.ld, a.f, a.lag, a.mean, a.range, a.S0, a.St, colsum,
           d_tiles, n_tiles, g, n_ranges * n_tiles, kc};
  const int vec = row_vec_width(a.X, a.ld);
  int rc;
#define DCG_TC_VARIANTS(P)                                                                          \
  rc = vec == 4 ? launch_variant<P, 4>(p, p.n_items, st) : vec == 2 ? launch_variant<P, 2>(p, p.n_items, st) \
                                                                    : launch_variant<P, 1>(p, p.n_items, st)
  if (prec == kPrecF16x3) { DCG_TC_VARIANTS(kPrecF16x3); }
  else if (prec == kPrecTf32x3) { DCG_TC_VARIANTS(kPrecTf32x3); }
  else { DCG_TC_VARIANTS(kPrecTf32x1); }
#undef DCG_TC_VARIANTS
  if (rc) return rc;
  if (a.colsum_lag && colsum) {
    tc_colsum_lag_kernel<<<(unsigned)ceil_div(a.f, 128), 128, 0, st>>>(
        a.X, a.n_rows, a.f, a.ld, a.lag, a.mean, a.range, a.colsum_t, a.colsum_lag);
    DCG_LAUNCH_CHECK();
  }
  return 0;
}

// the column sums ride on the diagonal S0 tiles: fused whenever S0 is computed
bool cov_tc_fuses_colsums(const CovArgs& a) {
  return a.n_rows - a.lag > 0 && a.S0 != nullptr;
}

}  // namespace dcg
