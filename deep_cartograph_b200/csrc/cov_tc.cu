// tcgen05 covariance engine (DCG_COV_TC_3XTF32 / DCG_COV_TC_1XTF32).
//
// Computes, for 128 x 128 tiles (I, J) of the feature axis and a range of frames,
//     S0[I,J]  += sum_t z_t[I] (x) z_t[J]         St[I,J] += sum_t z_t[I] (x) z_{t+lag}[J]
// as a dense contraction over the FRAME axis on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, M = N = 128, K = 8 frames per instruction, cta_group::1).
//
// Operands.  X is row-major (frames x features), so the contraction axis (frames) is the slow
// one in memory.  kind::tf32 takes K-major operands only (measured: tools_dev/tc_probe.cu; the
// MN-major SWIZZLE_NONE form returns zeros), so the producer warps transpose while staging: a
// thread owns ONE feature, reads 16 consecutive frames of it (each load instruction is a
// coalesced 128-byte row segment across the warp), standardises, splits into TF32 hi + lo, and
// writes 4-frame groups as 16-byte chunks into the K-major SWIZZLE_NONE canonical layout
//     [frame group (4 frames)][feature row 0..127][4 frames]      (core matrix = 8 rows x 16 B)
// Three operand tiles per 16-frame stage: A = z_t[I], B0 = z_t[J], Bt = z_{t+lag}[J], each as a
// hi and a lo plane (6 x 8 KB per stage, 4 stages).  Diagonal tiles (I == J) reuse A as B0.
//
// Precision.  3xTF32: D += Ahi*Bhi + Ahi*Blo + Alo*Bhi (hi = RN_tf32(z), lo = RN_tf32(z - hi)).
// The tensor core adds into its FP32 accumulator with truncation (measured: -5e-8 relative per
// MMA on same-sign sums), so accumulation is two-level: level-1 accumulators (TMEM columns
// [0,256)) take `kc` frames (default 128), then the epilogue warps add them with round-to-
// nearest into level-2 FP32 accumulators (TMEM columns [256,512)); at the end of a work item
// (<= 16384 frames) level 2 is added to the FP64 result with red.global.add.f64.
//
// Persistent CTAs (one per SM, static strided work-item schedule), warp-specialised roles joined
// by mbarrier pipelines:  warps 0-3 A | 4-7 B0 | 8-11 Bt producers | 12 MMA issuer | 13-16 epilogue.
//
// Roofline: tensor pipe.  Algorithmic work 3*F^2 FLOP per frame pair (2F^2 for St + F^2 for the
// upper triangle of S0); issued MMA FLOPs = 3x that for 3xTF32.  X is re-read once per tile row
// and column from L2 (work items of the same frame range run concurrently).
#include <cstdlib>
#include "dcg_common.cuh"
#include "cov_engines.cuh"
#include "tc_common.cuh"

namespace dcg {

using namespace tc;

namespace v1 {

constexpr int kTile = 128;           // UMMA M = N = 128 features
constexpr int kStage = 16;           // frames per pipeline stage (two K = 8 MMA steps)
constexpr int kNS = 4;               // pipeline stages
constexpr int kProdWarps = 12;       // 4 per operand tile (A, B0, Bt)
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = kProdWarps;
constexpr int kThreads = (kProdWarps + 1 + kEpiWarps) * 32;   // 544
constexpr int kGroupBytes = kTile * 16;                       // one 4-frame group of 128 rows
constexpr int kPlaneBytes = (kStage / 4) * kGroupBytes;       // 8 KB: one operand plane (hi or lo)
constexpr int kStageBytes = 6 * kPlaneBytes;                  // A_hi A_lo B0_hi B0_lo Bt_hi Bt_lo
constexpr size_t kSmemBytes = (size_t)kNS * kStageBytes + 1024;
constexpr int kMaxItemFrames = 16384;                         // level-2 FP32 accumulation span

__host__ __device__ inline bool tc_tile_needed(int i0, int j0, int f, int block, bool s0) {
  if (s0 && j0 + kTile - 1 < i0) return false;          // strictly-lower tile of the symmetric S0
  if (block <= 0) return true;
  const int ie = (i0 + kTile < f ? i0 + kTile : f) - 1, je = (j0 + kTile < f ? j0 + kTile : f) - 1;
  const int ia = i0 / block, ib = ie / block, ja = j0 / block, jb = je / block;
  return !(ib < ja || jb < ia);
}

constexpr int kFlagS0 = 1 << 29, kFlagSt = 1 << 30;

struct TcParams {
  const float* X;
  int64_t n_rows, ld;
  int f, lag;
  const float* mean;
  const float* range;
  double* S0;
  double* St;
  double* colsum;       // [f] sum_{t<M} z_t accumulated by the A producers of diagonal tiles (or null)
  const int* tiles;     // [n_tiles]: (ti * nt + tj) | flags
  int n_tiles, nt;
  int64_t granule;      // frames per work item (multiple of kc)
  int64_t n_items;      // n_tiles * n_ranges, range-major
  int kc;               // frames per level-1 chunk (multiple of kStage)
};

// Tile list: (ti * nt + tj) | flags, tiles needing both matrices first.  One thread; nt <= ~40.
__global__ void tc_plan_kernel(int* __restrict__ tiles, int nt, int f, int block, bool want_s0, bool want_st) {
  if (threadIdx.x != 0) return;
  int n = 0;
  for (int pass = 0; pass < 2; ++pass)
    for (int ti = 0; ti < nt; ++ti)
      for (int tj = 0; tj < nt; ++tj) {
        const bool s0 = want_s0 && tc_tile_needed(ti * kTile, tj * kTile, f, block, true);
        const bool stt = want_st && tc_tile_needed(ti * kTile, tj * kTile, f, block, false);
        if (!s0 && !stt) continue;
        if ((pass == 0) != (s0 && stt)) continue;
        tiles[n++] = (ti * nt + tj) | (s0 ? kFlagS0 : 0) | (stt ? kFlagSt : 0);
      }
}

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

template <bool X3, bool STD>
__global__ void __launch_bounds__(kThreads, 1) cov_tc_kernel(const TcParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  __shared__ uint64_t full_bar[kNS], empty_bar[kNS], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < kNS; ++i) { mbar_init(&full_bar[i], kProdWarps); mbar_init(&empty_bar[i], 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, kEpiWarps);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t l1_0 = tmem, l1_t = tmem + kTile, l2_0 = tmem + 2 * kTile, l2_t = tmem + 3 * kTile;

  const int64_t M = p.n_rows - p.lag;
  const uint32_t chunk_stages = (uint32_t)(p.kc / kStage);
  uint32_t gs = 0, gc = 0;      // pipeline stages / accumulator chunks consumed so far (all roles agree)

  for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
    // ---- decode the work item: (frame range, tile, which matrices) ---------------------------
    const int64_t range = item / p.n_tiles;
    const int code = p.tiles[item - range * p.n_tiles];
    const int tile = code & (kFlagS0 - 1);
    const bool do_s0 = code & kFlagS0, do_st = code & kFlagSt;
    const int64_t f0 = range * p.granule;
    const int64_t f1 = f0 + p.granule < M ? f0 + p.granule : M;
    const int i0 = (tile / p.nt) * kTile, j0 = (tile % p.nt) * kTile;
    const bool diag = i0 == j0;
    const uint32_t nS = (uint32_t)((f1 - f0 + kStage - 1) / kStage);
    const uint32_t nC = (nS + chunk_stages - 1) / chunk_stages;

    if (warp < kProdWarps) {
      // =============================== producers ==============================================
      const int op = warp >> 2;                       // 0 = A, 1 = B0, 2 = Bt
      const int row = (warp & 3) * 32 + lane;         // feature row of the tile
      const bool needed = op == 0 || (op == 1 ? (do_s0 && !diag) : do_st);
      const int col = (op == 0 ? i0 : j0) + row;
      const bool col_ok = col < p.f;
      const int shift = op == 2 ? p.lag : 0;
      // A rows beyond the item's range must be exactly zero; B rows only need to be in bounds
      const int64_t t_lim = op == 0 ? f1 : p.n_rows - shift;
      float mu = 0.f, ri = 1.f;
      if (STD && col_ok) { mu = p.mean[col]; ri = 1.0f / p.range[col]; }
      const bool want_sum = op == 0 && diag && p.colsum != nullptr;      // column sums ride on the diagonal tiles
      double zsum = 0.0;                             // FP32 within a 16-frame stage, FP64 across stages
      const uint32_t ld32 = (uint32_t)p.ld;          // host guarantees 16 * ld * 4 bytes < 2^32
      const float* pst = p.X + ((int64_t)shift + f0) * p.ld + col;      // this thread's column at frame f0
      const uint32_t plane_hi = smem_u32(smem) + (uint32_t)(2 * op) * kPlaneBytes + (uint32_t)row * 16;
      // 16 frames of this thread's feature -> registers (every load is a coalesced 128-byte row
      // segment across the warp)
      auto load_stage = [&](float (&x)[kStage], int64_t t0, const float* pt) {
        if (col_ok && t0 + kStage <= t_lim) {         // whole stage in bounds: no per-element predicates
#pragma unroll
          for (int j = 0; j < kStage; ++j) x[j] = __ldg(pt + (size_t)((uint32_t)j * ld32));
        } else {
#pragma unroll
          for (int j = 0; j < kStage; ++j)
            x[j] = (col_ok && t0 + j < t_lim) ? __ldg(pt + (size_t)((uint32_t)j * ld32)) : mu;   // mu -> z == 0 exactly
        }
      };
      // standardise, split, store K-major, publish the stage
      auto emit_stage = [&](const float (&x)[kStage], uint32_t g_stage) {
        const uint32_t slot = g_stage % kNS;
        mbar_wait(&empty_bar[slot], ((g_stage / kNS) & 1) ^ 1);
        const uint32_t dst = plane_hi + slot * kStageBytes;
        float ssum = 0.f;
#pragma unroll
        for (int g = 0; g < kStage / 4; ++g) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            // (x - mean) * RN(1/range): within 1 ulp of the reference's IEEE division (the
            // difference is a per-feature scale factor of at most 1 + 2^-24)
            const float z = STD ? (x[4 * g + v] - mu) * ri : x[4 * g + v];
            split_tf32_fast(z, hi[v], lo[v]);
            if (want_sum) ssum += z;
          }
          st_shared_v4(dst + g * kGroupBytes, hi[0], hi[1], hi[2], hi[3]);
          if (X3) st_shared_v4(dst + kPlaneBytes + g * kGroupBytes, lo[0], lo[1], lo[2], lo[3]);
        }
        if (want_sum) zsum += (double)ssum;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[slot]);
      };
      if (needed) {
        // software pipeline: the loads of stage s+1 are in flight while stage s is converted
        const size_t stage_stride = (size_t)kStage * ld32;
        float xa[kStage], xb[kStage];
        int64_t t0 = f0;
        load_stage(xa, t0, pst);
        uint32_t s = 0;
        for (; s + 1 < nS; s += 2) {
          load_stage(xb, t0 + kStage, pst + stage_stride);
          emit_stage(xa, gs++);
          if (s + 2 < nS) load_stage(xa, t0 + 2 * kStage, pst + 2 * stage_stride);
          emit_stage(xb, gs++);
          t0 += 2 * kStage;
          pst += 2 * stage_stride;
        }
        if (s < nS) emit_stage(xa, gs++);
        if (want_sum && col_ok) atomicAdd(p.colsum + col, zsum);
      } else {
        for (uint32_t s = 0; s < nS; ++s, ++gs) {
          const uint32_t slot = gs % kNS;
          mbar_wait(&empty_bar[slot], ((gs / kNS) & 1) ^ 1);
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_bar[slot]);
        }
      }
      gc += nC;
    } else if (warp == kMmaWarp) {
      // =============================== MMA issuer ============================================
      // The whole warp runs the (warp-uniform) control flow; one elected lane issues the tcgen05
      // instructions, so their operands stay in uniform registers.
      constexpr uint32_t idesc = make_idesc_tf32(kTile, kTile, 0, 0);   // both operands K-major
      const uint32_t sbase = smem_u32(smem);
      uint32_t in_chunk = 0;                                             // stages issued into the current chunk
      for (uint32_t s = 0; s < nS; ++s, ++gs) {
        const uint32_t slot = gs % kNS;
        const bool chunk_first = in_chunk == 0;
        const bool chunk_last = ++in_chunk == chunk_stages || s + 1 == nS;
        if (chunk_last) in_chunk = 0;
        if (chunk_first) mbar_wait(&acc_empty, (gc & 1) ^ 1);          // level-1 accumulators drained
        mbar_wait(&full_bar[slot], (gs / kNS) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t st_base = sbase + slot * kStageBytes;
          const uint64_t a_h0 = make_smem_desc(st_base + 0 * kPlaneBytes, kGroupBytes, 128);
          const uint64_t a_l0 = make_smem_desc(st_base + 1 * kPlaneBytes, kGroupBytes, 128);
          const uint64_t b_h0 = diag ? a_h0 : make_smem_desc(st_base + 2 * kPlaneBytes, kGroupBytes, 128);
          const uint64_t b_l0 = diag ? a_l0 : make_smem_desc(st_base + 3 * kPlaneBytes, kGroupBytes, 128);
          const uint64_t t_h0 = make_smem_desc(st_base + 4 * kPlaneBytes, kGroupBytes, 128);
          const uint64_t t_l0 = make_smem_desc(st_base + 5 * kPlaneBytes, kGroupBytes, 128);
#pragma unroll
          for (int h = 0; h < kStage / 8; ++h) {
            const uint64_t off = (uint64_t)((h * 2 * kGroupBytes) >> 4);     // K = 8 = two 4-frame groups
            const uint32_t acc = (chunk_first && h == 0) ? 0u : 1u;
            if (do_s0) {
              mma_tf32_ss(l1_0, a_h0 + off, b_h0 + off, idesc, acc);
              if (X3) { mma_tf32_ss(l1_0, a_h0 + off, b_l0 + off, idesc, 1); mma_tf32_ss(l1_0, a_l0 + off, b_h0 + off, idesc, 1); }
            }
            if (do_st) {
              mma_tf32_ss(l1_t, a_h0 + off, t_h0 + off, idesc, acc);
              if (X3) { mma_tf32_ss(l1_t, a_h0 + off, t_l0 + off, idesc, 1); mma_tf32_ss(l1_t, a_l0 + off, t_h0 + off, idesc, 1); }
            }
          }
          mma_commit(&empty_bar[slot]);                                    // stage consumed
          if (chunk_last) mma_commit(&acc_full);                           // chunk ready to drain
        }
        __syncwarp();
        if (chunk_last) ++gc;
      }
    } else {
      // =============================== epilogue ==============================================
      const int q = warp & 3;                          // TMEM lane quarter this warp may access
      const int row = q * 32 + lane;                   // tile row (feature i0 + row)
      const uint32_t lane_base = (uint32_t)(q * 32) << 16;
      for (uint32_t c = 0; c < nC; ++c, ++gc) {
        mbar_wait(&acc_full, gc & 1);
        tc_fence_after();
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
          if (which == 0 ? !do_s0 : !do_st) continue;
          const uint32_t src = (which == 0 ? l1_0 : l1_t) + lane_base;
          const uint32_t dst = (which == 0 ? l2_0 : l2_t) + lane_base;
#pragma unroll 1
          for (int c0 = 0; c0 < kTile; c0 += 32) {
            uint32_t v[32];
            tmem_ld_x32(src + c0, v);
            if (c != 0) {
              uint32_t u[32];
              tmem_ld_x32(dst + c0, u);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
            } else {
              tmem_ld_wait();
            }
            tmem_st_x32(dst + c0, v);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty);
      }
      // item done: level 2 -> FP64 result (red.global.add.f64, fire and forget)
      const int gi = i0 + row;
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {
        if (which == 0 ? !do_s0 : !do_st) continue;
        const uint32_t src = (which == 0 ? l2_0 : l2_t) + lane_base;
        double* out = (which == 0 ? p.S0 : p.St) + (size_t)gi * p.f + j0;
#pragma unroll 1
        for (int c0 = 0; c0 < kTile; c0 += 32) {
          uint32_t v[32];
          tmem_ld_x32(src + c0, v);
          tmem_ld_wait();
          if (gi < p.f) {
            const int jn = p.f - j0 - c0;               // valid columns from c0 on
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < jn) atomicAdd(out + c0 + j, (double)__uint_as_float(v[j]));
          }
        }
      }
      gs += nS;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}

}  // namespace v1
using namespace v1;

size_t cov_tc1_workspace_bytes(int64_t n_rows, int f, int lag, int block, int engine) {
  (void)n_rows; (void)lag; (void)block; (void)engine;
  const size_t nt = (size_t)ceil_div(f, kTile);
  return 256 + align_up(nt * nt * sizeof(int), 256);
}

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

int cov_tc1_launch(const CovArgs& a, cudaStream_t st) {
  int dev = 0, major = 0;
  DCG_CUDA_TRY(cudaGetDevice(&dev));
  DCG_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return DCG_E_ARCH;
  int kc = env_int("DCG_TC_KC", 128);
  kc = std::max(kStage, kc / kStage * kStage);

  const int64_t M = a.n_rows - a.lag;
  const int nt = (int)ceil_div(a.f, kTile);
  int* d_tiles = (int*)((char*)a.ws + 256);

  // needed tiles, the expensive ones (both matrices) first; the list itself is built on the device
  // (same enumeration) so the launch stays asynchronous
  int n_tiles = 0;
  for (int ti = 0; ti < nt; ++ti)
    for (int tj = 0; tj < nt; ++tj)
      n_tiles += (a.S0 && tc_tile_needed(ti * kTile, tj * kTile, a.f, a.block, true)) ||
                 (a.St && tc_tile_needed(ti * kTile, tj * kTile, a.f, a.block, false));
  if (n_tiles == 0 || M <= 0) return 0;
  tc_plan_kernel<<<1, 32, 0, st>>>(d_tiles, nt, a.f, a.block, a.S0 != nullptr, a.St != nullptr);
  DCG_LAUNCH_CHECK();

  // frames per work item: ~8 items per SM, a multiple of kc, at most kMaxItemFrames
  int64_t g = ceil_div(M * n_tiles, (int64_t)kNumSMs * 8);
  g = std::min<int64_t>(std::max<int64_t>(ceil_div(g, kc) * kc, kc), std::max(kc, kMaxItemFrames / kc * kc));
  const int64_t n_ranges = ceil_div(M, g);

  double* colsum = a.colsum_t ? a.colsum_t : a.colsum_lag;      // sum_{t<M} z_t lands here first
  TcParams p{a.X, a.n_rows, a.ld, a.f, a.lag, a.mean, a.range, a.S0, a.St, colsum,
             d_tiles, n_tiles, nt, g, n_ranges * n_tiles, kc};
  const int grid = (int)std::min<int64_t>(kNumSMs, p.n_items);
  const bool x3 = a.engine == DCG_COV_TC_3XTF32;
  const bool stdz = a.mean != nullptr;
#define DCG_TC_CASE(X3, STD)                                                                          \
  {                                                                                                   \
    auto kern = cov_tc_kernel<X3, STD>;                                                               \
    DCG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes)); \
    kern<<<grid, kThreads, kSmemBytes, st>>>(p);                                                      \
  }
  if (x3) { if (stdz) DCG_TC_CASE(true, true) else DCG_TC_CASE(true, false) }
  else { if (stdz) DCG_TC_CASE(false, true) else DCG_TC_CASE(false, false) }
#undef DCG_TC_CASE
  DCG_LAUNCH_CHECK();
  if (a.colsum_lag) {
    tc_colsum_lag_kernel<<<(unsigned)ceil_div(a.f, 128), 128, 0, st>>>(
        a.X, a.n_rows, a.f, a.ld, a.lag, a.mean, a.range, a.colsum_t, a.colsum_lag);
    DCG_LAUNCH_CHECK();
  }
  return 0;
}

bool cov_tc1_fuses_colsums(const CovArgs& a) { return a.n_rows - a.lag > 0; }

}  // namespace dcg
