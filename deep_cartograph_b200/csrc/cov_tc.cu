// tcgen05 covariance engine (DCG_COV_TC_3XTF32 / DCG_COV_TC_1XTF32).
//
// Computes, for 128 x 128 tiles (I, J) of the feature axis and a range of frames,
//     S0[I,J]  += sum_t z_t[I] (x) z_t[J]         St[I,J] += sum_t z_t[I] (x) z_{t+lag}[J]
// as a dense contraction over the FRAME axis on the 5th-generation tensor cores:
//   * D (FP32 accumulators for S0 and St) in TMEM, columns [0,128) and [128,256);
//   * A = z_t[I]^T (M = 128 features x K = 8 frames per MMA), standardised and split hi/lo by the
//     A-producer warps straight from global memory into TMEM (tcgen05.st; TS form) -- or, in the
//     SS variant, into shared memory;
//   * B = z_t[J] (N = 128 features), standardised and split by the B-producer warps into a shared-
//     memory RING of frames in the MN-major SWIZZLE_NONE canonical layout [feature/4][frame][4].
//     Frames are the K axis and advance linearly (16 bytes per frame) inside a feature group, so
//     the lag-shifted operand z_{t+lag}[J] is THE SAME ring read through a descriptor whose start
//     address is advanced by lag*16 bytes: S0 and St share every staged byte.
//   * split precision (3xTF32): D += Ahi*Bhi + Ahi*Blo + Alo*Bhi keeps ~2^-21 relative accuracy;
//   * every `kc` frames the FP32 accumulators are flushed into a CTA-private FP64 slab
//     (L2-resident), and at the end of a work item the slab is added to the FP64 result with
//     red.global.add.f64.
// Persistent CTAs (one per SM), dynamic work-item scheduler, warp-specialised roles connected by
// mbarrier pipelines:   warps 0-3 A producers | 4-7 B producers | 8-11 epilogue | 12 MMA issuer.
//
// Roofline: tensor pipe.  Algorithmic work 3*F^2 FLOP per frame pair (2F^2 for St + F^2 for the
// upper triangle of S0); issued MMA FLOPs = 3x that for 3xTF32.  X is re-read once per tile row/
// column from L2 (work items of the same frame range run concurrently).
#include <cstdlib>
#include <vector>
#include "dcg_common.cuh"
#include "cov_engines.cuh"
#include "tc_common.cuh"

namespace dcg {

using namespace tc;

constexpr int kTile = 128;          // UMMA M = N = 128 features
constexpr int kStage = 16;          // frames per pipeline stage (two K = 8 MMA steps)
constexpr int kThreads = 13 * 32;   // 4 A-producer + 4 B-producer + 4 epilogue + 1 MMA warp
constexpr int kSlabElems = kTile * kTile;

template <bool A_TMEM> struct TcCfg;
template <> struct TcCfg<true> {    // A operand in TMEM: shared memory holds only the B ring
  static constexpr int NA = 8;      // 8 stages x 32 TMEM columns (16 hi + 16 lo) = columns [256,512)
  static constexpr int NB = 10;     // 160-frame ring
};
template <> struct TcCfg<false> {   // A operand in shared memory as well
  static constexpr int NA = 4;
  static constexpr int NB = 6;      // 96-frame ring
};
template <bool A_TMEM> __host__ __device__ constexpr int tc_rb() { return TcCfg<A_TMEM>::NB * kStage; }
template <bool A_TMEM> __host__ __device__ constexpr int tc_rb_alloc() { return tc_rb<A_TMEM>() + 7; }   // odd, + wrap copies
template <bool A_TMEM> __host__ __device__ constexpr int tc_ra_alloc() { return A_TMEM ? 0 : TcCfg<A_TMEM>::NA * kStage + 1; }  // odd
// the ring must hold the frames [t, t + lag + 8) of the current stage plus two stages of run-ahead
template <bool A_TMEM> __host__ __device__ constexpr int tc_max_lag() { return tc_rb<A_TMEM>() - 3 * kStage; }
template <bool A_TMEM> __host__ __device__ constexpr size_t tc_smem_bytes() {
  return (size_t)2 * 32 * tc_rb_alloc<A_TMEM>() * 16 + (size_t)2 * 32 * tc_ra_alloc<A_TMEM>() * 16 + 1024;
}

__host__ __device__ inline bool tc_tile_needed(int i0, int j0, int f, int block, bool s0) {
  if (s0 && j0 + kTile - 1 < i0) return false;
  if (block <= 0) return true;
  const int ie = (i0 + kTile < f ? i0 + kTile : f) - 1, je = (j0 + kTile < f ? j0 + kTile : f) - 1;
  const int ia = i0 / block, ib = ie / block, ja = j0 / block, jb = je / block;
  return !(ib < ja || jb < ia);
}

struct TcPlan {
  int nt;            // tiles per axis
  int n_full;        // tiles needing S0 and St (cost 2)
  int n_half;        // tiles needing only one of them (cost 1)
  int per_round;     // items per round of two granules
  int64_t granule;   // frames per granule (multiple of kStage)
  int64_t rounds;
  int64_t n_items;
};

static TcPlan tc_make_plan(int64_t M, int f, int block, bool want_s0, bool want_st, int* full, int* half) {
  TcPlan p{};
  p.nt = (int)ceil_div(f, kTile);
  for (int ti = 0; ti < p.nt; ++ti)
    for (int tj = 0; tj < p.nt; ++tj) {
      const bool a = want_s0 && tc_tile_needed(ti * kTile, tj * kTile, f, block, true);
      const bool b = want_st && tc_tile_needed(ti * kTile, tj * kTile, f, block, false);
      if (a && b) { if (full) full[p.n_full] = ti * p.nt + tj; ++p.n_full; }
      else if (a || b) { if (half) half[p.n_half] = (ti * p.nt + tj) | (a ? 0 : (1 << 30)); ++p.n_half; }
    }
  p.per_round = 2 * p.n_full + p.n_half;
  if (p.per_round == 0) { p.n_items = 0; return p; }
  const int64_t target_items = (int64_t)kNumSMs * 16;
  int64_t rounds = std::max<int64_t>(1, target_items / p.per_round);
  int64_t g = ceil_div(ceil_div(M, 2 * rounds), kStage) * kStage;
  g = std::max<int64_t>(g, 4 * kStage);
  p.granule = g;
  p.rounds = ceil_div(M, 2 * g);
  p.n_items = p.rounds * p.per_round;
  return p;
}

struct TcParams {
  const float* X;
  int64_t n_rows, ld;
  int f, lag, block;
  const float* mean;
  const float* range;
  double* S0;
  double* St;
  const int* full_tiles;
  const int* half_tiles;
  int* counter;
  double* slabs;       // [grid][2][128*128], column-major per tile: slab[col*128 + row]
  TcPlan plan;
  int kc;              // frames between FP32 -> FP64 flushes (multiple of kStage)
};

__global__ void tc_reset_kernel(int* counter) { *counter = 0; }

template <bool A_TMEM, bool X3, bool STD, bool VEC4>
__global__ void __launch_bounds__(kThreads, 1) cov_tc_kernel(const TcParams p) {
  using Cfg = TcCfg<A_TMEM>;
  constexpr int NA = Cfg::NA, NB = Cfg::NB;
  constexpr int RB = tc_rb<A_TMEM>(), RBA = tc_rb_alloc<A_TMEM>(), RAA = tc_ra_alloc<A_TMEM>();

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  unsigned char* b_hi = smem;                                   // [32 groups][RBA rows][16 B]
  unsigned char* b_lo = b_hi + (size_t)32 * RBA * 16;
  unsigned char* a_hi_s = b_lo + (size_t)32 * RBA * 16;         // SS variant only: [32][RAA][16 B]
  unsigned char* a_lo_s = a_hi_s + (size_t)32 * RAA * 16;

  __shared__ uint64_t a_full[NA], a_empty[NA], b_full[NB], b_empty[NB], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_item;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], 4); mbar_init(&b_empty[i], 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t d0_t = tmem, dt_t = tmem + kTile, a_t = tmem + 2 * kTile;

  // running pipeline counters (identical in every role; they persist across work items)
  uint32_t ga = 0, gb = 0, gc = 0;    // A stages, B stages, accumulator chunks consumed so far
  const int64_t M = p.n_rows - p.lag;
  double* slab0 = p.slabs + (size_t)blockIdx.x * 2 * kSlabElems;
  double* slabt = slab0 + kSlabElems;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(p.counter, 1);
    __syncthreads();
    const int64_t item = s_item;
    if (item >= p.plan.n_items) break;

    // ---- decode the work item: (tile, which matrices, frame range) --------------------------
    const int64_t round = item / p.plan.per_round;
    const int rem = (int)(item - round * p.plan.per_round);
    int tile, do_s0, do_st;
    int64_t f0, f1;
    if (rem < 2 * p.plan.n_full) {
      const int g = rem / p.plan.n_full;
      tile = p.full_tiles[rem - g * p.plan.n_full];
      do_s0 = p.S0 != nullptr; do_st = p.St != nullptr;
      f0 = (2 * round + g) * p.plan.granule;
      f1 = f0 + p.plan.granule;
    } else {
      const int code = p.half_tiles[rem - 2 * p.plan.n_full];
      tile = code & ~(1 << 30);
      do_s0 = !(code >> 30); do_st = code >> 30;
      f0 = 2 * round * p.plan.granule;
      f1 = f0 + 2 * p.plan.granule;
    }
    if (f1 > M) f1 = M;
    if (f0 >= f1) continue;                         // uniform across the CTA
    const int i0 = (tile / p.plan.nt) * kTile, j0 = (tile % p.plan.nt) * kTile;
    const int lag = do_st ? p.lag : 0;              // S0-only items need no look-ahead
    const int64_t frames = f1 - f0;
    const uint32_t nA = (uint32_t)((frames + kStage - 1) / kStage);
    const uint32_t nB = nA + (uint32_t)((lag + kStage - 1) / kStage);   // covers every frame any MMA reads
    const uint32_t kc_stages = (uint32_t)(p.kc / kStage);
    const uint32_t nC = (nA + kc_stages - 1) / kc_stages;

    if (warp < 4) {
      // =============================== A producers ===========================================
      const int m = tid;                             // feature row of the tile == TMEM lane
      const int col = i0 + m;
      const bool col_ok = col < p.f;
      float mu = 0.f, rg = 1.f, ri = 1.f;
      if (STD && col_ok) { mu = p.mean[col]; rg = p.range[col]; ri = 1.0f / rg; }
      const float* xcol = p.X + col;
      for (uint32_t sa = 0; sa < nA; ++sa, ++ga) {
        const uint32_t slot = ga % NA;
        mbar_wait(&a_empty[slot], ((ga / NA) & 1) ^ 1);
        const int64_t t0 = f0 + (int64_t)sa * kStage;
        float x[kStage];
#pragma unroll
        for (int j = 0; j < kStage; ++j) {
          const int64_t t = t0 + j;
          x[j] = (col_ok && t < f1) ? __ldg(xcol + t * p.ld) : mu;     // mu -> z == 0 exactly
        }
        uint32_t hi[kStage], lo[kStage];
#pragma unroll
        for (int j = 0; j < kStage; ++j) {
          const float z = STD ? standardize1(x[j], mu, rg, ri) : x[j];
          split_tf32(z, hi[j], lo[j]);
        }
        if constexpr (A_TMEM) {
          tc_fence_after();
          const uint32_t base = a_t + ((uint32_t)(warp * 32) << 16) + slot * 32;
          tmem_st_x16(base, hi);
          if (X3) tmem_st_x16(base + 16, lo);
          tmem_st_wait();
          tc_fence_before();
        } else {
          // [group = m/4][row][m%4]
          float* ah = reinterpret_cast<float*>(a_hi_s) + ((size_t)(m >> 2) * RAA + slot * kStage) * 4 + (m & 3);
          float* al = reinterpret_cast<float*>(a_lo_s) + ((size_t)(m >> 2) * RAA + slot * kStage) * 4 + (m & 3);
#pragma unroll
          for (int j = 0; j < kStage; ++j) {
            ah[j * 4] = __uint_as_float(hi[j]);
            if (X3) al[j * 4] = __uint_as_float(lo[j]);
          }
          fence_proxy_async_smem();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[slot]);
      }
      gb += nB; gc += nC;
    } else if (warp < 8) {
      // =============================== B producers ===========================================
      const int w = warp - 4;                        // rows w, w+4, w+8, w+12 of every stage
      const int c = j0 + lane * 4;                   // 4 feature columns == one MN group
      float mu[4], rg[4], ri[4];
      bool ok[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        ok[v] = c + v < p.f;
        mu[v] = 0.f; rg[v] = 1.f;
        if (STD && ok[v]) { mu[v] = p.mean[c + v]; rg[v] = p.range[c + v]; }
        ri[v] = 1.0f / rg[v];
      }
      const int64_t t_last = f1 + lag;               // frames [f0, t_last) are staged; t_last <= n_rows
      for (uint32_t sb = 0; sb < nB; ++sb, ++gb) {
        const uint32_t slot = gb % NB;
        mbar_wait(&b_empty[slot], ((gb / NB) & 1) ^ 1);
        const int64_t t0 = f0 + (int64_t)sb * kStage;
        float x[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int64_t t = t0 + w + 4 * r;
          const float* xp = p.X + t * p.ld + c;
          if (t < t_last) {
            if (VEC4 && ok[3]) {
              const float4 q = ldg_stream4(xp);
              x[r][0] = q.x; x[r][1] = q.y; x[r][2] = q.z; x[r][3] = q.w;
            } else {
#pragma unroll
              for (int v = 0; v < 4; ++v) x[r][v] = ok[v] ? __ldg(xp + v) : mu[v];
            }
          } else {
#pragma unroll
            for (int v = 0; v < 4; ++v) x[r][v] = mu[v];
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float z = (STD && ok[v]) ? standardize1(x[r][v], mu[v], rg[v], ri[v]) : (ok[v] ? x[r][v] : 0.f);
            split_tf32(z, hi[v], lo[v]);
          }
          const uint32_t row = slot * kStage + w + 4 * r;
          const size_t off = ((size_t)lane * RBA + row) * 16;
          *reinterpret_cast<uint4*>(b_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (X3) *reinterpret_cast<uint4*>(b_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          if (row < 7) {                              // wrap copies behind the last ring row
            const size_t off2 = ((size_t)lane * RBA + RB + row) * 16;
            *reinterpret_cast<uint4*>(b_hi + off2) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (X3) *reinterpret_cast<uint4*>(b_lo + off2) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&b_full[slot]);
      }
      ga += nA; gc += nC;
    } else if (warp < 12) {
      // =============================== epilogue ==============================================
      const int q = warp - 8;                         // TMEM lane quarter == warp % 4
      const int row = q * 32 + lane;                  // tile row (feature i0 + row)
      const uint32_t lane_base = (uint32_t)(q * 32) << 16;
      for (uint32_t c = 0; c < nC; ++c, ++gc) {
        mbar_wait(&acc_full, gc & 1);
        tc_fence_after();
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
          if (which == 0 ? !do_s0 : !do_st) continue;
          double* slab = which == 0 ? slab0 : slabt;
          const uint32_t tbase = (which == 0 ? d0_t : dt_t) + lane_base;
#pragma unroll 1
          for (int c0 = 0; c0 < kTile; c0 += 32) {
            uint32_t v[32];
            tmem_ld_x32(tbase + c0, v);
            double* sp = slab + (size_t)c0 * kTile + row;
            if (c == 0) {
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) sp[(size_t)j * kTile] = (double)__uint_as_float(v[j]);
            } else {
              double old[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) old[j] = sp[(size_t)j * kTile];
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) sp[(size_t)j * kTile] = old[j] + (double)__uint_as_float(v[j]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty);
      }
      // item done: add the FP64 slab into the result (red.global.add.f64, fire and forget)
      const int gi = i0 + row;
      if (gi < p.f) {
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
          if (which == 0 ? !do_s0 : !do_st) continue;
          const double* slab = which == 0 ? slab0 : slabt;
          double* out = (which == 0 ? p.S0 : p.St) + (size_t)gi * p.f + j0;
          const int jn = min(kTile, p.f - j0);
#pragma unroll 4
          for (int j = 0; j < jn; ++j) atomicAdd(out + j, slab[(size_t)j * kTile + row]);
        }
      }
      ga += nA; gb += nB;
    } else {
      // =============================== MMA issuer ============================================
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc_tf32(kTile, kTile, A_TMEM ? 0 : 1, 1);
        const uint32_t bh = smem_u32(b_hi), bl = smem_u32(b_lo);
        const uint32_t ah = smem_u32(a_hi_s), al = smem_u32(a_lo_s);
        const uint32_t gb0 = gb;
        uint32_t b_waited = 0;                       // B stages of this item known to be full
        uint32_t in_chunk = 0;                       // stages issued into the current accumulator chunk
        for (uint32_t sa = 0; sa < nA; ++sa, ++ga) {
          const uint32_t slot = ga % NA;
          if (in_chunk == 0) mbar_wait(&acc_empty, (gc & 1) ^ 1);      // accumulators drained
          mbar_wait(&a_full[slot], (ga / NA) & 1);
          uint32_t need = (sa * kStage + kStage - 1 + lag) / kStage;   // last B stage this A stage reads
          if (need > nB - 1) need = nB - 1;
          while (b_waited <= need) {
            const uint32_t g = gb0 + b_waited;
            mbar_wait(&b_full[g % NB], (g / NB) & 1);
            ++b_waited;
          }
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t u = sa * kStage + h * 8;                    // frame offset inside the item
            const uint32_t first = (in_chunk == 0 && h == 0) ? 0u : 1u;
            const uint32_t r0 = ((gb0 + u / kStage) % NB) * kStage + (u % kStage);
            const uint32_t ul = u + lag;
            const uint32_t r1 = ((gb0 + ul / kStage) % NB) * kStage + (ul % kStage);
            const uint64_t b0h = make_smem_desc(bh + r0 * 16, 128, RBA * 16);
            const uint64_t b0l = make_smem_desc(bl + r0 * 16, 128, RBA * 16);
            const uint64_t bth = make_smem_desc(bh + r1 * 16, 128, RBA * 16);
            const uint64_t btl = make_smem_desc(bl + r1 * 16, 128, RBA * 16);
            if constexpr (A_TMEM) {
              const uint32_t a_h = a_t + slot * 32 + h * 8, a_l = a_h + 16;
              if (do_s0) {
                mma_tf32_ts(d0_t, a_h, b0h, idesc, first);
                if (X3) { mma_tf32_ts(d0_t, a_h, b0l, idesc, 1); mma_tf32_ts(d0_t, a_l, b0h, idesc, 1); }
              }
              if (do_st) {
                mma_tf32_ts(dt_t, a_h, bth, idesc, first);
                if (X3) { mma_tf32_ts(dt_t, a_h, btl, idesc, 1); mma_tf32_ts(dt_t, a_l, bth, idesc, 1); }
              }
            } else {
              const uint32_t ar = (slot * kStage + h * 8) * 16;
              const uint64_t a_h = make_smem_desc(ah + ar, 128, RAA * 16);
              const uint64_t a_l = make_smem_desc(al + ar, 128, RAA * 16);
              if (do_s0) {
                mma_tf32_ss(d0_t, a_h, b0h, idesc, first);
                if (X3) { mma_tf32_ss(d0_t, a_h, b0l, idesc, 1); mma_tf32_ss(d0_t, a_l, b0h, idesc, 1); }
              }
              if (do_st) {
                mma_tf32_ss(dt_t, a_h, bth, idesc, first);
                if (X3) { mma_tf32_ss(dt_t, a_h, btl, idesc, 1); mma_tf32_ss(dt_t, a_l, bth, idesc, 1); }
              }
            }
          }
          mma_commit(&a_empty[slot]);                                  // A stage consumed
          mma_commit(&b_empty[(gb0 + sa) % NB]);                       // frames < 16(sa+1) are dead
          if (++in_chunk == kc_stages || sa + 1 == nA) {
            mma_commit(&acc_full);                                     // accumulators ready to flush
            in_chunk = 0;
            ++gc;
          }
        }
        // release the look-ahead stages the producer filled beyond the last A stage
        for (uint32_t sb = nA; sb < nB; ++sb) {
          const uint32_t g = gb0 + sb;
          if (sb >= b_waited) mbar_wait(&b_full[g % NB], (g / NB) & 1);
          mma_commit(&b_empty[g % NB]);
        }
        {   // do not run ahead of (or exit before) the last asynchronous release
          const uint32_t g = gb0 + nB - 1;
          mbar_wait(&b_empty[g % NB], (g / NB) & 1);
        }
        gb = gb0 + nB;
      } else {
        ga += nA; gb += nB; gc += nC;
      }
      // keep the whole warp's counters identical
      ga = __shfl_sync(0xffffffffu, ga, 0);
      gb = __shfl_sync(0xffffffffu, gb, 0);
      gc = __shfl_sync(0xffffffffu, gc, 0);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem, 512);
}

size_t cov_tc_workspace_bytes(int64_t n_rows, int f, int lag, int block, int engine) {
  (void)n_rows; (void)lag; (void)block; (void)engine;
  const size_t nt = (size_t)ceil_div(f, kTile);
  return 256 + 2 * align_up(nt * nt * sizeof(int), 256) +
         (size_t)kNumSMs * 2 * kSlabElems * sizeof(double);
}

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

template <bool A_TMEM>
static int tc_launch_variant(const TcParams& p, bool x3, bool stdz, bool vec4, int grid, cudaStream_t st) {
  const size_t smem = tc_smem_bytes<A_TMEM>();
#define DCG_TC_CASE(X3, STD, V4)                                                                    \
  {                                                                                                 \
    auto kern = cov_tc_kernel<A_TMEM, X3, STD, V4>;                                                 \
    DCG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, kThreads, smem, st>>>(p);                                                          \
  }
  if (x3) {
    if (stdz) { if (vec4) DCG_TC_CASE(true, true, true) else DCG_TC_CASE(true, true, false) }
    else { if (vec4) DCG_TC_CASE(true, false, true) else DCG_TC_CASE(true, false, false) }
  } else {
    if (stdz) { if (vec4) DCG_TC_CASE(false, true, true) else DCG_TC_CASE(false, true, false) }
    else { if (vec4) DCG_TC_CASE(false, false, true) else DCG_TC_CASE(false, false, false) }
  }
#undef DCG_TC_CASE
  DCG_LAUNCH_CHECK();
  return 0;
}

int cov_tc_launch(const CovArgs& a, cudaStream_t st) {
  int dev = 0, major = 0;
  DCG_CUDA_TRY(cudaGetDevice(&dev));
  DCG_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return DCG_E_ARCH;
  const bool a_tmem = env_int("DCG_TC_A_TMEM", 1) != 0;
  const int max_lag = a_tmem ? tc_max_lag<true>() : tc_max_lag<false>();
  if (a.lag > max_lag) return cov_simt_launch(a, st);   // lag beyond the frame ring: CUDA-core engine
  int kc = env_int("DCG_TC_KC", 1024);
  kc = std::max(kStage, kc / kStage * kStage);

  const int64_t M = a.n_rows - a.lag;
  const size_t nt = (size_t)ceil_div(a.f, kTile);
  char* w = (char*)a.ws;
  int* counter = (int*)w; w += 256;
  int* full = (int*)w; w += align_up(nt * nt * sizeof(int), 256);
  int* half = (int*)w; w += align_up(nt * nt * sizeof(int), 256);
  double* slabs = (double*)w;

  std::vector<int> hfull(nt * nt), hhalf(nt * nt);
  TcPlan plan = tc_make_plan(M, a.f, a.block, a.S0 != nullptr, a.St != nullptr, hfull.data(), hhalf.data());
  if (plan.n_items == 0) return 0;
  if (plan.n_full) DCG_CUDA_TRY(cudaMemcpyAsync(full, hfull.data(), plan.n_full * sizeof(int), cudaMemcpyHostToDevice, st));
  if (plan.n_half) DCG_CUDA_TRY(cudaMemcpyAsync(half, hhalf.data(), plan.n_half * sizeof(int), cudaMemcpyHostToDevice, st));
  tc_reset_kernel<<<1, 1, 0, st>>>(counter);
  DCG_LAUNCH_CHECK();

  TcParams p{a.X, a.n_rows, a.ld, a.f, a.lag, a.block, a.mean, a.range, a.S0, a.St,
             full, half, counter, slabs, plan, kc};
  const int grid = (int)std::min<int64_t>(kNumSMs, plan.n_items);
  const bool x3 = a.engine == DCG_COV_TC_3XTF32;
  const bool stdz = a.mean != nullptr;
  const bool vec4 = row_vec_width(a.X, a.ld) == 4;
  return a_tmem ? tc_launch_variant<true>(p, x3, stdz, vec4, grid, st)
                : tc_launch_variant<false>(p, x3, stdz, vec4, grid, st);
}

}  // namespace dcg
