// Exact covariance engine (DCG_COV_TC_I8X3): integer tensor cores, TMA-staged operands.
//
//     S0[i,j] = sum_t z_t[i] z_t[j]        St[i,j] = sum_t z_t[i] z_{t+lag}[j]       (reference
//     cv_calculator.py:2244-2261: mlcolvar create_timelagged_dataset + TICA.compute's sums)
//
// The float32 engines (cov_tc.cu) accumulate in FP32 inside the tensor core; at C2 (1M frames x
// 1000 features) their accumulation error (1.6e-6 of the sums) moves the TICA eigenvectors by
// 2.7e-5 -- the eigenvector error is ~1000x the RANDOM part of the relative error of the sums --
// which misses the 1e-5 tolerance.  This engine makes the accumulation EXACT instead:
//
//   1. quantise (one HBM-bound pass, i8_quantize_kernel): every value becomes a 23-bit fixed-point
//      integer  q = rint((x - c_j) * 2^e_j),  c_j = mean_j (so x - c_j is the reference's own float32
//      subtraction, cv_calculator.py:834), 2^e_j the largest power of two that keeps |q| < 2^22 given
//      the column's min / max.  Two integer series are written, each as three balanced base-256
//      digits (int8 planes d0, d1, d2: value = d2 2^16 + d1 2^8 + d0, digits in [-128, 127]) in
//      FEATURE-major, frame-contiguous layout (the K-major operand layout of the contraction):
//          Z: q_t                    U: q_t + q_{t+lag}   (the lagged pair sum, |u| < 2^23)
//   2. contract (cov_i8_kernel): tcgen05.mma.cta_group::2.kind::i8 (int8 x int8 -> int32, exact;
//      twice the FP16 rate) on eight of the nine digit products,
//          A = d2 d2                 (weight 2^32)
//          B = d2 d1 + d1 d2         (weight 2^24)
//          C = d2 d0 + d1 d1 + d0 d2 (weight 2^16)
//          D = d1 d0 + d0 d1         (weight 2^8)         [only d0 d0 is left out: < 2^-30 of the scale.
//             Leaving D out as well costs 4e-9 rms of the sums at M = 1e6 -- and 1.3e-5 in the TICA
//             eigenvectors at C2, measured]
//      into four int32 TMEM accumulators per 256 x N tile (N <= 128: 4 N <= 512 TMEM columns), for up
//      to 32768 frames per work item with no intermediate rounding at all.  Both Grams are symmetric:
//      only tiles that touch the upper triangle are computed,
//          S0 = Z^T Z,      G = U^T U = S0 + S0' + St + St^T      (S0' = S0 shifted by `lag` frames)
//      and the symmetric part of St -- the only part the reference uses: mlcolvar symmetrises C_tau --
//      follows EXACTLY (same integers) as  (St + St^T)/2 = (G - 2 S0 + H - T)/2  with the 2 lag boundary
//      rows H = sum_{t<lag} z z^T, T = sum_{t>=M} z z^T (i8_finish_kernel).  This needs 2 x 20 instead of
//      20 + 28 tiles at F = 1000, and no operand is ever read at a frame offset: TMA boxes must start
//      16-byte aligned in global memory, which an arbitrary lag on 1-byte elements is not (measured:
//      cudaErrorIllegalInstruction, tools_dev/tma_probe.cu).
//      Operands are staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B boxes of 128 frames) straight from
//      the digit planes (128-byte rows: with 64-byte rows the TMA unit, ~2.3 clocks per box row, was the
//      bound): no thread touches an operand.  Warp roles: 0 = TMA producer, 1 = MMA issuer
//      (leader CTA), 2-5 = epilogue (int32 -> FP64: A 2^32 + B 2^24 + C 2^16 is exact in a double,
//      scaled by 2^-e_i 2^-e_j / (range_i range_j), red.global.add.f64).
//
// Error model: only the fixed-point rounding of the inputs remains, |z - z'| <= 2^-23 max|x - c| /
// range -- the float32 resolution of the data -- unbiased and independent between frames, so its
// effect on the sums is ~ sqrt(2/M) 2^-23: 1e-10 relative at M = 1e6, against a worst case of 1e-6.
//
// Frames are processed in windows (the planes of a window live in the caller's workspace).
// Roofline: tensor pipe (kind::i8); HBM for the quantise pass (4 + 6 bytes per element).
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include "dcg_common.cuh"
#include "tc_common.cuh"

namespace dcg {

using namespace tc;

namespace {

constexpr int kTileM = 256;            // tile rows (features I): UMMA M of the CTA pair
constexpr int kHalfM = 128;            // rows staged per CTA
constexpr int kMaxN = 128;             // tile columns (features J): 4 accumulators x N <= 512 TMEM columns
constexpr int kStageFrames = 128;      // frames per pipeline stage = one SWIZZLE_128B box row (a full 128-byte line)
constexpr int kPlaneA = kHalfM * kStageFrames;                 // 8 KB: one digit plane of the A half
constexpr int kMaxItemFrames = 32768;  // int32 headroom: 3 * 128 * 128 * 32768 < 2^31
constexpr int kKSteps = kStageFrames / 32;    // kind::i8 MMAs have K = 32
constexpr int kThreads = 192;          // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kQMax = 4161536;         // 2^22 - 2^15: |q_t + q_{t+lag}| <= 2^23 - 2^16, top digit in [-127, 127]

__host__ __device__ constexpr int plane_b_bytes(int N) { return (N / 2) * kStageFrames; }
__host__ __device__ constexpr int stage_bytes(int N) { return 3 * kPlaneA + 3 * plane_b_bytes(N); }
inline int num_stages(int N) { return std::min(8, (int)((227 * 1024 - 2048) / stage_bytes(N))); }

struct Tile8 {
  int ia0, ia1;        // first feature of the 128 rows staged by CTA 0 / CTA 1 of the pair (-1: none)
  int j0;              // first feature of the N columns
  int hi;              // one past the last valid feature (block / matrix edge), rows and columns
  int kind;            // 0 = S0 = Z^T Z, 1 = G = U^T U; only the upper triangle (j >= i) is written
  int tr0, tr1;        // the half computes a block BELOW the diagonal and writes its transpose
};

// Work tiles of one symmetric Gram per diagonal region (the whole matrix, or each hTICA block).
// N == 128: the region is a grid of 128 x 128 blocks of which the upper triangle incl. the diagonal is
// needed; a tile pairs two blocks that share their COLUMN block (the B operand of the pair MMA) -- a
// block (r, c) may also be computed as (c, r) and written transposed, so the odd block left over in
// every even column c = 4m (its diagonal block) is paired with the one of column 4m + 2, block (4m, 4m+2),
// computed as (4m+2, 4m): nb (nb + 1) / 2 blocks in ceil(that / 2) tiles (18 instead of 20 at nb = 8, 5
// instead of 6 at nb = 4).  Other N (narrow matrices): 256-row x N tiles that touch the upper triangle.
__host__ __device__ inline int enum_tiles8(int f, int block, int N, bool want_s0, bool want_st, Tile8* out) {
  int n = 0;
  const int w = block > 0 ? block : f;
  for (int b0 = 0; b0 < f; b0 += w) {
    const int b1 = b0 + w < f ? b0 + w : f;
    for (int kind = 0; kind < 2; ++kind) {
      if (!(kind == 0 ? want_s0 : want_st)) continue;
      if (N != kHalfM) {
        for (int i0 = b0; i0 < b1; i0 += kTileM)
          for (int j0 = b0; j0 < b1; j0 += N) {
            if (j0 + N <= i0) continue;
            if (out) out[n] = Tile8{i0, i0 + kHalfM < b1 ? i0 + kHalfM : -1, j0, b1, kind, 0, 0};
            ++n;
          }
        continue;
      }
      const int nb = (b1 - b0 + kHalfM - 1) / kHalfM;
      for (int c = 0; c < nb; ++c) {
        const int skip = (c & 1) ? -1 : ((c & 3) == 0 ? c : c - 2);      // the block left over in an even column
        int first = -1;
        for (int r = 0; r <= c; ++r) {
          if (r == skip) continue;
          if (first < 0) { first = r; continue; }
          if (out) out[n] = Tile8{b0 + first * kHalfM, b0 + r * kHalfM, b0 + c * kHalfM, b1, kind, 0, 0};
          ++n;
          first = -1;
        }
        if ((c & 3) == 0) {            // diagonal block (c, c) + block (c + 2, c), written transposed as (c, c + 2)
          if (out) out[n] = Tile8{b0 + c * kHalfM, c + 2 < nb ? b0 + (c + 2) * kHalfM : -1, b0 + c * kHalfM, b1, kind, 0, 1};
          ++n;
        }
      }
    }
  }
  return n;
}

__global__ void plan8_kernel(Tile8* tiles, int f, int block, int N, bool want_s0, bool want_st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) enum_tiles8(f, block, N, want_s0, want_st, tiles);
}

// ---- per-feature quantisation parameters -----------------------------------------------------------
// shift[j] = c_j, mul[j] = 2^e_j (float), scale[j] = 2^-e_j / range_j (double)
__global__ void i8_prep_kernel(int f, const float* __restrict__ mean, const float* __restrict__ range,
                               const float* __restrict__ xmin, const float* __restrict__ xmax,
                               float* __restrict__ shift, float* __restrict__ mul, double* __restrict__ scale) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= f) return;
  const float c = mean ? mean[j] : 0.f;
  const float r = range ? range[j] : 1.f;
  float amax = fmaxf(fabsf(xmax[j] - c), fabsf(xmin[j] - c));
  if (!(amax > 0.f) || !isfinite(amax)) amax = 1.f;
  // largest power of two with amax * 2^e <= kQMax (one ulp of slack for the float32 subtraction)
  int e = ilogbf((float)kQMax / (amax * 1.0000002f));
  e = e > 100 ? 100 : (e < -100 ? -100 : e);
  shift[j] = c;
  mul[j] = ldexpf(1.f, e);
  scale[j] = ldexp(1.0, -e) / (double)r;
}

// ---- quantise: X (frames x features, float32) -> int8 digit planes of Z and U, feature-major ---------
constexpr int kQF = 64;     // features per tile
constexpr int kQT = 128;    // frames per tile (one 128-byte line per feature and plane)
constexpr int kQRow = 33;   // words per shared-memory row (128 bytes + 4 padding)

__device__ __forceinline__ int quantize1(float x, float c, float m, int& clamped) {
  const float y = (x - c) * m;                         // the reference's float32 subtraction, then an exact scaling
  clamped += fabsf(y) > (float)kQMax + 0.5f;
  return max(-kQMax, min(kQMax, __float2int_rn(y)));
}

// byte K of a, b, c, d -> one word (a in the lowest byte)
template <int K>
__device__ __forceinline__ uint32_t pack_byte(int a, int b, int c, int d) {
  const uint32_t lo = __byte_perm((uint32_t)a, (uint32_t)b, ((4 + K) << 4) | K);
  const uint32_t hi = __byte_perm((uint32_t)c, (uint32_t)d, ((4 + K) << 4) | K);
  return __byte_perm(lo, hi, 0x5410);
}
// Balanced base-256 digits of four integers (|q| < 2^23), one word per digit:
//   d0 = byte 0 of q,  d1 = byte 0 of (q + 128) >> 8 = byte 1 of q + 128,
//   d2 = byte 0 of (((q + 128) >> 8) + 128) >> 8 = byte 2 of q + 128 + 32768      (floors nest)
__device__ __forceinline__ void digits4(const int (&q)[4], uint32_t& w0, uint32_t& w1, uint32_t& w2) {
  w0 = pack_byte<0>(q[0], q[1], q[2], q[3]);
  w1 = pack_byte<1>(q[0] + 128, q[1] + 128, q[2] + 128, q[3] + 128);
  w2 = pack_byte<2>(q[0] + 32896, q[1] + 32896, q[2] + 32896, q[3] + 32896);
}

// planes: [set (Z, U)][digit 0..2][feature][frame], frame stride 1, feature stride wpad, digit stride
// f * wpad, set stride 3 * f * wpad.  `pairs` frames of the window starting at row `row0` of X.
template <bool WITH_U>
__global__ void __launch_bounds__(256, 3)
i8_quantize_kernel(const float* __restrict__ X, int64_t row0, int64_t pairs, int lag, int f, int64_t ld,
                   const float* __restrict__ shift, const float* __restrict__ mul,
                   int8_t* __restrict__ planes, int64_t wpad, int tiles_per_block,
                   long long* __restrict__ qsum_t, long long* __restrict__ qsum_lag, int* __restrict__ info, int vec4) {
  extern __shared__ __align__(16) unsigned char qsmem[];
  uint32_t (*tile)[kQF][kQRow] = reinterpret_cast<uint32_t (*)[kQF][kQRow]>(qsmem);   // [3 or 6 planes][feature][word]
  __shared__ unsigned long long ssum[2][kQF];
  const int tid = threadIdx.x;
  // 4 features x 4 frames per thread and pass.  A warp covers 8 feature quads x 4 frame quads so that its
  // transposed shared-memory stores (word index 33 * feature + frame quad) hit 32 different banks; its
  // global loads are 128-byte row segments.
  const int fe4 = ((tid >> 5) & 1) * 8 + (tid & 7), frq = (tid >> 6) * 4 + ((tid >> 3) & 3);
  const int f0 = blockIdx.x * kQF;
  const int fc = f0 + 4 * fe4;
  float c[4], m[4];
  bool ok[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    ok[v] = fc + v < f;
    c[v] = ok[v] ? shift[fc + v] : 0.f;
    m[v] = ok[v] ? mul[fc + v] : 0.f;
  }
  if (tid < 2 * kQF) ssum[tid / kQF][tid % kQF] = 0ull;
  long long acc_t[4] = {0, 0, 0, 0}, acc_l[4] = {0, 0, 0, 0};
  int clamped = 0;
  const size_t plane_stride = (size_t)f * (size_t)wpad;
  const size_t lag_off = (size_t)lag * (size_t)ld;
  const int64_t tile0 = (int64_t)blockIdx.y * tiles_per_block;
  for (int tt = 0; tt < tiles_per_block; ++tt) {
    const int64_t t0 = (tile0 + tt) * kQT;               // first frame of the tile (window-local)
    if (t0 >= pairs) break;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      float x[4][4], xl[4][4];                           // [frame r][feature v]
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int64_t t = t0 + 64 * h + 4 * frq + r;
        const float* px = X + (size_t)(row0 + t) * (size_t)ld + fc;
        const bool live = t < pairs;
        if (live && vec4 && ok[3]) {
          const float4 a = ldg_stream4(px);
          x[r][0] = a.x; x[r][1] = a.y; x[r][2] = a.z; x[r][3] = a.w;
          if (WITH_U) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(px + lag_off));
            xl[r][0] = b.x; xl[r][1] = b.y; xl[r][2] = b.z; xl[r][3] = b.w;
          }
        } else {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            x[r][v] = (live && ok[v]) ? ldg_stream1(px + v) : c[v];
            if (WITH_U) xl[r][v] = (live && ok[v]) ? __ldg(px + lag_off + v) : c[v];
          }
        }
      }
      const int w = 16 * h + frq;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        int q[4], u[4];
        int ps = 0, pl = 0;                              // |q| < 2^22: four of them fit an int
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          q[r] = quantize1(x[r][v], c[v], m[v], clamped);
          ps += q[r];
          if (WITH_U) {
            const int ql = quantize1(xl[r][v], c[v], m[v], clamped);
            pl += ql;
            u[r] = q[r] + ql;
          }
        }
        acc_t[v] += ps;
        if (WITH_U) acc_l[v] += pl;
        const int fe = 4 * fe4 + v;
        uint32_t w0, w1, w2;
        digits4(q, w0, w1, w2);
        tile[0][fe][w] = w0; tile[1][fe][w] = w1; tile[2][fe][w] = w2;
        if (WITH_U) {
          digits4(u, w0, w1, w2);
          tile[3][fe][w] = w0; tile[4][fe][w] = w1; tile[5][fe][w] = w2;
        }
      }
    }
    __syncthreads();
    // (3 or 6) planes x 64 features x 128 bytes -> global, 16 bytes per thread and step, 128-byte lines
#pragma unroll
    for (int it = 0; it < (WITH_U ? 12 : 6); ++it) {
      const int chunk = it * 256 + tid;
      const int p = chunk >> 9, row = (chunk >> 3) & 63, cw = chunk & 7;
      if (f0 + row < f) {
        const uint32_t* src = &tile[p][row][4 * cw];
        const uint4 v4 = make_uint4(src[0], src[1], src[2], src[3]);
        int8_t* dst = planes + (size_t)p * plane_stride + (size_t)(f0 + row) * (size_t)wpad + (size_t)t0 + 16 * cw;
        *reinterpret_cast<uint4*>(dst) = v4;
      }
    }
    __syncthreads();
  }
  if (qsum_t) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      atomicAdd(&ssum[0][4 * fe4 + v], (unsigned long long)acc_t[v]);
      if (WITH_U) atomicAdd(&ssum[1][4 * fe4 + v], (unsigned long long)acc_l[v]);
    }
    __syncthreads();
    if (tid < kQF && f0 + tid < f) {
      atomicAdd((unsigned long long*)qsum_t + f0 + tid, ssum[0][tid]);
      if (WITH_U) atomicAdd((unsigned long long*)qsum_lag + f0 + tid, ssum[1][tid]);
    }
  }
  if (clamped && info) atomicAdd(info, clamped);
}

// Finish: column sums from the exact integer totals, and the symmetric part of St from the two Grams.
//   a = sum_{t<M} z_t, b = sum_{t>=lag} z_t = sum_{t<M} z_{t+lag}
//   (St + St^T)/2 = (G - 2 S0 + H - T)/2,  H = sum_{t<lag} z z^T,  T = sum_{t>=M} z z^T  (rows re-quantised
//   here with the same arithmetic, exact integers); written to BOTH triangles of St.  S0 keeps its upper
//   triangle (dcg.h).  Entries outside the diagonal blocks (block mode) are left as they are.
__global__ void i8_finish_sums_kernel(int f, const double* __restrict__ scale, const long long* __restrict__ qsum_t,
                                      const long long* __restrict__ qsum_lag, int lag,
                                      double* __restrict__ sum_t, double* __restrict__ sum_lag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= f) return;
  if (sum_t) sum_t[j] = (double)qsum_t[j] * scale[j];
  if (sum_lag) sum_lag[j] = (double)(lag > 0 ? qsum_lag[j] : qsum_t[j]) * scale[j];
}

__global__ void __launch_bounds__(256)
i8_finish_st_kernel(const float* __restrict__ X, int64_t n_rows, int f, int64_t ld, int lag, int block,
                    const float* __restrict__ shift, const float* __restrict__ mul, const double* __restrict__ scale,
                    const double* __restrict__ S0, double* __restrict__ St) {
  const int j = blockIdx.x * 16 + (threadIdx.x & 15), i = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (i >= f || j >= f || j < i) return;
  if (block > 0 && i / block != j / block) return;
  const int64_t M = n_rows - lag;
  const float ci = shift[i], mi = mul[i], cj = shift[j], mj = mul[j];
  long long d = 0;                                        // H - T in integer units
  int dummy = 0;
  for (int t = 0; t < lag; ++t) {
    const float* rh = X + (size_t)t * ld;
    const float* rt = X + (size_t)(M + t) * ld;
    d += (long long)quantize1(rh[i], ci, mi, dummy) * quantize1(rh[j], cj, mj, dummy);
    d -= (long long)quantize1(rt[i], ci, mi, dummy) * quantize1(rt[j], cj, mj, dummy);
  }
  const size_t ij = (size_t)i * f + j, ji = (size_t)j * f + i;
  const double v = 0.5 * (St[ij] - 2.0 * S0[ij] + (double)d * scale[i] * scale[j]);
  St[ij] = v;
  St[ji] = v;
}

// ---- contraction ------------------------------------------------------------------------------------
struct Params8 {
  const Tile8* tiles;
  int n_tiles;
  int64_t n_items;       // n_tiles * n_ranges, range-major
  int64_t granule;       // frames per work item (multiple of kStageFrames)
  int64_t pairs;         // lagged pairs in this window
  int N;                 // tile width (multiple of 32, <= kMaxN)
  int n_stages;
  int f;
  double* S0;
  double* St;
  const double* scale;   // per feature: 2^-e / range
  int dbg;               // diagnostics (DCG_I8_DBG): 1 = no TMA, 2 = no MMA, 4 = no TMEM loads
};

__device__ __forceinline__ uint32_t cluster_ctarank8() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync8() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote8(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster8(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 4-D tiled TMA load (frame, feature, digit, set) into this CTA's shared memory; completion (bytes) is
// signalled on the barrier at `bar_addr` -- a shared::cluster address, the LEADER CTA's barrier for both
// CTAs of the pair.  The box must start 16-byte aligned in global memory: frame coordinates are
// multiples of 64 here.
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar_addr,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tmem_alloc2_8(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
               ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2_8(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_i8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma2_commit_both8(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// kind::i8 instruction descriptor: signed 8-bit operands, both K-major, int32 accumulate, dense
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// K-major SWIZZLE_128B operand: rows of 128 bytes, 8-row swizzle atoms of 1024 bytes (= SBO)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t start) {
  uint64_t d = 0;
  d |= (uint64_t)((start >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                 // LBO (unused for swizzled K-major operands)
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
cov_i8_kernel(const __grid_constant__ CUtensorMap map_a,      // planes [set][digit][feature][frame], box 64 x 128
              const __grid_constant__ CUtensorMap map_b,      // same tensor, box 64 x N/2
              const Params8 p) {
  extern __shared__ unsigned char smem_raw8[];
  const uint32_t smem_base = (smem_u32(smem_raw8) + 1023u) & ~1023u;
  __shared__ uint64_t full_bar[8], empty_bar[8], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;
  __shared__ double epi_smem[4 * 32 * 9];          // per epilogue warp: 32 rows x 8 columns (+1 padding)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank8();
  const int N = p.N, NS = p.n_stages;
  const uint32_t pb = (uint32_t)plane_b_bytes(N), sb = (uint32_t)stage_bytes(N);

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, 2 * 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2_8(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync8();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  uint32_t gs = 0, gi = 0;       // stages / items consumed so far (all roles agree)
  const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  for (int64_t item = cluster_id; item < p.n_items; item += n_clusters, ++gi) {
    const int64_t range = item / p.n_tiles;
    const Tile8 td = p.tiles[item - range * p.n_tiles];
    const int64_t f0 = range * p.granule;
    const int64_t f1 = f0 + p.granule < p.pairs ? f0 + p.granule : p.pairs;
    const uint32_t nS = (uint32_t)((f1 - f0 + kStageFrames - 1) / kStageFrames);

    if (warp == 0) {
      // ===================== TMA producer (one lane per CTA) ===========================================
      if (lane == 0) {
        const int set = td.kind;                                   // 0: Z planes (S0), 1: U planes (G)
        // a CTA without a row block stages rows beyond the tensor extent: zero-filled by the TMA unit
        const int ia_t = rank ? td.ia1 : td.ia0;
        const int ia = ia_t < 0 ? p.f : ia_t, jb = td.j0 + (int)rank * (N / 2);
        for (uint32_t s = 0; s < nS; ++s) {
          const uint32_t g = gs + s, slot = g % NS;
          mbar_wait(&empty_bar[slot], ((g / NS) & 1) ^ 1);
          if (p.dbg & 1) { if (rank == 0) mbar_arrive(&full_bar[slot]); continue; }
          // both CTAs' copies complete on the leader's barrier (peer bit of the address cleared)
          const uint32_t bar = smem_u32(&full_bar[slot]) & 0xFEFFFFFFu;
          if (rank == 0) mbar_expect_tx(&full_bar[slot], 2 * sb);
          const uint32_t dst = smem_base + slot * sb;
          // items start at multiples of the stage length; frames at and beyond the window's pair count
          // are zero-filled by the TMA unit (tensor extent), so a last partial stage needs no masking
          const int t = (int)(f0 + (int64_t)s * kStageFrames);
#pragma unroll
          for (int pl = 0; pl < 3; ++pl) tma_load_4d(dst + pl * kPlaneA, &map_a, bar, t, ia, pl, set);
#pragma unroll
          for (int pl = 0; pl < 3; ++pl) tma_load_4d(dst + 3 * kPlaneA + pl * pb, &map_b, bar, t, jb, pl, set);
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (leader CTA) ===================================================
      if (rank == 0) {
        const uint32_t idesc = make_idesc_i8(kTileM, N);
        const uint32_t accA = tmem, accB = tmem + N, accC = tmem + 2 * N, accD = tmem + 3 * N;
        mbar_wait_cluster8(&acc_empty, (gi & 1) ^ 1);
        tc_fence_after();
        for (uint32_t s = 0; s < nS; ++s) {
          const uint32_t g = gs + s, slot = g % NS;
          mbar_wait_cluster8(&full_bar[slot], (g / NS) & 1);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t st = smem_base + slot * sb;
            // planes are stored d0, d1, d2 (plane index = digit index)
            const uint64_t a0 = make_smem_desc_sw128(st), a1 = make_smem_desc_sw128(st + kPlaneA),
                           a2 = make_smem_desc_sw128(st + 2 * kPlaneA);
            const uint32_t sbb = st + 3 * kPlaneA;
            const uint64_t b0 = make_smem_desc_sw128(sbb), b1 = make_smem_desc_sw128(sbb + pb),
                           b2 = make_smem_desc_sw128(sbb + 2 * pb);
#pragma unroll
            for (int h = 0; h < kKSteps && !(p.dbg & 2); ++h) {
              const uint64_t o = (uint64_t)(h * 32 >> 4);            // 32 frames = 32 bytes along K
              const uint32_t first = (s == 0 && h == 0) ? 0u : 1u;
              mma2_i8_ss(accA, a2 + o, b2 + o, idesc, first);
              mma2_i8_ss(accB, a2 + o, b1 + o, idesc, first);
              mma2_i8_ss(accB, a1 + o, b2 + o, idesc, 1u);
              mma2_i8_ss(accC, a2 + o, b0 + o, idesc, first);
              mma2_i8_ss(accC, a1 + o, b1 + o, idesc, 1u);
              mma2_i8_ss(accC, a0 + o, b2 + o, idesc, 1u);
              mma2_i8_ss(accD, a1 + o, b0 + o, idesc, first);
              mma2_i8_ss(accD, a0 + o, b1 + o, idesc, 1u);
            }
            mma2_commit_both8(&empty_bar[slot]);
            if (s + 1 == nS) mma2_commit_both8(&acc_full);
          }
          __syncwarp();
        }
      }
    } else {
      // ===================== epilogue: int32 accumulators -> FP64 result ================================
      const int q = warp & 3;                                  // TMEM lane quarter of this warp
      const uint32_t lane_base = (uint32_t)(q * 32) << 16;
      mbar_wait(&acc_full, gi & 1);
      tc_fence_after();
      // Row `32 q + lane` of this CTA's half of the tile sits in this lane.  The FP64 adds go out 8 columns
      // of 4 rows per instruction (64 contiguous bytes per row) through a small shared-memory transpose:
      // one row per lane would touch 32 different lines per instruction (measured: 20 us per work item).
      const int ia_t = rank ? td.ia1 : td.ia0;
      const bool transposed = (rank ? td.tr1 : td.tr0) != 0;
      const int row0 = ia_t + 32 * q;
      const int gi_row = row0 + lane;
      const bool row_ok = ia_t >= 0 && gi_row < td.hi;
      const double si = row_ok ? p.scale[gi_row] : 0.0;
      double* sE = epi_smem + (warp - 2) * (32 * 9);
      double* outm = td.kind == 0 ? p.S0 : p.St;
      for (int c0 = 0; c0 < N && !(p.dbg & 4) && ia_t >= 0; c0 += 16) {
        uint32_t va[16], vb[16], vc[16], vd[16];
        tmem_ld_x16(tmem + lane_base + c0, va);
        tmem_ld_x16(tmem + N + lane_base + c0, vb);
        tmem_ld_x16(tmem + 2 * N + lane_base + c0, vc);
        tmem_ld_x16(tmem + 3 * N + lane_base + c0, vd);
        tmem_ld_wait();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int cb = td.j0 + c0 + 8 * half;                   // first column of this group of 8
          if (cb >= td.hi) continue;                               // beyond the edge (warp-uniform)
          if (!transposed && cb + 7 < row0) continue;              // wholly below the diagonal (warp-uniform)
          double v8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int jj = 8 * half + j, gj = cb + j;
            // sum of the digit products down to weight 2^8 (only d0 d0 is left out): an integer below
            // 2^63 whose double is exact to 2^-53 relative
            const double v = (double)(int)va[jj] * 4294967296.0 + (double)(int)vb[jj] * 16777216.0 +
                             (double)(int)vc[jj] * 65536.0 + (double)(int)vd[jj] * 256.0;
            v8[j] = gj < td.hi ? v * si * p.scale[gj] : 0.0;
          }
          if (transposed) {
            // block below the diagonal: entry (gi, gj) goes to (gj, gi); for a fixed column the 32 lanes
            // write 32 consecutive doubles -- coalesced straight from the registers
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (row_ok && cb + j < td.hi) atomicAdd(outm + (size_t)(cb + j) * p.f + gi_row, v8[j]);
            continue;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) sE[lane * 9 + j] = v8[j];
          __syncwarp();
          const int col = lane & 7, gj = cb + col;
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            const int row = 4 * rr + (lane >> 3), gi = row0 + row;
            if (gi < td.hi && gj < td.hi && gj >= gi) atomicAdd(outm + (size_t)gi * p.f + gj, sE[row * 9 + col]);
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote8(&acc_empty, 0);
    }
    gs += nS;
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync8();
  if (warp == 1) tmem_dealloc2_8(tmem, 512);
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;          // resolved once; the driver entry point is process-wide
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// planes[2 sets][3 digits][f][wpad] int8; dims (frames, features, digits, sets); box (64 frames, rows, 1, 1)
int make_plane_map(CUtensorMap* map, const int8_t* planes, int64_t frames, int f, int64_t wpad, int n_sets, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return DCG_E_ARCH;
  const cuuint64_t dims[4] = {(cuuint64_t)frames, (cuuint64_t)f, 3, (cuuint64_t)n_sets};
  const cuuint64_t strides[3] = {(cuuint64_t)wpad, (cuuint64_t)wpad * (cuuint64_t)f, (cuuint64_t)wpad * (cuuint64_t)f * 3};
  const cuuint32_t box[4] = {(cuuint32_t)kStageFrames, (cuuint32_t)box_rows, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void*)planes, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : DCG_E_SHAPE;
}

int64_t env_i64(const char* name, int64_t dflt) {
  const char* s = getenv(name);
  return s ? atoll(s) : dflt;
}

// Tile width N (multiple of 32, <= 160).  Wider tiles read less shared memory per MMA (A 4 KB + B N/2 x 32
// bytes per 256 x N x 32 instruction), narrower ones waste less on the diagonal and on the last column
// tile; DCG_I8_N overrides.
int tile_width(int f, int block) {
  const int w = block > 0 ? std::min(block, f) : f;
  const int forced = (int)env_i64("DCG_I8_N", 0);
  if (forced >= 32 && forced <= kMaxN && forced % 32 == 0) return forced;
  int best = 32;
  double best_cost = 1e300;
  for (int N = 32; N <= kMaxN; N += 32) {
    // upper-triangle tiles of one diagonal block x (MMA cycles ~ N) / (shared-memory efficiency)
    int tiles = 0;
    for (int i0 = 0; i0 < w; i0 += kTileM)
      for (int j0 = 0; j0 < w; j0 += N) tiles += j0 + N > i0;
    const double smem_rate = (4096.0 + 32.0 * (N / 2)) / (0.5 * N) + 48.0;     // bytes per clock at full MMA rate
    const double cost = tiles * (double)N * std::max(1.0, smem_rate / 128.0);
    if (cost < best_cost) { best_cost = cost; best = N; }
  }
  return best;
}

// DCG_I8_DEBUG=1: synchronise after every launch and name the failing one on stderr (diagnostics only)
#define DCG_I8_STEP(name)                                                                    \
  do {                                                                                       \
    DCG_LAUNCH_CHECK();                                                                      \
    if (getenv("DCG_I8_DEBUG")) {                                                            \
      cudaError_t _e = cudaStreamSynchronize(st);                                            \
      fprintf(stderr, "[dcg i8] %s: %s\n", name, cudaGetErrorString(_e));                    \
      if (_e != cudaSuccess) return -(int)_e;                                                \
    }                                                                                        \
  } while (0)

// frames per window: the digit planes of a window (3 or 6 bytes per value) stay under ~6 GB
int64_t window_frames(int64_t n_rows, int f, int lag) {
  const int64_t cap = env_i64("DCG_I8_WINDOW_BYTES", (int64_t)6 << 30);
  const int sets = lag > 0 ? 2 : 1;
  int64_t w = std::max<int64_t>(cap / (3 * sets * (int64_t)f), 4096) / 128 * 128;
  return std::min<int64_t>(w, std::max<int64_t>(n_rows - lag, 1));
}

// ---- optional device timing of the two kernels (bench.py's roofline: CUDA events on the launching stream) --
struct Timing8 {
  bool on = false;
  int n = 0;                                 // contraction launches recorded since the last read
  cudaEvent_t q0[64], q1[64], c1[64];        // quantise start / quantise end = contraction start / contraction end
  bool created = false;
};
Timing8 g_timing;                            // diagnostics only: one stream, one thread

struct Layout8 {
  size_t tiles, shift, mul, scale, qsum, planes, total;
  int64_t wf, wpad;
  int n_tiles_max, n_sets;
};

Layout8 layout8(int64_t n_rows, int f, int lag, int block) {
  Layout8 L;
  const int N = tile_width(f, block);
  L.n_tiles_max = enum_tiles8(f, block, N, true, true, nullptr);
  L.n_sets = lag > 0 ? 2 : 1;
  L.wf = window_frames(n_rows, f, lag);
  L.wpad = (int64_t)align_up((size_t)L.wf, 128);
  size_t o = 256;
  L.tiles = o; o += align_up((size_t)L.n_tiles_max * sizeof(Tile8), 256);
  L.shift = o; o += align_up((size_t)f * 4, 256);
  L.mul = o; o += align_up((size_t)f * 4, 256);
  L.scale = o; o += align_up((size_t)f * 8, 256);
  L.qsum = o; o += align_up((size_t)f * 16, 256);
  L.planes = align_up(o, 1024); o = L.planes + (size_t)3 * L.n_sets * (size_t)f * (size_t)L.wpad;
  L.total = o + 1024;
  return L;
}

}  // namespace

}  // namespace dcg

using namespace dcg;

extern "C" int dcg_cov_i8_set_timing(int on) {
  if (on && !g_timing.created) {
    for (int i = 0; i < 64; ++i) {
      DCG_CUDA_TRY(cudaEventCreate(&g_timing.q0[i]));
      DCG_CUDA_TRY(cudaEventCreate(&g_timing.q1[i]));
      DCG_CUDA_TRY(cudaEventCreate(&g_timing.c1[i]));
    }
    g_timing.created = true;
  }
  g_timing.on = on != 0;
  g_timing.n = 0;
  return 0;
}

extern "C" int dcg_cov_i8_get_timing(float* quantize_ms, float* contract_ms, int* launches) {
  float q = 0.f, c = 0.f;
  for (int i = 0; i < g_timing.n; ++i) {
    float a = 0.f, b = 0.f;
    DCG_CUDA_TRY(cudaEventSynchronize(g_timing.c1[i]));
    DCG_CUDA_TRY(cudaEventElapsedTime(&a, g_timing.q0[i], g_timing.q1[i]));
    DCG_CUDA_TRY(cudaEventElapsedTime(&b, g_timing.q1[i], g_timing.c1[i]));
    q += a; c += b;
  }
  if (quantize_ms) *quantize_ms = q;
  if (contract_ms) *contract_ms = c;
  if (launches) *launches = g_timing.n;
  g_timing.n = 0;
  return 0;
}

extern "C" size_t dcg_cov_i8_workspace_bytes(int64_t n_rows, int f, int lag, int block) {
  if (n_rows <= 0 || f <= 0 || lag < 0 || lag >= n_rows || block < 0) return 0;
  return layout8(n_rows, f, lag, block).total;
}

extern "C" int dcg_cov_lag_i8_f32(const float* X, int64_t n_rows, int f, int64_t ld, int lag,
                                  const float* mean, const float* range, const float* xmin, const float* xmax,
                                  int block, double* S0, double* St, double* colsum_t, double* colsum_lag,
                                  int* info, void* ws, size_t ws_bytes, void* stream) {
  if (!X || (!S0 && !St) || !xmin || !xmax) return DCG_E_NULL;
  if ((mean == nullptr) != (range == nullptr)) return DCG_E_NULL;
  if (n_rows <= 0 || f <= 0 || ld < f || lag < 0 || lag >= n_rows || block < 0) return DCG_E_SHAPE;
  if (lag == 0) St = nullptr;
  if (St && !S0) return DCG_E_NULL;                  // the symmetric part of St is derived from S0 and G
  const Layout8 L = layout8(n_rows, f, lag, block);
  if (!ws || ws_bytes < L.total) return DCG_E_WORKSPACE;
  int dev = 0, major = 0;
  DCG_CUDA_TRY(cudaGetDevice(&dev));
  DCG_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return DCG_E_ARCH;
  cudaStream_t st = (cudaStream_t)stream;

  char* base = (char*)ws;
  Tile8* d_tiles = (Tile8*)(base + L.tiles);
  float* d_shift = (float*)(base + L.shift);
  float* d_mul = (float*)(base + L.mul);
  double* d_scale = (double*)(base + L.scale);
  long long* d_qsum = (long long*)(base + L.qsum);
  int8_t* d_planes = (int8_t*)(base + L.planes);

  const int64_t M = n_rows - lag;
  const int N = tile_width(f, block);
  const bool with_u = St != nullptr;                 // lag > 0 and the lagged sums are wanted
  const int n_tiles = enum_tiles8(f, block, N, S0 != nullptr, with_u, nullptr);
  if (S0) DCG_CUDA_TRY(cudaMemsetAsync(S0, 0, (size_t)f * f * sizeof(double), st));
  if (St) DCG_CUDA_TRY(cudaMemsetAsync(St, 0, (size_t)f * f * sizeof(double), st));
  if (info) DCG_CUDA_TRY(cudaMemsetAsync(info, 0, sizeof(int), st));
  DCG_CUDA_TRY(cudaMemsetAsync(d_qsum, 0, (size_t)f * 16, st));
  i8_prep_kernel<<<(unsigned)ceil_div(f, 128), 128, 0, st>>>(f, mean, range, xmin, xmax, d_shift, d_mul, d_scale);
  DCG_I8_STEP("prep");
  plan8_kernel<<<1, 32, 0, st>>>(d_tiles, f, block, N, S0 != nullptr, with_u);
  DCG_I8_STEP("plan");

  const int ns = num_stages(N);
  const size_t smem = (size_t)ns * stage_bytes(N) + 1024;
  DCG_CUDA_TRY(ensure_dynamic_smem((const void*)cov_i8_kernel, smem));
  DCG_CUDA_TRY(ensure_dynamic_smem((const void*)i8_quantize_kernel<true>, (size_t)6 * kQF * kQRow * 4));
  const int vec4 = row_vec_width(X, ld) == 4 ? 1 : 0;
  const int n_clusters = kNumSMs / 2;
  // the lagged column sum is needed whenever lag > 0, also when only S0 is asked for
  const bool lag_rows = lag > 0;

  for (int64_t w0 = 0; w0 < M; w0 += L.wf) {
    const int64_t pairs = std::min<int64_t>(L.wf, M - w0);
    // ---- quantise the window's pairs [w0, w0 + pairs): Z from row t, U from rows t and t + lag
    const int64_t n_ft = ceil_div(pairs, kQT);
    const int tpb = (int)std::max<int64_t>(1, std::min<int64_t>(32, n_ft / 64));
    dim3 qgrid((unsigned)ceil_div(f, kQF), (unsigned)ceil_div(n_ft, tpb));
    const size_t qsm = (size_t)(lag_rows ? 6 : 3) * kQF * kQRow * 4;
    const bool timed = g_timing.on && g_timing.n < 64;
    if (timed) cudaEventRecord(g_timing.q0[g_timing.n], st);
    if (lag_rows)
      i8_quantize_kernel<true><<<qgrid, 256, qsm, st>>>(X, w0, pairs, lag, f, ld, d_shift, d_mul, d_planes, L.wpad, tpb,
                                                      d_qsum, d_qsum + f, info, vec4);
    else
      i8_quantize_kernel<false><<<qgrid, 256, qsm, st>>>(X, w0, pairs, 0, f, ld, d_shift, d_mul, d_planes, L.wpad, tpb,
                                                       d_qsum, d_qsum + f, info, vec4);
    DCG_I8_STEP("quantize");
    if (timed) cudaEventRecord(g_timing.q1[g_timing.n], st);
    if (n_tiles == 0) continue;
    // ---- contract
    CUtensorMap ma, mb;
    int rc = make_plane_map(&ma, d_planes, pairs, f, L.wpad, L.n_sets, kHalfM);
    if (!rc) rc = make_plane_map(&mb, d_planes, pairs, f, L.wpad, L.n_sets, N / 2);
    if (rc) return rc;
    int64_t g = ceil_div(pairs * n_tiles, (int64_t)n_clusters * 8);
    g = std::min<int64_t>(std::max<int64_t>(ceil_div(g, kStageFrames) * kStageFrames, kStageFrames), kMaxItemFrames);
    g = std::max<int64_t>(kStageFrames, env_i64("DCG_I8_ITEM_FRAMES", g) / kStageFrames * kStageFrames);
    const int64_t n_ranges = ceil_div(pairs, g);
    Params8 p{d_tiles, n_tiles, n_ranges * n_tiles, g, pairs, N, ns, f, S0, St, d_scale,
              (int)env_i64("DCG_I8_DBG", 0)};
    const int nc = (int)std::min<int64_t>(n_clusters, p.n_items);
    cov_i8_kernel<<<2 * nc, kThreads, smem, st>>>(ma, mb, p);
    DCG_I8_STEP("contract");
    if (timed) cudaEventRecord(g_timing.c1[g_timing.n++], st);
  }
  if (colsum_t || colsum_lag) {
    i8_finish_sums_kernel<<<(unsigned)ceil_div(f, 128), 128, 0, st>>>(f, d_scale, d_qsum, d_qsum + f, lag, colsum_t,
                                                                     colsum_lag);
    DCG_I8_STEP("sums");
  }
  if (St) {
    dim3 fg((unsigned)ceil_div(f, 16), (unsigned)ceil_div(f, 16));
    i8_finish_st_kernel<<<fg, 256, 0, st>>>(X, n_rows, f, ld, lag, block, d_shift, d_mul, d_scale, S0, St);
    DCG_I8_STEP("finish");
  }
  return 0;
}
