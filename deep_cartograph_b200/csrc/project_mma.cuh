// A9/A10 (and hTICA level 1): fused standardise + projection as a streamed tall-skinny GEMM.
//
//   P[t, :] = ((x_t - mean) / range) @ W          (reference cv_calculator.py:918-991, 2363-2364)
//
// The pass is HBM-bound (4*f bytes per frame against 2*f*d FLOP, d <= 16), so the design goal is
// bytes in flight, not FLOP/s:
//   * X is streamed through a 6-stage shared-memory ring by `cp.async.bulk` row-segment copies
//     (2 KB each, completion on mbarriers): ~200 KB in flight per SM independent of the register
//     budget of the arithmetic;
//   * the arithmetic runs on the tensor cores as warp-level `mma.sync.m16n8k8` TF32 MMAs in split
//     precision (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, 11-bit pieces -> ~2^-22 per product), which
//     does the reduction over features inside the MMA: no warp shuffles, 8 accumulators per thread
//     instead of rows x d, ~0.2 warp instructions per matrix element;
//   * (x - mean) is an exact FP32 subtraction as in the reference; 1/range is folded into the
//     weights in FP64 before they are split (the B fragments live in registers for the lifetime of
//     a work item);
//   * the feature axis is cut into ranges of <= 1024 columns.  A work item is (range, block of
//     rows); the 8 consumer warps of a CTA each own 64 of the 512 features of a stage, their
//     partial 16 x 16 results are summed through shared memory in a fixed order (deterministic).
//     With one range (f <= 1024) P is written directly with per-CTA min/max; with several ranges
//     of a dense W each range writes a partial P that `pm_combine_kernel` sums in range order;
//     in block-diagonal mode (hTICA level 1) each range IS a diagonal block and writes its own
//     output columns, so all level-1 projections are one pass over X.
#pragma once
#include "dcg_common.cuh"
#include "tc_common.cuh"

namespace dcg {
namespace pm {

constexpr int kWarps = 8;                       // consumer warps
constexpr int kProducerWarps = 4;                // one warpgroup (setmaxnreg is per warpgroup)
constexpr int kThreads = (kWarps + kProducerWarps) * 32;
constexpr int kRows = 16;                       // rows per tile (MMA M)
constexpr int kSub = 512;                       // features per stage (64 per consumer warp)
constexpr int kPitch = kSub * 4 + 64;           // bytes; = 64 mod 128 -> conflict-free fragment reads
constexpr int kStageBytes = kRows * kPitch;
constexpr int kStages = 6;
constexpr int kRangeMax = 2 * kSub;             // widest copy window of a range
constexpr int kRedBytes = kWarps * kRows * 16 * 4;
constexpr int kSmemBytes = kStages * kStageBytes + 2 * kRedBytes + 2 * kStages * 8 + 16;

struct Geom {
  int f;        // valid features
  int rw;       // range width (dense: multiple of 16; blocks: the block width)
  int nr;       // number of ranges
  int bs;       // block-diagonal mode: output columns per block (0 = dense)
  int wld;      // row stride of W (floats)
  int wcol0;    // first column of W used (dense pass of <= 16 columns)
  int ncols;    // columns of this pass (dense)
};

__host__ __device__ __forceinline__ void range_of(const Geom& g, int c, int& vbeg, int& vend, int& nc) {
  vbeg = c * g.rw;
  vend = min(g.f, vbeg + g.rw);
  nc = g.bs ? min(g.bs, vend - vbeg) : g.ncols;
}

// Fragment-ordered operands: for (range c, warp w, group sq = 0..7, lane) 16 floats of B
// ([kstep][ntile][hi|lo][b0|b1]) and the 4 means of the lane's 4 features.
// Lane (g = lane / 4, t = lane % 4) of warp w owns, in group sq, features
//   cstart + 512 * (sq / 4) + 64 * w + 16 * (sq % 4) + 4 * t + e,  e = 0..3;
// MMA k-step ks contracts e = 2 ks (k index t) and e = 2 ks + 1 (k index t + 4).
__global__ void prepare_kernel(Geom g, const float* __restrict__ W, const float* __restrict__ mean,
                               const float* __restrict__ range, float* __restrict__ Bf, float* __restrict__ Mf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nr * kWarps * 8 * 32) return;
  const int lane = i & 31, sq = (i >> 5) & 7, w = (i >> 8) & 7, c = i >> 11;
  int vbeg, vend, nc;
  range_of(g, c, vbeg, vend, nc);
  const int cstart = vbeg & ~3;
  const int t = lane & 3, gid = lane >> 2;
  const int fb = cstart + (sq >> 2) * kSub + w * 64 + (sq & 3) * 16 + 4 * t;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int j = fb + e;
    Mf[(size_t)i * 4 + e] = (mean && j >= vbeg && j < vend) ? mean[j] : 0.f;
  }
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int j = fb + 2 * ks + b, col = 8 * nt + gid;
        double wv = 0.0;
        if (j >= vbeg && j < vend && col < nc) {
          wv = (double)W[(size_t)j * g.wld + g.wcol0 + col];
          if (range) wv /= (double)range[j];
        }
        const float hi = __uint_as_float(tc::cvt_rna_tf32((float)wv));
        const float lo = (float)(wv - (double)hi);      // the tensor core truncates it to 11 bits
        float* dst = Bf + (size_t)i * 16 + ((ks * 2 + nt) * 2) * 2 + b;
        dst[0] = hi;
        dst[2] = lo;
      }
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}"
               ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kWarps * 32) : "memory"); }

// NT = 8-column output tiles per pass (1: <= 8 columns, 2: <= 16).
// out element (range c, row r, column j) lives at out + c * out_cs + r * out_rs + j.
// part_min / part_max (mm_ld > 0): per-CTA column min / max at [blockIdx.x * mm_ld + j].
//
// Three warpgroups: two of consumers (8 warps) and one of producers.  The B fragments of a
// consumer thread (its 64 features x 16 columns, hi and lo) are 128 registers alone, and 12 warps
// = 3 per SM sub-partition allow only 168 each, so the producers hand their registers over with
// `setmaxnreg` (40 for them, 232 for the consumers: 2 * 232 + 40 = 504 <= 512 per lane).
template <int NT>
__global__ void __launch_bounds__(kThreads, 1)
project_mma_kernel(const float* __restrict__ X, int64_t n, int64_t ld, Geom g, int rows_per_item,
                   const float4* __restrict__ Bf, const float4* __restrict__ Mf,
                   float* __restrict__ out, int64_t out_rs, int64_t out_cs,
                   float* __restrict__ part_min, float* __restrict__ part_max, int mm_ld) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t stage0 = tc::smem_u32(smem);
  float* red = reinterpret_cast<float*>(smem + kStages * kStageBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + 2 * kRedBytes);
  uint64_t* empty = full + kStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&full[s], kProducerWarps); tc::mbar_init(&empty[s], kWarps); }
    tc::fence_barrier_init();
  }
  __syncthreads();

  const int64_t nsb = (n + rows_per_item - 1) / rows_per_item;
  const int64_t nitems = nsb * g.nr;

  if (warp >= kWarps) {
    // ---------------- producers: warp p copies rows 4p .. 4p+3 of every stage ----------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    const int pw = warp - kWarps;
    constexpr int kRowsPerProducer = kRows / kProducerWarps;
    uint32_t k = 0;
    for (int64_t it = blockIdx.x; it < nitems; it += gridDim.x) {
      const int c = (int)(it % g.nr);
      const int64_t r_beg = (it / g.nr) * rows_per_item, r_end = min(n, r_beg + rows_per_item);
      int vbeg, vend, nc;
      range_of(g, c, vbeg, vend, nc);
      const int cstart = vbeg & ~3;
      const int cend = (int)min(ld, (int64_t)((vend + 3) & ~3));
      const int nsub = (cend - cstart + kSub - 1) / kSub;
      for (int64_t row0 = r_beg; row0 < r_end; row0 += kRows) {
        const int64_t myrow0 = row0 + pw * kRowsPerProducer;
        const int nrows = (int)max((int64_t)0, min((int64_t)kRowsPerProducer, r_end - myrow0));
        for (int s = 0; s < nsub; ++s, ++k) {
          const uint32_t st = k % kStages, ph = (k / kStages) & 1;
          tc::mbar_wait(&empty[st], ph ^ 1);
          const int c0 = cstart + s * kSub;
          const uint32_t bytes = (uint32_t)min(kSub, cend - c0) * 4u;
          if (lane == 0) mbar_expect_tx(&full[st], bytes * nrows);
          __syncwarp();
          if (lane < nrows)
            bulk_g2s(stage0 + st * kStageBytes + (pw * kRowsPerProducer + lane) * kPitch,
                     X + (myrow0 + lane) * ld + c0, bytes, &full[st]);
        }
      }
    }
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");

  // ---------------- consumers ----------------
  const int t = lane & 3, gid = lane >> 2;
  const int tid = threadIdx.x;                       // 0..255
  const int orow = tid >> 4, ocol = tid & 15;        // the output element this thread finishes
  float mn = INFINITY, mx = -INFINITY;
  uint32_t k = 0, tile = 0;
  uint32_t bfr[8][NT * 8];                           // [group][ks][nt][hi|lo][b0|b1] (nt < NT)
  float mfr[8][4];
  int cur_c = -1;
  uint32_t fixmask = 0;
  int nvalid0 = 0;                                   // vend - (first feature of group 0)

  for (int64_t it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int c = (int)(it % g.nr);
    const int64_t r_beg = (it / g.nr) * rows_per_item, r_end = min(n, r_beg + rows_per_item);
    int vbeg, vend, nc;
    range_of(g, c, vbeg, vend, nc);
    const int cstart = vbeg & ~3;
    const int cend = (int)min(ld, (int64_t)((vend + 3) & ~3));
    const int nsub = (cend - cstart + kSub - 1) / kSub;
    if (c != cur_c) {
      cur_c = c;
      const size_t base = ((size_t)(c * kWarps + warp) * 8) * 32 + lane;
      fixmask = 0;
      nvalid0 = vend - (cstart + warp * 64 + 4 * t);
#pragma unroll
      for (int sq = 0; sq < 8; ++sq) {
        const float4* bp = Bf + (base + (size_t)sq * 32) * 4;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const float4 v = __ldg(bp + ks * 2 + nt);
            bfr[sq][(ks * NT + nt) * 4 + 0] = __float_as_uint(v.x);
            bfr[sq][(ks * NT + nt) * 4 + 1] = __float_as_uint(v.y);
            bfr[sq][(ks * NT + nt) * 4 + 2] = __float_as_uint(v.z);
            bfr[sq][(ks * NT + nt) * 4 + 3] = __float_as_uint(v.w);
          }
        const float4 m = __ldg(Mf + base + (size_t)sq * 32);
        mfr[sq][0] = m.x; mfr[sq][1] = m.y; mfr[sq][2] = m.z; mfr[sq][3] = m.w;
        // elements at or beyond vend are padding / not copied: they must read as exact zeros
        if (nvalid0 - ((sq >> 2) * kSub + (sq & 3) * 16) < 4) fixmask |= 1u << sq;
      }
    }

    for (int64_t row0 = r_beg; row0 < r_end; row0 += kRows, ++tile) {
      // separate chains for the main (hi*hi) and the correction (lo*hi + hi*lo) products: shorter
      // dependent MMA chains, and the small terms are summed among themselves
      float acc[NT][4], cor[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        cor[nt][0] = cor[nt][1] = cor[nt][2] = cor[nt][3] = 0.f;
      }

#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (s < nsub) {
          const uint32_t st = k % kStages, ph = (k / kStages) & 1;
          tc::mbar_wait(&full[st], ph);
          const uint32_t xa = stage0 + st * kStageBytes + gid * kPitch + (warp * 64 + 4 * t) * 4;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int sq = s * 4 + q;
            float4 va = lds128(xa + q * 64);
            float4 vb = lds128(xa + 8 * kPitch + q * 64);
            float a8[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
            if (fixmask & (1u << sq)) {
              const int nv = nvalid0 - (s * kSub + q * 16);
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (e >= nv) { a8[e] = 0.f; a8[4 + e] = 0.f; }
            }
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float z = a8[e] - mfr[sq][e & 3];
              hi[e] = tc::cvt_rna_tf32(z);
              lo[e] = __float_as_uint(z - __uint_as_float(hi[e]));
            }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              // a0 (row g, k t), a1 (row g+8, k t), a2 (row g, k t+4), a3 (row g+8, k t+4)
              const uint32_t ah[4] = {hi[2 * ks], hi[4 + 2 * ks], hi[2 * ks + 1], hi[4 + 2 * ks + 1]};
              const uint32_t al[4] = {lo[2 * ks], lo[4 + 2 * ks], lo[2 * ks + 1], lo[4 + 2 * ks + 1]};
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) {
                const uint32_t* b = &bfr[sq][(ks * NT + nt) * 4];   // hi b0, hi b1, lo b0, lo b1
                mma_tf32(acc[nt], ah, b[0], b[1]);
                mma_tf32(cor[nt], al, b[0], b[1]);
                mma_tf32(cor[nt], ah, b[2], b[3]);
              }
            }
          }
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&empty[st]);
          ++k;
        }
      }

      // fixed-order sum of the 8 warps' partial 16 x 16 tiles
      float* buf = red + (tile & 1) * (kRedBytes / 4);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        *reinterpret_cast<float2*>(&buf[(warp * kRows + gid) * 16 + 8 * nt + 2 * t]) =
            make_float2(acc[nt][0] + cor[nt][0], acc[nt][1] + cor[nt][1]);
        *reinterpret_cast<float2*>(&buf[(warp * kRows + gid + 8) * 16 + 8 * nt + 2 * t]) =
            make_float2(acc[nt][2] + cor[nt][2], acc[nt][3] + cor[nt][3]);
      }
      consumer_sync();
      if (ocol < NT * 8) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) v += buf[w * kRows * 16 + tid];
        if (row0 + orow < r_end && ocol < nc) {
          out[(int64_t)c * out_cs + (row0 + orow) * out_rs + ocol] = v;
          mn = fminf(mn, v);
          mx = fmaxf(mx, v);
        }
      }
    }
  }

  if (mm_ld > 0) {
    // per-CTA column min / max (CTAs without work publish +-inf)
    consumer_sync();
    float* smn = red;
    float* smx = red + 256;
    smn[tid] = mn;
    smx[tid] = mx;
    consumer_sync();
    if (tid < g.ncols) {
      float a = INFINITY, b = -INFINITY;
#pragma unroll
      for (int r = 0; r < 16; ++r) { a = fminf(a, smn[r * 16 + tid]); b = fmaxf(b, smx[r * 16 + tid]); }
      part_min[(size_t)blockIdx.x * mm_ld + tid] = a;
      part_max[(size_t)blockIdx.x * mm_ld + tid] = b;
    }
  }
}

// Dense W over several ranges: P[r, j] = sum_c part[c][r][j] in range order, with per-block
// column min / max partials.  One thread per row.
constexpr int kCombineThreads = 256;
__global__ void __launch_bounds__(kCombineThreads)
combine_kernel(const float* __restrict__ part, int nr, int64_t nb, int nc, int64_t part_cs,
               float* __restrict__ P, int64_t prs, float* __restrict__ part_min, float* __restrict__ part_max, int mm_ld) {
  float mn[16], mx[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { mn[j] = INFINITY; mx[j] = -INFINITY; }
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nb; r += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < nc) {
        float v = 0.f;
        for (int c = 0; c < nr; ++c) v += part[(int64_t)c * part_cs + r * nc + j];
        P[r * prs + j] = v;
        mn[j] = fminf(mn[j], v);
        mx[j] = fmaxf(mx[j], v);
      }
    }
  }
  if (mm_ld <= 0) return;
  __shared__ float sa[kCombineThreads / 32][16], sb[kCombineThreads / 32][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float a = mn[j], b = mx[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if (lane == 0) { sa[warp][j] = a; sb[warp][j] = b; }
  }
  __syncthreads();
  if (threadIdx.x < nc) {
    float a = INFINITY, b = -INFINITY;
#pragma unroll
    for (int w = 0; w < kCombineThreads / 32; ++w) { a = fminf(a, sa[w][threadIdx.x]); b = fmaxf(b, sb[w][threadIdx.x]); }
    part_min[(size_t)blockIdx.x * mm_ld + threadIdx.x] = a;
    part_max[(size_t)blockIdx.x * mm_ld + threadIdx.x] = b;
  }
}

}  // namespace pm
}  // namespace dcg
