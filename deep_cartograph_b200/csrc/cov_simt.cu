// A5/A6: fused standardise + lagged covariance sums -- C-ABI entry point, column sums, and the
// CUDA-core FP32 engine (DCG_COV_SIMT_F32).
// Replaces mlcolvar create_timelagged_dataset + TICA.compute's correlation sums as called from
// reference cv_calculator.py:2244-2261, 2306-2378 and the Gram matrix of sklearn PCA (:2204-2210).
//
// SIMT engine: classic register-tiled FP32 FMA contraction with exact FP32 products; each CTA owns
// a 64x64 tile of BOTH S0 and S_tau over a frame range, accumulates kSimtFlush frames in FP32
// registers and flushes with FP64 atomics.  It is the validation engine for the tcgen05 engine
// (cov_tc.cu) and the reference point for its speed-up; roofline: FP32 FMA pipe.
#include "dcg_common.cuh"
#include "cov_engines.cuh"

namespace dcg {

// ------------------------------------------------------------------------------------------
// column sums of the standardised rows:  a = sum_{t<M} z_t,  b = sum_{t>=lag} z_t
// ------------------------------------------------------------------------------------------
constexpr int kCsRows = 256;  // rows per block step (FP32 partial sums over 256/8 rows per thread)

__global__ void __launch_bounds__(256)
colsum_lag_kernel(const float* __restrict__ X, int64_t n_rows, int f, int64_t ld, int lag,
                  const float* __restrict__ mean, const float* __restrict__ range,
                  double* __restrict__ sum_t, double* __restrict__ sum_lag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 32 + lane;
  const int64_t M = n_rows - lag;
  const int64_t nblocks = (n_rows + kCsRows - 1) / kCsRows;
  double da = 0.0, db = 0.0;
  if (col < f) {
    float m = 0.f, rg = 1.f, ri = 1.f;
    if (mean) { m = mean[col]; rg = range[col]; ri = 1.0f / rg; }
    for (int64_t rb = blockIdx.y; rb < nblocks; rb += gridDim.y) {
      const int64_t r0 = rb * kCsRows, r1 = min(n_rows, r0 + kCsRows);
      float a = 0.f, b = 0.f;
      for (int64_t r = r0 + warp; r < r1; r += 8) {
        const float x = ldg_stream1(X + r * ld + col);
        const float z = mean ? standardize1(x, m, rg, ri) : x;
        if (r < M) a += z;
        if (r >= lag) b += z;
      }
      da += (double)a;
      db += (double)b;
    }
  }
  __shared__ double sa[8][32], sb[8][32];
  sa[warp][lane] = da;
  sb[warp][lane] = db;
  __syncthreads();
  if (warp == 0 && col < f) {
    double ta = 0.0, tb = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { ta += sa[w][lane]; tb += sb[w][lane]; }
    if (sum_t) atomicAdd(sum_t + col, ta);
    if (sum_lag) atomicAdd(sum_lag + col, tb);
  }
}

// ------------------------------------------------------------------------------------------
// SIMT FP32 engine
// ------------------------------------------------------------------------------------------
constexpr int kSimtTile = 64;     // output tile edge
constexpr int kSimtKS = 16;       // frames per smem stage
constexpr int kSimtFlush = 256;   // frames between FP64 flushes
constexpr int kSimtThreads = 256;

__device__ __forceinline__ bool tile_needed(int i0, int j0, int tile, int f, int block, bool s0) {
  if (s0 && j0 + tile - 1 < i0) return false;           // strictly-lower tile of the symmetric S0
  if (block <= 0) return true;
  const int ia = i0 / block, ib = (min(i0 + tile, f) - 1) / block;
  const int ja = j0 / block, jb = (min(j0 + tile, f) - 1) / block;
  return !(ib < ja || jb < ia);                          // row-block range meets column-block range
}

__global__ void __launch_bounds__(kSimtThreads)
cov_simt_kernel(const float* __restrict__ X, int64_t n_rows, int f, int64_t ld, int lag,
                const float* __restrict__ mean, const float* __restrict__ range, int block,
                double* __restrict__ S0, double* __restrict__ St, int nt, int64_t frames_per_cta) {
  const int ti = blockIdx.x / nt, tj = blockIdx.x % nt;
  const int i0 = ti * kSimtTile, j0 = tj * kSimtTile;
  const bool do_s0 = S0 && tile_needed(i0, j0, kSimtTile, f, block, true);
  const bool do_st = St && tile_needed(i0, j0, kSimtTile, f, block, false);
  if (!do_s0 && !do_st) return;
  const int64_t M = n_rows - lag;
  const int64_t t_begin = (int64_t)blockIdx.y * frames_per_cta;
  const int64_t t_end = min(M, t_begin + frames_per_cta);
  if (t_begin >= t_end) return;

  __shared__ __align__(16) float As[kSimtKS][kSimtTile];
  __shared__ __align__(16) float B0s[kSimtKS][kSimtTile];
  __shared__ __align__(16) float Bts[kSimtKS][kSimtTile];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, 4 x 4 outputs each
  // loader role: thread -> (row kr in stage, 4 consecutive columns c4)
  const int kr = tid >> 4, c4 = (tid & 15) * 4;
  float am[4], ar[4], ai[4], bm[4], br[4], bi[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const int ca = i0 + c4 + v, cb = j0 + c4 + v;
    am[v] = (mean && ca < f) ? mean[ca] : 0.f;
    ar[v] = (range && ca < f) ? range[ca] : 1.f;
    bm[v] = (mean && cb < f) ? mean[cb] : 0.f;
    br[v] = (range && cb < f) ? range[cb] : 1.f;
    ai[v] = 1.0f / ar[v];
    bi[v] = 1.0f / br[v];
  }
  const bool stdz = mean != nullptr;

  float acc0[4][4], acct[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) { acc0[a][b] = 0.f; acct[a][b] = 0.f; }

  auto flush = [&]() {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int gi = i0 + ty * 4 + a;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int gj = j0 + tx * 4 + b;
        if (gi < f && gj < f) {
          if (do_s0) atomicAdd(S0 + (size_t)gi * f + gj, (double)acc0[a][b]);
          if (do_st) atomicAdd(St + (size_t)gi * f + gj, (double)acct[a][b]);
        }
        acc0[a][b] = 0.f;
        acct[a][b] = 0.f;
      }
    }
  };

  float ra[4], rb0[4], rbt[4];
  auto load_stage = [&](int64_t t) {
    const int64_t row = t + kr;
    const bool ok = row < t_end;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int ca = i0 + c4 + v, cb = j0 + c4 + v;
      float xa = 0.f, xb = 0.f, xt = 0.f;
      if (ok && ca < f) { xa = __ldg(X + row * ld + ca); xa = stdz ? standardize1(xa, am[v], ar[v], ai[v]) : xa; }
      if (ok && cb < f) {
        if (do_s0) { xb = __ldg(X + row * ld + cb); xb = stdz ? standardize1(xb, bm[v], br[v], bi[v]) : xb; }
        if (do_st) { xt = __ldg(X + (row + lag) * ld + cb); xt = stdz ? standardize1(xt, bm[v], br[v], bi[v]) : xt; }
      }
      ra[v] = xa; rb0[v] = xb; rbt[v] = xt;
    }
  };

  load_stage(t_begin);
  int since_flush = 0;
  for (int64_t t = t_begin; t < t_end; t += kSimtKS) {
    *reinterpret_cast<float4*>(&As[kr][c4]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    *reinterpret_cast<float4*>(&B0s[kr][c4]) = make_float4(rb0[0], rb0[1], rb0[2], rb0[3]);
    *reinterpret_cast<float4*>(&Bts[kr][c4]) = make_float4(rbt[0], rbt[1], rbt[2], rbt[3]);
    __syncthreads();
    if (t + kSimtKS < t_end) load_stage(t + kSimtKS);
#pragma unroll
    for (int k = 0; k < kSimtKS; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&B0s[k][tx * 4]);
      const float4 t4 = *reinterpret_cast<const float4*>(&Bts[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
      const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          acc0[a][b] = fmaf(av[a], bv[b], acc0[a][b]);
          acct[a][b] = fmaf(av[a], tv[b], acct[a][b]);
        }
    }
    __syncthreads();
    since_flush += kSimtKS;
    if (since_flush >= kSimtFlush) { flush(); since_flush = 0; }
  }
  flush();
}

int cov_simt_launch(const CovArgs& a, cudaStream_t st) {
  const int nt = (int)ceil_div(a.f, kSimtTile);
  const int64_t M = a.n_rows - a.lag;
  // enough frame splits for ~16 CTAs per SM over the needed tiles
  int64_t splits = std::max<int64_t>(1, (int64_t)kNumSMs * 16 / std::max(1, nt * nt / 2));
  int64_t fpc = ceil_div(M, splits);
  fpc = std::max<int64_t>(kSimtFlush, ceil_div(fpc, kSimtFlush) * kSimtFlush);
  splits = ceil_div(M, fpc);
  if (splits > 65535) { fpc = ceil_div(ceil_div(M, 65535), kSimtFlush) * kSimtFlush; splits = ceil_div(M, fpc); }
  dim3 grid((unsigned)(nt * nt), (unsigned)splits);
  cov_simt_kernel<<<grid, kSimtThreads, 0, st>>>(a.X, a.n_rows, a.f, a.ld, a.lag, a.mean, a.range,
                                                 a.block, a.S0, a.St, nt, fpc);
  DCG_LAUNCH_CHECK();
  return 0;
}

}  // namespace dcg

using namespace dcg;

extern "C" size_t dcg_cov_workspace_bytes(int64_t n_rows, int f, int lag, int block, int engine) {
  if (n_rows <= 0 || f <= 0) return 0;
  if (engine == DCG_COV_SIMT_F32) return 256;
  return cov_tc_workspace_bytes(n_rows, f, lag, block, engine);
}

extern "C" int dcg_cov_lag_f32(const float* X, int64_t n_rows, int f, int64_t ld, int lag,
                               const float* mean, const float* range, int block,
                               double* S0, double* St, double* colsum_t, double* colsum_lag,
                               int engine, void* ws, size_t ws_bytes, void* stream) {
  if (!X || (!S0 && !St)) return DCG_E_NULL;
  if ((mean == nullptr) != (range == nullptr)) return DCG_E_NULL;
  if (n_rows <= 0 || f <= 0 || ld < f || lag < 0 || lag >= n_rows || block < 0) return DCG_E_SHAPE;
  if (engine != DCG_COV_SIMT_F32 && engine != DCG_COV_TC_3XTF32 && engine != DCG_COV_TC_1XTF32 &&
      engine != DCG_COV_TC_3XF16)
    return DCG_E_MODE;
  if (!ws || ws_bytes < dcg_cov_workspace_bytes(n_rows, f, lag, block, engine)) return DCG_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  if (lag == 0) St = nullptr;   // S_tau == S0 when lag == 0: not computed
  CovArgs a{X, n_rows, f, ld, lag, mean, range, block, S0, St, colsum_t, colsum_lag, engine, ws, ws_bytes};

  if (colsum_t) DCG_CUDA_TRY(cudaMemsetAsync(colsum_t, 0, (size_t)f * sizeof(double), st));
  if (colsum_lag) DCG_CUDA_TRY(cudaMemsetAsync(colsum_lag, 0, (size_t)f * sizeof(double), st));
  // the tcgen05 engine accumulates the column sums inside its diagonal tiles (no extra pass over X)
  if ((colsum_t || colsum_lag) && (engine == DCG_COV_SIMT_F32 || !cov_tc_fuses_colsums(a))) {
    const int64_t gy = std::min<int64_t>(ceil_div(n_rows, kCsRows), 16384);
    colsum_lag_kernel<<<dim3((unsigned)ceil_div(f, 32), (unsigned)gy), 256, 0, st>>>(
        X, n_rows, f, ld, lag, mean, range, colsum_t, colsum_lag);
    DCG_LAUNCH_CHECK();
  }
  if (S0) DCG_CUDA_TRY(cudaMemsetAsync(S0, 0, (size_t)f * f * sizeof(double), st));
  if (St) DCG_CUDA_TRY(cudaMemsetAsync(St, 0, (size_t)f * f * sizeof(double), st));
  if (engine == DCG_COV_SIMT_F32) return cov_simt_launch(a, st);
  return cov_tc_launch(a, st);
}
