// N4: free-energy surface of the projected frames by binned kernel density estimation.
// Replaces mlcolvar.utils.fes.compute_fes(backend="KDEpy") as called from the reference's
// modules/figures/figures.py:95 (driven from tools/train_colvars/train_colvars_workflow.py:146-182:
// one 1-D FES per CV component with 100 blocks, one 2-D FES per CV pair with 1 block).
//
// KDEpy's FFTKDE = linear binning of the samples onto the evaluation grid, then a convolution of the
// bin weights with the kernel sampled at the grid offsets.  Two kernels:
//   fes_bin_kernel     the pass over the N x d projection (HBM-bound, 4 d bytes per frame): every frame
//                      adds (1-a)(1-b), a(1-b), (1-a)b, ab to the four grid nodes around it (two in 1-D).
//                      Frames are split into `blocks` contiguous blocks (numpy.array_split) with one
//                      grid per block; a CTA owns a contiguous frame range and keeps the grid of its
//                      current block in shared memory (FP32 native atomics), flushed with FP64 global
//                      atomics at block boundaries.
//   fes_smooth_kernel  separable Gaussian convolution of the small grids in FP64 (density on the grid).
// The logarithm, the block average and its error are O(blocks x grid) and stay in the caller.
#include "dcg_common.cuh"

namespace dcg {
namespace {

constexpr int kBinThreads = 512;

__device__ __forceinline__ int64_t block_of_frame(int64_t t, int64_t q, int64_t r) {
  // numpy.array_split: the first r blocks have q + 1 frames, the others q
  const int64_t head = r * (q + 1);
  return t < head ? t / (q + 1) : r + (t - head) / (q > 0 ? q : 1);
}
__device__ __forceinline__ int64_t block_end(int64_t b, int64_t q, int64_t r) {
  return b < r ? (b + 1) * (q + 1) : r * (q + 1) + (b + 1 - r) * q;
}

template <bool TWO_D, bool SMEM>
__global__ void __launch_bounds__(kBinThreads)
fes_bin_kernel(const float* __restrict__ P, int64_t n, int64_t ld, int c0, int c1,
               double lo0, double inv0, double lo1, double inv1, int G, int blocks,
               double* __restrict__ hist, unsigned long long* __restrict__ n_outside, int64_t frames_per_cta) {
  extern __shared__ float sh[];
  const int cells = TWO_D ? G * G : G;
  const int64_t q = n / blocks, r = n % blocks;
  const int64_t t_begin = (int64_t)blockIdx.x * frames_per_cta;
  const int64_t t_end = min(n, t_begin + frames_per_cta);
  if (t_begin >= t_end) return;
  unsigned long long outside = 0;
  int64_t t0 = t_begin;
  while (t0 < t_end) {
    const int64_t b = block_of_frame(t0, q, r);
    const int64_t t1 = min(t_end, block_end(b, q, r));
    double* hb = hist + (size_t)b * cells;
    if (SMEM) {
      for (int i = threadIdx.x; i < cells; i += kBinThreads) sh[i] = 0.f;
      __syncthreads();
    }
    for (int64_t t = t0 + threadIdx.x; t < t1; t += kBinThreads) {
      const float* row = P + (size_t)t * ld;
      const double fx = ((double)__ldg(row + c0) - lo0) * inv0;       // grid coordinate in [0, G-1]
      double fy = 0.0;
      if (TWO_D) fy = ((double)__ldg(row + c1) - lo1) * inv1;
      if (!(fx >= 0.0 && fx <= (double)(G - 1) && fy >= 0.0 && fy <= (double)(G - 1))) { ++outside; continue; }
      int ix = min((int)fx, G - 2 >= 0 ? G - 2 : 0);
      const float ax = (float)(fx - ix);
      if (TWO_D) {
        int iy = min((int)fy, G - 2 >= 0 ? G - 2 : 0);
        const float ay = (float)(fy - iy);
        const int o = iy * G + ix;
        if (SMEM) {
          atomicAdd(&sh[o], (1.f - ax) * (1.f - ay));
          atomicAdd(&sh[o + 1], ax * (1.f - ay));
          atomicAdd(&sh[o + G], (1.f - ax) * ay);
          atomicAdd(&sh[o + G + 1], ax * ay);
        } else {
          atomicAdd(&hb[o], (double)((1.f - ax) * (1.f - ay)));
          atomicAdd(&hb[o + 1], (double)(ax * (1.f - ay)));
          atomicAdd(&hb[o + G], (double)((1.f - ax) * ay));
          atomicAdd(&hb[o + G + 1], (double)(ax * ay));
        }
      } else {
        if (SMEM) {
          atomicAdd(&sh[ix], 1.f - ax);
          atomicAdd(&sh[ix + 1], ax);
        } else {
          atomicAdd(&hb[ix], (double)(1.f - ax));
          atomicAdd(&hb[ix + 1], (double)ax);
        }
      }
    }
    if (SMEM) {
      __syncthreads();
      for (int i = threadIdx.x; i < cells; i += kBinThreads) {
        const float v = sh[i];
        if (v != 0.f) atomicAdd(&hb[i], (double)v);
      }
      __syncthreads();
    }
    t0 = t1;
  }
  if (outside && n_outside) atomicAdd(n_outside, outside);
}

// density[b][iy][ix] = sum_j w[b][jy][jx] K(x_ix - x_jx) K(y_iy - y_jy) / n_b, K = N(0, h^2) per axis.
// pass 0: along x into tmp; pass 1: along y into out (1-D: pass 0 only, straight into out).
__global__ void fes_smooth_kernel(const double* __restrict__ in, double* __restrict__ out, int blocks, int G, int rows,
                                  int axis, double step, double h, const double* __restrict__ norm /* per block or null */) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_block = (int64_t)rows * G;
  if (idx >= (int64_t)blocks * per_block) return;
  const int b = (int)(idx / per_block);
  const int rem = (int)(idx - (int64_t)b * per_block);
  const int iy = rem / G, ix = rem - iy * G;
  const double* base = in + (size_t)b * per_block;
  const double c = -0.5 / (h * h), pref = 0.3989422804014327 / h;
  double acc = 0.0;
  if (axis == 0) {
    const double* rowp = base + (size_t)iy * G;
    for (int j = 0; j < G; ++j) { const double dlt = (double)(ix - j) * step; acc += rowp[j] * exp(c * dlt * dlt); }
  } else {
    for (int j = 0; j < rows; ++j) { const double dlt = (double)(iy - j) * step; acc += base[(size_t)j * G + ix] * exp(c * dlt * dlt); }
  }
  acc *= pref;
  if (norm) acc /= norm[b];
  out[idx] = acc;
}

}  // namespace
}  // namespace dcg

using namespace dcg;

extern "C" int dcg_fes_bin_f32(const float* P, int64_t n, int64_t ld, int c0, int c1,
                               double lo0, double hi0, double lo1, double hi1, int G, int blocks,
                               double* hist, int64_t* n_outside, void* stream) {
  if (!P || !hist) return DCG_E_NULL;
  if (n <= 0 || ld <= 0 || c0 < 0 || c0 >= ld || c1 >= ld || G < 2 || G > 4096 || blocks < 1 || blocks > n ||
      !(hi0 > lo0) || (c1 >= 0 && !(hi1 > lo1)))
    return DCG_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const bool two_d = c1 >= 0;
  const size_t cells = two_d ? (size_t)G * G : (size_t)G;
  DCG_CUDA_TRY(cudaMemsetAsync(hist, 0, cells * (size_t)blocks * sizeof(double), st));
  if (n_outside) DCG_CUDA_TRY(cudaMemsetAsync(n_outside, 0, sizeof(int64_t), st));
  const double inv0 = (double)(G - 1) / (hi0 - lo0), inv1 = two_d ? (double)(G - 1) / (hi1 - lo1) : 0.0;
  // contiguous frame ranges: ~2 CTAs per SM, at least 4096 frames each
  int64_t fpc = std::max<int64_t>(4096, ceil_div(n, (int64_t)2 * kNumSMs));
  const unsigned grid = (unsigned)ceil_div(n, fpc);
  const size_t smem = cells * sizeof(float);
  const bool use_smem = smem <= 200 * 1024;
  unsigned long long* no = (unsigned long long*)n_outside;
#define DCG_FES_LAUNCH(TD, SM)                                                                         \
  do {                                                                                                 \
    auto kern = fes_bin_kernel<TD, SM>;                                                                \
    if (SM) DCG_CUDA_TRY(ensure_dynamic_smem((const void*)kern, smem));                                \
    kern<<<grid, kBinThreads, SM ? smem : 0, st>>>(P, n, ld, c0, two_d ? c1 : 0, lo0, inv0, lo1, inv1, G, blocks, \
                                                   hist, no, fpc);                                     \
  } while (0)
  if (two_d) { if (use_smem) DCG_FES_LAUNCH(true, true); else DCG_FES_LAUNCH(true, false); }
  else { if (use_smem) DCG_FES_LAUNCH(false, true); else DCG_FES_LAUNCH(false, false); }
#undef DCG_FES_LAUNCH
  DCG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dcg_fes_smooth_f64(const double* hist, int blocks, int G, int dim, double step0, double step1,
                                  double bandwidth, const double* block_frames, double* density, double* tmp,
                                  void* stream) {
  if (!hist || !density || (dim == 2 && !tmp)) return DCG_E_NULL;
  if (blocks < 1 || G < 2 || (dim != 1 && dim != 2) || !(bandwidth > 0.0) || !(step0 > 0.0) || (dim == 2 && !(step1 > 0.0)))
    return DCG_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = dim == 2 ? G : 1;
  const int64_t total = (int64_t)blocks * rows * G;
  const unsigned grid = (unsigned)ceil_div(total, 256);
  if (dim == 1) {
    fes_smooth_kernel<<<grid, 256, 0, st>>>(hist, density, blocks, G, 1, 0, step0, bandwidth, block_frames);
  } else {
    fes_smooth_kernel<<<grid, 256, 0, st>>>(hist, tmp, blocks, G, rows, 0, step0, bandwidth, nullptr);
    fes_smooth_kernel<<<grid, 256, 0, st>>>(tmp, density, blocks, G, rows, 1, step1, bandwidth, block_frames);
  }
  DCG_LAUNCH_CHECK();
  return 0;
}
