// A11: DeepTICA minibatch correlation sums (weighted), FP64 outputs.
// Replaces the correlation sums inside mlcolvar DeepTICA.training_step / TICA.compute as driven
// from reference cv_calculator.py:1508-1524.  B x d network outputs (d <= 32): a memory/latency-
// bound reduction, not a tensor-core contraction.
#include "dcg_common.cuh"

namespace dcg {

// out layout (doubles): [0]=sum w, [1]=sum wl, [2..2+d)=sum w f, then d*d sum w f f^T,
// d*d sum wl f g^T, d sum wl f, d sum wl g.
// A CTA stages kTcRowsPerCta rows (f row, g row, w, wl) in shared memory; thread e < d*d owns the
// matrix entry (i, j) = (e / d, e % d) of BOTH matrices, thread d*d + i the three vector entries of
// component i, thread d*d + d the two weight sums -- uniform code inside each class, rows unrolled
// by 4, FP32 within the CTA's rows, FP64 across CTAs (one atomic per entry and CTA).
constexpr int kTcThreads = 256;
constexpr int kTcRowsPerCta = 128;

__global__ void __launch_bounds__(kTcThreads)
ticacov_kernel(const float* __restrict__ f, const float* __restrict__ g, const float* __restrict__ w,
               const float* __restrict__ wl, int64_t B, int d, double* __restrict__ out) {
  extern __shared__ float sm[];              // [rows][2*d + 2] : f row, g row, w, wl
  const int stride = 2 * d + 2;
  const int64_t r0 = (int64_t)blockIdx.x * kTcRowsPerCta;
  const int rows = (int)min((int64_t)kTcRowsPerCta, B - r0);
  for (int i = threadIdx.x; i < rows * d; i += kTcThreads) {
    const int r = i / d, q = i - r * d;
    sm[r * stride + q] = f[(r0 + r) * d + q];
    sm[r * stride + d + q] = g[(r0 + r) * d + q];
  }
  for (int r = threadIdx.x; r < rows; r += kTcThreads) {
    sm[r * stride + 2 * d] = w ? w[r0 + r] : 1.f;
    sm[r * stride + 2 * d + 1] = wl ? wl[r0 + r] : 1.f;
  }
  __syncthreads();
  const int dd = d * d;
  double* o_swf = out + 2;
  double* o_sff = out + 2 + d;
  double* o_sfg = o_sff + dd;
  double* o_slf = o_sfg + dd;
  double* o_slg = o_slf + d;
  for (int e = threadIdx.x; e < dd + d + 1; e += kTcThreads) {
    if (e < dd) {
      const int i = e / d, j = e - i * d;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
      for (int r = 0; r < rows; ++r) {
        const float* row = sm + r * stride;
        const float fi = row[i];
        a0 = fmaf(row[2 * d] * fi, row[j], a0);
        a1 = fmaf(row[2 * d + 1] * fi, row[d + j], a1);
      }
      atomicAdd(o_sff + e, (double)a0);
      atomicAdd(o_sfg + e, (double)a1);
    } else if (e < dd + d) {
      const int i = e - dd;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 4
      for (int r = 0; r < rows; ++r) {
        const float* row = sm + r * stride;
        a0 = fmaf(row[2 * d], row[i], a0);
        a1 = fmaf(row[2 * d + 1], row[i], a1);
        a2 = fmaf(row[2 * d + 1], row[d + i], a2);
      }
      atomicAdd(o_swf + i, (double)a0);
      atomicAdd(o_slf + i, (double)a1);
      atomicAdd(o_slg + i, (double)a2);
    } else {
      float a0 = 0.f, a1 = 0.f;
      for (int r = 0; r < rows; ++r) { a0 += sm[r * stride + 2 * d]; a1 += sm[r * stride + 2 * d + 1]; }
      atomicAdd(out + 0, (double)a0);
      atomicAdd(out + 1, (double)a1);
    }
  }
}

}  // namespace dcg

using namespace dcg;

extern "C" size_t dcg_ticacov_out_doubles(int d) {
  if (d < 1 || d > 32) return 0;
  return (size_t)(2 + d + 2 * d * d + 2 * d);
}

extern "C" int dcg_ticacov_f32(const float* f, const float* g, const float* w, const float* wl,
                               int64_t B, int d, double* out, void* stream) {
  if (!f || !g || !out) return DCG_E_NULL;
  if (B <= 0 || d < 1 || d > 32) return DCG_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  DCG_CUDA_TRY(cudaMemsetAsync(out, 0, dcg_ticacov_out_doubles(d) * sizeof(double), st));
  const size_t smem = (size_t)kTcRowsPerCta * (2 * d + 2) * sizeof(float);
  DCG_CUDA_TRY(ensure_dynamic_smem((const void*)ticacov_kernel, (size_t)(smem)));
  ticacov_kernel<<<(unsigned)ceil_div(B, kTcRowsPerCta), kTcThreads, smem, st>>>(f, g, w, wl, B, d, out);
  DCG_LAUNCH_CHECK();
  return 0;
}
