// A11: DeepTICA minibatch correlation sums (weighted), FP64 outputs.
// Replaces the correlation sums inside mlcolvar DeepTICA.training_step / TICA.compute as driven
// from reference cv_calculator.py:1508-1524.  B x d network outputs (d <= 32): a memory/latency-
// bound reduction, not a tensor-core contraction.
#include "dcg_common.cuh"

namespace dcg {

// out layout (doubles): [0]=sum w, [1]=sum wl, [2..2+d)=sum w f, then d*d sum w f f^T,
// d*d sum wl f g^T, d sum wl f, d sum wl g.
// One warp handles a strip of rows; lane pair (i, j) ownership: lane l accumulates entries
// e = l, l+32, ... of the d*d matrices in FP32 over <= 64 rows, then FP64.
constexpr int kTcThreads = 256;
constexpr int kTcRowsPerCta = 512;

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
  const int n_out = 2 + d + 2 * d * d + 2 * d;
  for (int e = threadIdx.x; e < n_out; e += kTcThreads) {
    double acc = 0.0;
    // decode entry
    int kind, i = 0, j = 0;
    if (e == 0) kind = 0;
    else if (e == 1) kind = 1;
    else if (e < 2 + d) { kind = 2; i = e - 2; }
    else if (e < 2 + d + d * d) { kind = 3; i = (e - 2 - d) / d; j = (e - 2 - d) % d; }
    else if (e < 2 + d + 2 * d * d) { kind = 4; i = (e - 2 - d - d * d) / d; j = (e - 2 - d - d * d) % d; }
    else if (e < 2 + 2 * d + 2 * d * d) { kind = 5; i = e - (2 + d + 2 * d * d); }
    else { kind = 6; i = e - (2 + 2 * d + 2 * d * d); }
    for (int rb = 0; rb < rows; rb += 64) {
      float p = 0.f;
      const int re = min(rows, rb + 64);
      for (int r = rb; r < re; ++r) {
        const float* row = sm + r * stride;
        const float ww = row[2 * d], wwl = row[2 * d + 1];
        float v;
        switch (kind) {
          case 0: v = ww; break;
          case 1: v = wwl; break;
          case 2: v = ww * row[i]; break;
          case 3: v = ww * row[i] * row[j]; break;
          case 4: v = wwl * row[i] * row[d + j]; break;
          case 5: v = wwl * row[i]; break;
          default: v = wwl * row[d + i]; break;
        }
        p += v;
      }
      acc += (double)p;
    }
    atomicAdd(out + e, acc);
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
  DCG_CUDA_TRY(cudaFuncSetAttribute(ticacov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ticacov_kernel<<<(unsigned)ceil_div(B, kTcRowsPerCta), kTcThreads, smem, st>>>(f, g, w, wl, B, d, out);
  DCG_LAUNCH_CHECK();
  return 0;
}
