// Internal interface between the dcg_cov_lag_f32 entry point and its contraction engines.
#pragma once
#include "dcg_common.cuh"

namespace dcg {

struct CovArgs {
  const float* X;
  int64_t n_rows;
  int f;
  int64_t ld;
  int lag;
  const float* mean;    // may be null (no standardisation)
  const float* range;
  int block;            // 0 = dense, >0 = diagonal blocks of this width
  double* S0;           // may be null
  double* St;           // may be null
  double* colsum_t;     // may be null; sum_{t<M} z_t   (tcgen05 engine: fused into the diagonal tiles)
  double* colsum_lag;   // may be null; sum_{t>=lag} z_t
  int engine;
  void* ws;
  size_t ws_bytes;
};

// S0 / St are zero-filled by the caller; engines accumulate into them.
int cov_simt_launch(const CovArgs& a, cudaStream_t st);
int cov_tc_launch(const CovArgs& a, cudaStream_t st);
bool cov_tc_fuses_colsums(const CovArgs& a);
size_t cov_tc_workspace_bytes(int64_t n_rows, int f, int lag, int block, int engine);

}  // namespace dcg
