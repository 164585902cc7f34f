// Standalone tcgen05 probe (diagnostics only, not part of the product library).
// Answers, on a real B200, the hardware questions the covariance engine's design rests on:
//   1. kind::tf32 with MN-major, SWIZZLE_NONE shared-memory operands: layout + descriptor semantics,
//      including a descriptor start address shifted by an arbitrary number of K rows (lag shift);
//   2. A operand from TMEM (TS form) written with tcgen05.st.32x32b;
//   3. issue rate of M=128 x N={128,256} x K=8 MMAs in SS and TS form;
//   4. rounding behaviour of the FP32 accumulation over long K (truncation vs round-to-nearest);
//   5. what the tensor core does with the low 13 mantissa bits of an fp32 operand.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc_probe tc_probe.cu -lcuda
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "../deep_cartograph_b200/csrc/tc_common.cuh"

using namespace dcg::tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int KROWS = 64;       // K rows staged in smem (8 K-steps)
constexpr int PADROWS = 40;     // extra rows so a shifted window stays inside the buffer

struct Params {
  const float* A;   // [KROWS][128]
  const float* B;   // [KROWS + PADROWS][N]
  float* D;         // [128][N]
  long long* cycles;
  int N;
  int mode;         // 0 = SS, 1 = TS
  int shift;        // B window starts at row `shift`
  int reps;         // the 8 K-steps are issued `reps` times (accumulating)
  int a_mn;         // SS only: 1 = A stored MN-major, 0 = K-major
  int b_mn;         // 1 = B stored MN-major, 0 = K-major
  int roundtrip;    // 1 = no MMA: tcgen05.st a pattern into D, read it back
  int a_off;        // TS: extra TMEM column offset of the A operand (alignment probe)
};

__global__ void __launch_bounds__(128, 1) probe_kernel(Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N;
  const int RB = KROWS + PADROWS;                 // rows per MN-group in the B buffer
  float* As = reinterpret_cast<float*>(smem);                       // [32 groups][KROWS][4]
  float* Bs = reinterpret_cast<float*>(smem + 32 * KROWS * 16);     // [N/4 groups][RB][4]

  // stage operands in the MN-major SWIZZLE_NONE canonical layout
  // MN-major: [mn/4][k][mn%4]          (core matrix = 4 mn x 8 k, 128 B; SBO = mn-group stride)
  // K-major : [k/4][mn][k%4]           (core matrix = 8 mn x 4 k, 128 B; LBO = k-group stride, SBO = 8-row stride = 128)
  for (int i = tid; i < KROWS * 128; i += 128) {
    const int k = i / 128, m = i % 128;
    if (p.a_mn) As[((m >> 2) * KROWS + k) * 4 + (m & 3)] = p.A[i];
    else As[((k >> 2) * 128 + m) * 4 + (k & 3)] = p.A[i];
  }
  for (int i = tid; i < RB * N; i += 128) {
    const int k = i / N, n = i % N;
    if (p.b_mn) Bs[((n >> 2) * RB + k) * 4 + (n & 3)] = p.B[i];
    else Bs[((k >> 2) * N + n) * 4 + (k & 3)] = p.B[i];
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t a_tmem = tmem + 256 + p.a_off;   // TS: A lives at columns [256 + a_off, ... + KROWS)

  if (p.mode == 1) {
    // thread m (lane 32*warp + lane of TMEM) holds A[k][m] for all k
    for (int ks = 0; ks < KROWS / 8; ++ks) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(p.A[(ks * 8 + j) * 128 + tid]);
      tmem_st_x8(a_tmem + ((uint32_t)(warp * 32) << 16) + ks * 8, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (p.roundtrip) {
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __float_as_uint((float)(tid * 1000 + c0 + j));
      tmem_st_x8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  long long t0 = 0, t1 = 0;
  if (tid == 0 && !p.roundtrip) {
    const uint32_t idesc = make_idesc_tf32(128, N, p.mode == 0 ? p.a_mn : 0, p.b_mn);
    const uint32_t a_base = smem_u32(As), b_base = smem_u32(Bs) + p.shift * 16;
    // descriptors precomputed: the timed loop only issues
    uint64_t bd[KROWS / 8], ad[KROWS / 8];
#pragma unroll
    for (int ks = 0; ks < KROWS / 8; ++ks) {
      // MN-major: K step = 8 rows x 16 B = 128 B, SBO = group stride
      // K-major : K step = 2 k-groups; LBO = k-group stride, SBO = 128 (8 rows x 16 B); shift must be a multiple of 4
      bd[ks] = p.b_mn ? make_smem_desc(b_base + ks * 128, 128, RB * 16)
                      : make_smem_desc(smem_u32(Bs) + (ks * 2 + p.shift / 4) * N * 16, N * 16, 128);
      ad[ks] = p.a_mn ? make_smem_desc(a_base + ks * 128, 128, KROWS * 16)
                      : make_smem_desc(a_base + ks * 2 * 128 * 16, 128 * 16, 128);
    }
    t0 = clock64();
    if (p.mode == 0) {
      for (int r = 0; r < p.reps; ++r) {
#pragma unroll
        for (int ks = 0; ks < KROWS / 8; ++ks) mma_tf32_ss(tmem, ad[ks], bd[ks], idesc, (r | ks) ? 1u : 0u);
      }
    } else {
      for (int r = 0; r < p.reps; ++r) {
#pragma unroll
        for (int ks = 0; ks < KROWS / 8; ++ks) mma_tf32_ts(tmem, a_tmem + ks * 8, bd[ks], idesc, (r | ks) ? 1u : 0u);
      }
    }
    mma_commit(&bar);
  }
  if (!p.roundtrip) mbar_wait(&bar, 0);
  if (tid == 0) { t1 = clock64(); p.cycles[0] = t1 - t0; }
  tc_fence_after();
  // epilogue: lane (= row m) x 32 columns at a time
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) p.D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// TMEM <-> register bandwidth: 4 warps (one per SMSP / lane quarter).
//  mode 0: tcgen05.ld x32 only; 1: tcgen05.st x32 only; 2: the level-2 drain loop (ld L1, ld L2, add, st L2)
__global__ void __launch_bounds__(128, 1) tmem_bw_kernel(int mode, int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s + ((uint32_t)(warp * 32) << 16);
  uint32_t v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __float_as_uint((float)(tid + j));
  for (int c0 = 0; c0 < 512; c0 += 32) tmem_st_x32(tmem + c0, v);
  tmem_st_wait();
  __syncthreads();
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 32) { tmem_ld_x32(tmem + c0, v); tmem_ld_wait(); acc += __uint_as_float(v[it & 31]); }
    } else if (mode == 1) {
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 32) tmem_st_x32(tmem + c0, v);
      tmem_st_wait();
    } else {
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t u[32];
        tmem_ld_x32(tmem + c0, v);
        tmem_ld_x32(tmem + 256 + c0, u);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        tmem_st_x32(tmem + 256 + c0, v);
      }
      tmem_st_wait();
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) cycles[0] = t1 - t0;
  sink[tid] = acc + __uint_as_float(v[0]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

static float tf32_rn(float x) {   // round-to-nearest-even-ish (rna = ties away) to 10-bit mantissa
  uint32_t u; memcpy(&u, &x, 4);
  u += 0x1000u; u &= 0xFFFFE000u;
  float r; memcpy(&r, &u, 4); return r;
}
static float tf32_trunc(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  float r; memcpy(&r, &u, 4); return r;
}

struct Result { double max_abs_err, mean_rel_err, max_rel_err; long long cycles; };

static int g_a_mn = 1, g_b_mn = 1, g_roundtrip = 0, g_verbose = 0, g_a_off = 0;
static Result run(int N, int mode, int shift, int reps, const std::vector<float>& A, const std::vector<float>& B,
                  int ref_round /*0 exact inputs, 1 rn, 2 trunc*/) {
  const int RB = KROWS + PADROWS;
  float *dA, *dB, *dD; long long* dC;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, 128 * N * 4));
  Params p{dA, dB, dD, dC, N, mode, shift, reps, g_a_mn, g_b_mn, g_roundtrip, g_a_off};
  const size_t smem = 32 * KROWS * 16 + (size_t)(N / 4) * RB * 16;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(p);
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * N);
  long long cyc;
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  Result res{0, 0, 0, cyc};
  if (g_roundtrip) {
    double bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) bad += D[m * N + n] != (float)(m * 1000 + n);
    res.max_abs_err = bad;
    return res;
  }
  if (g_verbose) {
    int exact = 0, nz = 0;
    std::vector<double> ref(128 * N);
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
      double s2 = 0; for (int k = 0; k < KROWS; ++k) s2 += (double)A[k * 128 + m] * (double)B[(k + shift) * N + n];
      ref[m * N + n] = s2 * reps; exact += D[m * N + n] == (float)(s2 * reps); nz += D[m * N + n] != 0.f;
    }
    printf("   exact=%d/%d nonzero=%d  D[0][0..5]= %g %g %g %g %g %g | ref %g %g %g %g %g %g | D[1][0]=%g ref %g D[5][7]=%g ref %g\n",
           exact, 128 * N, nz, D[0], D[1], D[2], D[3], D[4], D[5], ref[0], ref[1], ref[2], ref[3], ref[4], ref[5],
           D[N], ref[N], D[5 * N + 7], ref[5 * N + 7]);
  }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < KROWS; ++k) {
        float a = A[k * 128 + m], b = B[(k + shift) * N + n];
        if (ref_round == 1) { a = tf32_rn(a); b = tf32_rn(b); }
        if (ref_round == 2) { a = tf32_trunc(a); b = tf32_trunc(b); }
        s += (double)a * (double)b;
      }
      s *= reps;
      const double e = (double)D[m * N + n] - s;
      res.max_abs_err = fmax(res.max_abs_err, fabs(e));
      if (s != 0) { res.mean_rel_err += e / s; res.max_rel_err = fmax(res.max_rel_err, fabs(e / s)); }
    }
  res.mean_rel_err /= 128.0 * N;
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return res;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s cc %d.%d sms %d clock %d kHz\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.clockRate);
  const int RB = KROWS + PADROWS;
  srand(1);
  // 0: TMEM st -> ld round trip (no MMA)
  {
    std::vector<float> A(KROWS * 128, 0.f), B((size_t)RB * 128, 0.f);
    g_roundtrip = 1;
    Result r = run(128, 0, 0, 1, A, B, 0);
    printf("tmem st/ld roundtrip: mismatches=%g\n", r.max_abs_err);
    g_roundtrip = 0;
  }
  // 0b: layout controls on integer data, verbose
  g_verbose = 1;
  {
    const int N = 128;
    std::vector<float> A(KROWS * 128), B((size_t)RB * N);
    for (auto& v : A) v = (float)(rand() % 17 - 8);
    for (auto& v : B) v = (float)(rand() % 17 - 8);
    for (int amn : {0, 1}) for (int bmn : {0, 1}) {
      g_a_mn = amn; g_b_mn = bmn;
      Result r = run(N, 0, 0, 1, A, B, 0);
      printf("layout SS a_mn=%d b_mn=%d : max_abs_err=%g\n", amn, bmn, r.max_abs_err);
    }
    for (int bmn : {0, 1}) {
      g_a_mn = 1; g_b_mn = bmn;
      Result r = run(N, 1, 0, 1, A, B, 0);
      printf("layout TS b_mn=%d : max_abs_err=%g\n", bmn, r.max_abs_err);
    }
    g_a_mn = 1; g_b_mn = 1;
    // single hot element: A[k=0][m=0] = 1, B[k=0][n] = n + 1
    std::fill(A.begin(), A.end(), 0.f); std::fill(B.begin(), B.end(), 0.f);
    A[0] = 1.f; for (int n = 0; n < N; ++n) B[n] = (float)(n + 1);
    for (int bmn : {0, 1}) { g_b_mn = bmn; Result r = run(N, 0, 0, 1, A, B, 0); printf("hot SS b_mn=%d a_mn=1: err=%g\n", bmn, r.max_abs_err); }
    g_b_mn = 1;
  }
  // 0c: TMEM <-> register bandwidth (bytes moved per iteration: 128 lanes x 256 columns x 4 B = 128 KB)
  {
    long long* dC; float* dS; CK(cudaMalloc(&dC, 8)); CK(cudaMalloc(&dS, 128 * 4));
    const char* names[3] = {"ld x32 (256 cols)", "st x32 (256 cols)", "drain: ld L1 + ld L2 + add + st L2 (256 cols)"};
    for (int mode = 0; mode < 3; ++mode) {
      tmem_bw_kernel<<<1, 128>>>(mode, 64, dC, dS);
      CK(cudaDeviceSynchronize());
      long long cyc; CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
      printf("tmem bw %-48s : %.0f cycles per 128 KB pass\n", names[mode], (double)cyc / 64);
    }
  }
  // 0d: TS A operand at unaligned TMEM columns
  {
    const int N = 128;
    std::vector<float> A(KROWS * 128), B((size_t)RB * N);
    for (auto& v : A) v = (float)(rand() % 17 - 8);
    for (auto& v : B) v = (float)(rand() % 17 - 8);
    g_a_mn = 0; g_b_mn = 0;
    for (int off : {0, 1, 2, 3, 4, 5, 10}) {
      g_a_off = off;
      Result r = run(N, 1, 0, 1, A, B, 0);
      printf("TS a_off=%2d : max_abs_err=%g\n", off, r.max_abs_err);
    }
    g_a_off = 0;
  }
  g_verbose = 0;
  g_a_mn = 0; g_b_mn = 0;
  // 1/2: exact integer data: layout + descriptor + shift + TS
  for (int N : {128, 256}) {
    std::vector<float> A(KROWS * 128), B((size_t)RB * N);
    for (auto& v : A) v = (float)(rand() % 17 - 8);
    for (auto& v : B) v = (float)(rand() % 17 - 8);
    for (int mode : {0, 1})
      for (int shift : {0, 4, 8, 36}) {
        Result r = run(N, mode, shift, 1, A, B, 0);
        printf("exact  N=%d mode=%s shift=%2d : max_abs_err=%g  cycles=%lld\n", N, mode ? "TS" : "SS", shift, r.max_abs_err, r.cycles);
      }
  }
  // 3: issue rate (reps x 8 MMAs)
  for (int N : {128, 256}) {
    std::vector<float> A(KROWS * 128, 1.0f), B((size_t)RB * N, 1.0f);
    for (int mode : {0, 1})
      for (int reps : {16, 256}) {
        Result r = run(N, mode, 0, reps, A, B, 0);
        printf("rate   N=%d mode=%s mmas=%5d : cycles=%lld  cycles/mma=%.1f  max_abs_err=%g\n", N, mode ? "TS" : "SS",
               reps * 8, r.cycles, (double)r.cycles / (reps * 8), r.max_abs_err);
      }
  }
  // 4: accumulation rounding: positive tf32-exact operands, long K
  {
    const int N = 128;
    std::vector<float> A(KROWS * 128), B((size_t)RB * N);
    for (auto& v : A) v = tf32_trunc(0.5f + (float)rand() / RAND_MAX);
    for (auto& v : B) v = tf32_trunc(0.5f + (float)rand() / RAND_MAX);
    for (int reps : {1, 4, 16, 64, 256, 1024}) {
      Result r = run(N, 0, 0, reps, A, B, 0);
      printf("accum  K=%6d : mean_rel_err=%+.3e max_rel_err=%.3e\n", reps * KROWS, r.mean_rel_err, r.max_rel_err);
    }
    // zero-mean operands (what standardised features look like)
    for (auto& v : A) v = tf32_trunc(2.0f * (float)rand() / RAND_MAX - 1.0f);
    for (auto& v : B) v = tf32_trunc(2.0f * (float)rand() / RAND_MAX - 1.0f);
    for (int reps : {1, 16, 256}) {
      Result r = run(N, 0, 0, reps, A, B, 0);
      printf("accum0 K=%6d : max_abs_err=%.3e (|D| ~ %g)\n", reps * KROWS, r.max_abs_err, sqrt((double)reps * KROWS) / 3);
    }
  }
  // 5: low mantissa bits: full fp32 operands vs references with rn / truncated inputs
  {
    const int N = 128;
    std::vector<float> A(KROWS * 128), B((size_t)RB * N);
    for (auto& v : A) v = 0.5f + (float)rand() / RAND_MAX;
    for (auto& v : B) v = 0.5f + (float)rand() / RAND_MAX;
    Result r0 = run(N, 0, 0, 1, A, B, 0), r1 = run(N, 0, 0, 1, A, B, 1), r2 = run(N, 0, 0, 1, A, B, 2);
    printf("lowbits: max_rel_err vs exact-input ref %.3e | vs rn-input ref %.3e | vs trunc-input ref %.3e\n",
           r0.max_rel_err, r1.max_rel_err, r2.max_rel_err);
    Result t2 = run(N, 1, 0, 1, A, B, 2);
    printf("lowbits(TS): vs trunc-input ref %.3e\n", t2.max_rel_err);
  }
  printf("probe done\n");
  return 0;
}
