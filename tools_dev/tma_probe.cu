// Standalone probe: one 3-D tiled TMA load of a uint8 tensor [3][rows][cols] with variants.
// usage: tma_probe <swizzle 0|1(32)|2(64)|3(128)> <rank 2|3> <cluster 0|1> <entry 0=link 1=runtime>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK, int CL>
__global__ void probe_kernel(const __grid_constant__ CUtensorMap map, uint8_t* out, int box_rows, int c0, int c1, int c2) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = 64 * box_rows;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(base), "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(base), "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(c0), "r"(c1) : "memory");
  }
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  const unsigned char* s = smem + (base - smem_u32(smem));
  for (int i = threadIdx.x; i < 64 * box_rows; i += blockDim.x) out[i] = s[i];
}

int main(int argc, char** argv) {
  const int sw = argc > 1 ? atoi(argv[1]) : 2, rank = argc > 2 ? atoi(argv[2]) : 3, cl = argc > 3 ? atoi(argv[3]) : 0;
  const int rows = 200, cols = 1024;
  const int box_rows = argc > 5 ? atoi(argv[5]) : 128, crow = argc > 4 ? atoi(argv[4]) : 40, ccol = argc > 6 ? atoi(argv[6]) : 64;
  std::vector<uint8_t> h(3 * rows * cols);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 10));
  uint8_t *d, *o;
  cudaMalloc(&d, h.size()); cudaMalloc(&o, 64 * box_rows);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 3};
  cuuint64_t strides[2] = {(cuuint64_t)cols, (cuuint64_t)cols * rows};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUtensorMapSwizzle swz[4] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B};
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  CUresult r = ((Fn)ptr)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swz[sw], CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d (query %d)\n", (int)r, (int)q);
  const size_t smem = 64 * box_rows + 2048;
  if (rank == 3) {
    cudaFuncSetAttribute(probe_kernel<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cl) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      cfg.attrs = &at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, probe_kernel<3, 0>, map, o, box_rows, ccol, crow, 1);
    } else
    probe_kernel<3, 0><<<1, 128, smem>>>(map, o, box_rows, ccol, crow, 1);
  } else {
    cudaFuncSetAttribute(probe_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<2, 0><<<1, 128, smem>>>(map, o, box_rows, 64, 40, 0);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("sw=%d rank=%d: %s\n", sw, rank, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<uint8_t> ho(64 * box_rows);
    cudaMemcpy(ho.data(), o, ho.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    const int pl = rank == 3 ? 1 : 0;
    for (int rr = 0; rr < box_rows; ++rr)
      for (int c = 0; c < 64; ++c) {
        int chunk = c / 16;
        int pc = chunk;
        if (sw == 2) pc = chunk ^ ((rr >> 1) & 3);
        if (sw == 1) pc = chunk ^ ((rr >> 2) & 1);
        if (sw == 3) pc = chunk ^ (rr & 7);      // only chunks 0..3 used with a 64-byte box (see printout)
        const uint8_t got = ho[rr * 64 + pc * 16 + c % 16];
        const size_t src = (size_t)pl * rows * cols + (size_t)(crow + rr) * cols + ccol + c;
        const uint8_t exp = (crow + rr) < rows ? h[src] : 0;
        if (got != exp) ++bad;
      }
    printf("mismatches vs expected swizzle pattern: %d of %d\n", bad, 64 * box_rows);
  }
  return 0;
}
