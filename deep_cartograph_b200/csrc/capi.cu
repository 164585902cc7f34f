// Library-level entry points: version, error strings, device info.
#include "dcg_common.cuh"

extern "C" int dcg_version(void) { return DCG_VERSION; }

extern "C" const char* dcg_error_string(int code) {
  if (code == 0) return "ok";
  switch (code) {
    case DCG_E_NULL: return "dcg: required pointer is NULL";
    case DCG_E_SHAPE: return "dcg: shape / size argument out of range";
    case DCG_E_WORKSPACE: return "dcg: workspace missing or too small";
    case DCG_E_ALIGN: return "dcg: pointer alignment";
    case DCG_E_ARCH: return "dcg: device is not sm_100 (tcgen05 engine needs a B200)";
    case DCG_E_MODE: return "dcg: unknown engine / dtype flag";
    default: break;
  }
  if (code < 0 && code > -1000) return cudaGetErrorString((cudaError_t)(-code));
  return "dcg: unknown error";
}

extern "C" int dcg_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  DCG_CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp p;
  DCG_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}

// ---- per-device, per-kernel attribute caches (see dcg_common.cuh) --------------------------------
#include <mutex>
namespace dcg {
namespace {
struct FuncEntry { const void* func; int device; size_t smem_set; int occ_threads; size_t occ_smem; int occ; };
constexpr int kMaxFuncEntries = 512;
FuncEntry g_funcs[kMaxFuncEntries];
int g_n_funcs = 0;
std::mutex g_funcs_mutex;

FuncEntry* find_entry(const void* func, int device) {
  for (int i = 0; i < g_n_funcs; ++i)
    if (g_funcs[i].func == func && g_funcs[i].device == device) return &g_funcs[i];
  if (g_n_funcs == kMaxFuncEntries) return nullptr;
  g_funcs[g_n_funcs] = FuncEntry{func, device, 0, 0, 0, 0};
  return &g_funcs[g_n_funcs++];
}
}  // namespace

cudaError_t ensure_dynamic_smem(const void* func, size_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(g_funcs_mutex);
  FuncEntry* en = find_entry(func, dev);
  if (en && en->smem_set >= bytes && en->smem_set > 0) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess && en) en->smem_set = bytes;
  return e;
}

cudaError_t cached_occupancy(int* per_sm, const void* func, int threads, size_t smem) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(g_funcs_mutex);
  FuncEntry* en = find_entry(func, dev);
  if (en && en->occ > 0 && en->occ_threads == threads && en->occ_smem == smem) { *per_sm = en->occ; return cudaSuccess; }
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, func, threads, smem);
  if (e == cudaSuccess && en) { en->occ = *per_sm; en->occ_threads = threads; en->occ_smem = smem; }
  return e;
}
}  // namespace dcg
