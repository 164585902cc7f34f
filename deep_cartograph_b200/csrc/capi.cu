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
