// Small all-reduce over NVLink peer memory: the frame-sharded hot path reduces a few KB per
// Lloyd iteration ([sums | counts | stats]), per projection pass ([-min | max] of the CVs) and per KMeans
// set-up.  At that size an NCCL all-reduce is pure latency (~25 us on 8 GPUs, ten of them per C2 step), so
// these exchanges are ONE kernel of one CTA per GPU that works on peer memory directly:
//
//   push:    every rank stores its vector into slot `rank` of every peer's inbox (remote stores over
//            NVLink / NVSwitch, coalesced), fences at system scope and releases flag[rank] = seq on every peer;
//   wait:    until all `world` flags of its own inbox show seq (acquire, system scope);
//   reduce:  out[i] = op over the slots in rank order -- the same order on every rank, so all ranks hold
//            bitwise identical results (an atomics-based reduction would not).
//
// Two inbox sets alternate with the parity of seq: a rank can start exchange s + 1 (and overwrite its slot on
// a peer) while that peer still sums exchange s, but not s + 2 -- it cannot finish s + 1 without the peer's
// flag for s + 1, which the peer sends after its kernel for s has ended (stream order).
// The buffers are plain cudaMalloc allocations shared through CUDA IPC handles (one process per GPU).
#include <cstring>
#include "dcg_common.cuh"

namespace dcg {
namespace {

constexpr int kP2PMaxWorld = 16;
constexpr int kP2PThreads = 1024;
constexpr size_t kP2PHeaderBytes = 4096;          // flags: [parity 2][world <= 16] x 128 bytes

struct P2PPeers { double* base[kP2PMaxWorld]; };

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long* flag_of(double* base, int parity, int r) {
  return reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(base) + ((size_t)parity * kP2PMaxWorld + r) * 128);
}
__device__ __forceinline__ double* slot_of(double* base, int parity, int r, int world, size_t slot) {
  return reinterpret_cast<double*>(reinterpret_cast<char*>(base) + kP2PHeaderBytes) + ((size_t)parity * world + r) * slot;
}

// op: 0 = sum, 1 = max
__global__ void __launch_bounds__(kP2PThreads, 1)
p2p_allreduce_kernel(const double* __restrict__ v, double* __restrict__ out, int n, int op, int rank, int world,
                     P2PPeers peers, size_t slot, unsigned long long seq) {
  const int tid = threadIdx.x, parity = (int)(seq & 1ull);
  for (int p = 0; p < world; ++p) {
    double* dst = slot_of(peers.base[p], parity, rank, world, slot);
    for (int i = tid; i < n; i += kP2PThreads) dst[i] = v[i];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < world) st_release_sys(flag_of(peers.base[tid], parity, rank), seq);
  if (tid < world) {
    const unsigned long long* f = flag_of(peers.base[rank], parity, tid);
    while (ld_acquire_sys(f) < seq) __nanosleep(100);
  }
  __syncthreads();
  __threadfence_system();
  const double* mine = slot_of(peers.base[rank], parity, 0, world, slot);
  for (int i = tid; i < n; i += kP2PThreads) {
    double a = __ldcv(mine + i);                    // written by peers: never from a stale L1 line
    for (int r = 1; r < world; ++r) {
      const double b = __ldcv(mine + (size_t)r * slot + i);
      a = op == 0 ? a + b : fmax(a, b);
    }
    out[i] = a;
  }
}

}  // namespace
}  // namespace dcg

using namespace dcg;

extern "C" size_t dcg_p2p_buffer_bytes(int world, int64_t slot_doubles) {
  if (world < 1 || world > kP2PMaxWorld || slot_doubles < 1) return 0;
  return kP2PHeaderBytes + (size_t)2 * world * (size_t)slot_doubles * sizeof(double);
}

extern "C" int dcg_p2p_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  if (!ptr || !handle64 || bytes == 0) return DCG_E_NULL;
  void* p = nullptr;
  DCG_CUDA_TRY(cudaMalloc(&p, bytes));
  DCG_CUDA_TRY(cudaMemset(p, 0, bytes));
  DCG_CUDA_TRY(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  DCG_CUDA_TRY(cudaIpcGetMemHandle(&h, p));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  *ptr = p;
  return 0;
}

extern "C" int dcg_p2p_open(const unsigned char* handle64, void** ptr) {
  if (!ptr || !handle64) return DCG_E_NULL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  DCG_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int dcg_p2p_close(void* ptr, int opened) {
  if (!ptr) return 0;
  if (opened) DCG_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  else DCG_CUDA_TRY(cudaFree(ptr));
  return 0;
}

extern "C" int dcg_p2p_allreduce_f64(const double* v, double* out, int n, int op, int rank, int world,
                                     void* const* peer_bases, int64_t slot_doubles, uint64_t seq, void* stream) {
  if (!v || !out || !peer_bases) return DCG_E_NULL;
  if (world < 1 || world > kP2PMaxWorld || rank < 0 || rank >= world || n < 1 || n > slot_doubles || seq == 0)
    return DCG_E_SHAPE;
  if (op != 0 && op != 1) return DCG_E_MODE;
  P2PPeers peers;
  for (int r = 0; r < kP2PMaxWorld; ++r) peers.base[r] = r < world ? (double*)peer_bases[r] : nullptr;
  for (int r = 0; r < world; ++r)
    if (!peers.base[r]) return DCG_E_NULL;
  p2p_allreduce_kernel<<<1, kP2PThreads, 0, (cudaStream_t)stream>>>(v, out, n, op, rank, world, peers, (size_t)slot_doubles,
                                                                   (unsigned long long)seq);
  DCG_LAUNCH_CHECK();
  return 0;
}
