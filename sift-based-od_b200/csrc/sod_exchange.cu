// K3 over peer memory: the merge of the shard-local top-2 lists of a database sharded over G GPUs
//  without a collective library call.  Replaces, for small query batches, the NCCL form of
// the exchange (sod_top2_keys -> all-to-all -> sod_top2_merge_keys -> all-gather -> sod_top2_from_keys);
// there is no reference counterpart (the reference matches one database in one cv2 call, main.py:70-71).
//
// Every rank owns one exchange buffer (sod_exchange_alloc: plain cudaMalloc memory, exported with
// cudaIpcGetMemHandle and mapped by all peers of the node):
//     header   flags[r] = number of the last call whose keys rank r has finished storing here,
//              epoch    = number of this rank's last call (device-resident, so that a captured CUDA
//                         graph advances it on every replay), block counter of the push kernel
//     keys     [2 parities][G ranks][max_query] key pairs (d2 << 32 | global row, 16 B per row)
// A call is two kernels on the caller's stream:
//   xchg_push_kernel   packs this rank's lists into keys and stores them into slot `rank` of EVERY
//                      rank's buffer (16-byte stores over NVLink for the peers); the last block to finish
//                      publishes the call number in every buffer's flags[rank] (system-scope release)
//   xchg_merge_kernel  waits until all G flags of its own buffer have reached the call number (acquire),
//                      then merges the G key pairs of every row and applies the ratio test
// The key slots alternate with the call parity: a rank can only start call e+2 after its merge of call
// e+1, which waited for every peer's push of e+1, which that peer issued after its own merge of call e -
// so nobody overwrites a slot that is still being read.
#include <cstring>

#include "sod_common.cuh"
#include "sod_top2.cuh"

namespace sod {
namespace {

constexpr int kMaxRanks = SOD_EXCHANGE_MAX_RANKS;
constexpr size_t kHeaderBytes = 512;

struct XchgHeader {
  uint32_t flags[kMaxRanks];  // written by the peers
  uint32_t epoch;             // this rank's call counter
  uint32_t done_blocks;
};
static_assert(sizeof(XchgHeader) <= kHeaderBytes, "header does not fit");

struct XchgArgs {
  void* peer[kMaxRanks];  // every rank's exchange buffer as mapped in this process (own one included)
  int rank, world;
  int64_t max_query, nq;
};

__device__ __forceinline__ longlong2* key_slot(void* buf, int parity, int src_rank, int world, int64_t max_query) {
  return reinterpret_cast<longlong2*>(static_cast<char*>(buf) + kHeaderBytes) +
         (static_cast<int64_t>(parity) * world + src_rank) * max_query;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) xchg_push_kernel(const XchgArgs a, const int32_t* __restrict__ idx,
                                                        const uint32_t* __restrict__ d2) {
  XchgHeader* hdr = static_cast<XchgHeader*>(a.peer[a.rank]);
  const uint32_t epoch = hdr->epoch + 1;  // only rewritten by the last block, after every block has read it
  const int parity = static_cast<int>(epoch & 1u);
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; row < a.nq;
       row += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const longlong2 k = top2_pack_keys(*reinterpret_cast<const int2*>(idx + row * 2),
                                       *reinterpret_cast<const uint2*>(d2 + row * 2));
    SOD_DCHECK(row < a.max_query);
    for (int p = 0; p < a.world; ++p) key_slot(a.peer[p], parity, a.rank, a.world, a.max_query)[row] = k;
  }
  __threadfence_system();  // this thread's peer stores are visible system-wide before the block reports
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(&hdr->done_blocks, 1u) == gridDim.x - 1) {
      hdr->done_blocks = 0;
      hdr->epoch = epoch;
      __threadfence_system();
      for (int p = 0; p < a.world; ++p)
        st_release_sys(&static_cast<XchgHeader*>(a.peer[p])->flags[a.rank], epoch);
    }
  }
}

#ifndef SOD_EXCHANGE_WATCHDOG_CYCLES
#define SOD_EXCHANGE_WATCHDOG_CYCLES 8000000000ll  // ~4 s: a rank that never calls is a protocol error
#endif

__global__ void __launch_bounds__(256) xchg_merge_kernel(const XchgArgs a, int32_t* __restrict__ out_idx,
                                                         uint32_t* __restrict__ out_d2, float* __restrict__ out_dist,
                                                         uint8_t* __restrict__ out_pass, double ratio) {
  void* own = a.peer[a.rank];
  const XchgHeader* hdr = static_cast<const XchgHeader*>(own);
  const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(&hdr->epoch);  // set by the push kernel before
  if (threadIdx.x < a.world) {
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(&hdr->flags[threadIdx.x]) - epoch) < 0) {
      __nanosleep(64);
      if (clock64() - t0 > SOD_EXCHANGE_WATCHDOG_CYCLES) __trap();
    }
  }
  __syncthreads();
  const int parity = static_cast<int>(epoch & 1u);
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; row < a.nq;
       row += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    long long k1 = kNoneKey, k2 = kNoneKey;
    for (int r = 0; r < a.world; ++r)  // L2 reads: the lines were written by other GPUs during this launch
      top2_fold_keys(k1, k2, __ldcg(key_slot(own, parity, r, a.world, a.max_query) + row));
    write_top2_keys_row(row, k1, k2, out_idx, out_d2, out_dist, out_pass, ratio);
  }
}

}  // namespace
}  // namespace sod

using namespace sod;

extern "C" {

size_t sod_exchange_bytes(int64_t max_query, int32_t world) {
  if (max_query < 0 || world < 1 || world > kMaxRanks) return 0;
  return kHeaderBytes + static_cast<size_t>(2) * world * static_cast<size_t>(max_query) * sizeof(longlong2);
}

int sod_exchange_alloc(size_t bytes, void** buffer_out) {
  SOD_CHECK_ARG(buffer_out && bytes >= kHeaderBytes, "null output or size below the header");
  void* p = nullptr;
  SOD_CHECK_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("clearing the exchange buffer failed: %s", cudaGetErrorString(e));
    return SOD_ERR_CUDA;
  }
  *buffer_out = p;
  return SOD_OK;
}

int sod_exchange_free(void* buffer) {
  if (buffer) SOD_CHECK_CUDA(cudaFree(buffer));
  return SOD_OK;
}

int sod_ipc_export(const void* buffer, uint8_t* handle_host) {
  SOD_CHECK_ARG(buffer && handle_host, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == SOD_IPC_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h;
  SOD_CHECK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(buffer)));
  memcpy(handle_host, &h, sizeof(h));
  return SOD_OK;
}

int sod_ipc_open(const uint8_t* handle_host, void** buffer_out) {
  SOD_CHECK_ARG(handle_host && buffer_out, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  SOD_CHECK_CUDA(cudaIpcOpenMemHandle(buffer_out, h, cudaIpcMemLazyEnablePeerAccess));
  return SOD_OK;
}

int sod_ipc_close(void* buffer) {
  if (buffer) SOD_CHECK_CUDA(cudaIpcCloseMemHandle(buffer));
  return SOD_OK;
}

int sod_top2_exchange_peer(const int32_t* idx, const uint32_t* d2, int64_t n_query, int32_t rank, int32_t world,
                           void* const* peer_buffers_host, int64_t max_query, int32_t* out_idx, uint32_t* out_d2,
                           float* out_dist, uint8_t* out_pass, double ratio, sod_stream_t stream) {
  SOD_CHECK_ARG(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "bad rank / world");
  SOD_CHECK_ARG(n_query >= 0 && n_query <= max_query, "n_query exceeds the capacity of the exchange buffers");
  SOD_CHECK_ARG(peer_buffers_host, "null buffer table");
  if (n_query == 0) return SOD_OK;  // nothing to exchange: every rank skips the call alike
  SOD_CHECK_ARG(idx && d2 && out_idx && out_d2, "null pointer");
  XchgArgs a;
  for (int p = 0; p < kMaxRanks; ++p) a.peer[p] = p < world ? peer_buffers_host[p] : nullptr;
  for (int p = 0; p < world; ++p) SOD_CHECK_ARG(a.peer[p], "null exchange buffer of rank %d", p);
  a.rank = rank; a.world = world; a.max_query = max_query; a.nq = n_query;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 256;
  int64_t blocks = (n_query + threads - 1) / threads;
  const int sms = device_sm_count();
  if (sms <= 0) return SOD_ERR_CUDA;
  if (blocks > static_cast<int64_t>(sms) * 4) blocks = static_cast<int64_t>(sms) * 4;
  xchg_push_kernel<<<static_cast<unsigned>(blocks), threads, 0, st>>>(a, idx, d2);
  SOD_CHECK_LAUNCH("xchg_push_kernel");
  xchg_merge_kernel<<<static_cast<unsigned>(blocks), threads, 0, st>>>(a, out_idx, out_d2, out_dist, out_pass, ratio);
  SOD_CHECK_LAUNCH("xchg_merge_kernel");
  return SOD_OK;
}

}  // extern "C"
