// comm.cuh -- multi-GPU plumbing of one rank (one process per GPU).
//
// Two transports:
//  * NCCL (resolved with dlopen at communicator creation, so the single-GPU library has no
//    link-time dependency on it) for the large once-per-trial sums;
//  * a peer-memory exchange over NVLink / NVSwitch for the small per-PCG-iteration sum: every
//    rank owns one device region, mapped into every other rank with CUDA IPC.  A kernel PUSHES
//    its partial vector straight into all peers' regions (plain remote stores), publishes a
//    sequence number behind a system-scope fence, and the consumer kernel on every rank spins on
//    its LOCAL flags and adds the `world` partials in rank order -- the all-reduce is the
//    epilogue of the producing kernel and the prologue of the consuming one, no collective
//    launch, no host involvement, so the whole PCG loop stays one device-side WHILE graph.
#pragma once
#include "common.cuh"

#define ISFM_MAX_PEERS 8

namespace isfm {

// Device-visible view of the exchange, passed BY VALUE to kernels.
// Region layout (identical on every rank):
//   [0, 64)      uint32 flags[2][ISFM_MAX_PEERS]   flags[parity][src] = last sequence number pushed by src
//   [512, 516)   uint32 seq                         number of exchanges completed by THIS rank (local only)
//   [1024, ...)  data[2][world][slot_bytes]         data[parity][src] = partial vector of rank src
struct PeerExchange {
  unsigned char* base[ISFM_MAX_PEERS];
  int world;
  int rank;
  unsigned long long slot_bytes;
};
constexpr size_t PEER_HEADER_BYTES = 1024;
constexpr size_t PEER_SEQ_OFFSET = 512;

}  // namespace isfm

struct isfm_comm {
  void* nccl_comm = nullptr;  // ncclComm_t
  int rank = 0;
  int world = 1;
  // peer-memory exchange (comm_peer_ensure)
  bool peer_ready = false;
  bool peer_failed = false;       // set once: IPC not available on this box, stay on NCCL
  isfm::PeerExchange px{};
  void* local_region = nullptr;
  size_t region_bytes = 0;
};

namespace isfm {

// In-place sum all-reduce of `count` elements (is_double ? f64 : f32) on `stream`.
// No-op when comm is NULL or world == 1.
void comm_allreduce_sum(isfm_comm* comm, void* buf, size_t count, bool is_double, cudaStream_t stream);
inline int comm_world(const isfm_comm* c) { return c ? c->world : 1; }
inline int comm_rank(const isfm_comm* c) { return c ? c->rank : 0; }

// Reduce-scatter over uneven ranges: range r of `buf` ([elem_off[r], elem_off[r] + elem_cnt[r])
// elements) is summed over all ranks into rank r's `own_out` (elem_cnt[rank] elements); `buf` is
// not modified (one grouped ncclReduce per range).
void comm_reduce_ranges(isfm_comm* comm, const void* buf, void* own_out, const size_t* elem_off, const size_t* elem_cnt,
                        bool is_double, cudaStream_t stream);
// recv[r * bytes_per_rank ...] = rank r's `send` (ncclAllGather)
void comm_allgather_bytes(isfm_comm* comm, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t stream);

// COLLECTIVE (every rank, same argument): makes sure the peer exchange exists with at least
// `slot_bytes` per (parity, source) slot.  Returns false -- on every rank alike -- when peer
// memory is unavailable (world > ISFM_MAX_PEERS, IPC refused, ISFM_NO_PEER set); the caller then
// uses comm_allreduce_sum.
bool comm_peer_ensure(isfm_comm* comm, size_t slot_bytes, cudaStream_t stream);

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t* peer_flags(const PeerExchange& px, int dst, int parity) {
  return reinterpret_cast<uint32_t*>(px.base[dst]) + parity * ISFM_MAX_PEERS;
}
__device__ __forceinline__ uint32_t* peer_seq(const PeerExchange& px) {
  return reinterpret_cast<uint32_t*>(px.base[px.rank] + PEER_SEQ_OFFSET);
}
template <typename T>
__device__ __forceinline__ T* peer_slot(const PeerExchange& px, int dst, int parity, int src) {
  return reinterpret_cast<T*>(px.base[dst] + PEER_HEADER_BYTES + ((size_t)parity * px.world + src) * px.slot_bytes);
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Consumer side: wait until every rank has published sequence number `seq` in this rank's own
// flags (threads 0..world-1 poll one flag each).  Bounded: returns false after ~10 s so that a
// lost peer cannot hang the GPU.  Must be called by the whole CTA.
__device__ __forceinline__ bool peer_wait_all(const PeerExchange& px, int parity, uint32_t seq) {
  __shared__ int ok__;
  if (threadIdx.x == 0) ok__ = 1;
  __syncthreads();
  if ((int)threadIdx.x < px.world) {
    const uint32_t* f = peer_flags(px, px.rank, parity) + threadIdx.x;
    long long spins = 0;
    while ((int32_t)(ld_acquire_sys(f) - seq) < 0) {
      if (++spins > (1ll << 24)) { ok__ = 0; break; }
      __nanosleep(spins < 64 ? 0 : 200);
    }
  }
  __syncthreads();
  return ok__ != 0;
}
#endif

}  // namespace isfm
