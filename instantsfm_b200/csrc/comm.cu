// comm.cu -- NCCL via dlopen and the CUDA-IPC peer-memory exchange; see comm.cuh.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <vector>

#include "comm.cuh"

namespace isfm {
namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& api() {
  static NcclApi a;
  if (a.lib) return a;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.lib) break;
  }
  if (!a.lib) throw IsfmError(ISFM_ENCCL, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
#define LOAD(field, sym)                                                                   \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, sym));                        \
  if (!a.field) throw IsfmError(ISFM_ENCCL, "missing NCCL symbol " sym)
  LOAD(GetUniqueId, "ncclGetUniqueId");
  LOAD(CommInitRank, "ncclCommInitRank");
  LOAD(CommDestroy, "ncclCommDestroy");
  LOAD(AllReduce, "ncclAllReduce");
  LOAD(AllGather, "ncclAllGather");
  LOAD(Reduce, "ncclReduce");
  LOAD(GroupStart, "ncclGroupStart");
  LOAD(GroupEnd, "ncclGroupEnd");
  LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
  return a;
}

void check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) throw IsfmError(ISFM_ENCCL, std::string(what) + ": " + api().GetErrorString(r));
}

}  // namespace

void comm_allreduce_sum(isfm_comm* comm, void* buf, size_t count, bool is_double, cudaStream_t stream) {
  if (!comm || comm->world <= 1 || count == 0) return;
  g_launch_count++;
  check(api().AllReduce(buf, buf, count, is_double ? ncclDouble : ncclFloat, ncclSum,
                        static_cast<ncclComm_t>(comm->nccl_comm), stream), "ncclAllReduce");
}

void comm_reduce_ranges(isfm_comm* comm, const void* buf, void* own_out, const size_t* elem_off, const size_t* elem_cnt,
                        bool is_double, cudaStream_t stream) {
  if (!comm || comm->world <= 1) return;
  const size_t es = is_double ? 8 : 4;
  g_launch_count++;
  check(api().GroupStart(), "ncclGroupStart");
  for (int r = 0; r < comm->world; ++r) {
    if (elem_cnt[r] == 0) continue;
    const void* p = static_cast<const unsigned char*>(buf) + elem_off[r] * es;
    check(api().Reduce(p, own_out, elem_cnt[r], is_double ? ncclDouble : ncclFloat, ncclSum, r,
                       static_cast<ncclComm_t>(comm->nccl_comm), stream), "ncclReduce");   // own_out is read at the root only
  }
  check(api().GroupEnd(), "ncclGroupEnd");
}

void comm_allgather_bytes(isfm_comm* comm, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t stream) {
  ISFM_REQUIRE(comm && comm->world > 1, ISFM_EINVAL, "comm_allgather_bytes without a communicator");
  g_launch_count++;
  check(api().AllGather(send, recv, bytes_per_rank, ncclChar, static_cast<ncclComm_t>(comm->nccl_comm), stream), "ncclAllGather");
}

namespace {

void peer_teardown(isfm_comm* comm) {
  if (!comm) return;
  for (int r = 0; r < comm->world && r < ISFM_MAX_PEERS; ++r)
    if (r != comm->rank && comm->px.base[r]) cudaIpcCloseMemHandle(comm->px.base[r]);
  if (comm->local_region) cudaFree(comm->local_region);
  comm->local_region = nullptr;
  comm->region_bytes = 0;
  comm->peer_ready = false;
  comm->px = PeerExchange{};
}

// sum over ranks of one float per rank (a consensus vote), host result
float vote_sum(isfm_comm* comm, float mine, float* d_scratch, cudaStream_t s) {
  ISFM_CUDA(cudaMemcpyAsync(d_scratch, &mine, sizeof(float), cudaMemcpyHostToDevice, s));
  check(api().AllReduce(d_scratch, d_scratch, 1, ncclFloat, ncclSum, static_cast<ncclComm_t>(comm->nccl_comm), s), "ncclAllReduce");
  float out = 0.f;
  ISFM_CUDA(cudaMemcpyAsync(&out, d_scratch, sizeof(float), cudaMemcpyDeviceToHost, s));
  ISFM_CUDA(cudaStreamSynchronize(s));
  return out;
}

}  // namespace

bool comm_peer_ensure(isfm_comm* comm, size_t slot_bytes, cudaStream_t s) {
  if (!comm || comm->world <= 1) return false;
  if (comm->peer_failed || comm->world > ISFM_MAX_PEERS || getenv("ISFM_NO_PEER")) return false;
  slot_bytes = (slot_bytes + 255) / 256 * 256;
  if (comm->peer_ready && slot_bytes <= comm->px.slot_bytes) return true;
  const int world = comm->world, rank = comm->rank;
  // everything in flight on this device (kernels of an earlier handle that still push / poll)
  // must be finished on EVERY rank before the old region goes away: sync + vote as a barrier
  ISFM_CUDA(cudaDeviceSynchronize());
  float* d_vote = nullptr;
  ISFM_CUDA(cudaMalloc(&d_vote, sizeof(float)));
  vote_sum(comm, 1.f, d_vote, s);
  peer_teardown(comm);
  // a generous slot so that later, larger camera systems rarely re-map (2 x world slots per region)
  slot_bytes = std::max<size_t>(slot_bytes * 2, (size_t)1 << 20);
  const size_t bytes = PEER_HEADER_BYTES + 2 * (size_t)world * slot_bytes;
  bool ok = true;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof mine);
  if (cudaMalloc(&comm->local_region, bytes) != cudaSuccess) { cudaGetLastError(); ok = false; comm->local_region = nullptr; }
  if (ok) ok = cudaMemset(comm->local_region, 0, bytes) == cudaSuccess;
  if (ok) ok = cudaIpcGetMemHandle(&mine, comm->local_region) == cudaSuccess;
  if (!ok) cudaGetLastError();
  // exchange the handles (64 bytes each) with an NCCL all-gather
  unsigned char* d_handles = nullptr;
  ISFM_CUDA(cudaMalloc(&d_handles, (size_t)world * sizeof(cudaIpcMemHandle_t)));
  ISFM_CUDA(cudaMemcpyAsync(d_handles + (size_t)rank * sizeof mine, &mine, sizeof mine, cudaMemcpyHostToDevice, s));
  check(api().AllGather(d_handles + (size_t)rank * sizeof mine, d_handles, sizeof mine, ncclChar,
                        static_cast<ncclComm_t>(comm->nccl_comm), s), "ncclAllGather");
  std::vector<cudaIpcMemHandle_t> all(world);
  ISFM_CUDA(cudaMemcpyAsync(all.data(), d_handles, (size_t)world * sizeof mine, cudaMemcpyDeviceToHost, s));
  ISFM_CUDA(cudaStreamSynchronize(s));
  ok = vote_sum(comm, ok ? 1.f : 0.f, d_vote, s) == (float)world;   // every rank produced a handle
  comm->px.world = world; comm->px.rank = rank; comm->px.slot_bytes = slot_bytes;
  if (ok) {
    comm->px.base[rank] = static_cast<unsigned char*>(comm->local_region);
    for (int r = 0; r < world && ok; ++r) {
      if (r == rank) continue;
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
      comm->px.base[r] = static_cast<unsigned char*>(p);
    }
  }
  ok = vote_sum(comm, ok ? 1.f : 0.f, d_vote, s) == (float)world;   // every rank mapped every peer
  cudaFree(d_handles);
  cudaFree(d_vote);
  if (!ok) {
    peer_teardown(comm);
    comm->peer_failed = true;
    return false;
  }
  comm->region_bytes = bytes;
  comm->peer_ready = true;
  return true;
}

}  // namespace isfm

using namespace isfm;

extern "C" int isfm_comm_unique_id(uint8_t id_out[128]) {
  try {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    check(api().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(id_out, &id, 128);
    return ISFM_OK;
  } catch (const IsfmError& e) { set_last_error(e.what()); return e.code; }
}

extern "C" int isfm_comm_create(const uint8_t id_in[128], int rank, int world, isfm_comm** out) {
  try {
    ISFM_REQUIRE(out && id_in && world >= 1 && rank >= 0 && rank < world, ISFM_EINVAL, "isfm_comm_create");
    ncclUniqueId id;
    memcpy(&id, id_in, 128);
    ncclComm_t c;
    check(api().CommInitRank(&c, world, id, rank), "ncclCommInitRank");
    isfm_comm* h = new isfm_comm();
    h->nccl_comm = c; h->rank = rank; h->world = world;
    *out = h;
    return ISFM_OK;
  } catch (const IsfmError& e) { set_last_error(e.what()); return e.code; }
}

extern "C" int isfm_comm_peer_enabled(const isfm_comm* comm) { return comm && comm->peer_ready ? 1 : 0; }

extern "C" void isfm_comm_destroy(isfm_comm* comm) {
  if (!comm) return;
  cudaDeviceSynchronize();
  peer_teardown(comm);
  cudaGetLastError();
  try { if (comm->nccl_comm) api().CommDestroy(static_cast<ncclComm_t>(comm->nccl_comm)); } catch (...) {}
  delete comm;
}
