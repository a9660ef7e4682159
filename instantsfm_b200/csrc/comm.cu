// comm.cu -- NCCL via dlopen; see comm.cuh.
#include <dlfcn.h>
#include <nccl.h>

#include "comm.cuh"

namespace isfm {
namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
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

extern "C" void isfm_comm_destroy(isfm_comm* comm) {
  if (!comm) return;
  try { if (comm->nccl_comm) api().CommDestroy(static_cast<ncclComm_t>(comm->nccl_comm)); } catch (...) {}
  delete comm;
}
