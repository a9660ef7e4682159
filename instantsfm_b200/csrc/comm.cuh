// comm.cuh -- thin NCCL wrapper (one communicator per process / GPU).
// NCCL is resolved with dlopen at communicator creation so that the single-GPU library has
// no link-time dependency on it; torch ships libnccl.so.2 and has it loaded already.
#pragma once
#include "common.cuh"

struct isfm_comm {
  void* nccl_comm = nullptr;  // ncclComm_t
  int rank = 0;
  int world = 1;
};

namespace isfm {

// In-place sum all-reduce of `count` elements (is_double ? f64 : f32) on `stream`.
// No-op when comm is NULL or world == 1.
void comm_allreduce_sum(isfm_comm* comm, void* buf, size_t count, bool is_double, cudaStream_t stream);
inline int comm_world(const isfm_comm* c) { return c ? c->world : 1; }
inline int comm_rank(const isfm_comm* c) { return c ? c->rank : 0; }

}  // namespace isfm
