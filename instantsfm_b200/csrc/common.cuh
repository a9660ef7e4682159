// common.cuh -- error handling, device buffers, launch accounting, block reductions.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <map>
#include <mutex>
#include <utility>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/isfm_b200.h"

namespace isfm {

// thread-local last error (isfm_last_error)
void set_last_error(const std::string& msg);
const char* get_last_error();

struct IsfmError : std::runtime_error {
  int code;
  IsfmError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define ISFM_CUDA(call)                                                                    \
  do {                                                                                     \
    cudaError_t err__ = (call);                                                            \
    if (err__ != cudaSuccess) {                                                            \
      char buf__[512];                                                                     \
      snprintf(buf__, sizeof buf__, "%s at %s:%d: %s", #call, __FILE__, __LINE__,          \
               cudaGetErrorString(err__));                                                 \
      throw ::isfm::IsfmError(ISFM_ECUDA, buf__);                                          \
    }                                                                                      \
  } while (0)

#define ISFM_REQUIRE(cond, code, msg)                                                      \
  do {                                                                                     \
    if (!(cond)) throw ::isfm::IsfmError(code, std::string(msg) + " (" #cond ")");         \
  } while (0)

// global launch counter (isfm_launch_count) -- the bench reports it as gpu_launches
extern int64_t g_launch_count;

enum Timer {
  T_LINEARIZE = 0, T_POINT_BLOCKS, T_POINT_SOLVE, T_CAMERA_BLOCKS, T_SCHUR_OFFDIAG, T_PRECOND,
  T_PCG_SPMV, T_PCG_VEC, T_BACKSUB, T_UPDATE, T_COST, T_INDEX_PREP, T_REDUCE, T_COMM, T_MISC, T_SPARE
};
static_assert(T_SPARE + 1 == ISFM_N_TIMERS, "timer table size");

// Optional per-kernel-family CUDA-event timing on the handle's stream.
struct KernelTimers {
  bool enabled = false;
  cudaStream_t stream = nullptr;
  double ms[ISFM_N_TIMERS] = {0};
  int64_t launches[ISFM_N_TIMERS] = {0};
  struct Pending { cudaEvent_t a, b; int id; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;

  cudaEvent_t get_event() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; ISFM_CUDA(cudaEventCreate(&e)); return e;
  }
  void begin(int id) {
    launches[id]++;
    g_launch_count++;
    if (!enabled) return;
    Pending p{get_event(), get_event(), id};
    ISFM_CUDA(cudaEventRecord(p.a, stream));
    pending.push_back(p);
  }
  void end() {
    if (!enabled) return;
    ISFM_CUDA(cudaEventRecord(pending.back().b, stream));
  }
  void resolve() {
    if (pending.empty()) return;
    ISFM_CUDA(cudaStreamSynchronize(stream));
    for (auto& p : pending) {
      float t = 0.f;
      ISFM_CUDA(cudaEventElapsedTime(&t, p.a, p.b));
      ms[p.id] += t;
      pool.push_back(p.a); pool.push_back(p.b);
    }
    pending.clear();
  }
  void reset(bool enable) {
    resolve();
    for (int i = 0; i < ISFM_N_TIMERS; ++i) { ms[i] = 0; launches[i] = 0; }
    enabled = enable;
  }
  ~KernelTimers() {
    for (auto& p : pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : pool) cudaEventDestroy(e);
  }
};

// RAII scope: KT(timers, id); kernel<<<>>>(); -- end recorded at scope exit
struct TimerScope {
  KernelTimers& t;
  TimerScope(KernelTimers& t_, int id) : t(t_) { t.begin(id); }
  ~TimerScope() { t.end(); }
};

// Device memory comes from the device's default stream-ordered pool with an unlimited release
// threshold: buffers freed by one handle are re-used by the next (the reference's pipeline
// creates three BA solvers back to back) instead of going back to the driver.
inline void ensure_pool_configured() {
  static bool done = false;
  if (done) return;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t threshold = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
  done = true;
}

// Process-wide cache of large device blocks.  The pipeline creates solver after solver of (almost)
// the same size (three BA rounds, the retriangulation loop); handing the freed blocks of one
// handle straight to the next avoids the pool's sub-allocation of big free chunks by small
// requests, which fragments it and makes later handles map new physical memory -- measured as
// set-up stalls of up to 0.9 s on a 5 M-observation problem.  Blocks of >= 1 MB are kept (up to
// 16 GB in total, then the largest go back to the pool) and re-used for requests of 88..100 % of
// their size.  Safe because every handle works on one stream and synchronises it before it dies.
struct BufferCache {
  std::mutex m;
  std::multimap<std::pair<int, size_t>, void*> blocks;   // (device, bytes) -> block: never handed across devices
  size_t cached = 0;
  static constexpr size_t MIN_BYTES = (size_t)1 << 20, LIMIT = (size_t)16 << 30;
  bool enabled = getenv("ISFM_NO_BUFFER_CACHE") == nullptr;
  void* take(size_t bytes, size_t* got) {
    if (!enabled || bytes < MIN_BYTES) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(m);
    auto it = blocks.lower_bound(std::make_pair(dev, bytes));
    if (it == blocks.end() || it->first.first != dev || it->first.second > bytes + bytes / 8) return nullptr;
    void* p = it->second;
    *got = it->first.second;
    cached -= it->first.second;
    blocks.erase(it);
    return p;
  }
  void give(void* p, size_t bytes) {
    int dev = 0;
    if (enabled && bytes >= MIN_BYTES && cudaGetDevice(&dev) == cudaSuccess) {
      // (blocks are released under the device they were allocated on: a handle lives on one device)
      std::lock_guard<std::mutex> lk(m);
      blocks.emplace(std::make_pair(dev, bytes), p);
      cached += bytes;
      // over the limit: the largest blocks of THIS device go back to the pool
      while (cached > LIMIT) {
        auto it = blocks.lower_bound(std::make_pair(dev + 1, (size_t)0));
        if (it == blocks.begin()) break;
        --it;
        if (it->first.first != dev) break;
        cudaFreeAsync(it->second, 0);
        cached -= it->first.second;
        blocks.erase(it);
      }
      return;
    }
    cudaFreeAsync(p, 0);
  }
};
inline BufferCache& buffer_cache() { static BufferCache* c = new BufferCache(); return *c; }   // never destroyed: outlives every handle

template <typename T>
struct DeviceBuffer {
  T* ptr = nullptr;
  size_t count = 0;
  size_t block_bytes = 0;   // size of the underlying block (>= count * sizeof(T) when it came from the cache)
  DeviceBuffer() {}
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  ~DeviceBuffer() { release(); }
  void release() { if (ptr) buffer_cache().give(ptr, block_bytes); ptr = nullptr; count = 0; block_bytes = 0; }
  void alloc(size_t n) {
    if (n <= count && ptr) return;
    release();
    if (n == 0) n = 1;
    ensure_pool_configured();
    size_t got = 0;
    if (void* p = buffer_cache().take(n * sizeof(T), &got)) {
      ptr = static_cast<T*>(p); block_bytes = got; count = n;
      return;
    }
    ISFM_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ptr), n * sizeof(T), 0));
    count = n; block_bytes = n * sizeof(T);
  }
  void zero(cudaStream_t s) { if (ptr) ISFM_CUDA(cudaMemsetAsync(ptr, 0, count * sizeof(T), s)); }
  T* get() const { return ptr; }
  void swap(DeviceBuffer& o) { std::swap(ptr, o.ptr); std::swap(count, o.count); std::swap(block_bytes, o.block_bytes); }
};

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
// Asynchronous global -> shared copies (LDGSTS): 16 bytes through L2 only for streamed data,
// 4 / 8 bytes through L1 for gathered vector entries.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
template <int BYTES> __device__ __forceinline__ void cp_async_small(void* smem_dst, const void* gmem_src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
}
// L2 eviction-priority policies (126 MB L2): streamed-once data is marked evict_first so that
// small producer -> consumer buffers marked evict_last survive until the next kernel reads them.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ void cp_async16_hint(void* smem_dst, const void* gmem_src, uint64_t policy) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "l"(policy) : "memory");
}
__device__ __forceinline__ void st_global_hint(float* p, float v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ void st_global_hint(double* p, double v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Bulk asynchronous copy (TMA, UBLKCP): one thread moves a contiguous, 16-byte aligned run of
// global memory into shared memory; completion is signalled on an mbarrier by byte count.
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(d), "l"(gmem_src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(a), "r"(parity) : "memory");
}

// block-wide sum in double; result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh__[32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh__[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh__[threadIdx.x] : 0.0;
  if (w == 0)
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic grid-wide sum of per-block partials: every block re-reduces `n` doubles.
__device__ __forceinline__ double reduce_partials(const double* __restrict__ part, int n) {
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += part[i];
  v = block_sum(v);
  __shared__ double bc__;
  if (threadIdx.x == 0) bc__ = v;
  __syncthreads();
  return bc__;
}
#endif

}  // namespace isfm
