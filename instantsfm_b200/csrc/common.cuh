// common.cuh -- error handling, device buffers, launch accounting, block reductions.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <atomic>
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
extern std::atomic<int64_t> g_launch_count;
// false while a thread records kernels into a CUDA graph (captured, not executed)
inline bool& tl_count_launches() { static thread_local bool on = true; return on; }

enum Timer {
  T_LINEARIZE = 0, T_POINT_BLOCKS, T_POINT_SOLVE, T_CAMERA_BLOCKS, T_SCHUR_OFFDIAG, T_PRECOND,
  T_PCG_SPMV, T_PCG_VEC, T_BACKSUB, T_UPDATE, T_COST, T_INDEX_PREP, T_REDUCE, T_COMM, T_MISC, T_COARSE
};
static_assert(T_COARSE + 1 == ISFM_N_TIMERS, "timer table size");

// Optional per-kernel-family CUDA-event timing on the handle's stream.
struct KernelTimers {
  bool enabled = false;
  bool fine = false;   // per-iteration kernels instead of the persistent PCG kernel (ISFM_TIMERS_FINE=1)
  bool enabled_fine() const { return enabled && fine; }
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
    if (tl_count_launches()) g_launch_count++;
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
    fine = getenv("ISFM_TIMERS_FINE") != nullptr;
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

// ---------------------------------------------------------------------------------------
// Device / stream context of the calling host thread.  Every C entry point (destroy included)
// opens a CtxGuard with the handle's device and stream: the device becomes current for the call
// and every DeviceBuffer allocated or released inside it is ordered on that stream.
// ---------------------------------------------------------------------------------------
struct DeviceCtx {
  int dev = -1;
  cudaStream_t stream = nullptr;
  bool synced = false;   // the stream has been synchronised and receives no more work (handle teardown)
};
inline DeviceCtx& tl_ctx() { static thread_local DeviceCtx c; return c; }

struct CtxGuard {
  DeviceCtx saved;
  int prev_dev = -1, dev = -1;
  CtxGuard(int dev_, cudaStream_t s) : dev(dev_) {
    saved = tl_ctx();
    if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; } }
    if (cudaGetDevice(&prev_dev) != cudaSuccess) { cudaGetLastError(); prev_dev = -1; }
    if (prev_dev != dev) cudaSetDevice(dev);
    tl_ctx() = DeviceCtx{dev, s, false};
  }
  void mark_synced() { tl_ctx().synced = true; }
  ~CtxGuard() {
    tl_ctx() = saved;
    if (prev_dev >= 0 && prev_dev != dev) cudaSetDevice(prev_dev);
  }
  CtxGuard(const CtxGuard&) = delete;
  CtxGuard& operator=(const CtxGuard&) = delete;
};

// Device memory comes from a PRIVATE stream-ordered pool per device (never the device's default
// pool, which torch's cudaMallocAsync backend shares): its release threshold is unlimited so that
// buffers freed by one handle are re-used by the next (the reference's pipeline creates three BA
// solvers back to back), and isfm_trim_cache() hands everything back to the driver.
constexpr int ISFM_MAX_DEVICES = 64;
struct DevicePools {
  std::mutex m;
  cudaMemPool_t pool[ISFM_MAX_DEVICES] = {nullptr};
  cudaMemPool_t get(int dev) {
    if (dev < 0 || dev >= ISFM_MAX_DEVICES) return nullptr;
    std::lock_guard<std::mutex> lk(m);
    if (pool[dev]) return pool[dev];
    cudaMemPoolProps props;
    memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t p = nullptr;
    if (cudaMemPoolCreate(&p, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    uint64_t threshold = ~0ull;
    cudaMemPoolSetAttribute(p, cudaMemPoolAttrReleaseThreshold, &threshold);
    pool[dev] = p;
    return p;
  }
};
inline DevicePools& device_pools() { static DevicePools* p = new DevicePools(); return *p; }

inline void* pool_alloc(int dev, size_t bytes, cudaStream_t s) {
  void* p = nullptr;
  cudaMemPool_t pool = device_pools().get(dev);
  if (pool) ISFM_CUDA(cudaMallocFromPoolAsync(&p, bytes, pool, s));
  else ISFM_CUDA(cudaMallocAsync(&p, bytes, s));
  return p;
}

// Process-wide cache of large device blocks.  The pipeline creates solver after solver of (almost)
// the same size (three BA rounds, the retriangulation loop); handing the freed blocks of one
// handle straight to the next avoids the pool's sub-allocation of big free chunks by small
// requests, which fragments it and makes later handles map new physical memory -- measured as
// set-up stalls of up to 0.9 s on a 5 M-observation problem.  Blocks of >= 1 MB are kept (up to
// `limit` bytes per process, default 4 GB, ISFM_CACHE_LIMIT_MB / isfm_set_cache_limit; beyond it
// the largest go back to the pool) and re-used for requests of 88..100 % of their size.
// Stream safety: a block is filed under the device it was allocated on together with an event
// recorded on the stream that released it; whoever takes it makes its own stream wait on that
// event first, so a block handed to another handle / stream is never written while work of the
// previous owner is still in flight.  Blocks released during handle teardown (stream already
// synchronised) carry no event.
struct BufferCache {
  struct Block { void* p; cudaEvent_t ev; cudaStream_t stream; };
  std::mutex m;
  std::multimap<std::pair<int, size_t>, Block> blocks;   // (device, bytes) -> block
  size_t cached = 0;
  static constexpr size_t MIN_BYTES = (size_t)1 << 20;
  size_t limit = (size_t)4 << 30;
  bool enabled = getenv("ISFM_NO_BUFFER_CACHE") == nullptr;
  BufferCache() { if (const char* e = getenv("ISFM_CACHE_LIMIT_MB")) limit = (size_t)atoll(e) << 20; }
  void* take(int dev, size_t bytes, cudaStream_t s, size_t* got) {
    if (!enabled || bytes < MIN_BYTES) return nullptr;
    Block b;
    {
      std::lock_guard<std::mutex> lk(m);
      auto it = blocks.lower_bound(std::make_pair(dev, bytes));
      if (it == blocks.end() || it->first.first != dev || it->first.second > bytes + bytes / 8) return nullptr;
      b = it->second;
      *got = it->first.second;
      cached -= it->first.second;
      blocks.erase(it);
    }
    if (b.ev) {
      if (b.stream != s) cudaStreamWaitEvent(s, b.ev, 0);   // same stream: already ordered
      cudaEventDestroy(b.ev);
    }
    return b.p;
  }
  // `s` = stream the block was used on; `synced` = that stream is idle for good (no event needed)
  void give(int dev, void* p, size_t bytes, cudaStream_t s, bool synced) {
    if (enabled && bytes >= MIN_BYTES && limit > 0) {
      Block b{p, nullptr, s};
      if (!synced) {
        if (cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(b.ev, s) != cudaSuccess) {
          cudaGetLastError();
          if (b.ev) cudaEventDestroy(b.ev);
          cudaFreeAsync(p, s);
          return;
        }
      }
      std::vector<Block> evicted;
      {
        std::lock_guard<std::mutex> lk(m);
        blocks.emplace(std::make_pair(dev, bytes), b);
        cached += bytes;
        // over the limit: the largest blocks of THIS device go back to the pool
        while (cached > limit) {
          auto it = blocks.lower_bound(std::make_pair(dev + 1, (size_t)0));
          if (it == blocks.begin()) break;
          --it;
          if (it->first.first != dev) break;
          evicted.push_back(it->second);
          cached -= it->first.second;
          blocks.erase(it);
        }
      }
      for (auto& e : evicted) free_block(e, s);
      return;
    }
    if (synced) cudaFree(p); else cudaFreeAsync(p, s);
  }
  // frees a cached block of the CURRENT device on stream `s`, after the work of its last user
  static void free_block(const Block& b, cudaStream_t s) {
    if (b.ev) { if (b.stream != s) cudaStreamWaitEvent(s, b.ev, 0); cudaEventDestroy(b.ev); }
    cudaFreeAsync(b.p, s);
  }
  // every cached block back to its pool, the pools back to the driver (isfm_trim_cache)
  void trim() {
    std::multimap<std::pair<int, size_t>, Block> all;
    { std::lock_guard<std::mutex> lk(m); all.swap(blocks); cached = 0; }
    int prev = -1;
    cudaGetDevice(&prev);
    for (auto& kv : all) {
      cudaSetDevice(kv.first.first);
      if (kv.second.ev) { cudaEventSynchronize(kv.second.ev); cudaEventDestroy(kv.second.ev); }
      cudaFree(kv.second.p);
    }
    for (int d = 0; d < ISFM_MAX_DEVICES; ++d) {
      cudaMemPool_t p;
      { std::lock_guard<std::mutex> lk(device_pools().m); p = device_pools().pool[d]; }
      if (p) { cudaSetDevice(d); cudaDeviceSynchronize(); cudaMemPoolTrimTo(p, 0); }
    }
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
  }
};
inline BufferCache& buffer_cache() { static BufferCache* c = new BufferCache(); return *c; }   // never destroyed: outlives every handle

template <typename T>
struct DeviceBuffer {
  T* ptr = nullptr;
  size_t count = 0;
  size_t block_bytes = 0;   // size of the underlying block (>= count * sizeof(T) when it came from the cache)
  int dev = -1;             // device and stream the block was allocated under (the handle's)
  cudaStream_t stream = nullptr;
  DeviceBuffer() {}
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  ~DeviceBuffer() { release(); }
  void release() {
    if (ptr) {
      const DeviceCtx& c = tl_ctx();
      const bool synced = c.synced && c.dev == dev && c.stream == stream;
      int cur = -1;
      const bool switch_dev = cudaGetDevice(&cur) == cudaSuccess && cur != dev && dev >= 0;
      if (switch_dev) cudaSetDevice(dev);
      buffer_cache().give(dev, ptr, block_bytes, stream, synced);
      if (switch_dev) cudaSetDevice(cur);
    }
    ptr = nullptr; count = 0; block_bytes = 0;
  }
  void alloc(size_t n) {
    if (n <= count && ptr) return;
    release();
    if (n == 0) n = 1;
    const DeviceCtx& c = tl_ctx();
    dev = c.dev; stream = c.stream;
    if (dev < 0) { ISFM_CUDA(cudaGetDevice(&dev)); }
    size_t got = 0;
    if (void* p = buffer_cache().take(dev, n * sizeof(T), stream, &got)) {
      ptr = static_cast<T*>(p); block_bytes = got; count = n;
      return;
    }
    ptr = static_cast<T*>(pool_alloc(dev, n * sizeof(T), stream));
    count = n; block_bytes = n * sizeof(T);
  }
  void zero(cudaStream_t s) { if (ptr) ISFM_CUDA(cudaMemsetAsync(ptr, 0, count * sizeof(T), s)); }
  T* get() const { return ptr; }
  void swap(DeviceBuffer& o) {
    std::swap(ptr, o.ptr); std::swap(count, o.count); std::swap(block_bytes, o.block_bytes);
    std::swap(dev, o.dev); std::swap(stream, o.stream);
  }
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
