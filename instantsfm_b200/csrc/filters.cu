// filters.cu -- the inter-BA track filters of instantsfm/processors/track_filter.py as streaming
// kernels.  The reference computes in numpy fp64; these kernels compute in
// fp64 with the reference's order of operations and WITHOUT fused multiply-adds (explicit
// __dmul_rn / __dadd_rn), so that a comparison against a threshold flips only where numpy's own
// einsum / matmul summation order would.  All three are HBM-bound gathers:
//   per observation: image id 4 + track index 4 + bearing 24 read, 1 byte written; the 3x4 pose
//   (96 B / image) and the point (24 B / track) come from L2.
#include "common.cuh"
#include "filter_math.cuh"

namespace isfm {
namespace {

constexpr int FILTER_TPB = 256;

// MODE 0 / 1: see filter_math.cuh::filter_keep
template <int MODE>
__global__ void __launch_bounds__(FILTER_TPB)
filter_observations_kernel(int64_t n_obs, const double* __restrict__ world2cam, const double* __restrict__ xyz,
                           const double* __restrict__ feat, const int32_t* __restrict__ image_ids,
                           const int32_t* __restrict__ track_idx, double thr, uint8_t* __restrict__ valid_out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_obs; i += (int64_t)gridDim.x * blockDim.x) {
    const double* __restrict__ M = world2cam + (size_t)image_ids[i] * 16;
    const double* __restrict__ X = xyz + (size_t)track_idx[i] * 3;
    valid_out[i] = filter_keep<MODE>(M, X[0], X[1], X[2], feat[3 * i], feat[3 * i + 1], feat[3 * i + 2], thr) ? 1 : 0;
  }
}

// FilterTracksTriangulationAngle (track_filter.py:116-137): a track is removed iff every pair of
// its viewing directions (point - camera centre, normalised with norm + EPS) has dot product >
// cos(min_angle) -- including each direction with itself, whose dot product n^2/(n+EPS)^2 is
// compared like any other, so an empty track (np.all of nothing) and a single-view track are
// removed.  The reference de-duplicates image ids with np.unique first; duplicates only repeat
// dot products that are already in the matrix, so the verdict is the same without it.
// One warp per track; directions staged in shared memory in chunks, all pairs tested.
constexpr int TRI_WARPS = 8;
constexpr int TRI_MAXK = 64;   // directions held in shared memory per warp; longer tracks stream the second operand

__global__ void __launch_bounds__(TRI_WARPS * 32)
filter_triangulation_kernel(int64_t n_trk, const int64_t* __restrict__ track_off, const int32_t* __restrict__ image_ids,
                            const double* __restrict__ centers, const double* __restrict__ xyz, double thr,
                            uint8_t* __restrict__ remove_out) {
  __shared__ double dir[TRI_WARPS][TRI_MAXK][3];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t t = blockIdx.x * (int64_t)TRI_WARPS + w; t < n_trk; t += (int64_t)gridDim.x * TRI_WARPS) {
    const int64_t beg = track_off[t];
    const int k = (int)(track_off[t + 1] - beg);
    const double X0 = xyz[3 * t], X1 = xyz[3 * t + 1], X2 = xyz[3 * t + 2];
    auto direction = [&](int a, double* d) {
      const double* c = centers + (size_t)image_ids[beg + a] * 3;
      const double v0 = __dadd_rn(X0, -c[0]), v1 = __dadd_rn(X1, -c[1]), v2 = __dadd_rn(X2, -c[2]);
      const double n = __dadd_rn(sqrt(dot3_nofma(v0, v1, v2, v0, v1, v2)), FILTER_EPS);
      d[0] = __ddiv_rn(v0, n); d[1] = __ddiv_rn(v1, n); d[2] = __ddiv_rn(v2, n);
    };
    bool all_small = true;   // every tested pair so far has dot > thr
    for (int a0 = 0; a0 < k && all_small; a0 += TRI_MAXK) {
      const int na = min(TRI_MAXK, k - a0);
      __syncwarp();
      for (int a = lane; a < na; a += 32) direction(a0 + a, dir[w][a]);
      __syncwarp();
      // pairs inside the chunk (a <= b covers the diagonal too)
      for (int idx = lane; idx < na * na; idx += 32) {
        const int a = idx / na, b = idx % na;
        if (a <= b && !(dot3_nofma(dir[w][a][0], dir[w][a][1], dir[w][a][2], dir[w][b][0], dir[w][b][1], dir[w][b][2]) > thr))
          all_small = false;
      }
      // pairs of this chunk with later directions
      for (int b = a0 + na + lane; b < k; b += 32) {
        double d[3];
        direction(b, d);
        for (int a = 0; a < na; ++a)
          if (!(dot3_nofma(dir[w][a][0], dir[w][a][1], dir[w][a][2], d[0], d[1], d[2]) > thr)) all_small = false;
      }
      all_small = __all_sync(0xffffffffu, all_small);
    }
    if (lane == 0) remove_out[t] = all_small ? 1 : 0;
  }
}

// Host or device pointer -> device pointer on stream s (staged copy when it is host memory).
template <typename U>
struct Staged {
  DeviceBuffer<U> buf;
  const U* ptr = nullptr;
  void in(const U* src, size_t count, cudaStream_t s) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeDevice) { ptr = src; return; }
    cudaGetLastError();
    buf.alloc(count);
    ISFM_CUDA(cudaMemcpyAsync(buf.get(), src, count * sizeof(U), cudaMemcpyDefault, s));
    ptr = buf.get();
  }
};

void require_device() {
  int n = 0;
  ISFM_CUDA(cudaGetDeviceCount(&n));
  ISFM_REQUIRE(n > 0, ISFM_ECUDA, "no CUDA device: this library has no CPU path");
}

}  // namespace
}  // namespace isfm

using namespace isfm;

#define ISFM_TRY try {
#define ISFM_CATCH                                                                         \
  return ISFM_OK; }                                                                        \
  catch (const IsfmError& e) { set_last_error(e.what()); return e.code; }                  \
  catch (const std::exception& e) { set_last_error(e.what()); return ISFM_ECUDA; }

extern "C" int isfm_filter_observations(int32_t mode, int64_t n_obs, int64_t n_img, int64_t n_trk, const double* world2cam,
                                        const double* xyz, const double* features_undist, const int32_t* image_ids,
                                        const int32_t* track_idx, double threshold, uint8_t* valid_out, void* stream) {
  ISFM_TRY
  ISFM_REQUIRE(mode == ISFM_FILTER_ANGLE || mode == ISFM_FILTER_REPROJECTION_NORMALIZED, ISFM_EINVAL, "isfm_filter_observations: mode");
  ISFM_REQUIRE(n_obs >= 0 && n_img >= 0 && n_trk >= 0, ISFM_EINVAL, "isfm_filter_observations: sizes");
  if (n_obs == 0) return ISFM_OK;
  ISFM_REQUIRE(world2cam && xyz && features_undist && image_ids && track_idx && valid_out, ISFM_EINVAL, "isfm_filter_observations: null");
  require_device();
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CtxGuard guard__(-1, s);   // buffers below are ordered on the caller's stream, on the current device
  Staged<double> M, X, F; Staged<int32_t> I, T;
  M.in(world2cam, (size_t)n_img * 16, s); X.in(xyz, (size_t)n_trk * 3, s); F.in(features_undist, (size_t)n_obs * 3, s);
  I.in(image_ids, (size_t)n_obs, s); T.in(track_idx, (size_t)n_obs, s);
  cudaPointerAttributes at;
  const bool out_dev = cudaPointerGetAttributes(&at, valid_out) == cudaSuccess && at.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  DeviceBuffer<uint8_t> out_buf;
  uint8_t* out = valid_out;
  if (!out_dev) { out_buf.alloc((size_t)n_obs); out = out_buf.get(); }
  const int grid = (int)std::min<int64_t>(div_up(n_obs, FILTER_TPB), 148 * 16);
  g_launch_count++;
  if (mode == ISFM_FILTER_ANGLE)
    filter_observations_kernel<0><<<grid, FILTER_TPB, 0, s>>>(n_obs, M.ptr, X.ptr, F.ptr, I.ptr, T.ptr, threshold, out);
  else
    filter_observations_kernel<1><<<grid, FILTER_TPB, 0, s>>>(n_obs, M.ptr, X.ptr, F.ptr, I.ptr, T.ptr, threshold, out);
  ISFM_CUDA(cudaGetLastError());
  if (!out_dev) ISFM_CUDA(cudaMemcpyAsync(valid_out, out, (size_t)n_obs, cudaMemcpyDeviceToHost, s));
  ISFM_CUDA(cudaStreamSynchronize(s));
  ISFM_CATCH
}

extern "C" int isfm_filter_triangulation_angle(int64_t n_trk, int64_t n_obs, int64_t n_img, const int64_t* track_off,
                                               const int32_t* image_ids, const double* centers, const double* xyz,
                                               double cos_threshold, uint8_t* remove_out, void* stream) {
  ISFM_TRY
  ISFM_REQUIRE(n_trk >= 0 && n_obs >= 0 && n_img >= 0, ISFM_EINVAL, "isfm_filter_triangulation_angle: sizes");
  if (n_trk == 0) return ISFM_OK;
  ISFM_REQUIRE(track_off && centers && xyz && remove_out && (image_ids || n_obs == 0), ISFM_EINVAL, "isfm_filter_triangulation_angle: null");
  require_device();
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CtxGuard guard__(-1, s);   // buffers below are ordered on the caller's stream, on the current device
  Staged<int64_t> O; Staged<int32_t> I; Staged<double> C, X;
  O.in(track_off, (size_t)n_trk + 1, s); I.in(image_ids, (size_t)std::max<int64_t>(n_obs, 1), s);
  C.in(centers, (size_t)std::max<int64_t>(n_img, 1) * 3, s); X.in(xyz, (size_t)n_trk * 3, s);
  cudaPointerAttributes at;
  const bool out_dev = cudaPointerGetAttributes(&at, remove_out) == cudaSuccess && at.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  DeviceBuffer<uint8_t> out_buf;
  uint8_t* out = remove_out;
  if (!out_dev) { out_buf.alloc((size_t)n_trk); out = out_buf.get(); }
  const int grid = (int)std::min<int64_t>(div_up(n_trk, TRI_WARPS), 148 * 8);
  g_launch_count++;
  filter_triangulation_kernel<<<grid, TRI_WARPS * 32, 0, s>>>(n_trk, O.ptr, I.ptr, C.ptr, X.ptr, cos_threshold, out);
  ISFM_CUDA(cudaGetLastError());
  if (!out_dev) ISFM_CUDA(cudaMemcpyAsync(remove_out, out, (size_t)n_trk, cudaMemcpyDeviceToHost, s));
  ISFM_CUDA(cudaStreamSynchronize(s));
  ISFM_CATCH
}
