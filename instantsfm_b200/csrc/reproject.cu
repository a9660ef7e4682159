// reproject.cu -- batched reprojection test of candidate observations:
// the arithmetic of complete_tracks (instantsfm/processors/track_retriangulation.py:81-91),
//   y = rotate_quat(X, cam[:7]);  valid = y.z > EPSILON;
//   e = || reproject_<model>(X, cam, pp) - observed ||;  pass = (e <= threshold) & valid,
// with the same nine camera models as K1 (math.cuh) in fp64 like the reference's torch.float64.
// One thread per candidate; HBM-bound gather: observed 16 + indices 8 read, 1 (+8) written, the
// camera row (<= 19 doubles) and the point (24 B) come from L2.
#include "common.cuh"
#include "math.cuh"

namespace isfm {
namespace {

constexpr int RT_TPB = 256;

template <int MODEL>
__global__ void __launch_bounds__(RT_TPB)
reprojection_test_kernel(int64_t n_obs, const double* __restrict__ cam, const double* __restrict__ pp,
                         const double* __restrict__ pts, const double* __restrict__ obs, const int32_t* __restrict__ cam_idx,
                         const int32_t* __restrict__ pt_idx, double max_error, double min_depth, uint8_t* __restrict__ pass_out,
                         double* __restrict__ err_out) {
  constexpr int CW = 7 + ModelTraits<MODEL>::NI;
  for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < n_obs; a += (int64_t)gridDim.x * blockDim.x) {
    const int c = cam_idx[a], p = pt_idx[a];
    double cr[CW], ppv[2], X[3], o[2], r[2], R[9], y[3];
#pragma unroll
    for (int i = 0; i < CW; ++i) cr[i] = __ldg(cam + (size_t)c * CW + i);
    ppv[0] = __ldg(pp + 2 * (size_t)c); ppv[1] = __ldg(pp + 2 * (size_t)c + 1);
#pragma unroll
    for (int i = 0; i < 3; ++i) X[i] = __ldg(pts + 3 * (size_t)p + i);
    o[0] = obs[2 * a]; o[1] = obs[2 * a + 1];
    transform_point(cr, X, R, y);
    ba_residual<MODEL, double>(cr, ppv, X, o, r);
    const double e = sqrt(r[0] * r[0] + r[1] * r[1]);
    pass_out[a] = (e <= max_error && y[2] > min_depth) ? 1 : 0;   // NaN error (point on the camera plane) fails
    if (err_out) err_out[a] = e;
  }
}

template <typename U>
const U* stage_in(DeviceBuffer<U>& buf, const U* src, size_t count, cudaStream_t s) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeDevice) return src;
  cudaGetLastError();
  buf.alloc(count);
  ISFM_CUDA(cudaMemcpyAsync(buf.get(), src, count * sizeof(U), cudaMemcpyDefault, s));
  return buf.get();
}

template <typename U>
bool is_device(const U* p) {
  cudaPointerAttributes at;
  const bool dev = cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  return dev;
}

}  // namespace
}  // namespace isfm

using namespace isfm;

extern "C" int isfm_reprojection_test(int32_t model_id, int64_t n_obs, int64_t n_cam, int64_t n_pt, const double* cam,
                                      const double* pp, const double* pts, const double* obs, const int32_t* cam_idx,
                                      const int32_t* pt_idx, double max_error, double min_depth, uint8_t* pass_out,
                                      double* err_out, void* stream) {
  try {
    const int ni = model_n_intr(model_id);
    if (ni < 0) throw IsfmError(ISFM_EUNSUPPORTED_MODEL, "Unsupported camera model");
    ISFM_REQUIRE(n_obs >= 0 && n_cam >= 0 && n_pt >= 0, ISFM_EINVAL, "isfm_reprojection_test: sizes");
    if (n_obs == 0) return ISFM_OK;
    ISFM_REQUIRE(cam && pp && pts && obs && cam_idx && pt_idx && pass_out, ISFM_EINVAL, "isfm_reprojection_test: null");
    int n_dev = 0;
    ISFM_CUDA(cudaGetDeviceCount(&n_dev));
    ISFM_REQUIRE(n_dev > 0, ISFM_ECUDA, "no CUDA device: this library has no CPU path");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
  CtxGuard guard__(-1, s);   // buffers below are ordered on the caller's stream, on the current device
    DeviceBuffer<double> b_cam, b_pp, b_pts, b_obs, b_err; DeviceBuffer<int32_t> b_ci, b_pi; DeviceBuffer<uint8_t> b_pass;
    const double* d_cam = stage_in(b_cam, cam, (size_t)n_cam * (7 + ni), s);
    const double* d_pp = stage_in(b_pp, pp, (size_t)n_cam * 2, s);
    const double* d_pts = stage_in(b_pts, pts, (size_t)n_pt * 3, s);
    const double* d_obs = stage_in(b_obs, obs, (size_t)n_obs * 2, s);
    const int32_t* d_ci = stage_in(b_ci, cam_idx, (size_t)n_obs, s);
    const int32_t* d_pi = stage_in(b_pi, pt_idx, (size_t)n_obs, s);
    const bool pass_dev = is_device(pass_out), err_dev = err_out && is_device(err_out);
    uint8_t* d_pass = pass_out;
    double* d_err = err_out;
    if (!pass_dev) { b_pass.alloc((size_t)n_obs); d_pass = b_pass.get(); }
    if (err_out && !err_dev) { b_err.alloc((size_t)n_obs); d_err = b_err.get(); }
    const int grid = (int)std::min<int64_t>(div_up(n_obs, RT_TPB), 148 * 16);
    g_launch_count++;
#define ISFM_RT(M) reprojection_test_kernel<M><<<grid, RT_TPB, 0, s>>>(n_obs, d_cam, d_pp, d_pts, d_obs, d_ci, d_pi, max_error, min_depth, d_pass, d_err)
    switch (model_id) {
      case 0: ISFM_RT(0); break; case 1: ISFM_RT(1); break; case 2: ISFM_RT(2); break; case 3: ISFM_RT(3); break;
      case 4: ISFM_RT(4); break; case 5: ISFM_RT(5); break; case 6: ISFM_RT(6); break; case 8: ISFM_RT(8); break;
      case 9: ISFM_RT(9); break;
      default: throw IsfmError(ISFM_EUNSUPPORTED_MODEL, "Unsupported camera model");
    }
#undef ISFM_RT
    ISFM_CUDA(cudaGetLastError());
    if (!pass_dev) ISFM_CUDA(cudaMemcpyAsync(pass_out, d_pass, (size_t)n_obs, cudaMemcpyDeviceToHost, s));
    if (err_out && !err_dev) ISFM_CUDA(cudaMemcpyAsync(err_out, d_err, (size_t)n_obs * sizeof(double), cudaMemcpyDeviceToHost, s));
    ISFM_CUDA(cudaStreamSynchronize(s));
    return ISFM_OK;
  } catch (const IsfmError& e) { set_last_error(e.what()); return e.code; }
  catch (const std::exception& e) { set_last_error(e.what()); return ISFM_ECUDA; }
}
