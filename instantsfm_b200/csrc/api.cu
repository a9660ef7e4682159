// api.cu -- extern "C" entry points declared in include/isfm_b200.h.
#include <cstdlib>

#include "ba_solver.cuh"
#include "gp_solver.cuh"

namespace isfm {
std::atomic<int64_t> g_launch_count{0};
static thread_local std::string tl_error;
void set_last_error(const std::string& msg) { tl_error = msg; }
const char* get_last_error() { return tl_error.c_str(); }
}  // namespace isfm

using namespace isfm;

// A handle is bound to the device that is current at creation and to one stream (the caller's,
// or an own blocking stream when the caller passes the legacy default stream, which cannot be
// captured into a CUDA graph).  Every entry point -- destroy included -- runs under a CtxGuard:
// the handle's device is made current for the call and device buffers are ordered on its stream.
struct HandleCtx {
  int dev = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  void init(void* caller_stream) {
    ISFM_CUDA(cudaGetDevice(&dev));
    stream = static_cast<cudaStream_t>(caller_stream);
    if (stream == nullptr) {
      ISFM_CUDA(cudaStreamCreate(&own_stream));
      stream = own_stream;
    }
  }
  template <typename Impl> void destroy(Impl*& impl) {
    CtxGuard g(dev, stream);
    cudaStreamSynchronize(stream);
    g.mark_synced();          // buffers released below need no stream ordering any more
    delete impl;
    impl = nullptr;
    if (own_stream) cudaStreamDestroy(own_stream);
    cudaGetLastError();
  }
};
struct isfm_ba { BASolverBase* impl = nullptr; HandleCtx ctx; };
struct isfm_gp { GPSolverBase* impl = nullptr; HandleCtx ctx; };
#define ISFM_GUARD(h) ISFM_REQUIRE(h && h->impl, ISFM_EINVAL, "null handle"); CtxGuard guard__(h->ctx.dev, h->ctx.stream)

#define ISFM_TRY try {
#define ISFM_CATCH                                                                         \
  return ISFM_OK; }                                                                        \
  catch (const IsfmError& e) { set_last_error(e.what()); return e.code; }                  \
  catch (const std::exception& e) { set_last_error(e.what()); return ISFM_ECUDA; }

static const char* kTimerNames[ISFM_N_TIMERS] = {
    "linearize", "point_blocks", "point_solve", "camera_blocks", "schur_offdiag", "precond", "pcg_spmv", "pcg_vec",
    "backsub", "update", "cost", "index_prep", "reduce", "comm", "misc", "coarse"};

extern "C" {

const char* isfm_version(void) { return "isfm_b200 0.1.0 (sm_100a)"; }
const char* isfm_last_error(void) { return get_last_error(); }
int64_t isfm_launch_count(void) { return g_launch_count.load(); }
void isfm_trim_cache(void) { buffer_cache().trim(); }
void isfm_set_cache_limit(uint64_t bytes) { std::lock_guard<std::mutex> lk(buffer_cache().m); buffer_cache().limit = (size_t)bytes; }
const char* isfm_timer_name(int32_t i) { return (i >= 0 && i < ISFM_N_TIMERS) ? kTimerNames[i] : ""; }

int isfm_partition_points(const int64_t* point_offsets, int64_t n_pt, int world, int64_t* part_begin_out) {
  ISFM_TRY
  ISFM_REQUIRE(point_offsets && part_begin_out && n_pt >= 0 && world >= 1, ISFM_EINVAL, "isfm_partition_points");
  const int64_t n_obs = point_offsets[n_pt];
  int64_t p = 0;
  for (int g = 0; g <= world; ++g) {
    // first point whose starting offset >= g * n_obs / world (integer arithmetic, exact)
    const __int128 target_num = (__int128)g * n_obs;
    while (p < n_pt && (__int128)point_offsets[p] * world < target_num) ++p;
    part_begin_out[g] = (g == world) ? n_pt : p;
  }
  ISFM_CATCH
}

void isfm_ba_default_desc(isfm_ba_desc* d) {
  if (!d) return;
  memset(d, 0, sizeof *d);
  d->dtype = 0; d->model_id = 3; d->optimize_poses = 1; d->reject = 30;
  d->huber_delta = 1.0; d->tr_radius = 1e4; d->tr_max = 1e10; d->tr_up = 2.0; d->tr_down = 0.0625;
  d->pcg_tol = 1e-6; d->pcg_max_iter = 0;
}

int isfm_ba_create(const isfm_ba_desc* desc, isfm_ba** out) {
  ISFM_TRY
  ISFM_REQUIRE(desc && out, ISFM_EINVAL, "isfm_ba_create: null argument");
  ISFM_REQUIRE(desc->dtype == 0 || desc->dtype == 1, ISFM_EINVAL, "dtype must be 0 (f32) or 1 (f64)");
  ISFM_REQUIRE(desc->huber_delta > 0 && desc->tr_radius > 0 && desc->pcg_tol > 0, ISFM_EINVAL, "bad optimiser options");
  if (model_n_intr(desc->model_id) < 0) throw IsfmError(ISFM_EUNSUPPORTED_MODEL, "Unsupported camera model");
  int dev_count = 0;
  ISFM_CUDA(cudaGetDeviceCount(&dev_count));
  ISFM_REQUIRE(dev_count > 0, ISFM_ECUDA, "no CUDA device: this library has no CPU path");
  isfm_ba* h = new isfm_ba();
  try {
    h->ctx.init(desc->stream);
    CtxGuard guard__(h->ctx.dev, h->ctx.stream);
    isfm_ba_desc d = *desc;
    d.stream = h->ctx.stream;
    h->impl = d.dtype == 0 ? make_ba_solver_f32(d) : make_ba_solver_f64(d);
  } catch (...) {
    if (h->ctx.own_stream) cudaStreamDestroy(h->ctx.own_stream);
    delete h;
    throw;
  }
  *out = h;
  ISFM_CATCH
}

void isfm_ba_destroy(isfm_ba* h) {
  if (!h) return;
  h->ctx.destroy(h->impl);
  delete h;
}

int isfm_ba_get_matvec_units(isfm_ba* h, int64_t* owned_out, int64_t* total_out) {
  ISFM_TRY
  ISFM_GUARD(h);
  if (owned_out) *owned_out = h->impl->matvec_units_owned;
  if (total_out) *total_out = h->impl->matvec_units_total;
  ISFM_CATCH
}

int isfm_ba_set_problem(isfm_ba* h, int64_t n_cam, int64_t n_pt, int64_t n_obs, const void* cam, const void* pp,
                        const void* pts, const void* obs, const int32_t* cam_idx, const int32_t* pt_idx) {
  ISFM_TRY
  ISFM_GUARD(h);
  h->impl->set_problem(n_cam, n_pt, n_obs, cam, pp, pts, obs, cam_idx, pt_idx);
  ISFM_CATCH
}

int isfm_ba_step(isfm_ba* h, double* loss_out, isfm_step_stats* stats) {
  ISFM_TRY
  ISFM_GUARD(h);
  h->impl->step(loss_out, stats);
  ISFM_CATCH
}

// bundle_adjustment.py:128-141
static bool should_stop(const double* hist, int n, double ftol, bool identical_test) {
  const int w = 4;
  if (n < 2 * w) return false;
  double recent = 0, prev = 0;
  for (int i = 0; i < w; ++i) { recent += hist[n - 1 - i]; prev += hist[n - 1 - w - i]; }
  recent /= w; prev /= w;
  double improvement = (prev - recent) / prev;
  if (std::fabs(improvement) < ftol) return true;
  return identical_test && hist[n - 1] == hist[n - 2];
}

int isfm_ba_solve(isfm_ba* h, int32_t max_iterations, double function_tolerance, double* hist, int32_t* n_out) {
  ISFM_TRY
  ISFM_REQUIRE(h && hist && n_out && max_iterations >= 0, ISFM_EINVAL, "isfm_ba_solve");
  ISFM_GUARD(h);
  int n = 0;
  for (int it = 0; it < max_iterations; ++it) {
    h->impl->step(&hist[n], nullptr);
    ++n;
    if (should_stop(hist, n, function_tolerance, true)) break;
  }
  *n_out = n;
  ISFM_CATCH
}

int isfm_ba_get_params(isfm_ba* h, void* cam_out, void* pts_out) {
  ISFM_TRY ISFM_GUARD(h); h->impl->get_params(cam_out, pts_out); ISFM_CATCH
}
int isfm_ba_set_params(isfm_ba* h, const void* cam, const void* pts) {
  ISFM_TRY ISFM_GUARD(h); h->impl->set_params(cam, pts); ISFM_CATCH
}
int isfm_ba_cost(isfm_ba* h, double* robust, double* sq) {
  ISFM_TRY ISFM_GUARD(h); h->impl->cost(robust, sq); ISFM_CATCH
}
int isfm_ba_get_structure(isfm_ba* h, int32_t* obs_perm, int64_t* point_offsets, int32_t* cam_perm, int64_t* cam_offsets) {
  ISFM_TRY ISFM_GUARD(h); h->impl->get_structure(obs_perm, point_offsets, cam_perm, cam_offsets); ISFM_CATCH
}
int isfm_ba_get_schur_pattern(isfm_ba* h, int64_t* nnzb, int64_t* n_pairs, int64_t* row_ptr, int32_t* col_idx) {
  ISFM_TRY ISFM_GUARD(h); h->impl->get_schur_pattern(nnzb, n_pairs, row_ptr, col_idx); ISFM_CATCH
}
int isfm_ba_debug_get(isfm_ba* h, int32_t what, void* dst) {
  ISFM_TRY ISFM_GUARD(h); h->impl->debug_get(what, dst); ISFM_CATCH
}
int isfm_ba_get_timers(isfm_ba* h, double ms_out[ISFM_N_TIMERS], int64_t launches_out[ISFM_N_TIMERS]) {
  ISFM_TRY
  ISFM_GUARD(h);
  h->impl->timers.resolve();
  for (int i = 0; i < ISFM_N_TIMERS; ++i) {
    if (ms_out) ms_out[i] = h->impl->timers.ms[i];
    if (launches_out) launches_out[i] = h->impl->timers.launches[i];
  }
  ISFM_CATCH
}
int isfm_ba_get_pcg_phases(isfm_ba* h, double ms_out[8], int64_t* solves_out, int32_t* two_level_out) {
  ISFM_TRY ISFM_GUARD(h); h->impl->get_pcg_phases(ms_out, solves_out, two_level_out); ISFM_CATCH
}
int isfm_ba_reset_timers(isfm_ba* h, int32_t enable) {
  ISFM_TRY ISFM_GUARD(h); h->impl->timers.reset(enable != 0); ISFM_CATCH
}

// ---------------------------------------------------------------------------------------
void isfm_gp_default_desc(isfm_gp_desc* d) {
  if (!d) return;
  memset(d, 0, sizeof *d);
  d->dtype = 0; d->reject = 30; d->huber_delta = 0.1; d->tr_radius = 1e3; d->tr_max = 1e8; d->tr_up = 2.0;
  d->tr_down = 0.0625; d->pcg_tol = 1e-6; d->pcg_max_iter = 0; d->optimize_scales = 1;
}

int isfm_gp_create(const isfm_gp_desc* desc, isfm_gp** out) {
  ISFM_TRY
  ISFM_REQUIRE(desc && out, ISFM_EINVAL, "isfm_gp_create: null argument");
  ISFM_REQUIRE(desc->dtype == 0 || desc->dtype == 1, ISFM_EINVAL, "dtype must be 0 (f32) or 1 (f64)");
  int dev_count = 0;
  ISFM_CUDA(cudaGetDeviceCount(&dev_count));
  ISFM_REQUIRE(dev_count > 0, ISFM_ECUDA, "no CUDA device: this library has no CPU path");
  isfm_gp* h = new isfm_gp();
  try {
    h->ctx.init(desc->stream);
    CtxGuard guard__(h->ctx.dev, h->ctx.stream);
    isfm_gp_desc d = *desc;
    d.stream = h->ctx.stream;
    h->impl = d.dtype == 0 ? make_gp_solver_f32(d) : make_gp_solver_f64(d);
  } catch (...) {
    if (h->ctx.own_stream) cudaStreamDestroy(h->ctx.own_stream);
    delete h;
    throw;
  }
  *out = h;
  ISFM_CATCH
}
void isfm_gp_destroy(isfm_gp* h) {
  if (!h) return;
  h->ctx.destroy(h->impl);
  delete h;
}
int isfm_gp_set_problem(isfm_gp* h, int64_t n_cam, int64_t n_pt, int64_t n_obs, const void* centres, const void* pts,
                        const void* scales, const void* rays, const int32_t* cam_idx, const int32_t* pt_idx,
                        const uint8_t* is_calibrated, const uint8_t* scale_fixed) {
  ISFM_TRY
  ISFM_GUARD(h);
  h->impl->set_problem(n_cam, n_pt, n_obs, centres, pts, scales, rays, cam_idx, pt_idx, is_calibrated, scale_fixed);
  ISFM_CATCH
}
int isfm_gp_step(isfm_gp* h, double* loss_out, isfm_step_stats* stats) {
  ISFM_TRY ISFM_GUARD(h); h->impl->step(loss_out, stats); ISFM_CATCH
}
int isfm_gp_solve(isfm_gp* h, int32_t max_iterations, double function_tolerance, double* hist, int32_t* n_out) {
  ISFM_TRY
  ISFM_REQUIRE(h && hist && n_out && max_iterations >= 0, ISFM_EINVAL, "isfm_gp_solve");
  ISFM_GUARD(h);
  int n = 0;
  for (int it = 0; it < max_iterations; ++it) {
    h->impl->step(&hist[n], nullptr);
    ++n;
    if (should_stop(hist, n, function_tolerance, false)) break;  // global_positioning.py:178-183
  }
  *n_out = n;
  ISFM_CATCH
}
int isfm_gp_get_params(isfm_gp* h, void* centres_out, void* pts_out, void* scales_out) {
  ISFM_TRY ISFM_GUARD(h); h->impl->get_params(centres_out, pts_out, scales_out); ISFM_CATCH
}
int isfm_gp_cost(isfm_gp* h, double* robust, double* sq) {
  ISFM_TRY ISFM_GUARD(h); h->impl->cost(robust, sq); ISFM_CATCH
}
int isfm_gp_get_timers(isfm_gp* h, double ms_out[ISFM_N_TIMERS], int64_t launches_out[ISFM_N_TIMERS]) {
  ISFM_TRY
  ISFM_GUARD(h);
  h->impl->timers.resolve();
  for (int i = 0; i < ISFM_N_TIMERS; ++i) {
    if (ms_out) ms_out[i] = h->impl->timers.ms[i];
    if (launches_out) launches_out[i] = h->impl->timers.launches[i];
  }
  ISFM_CATCH
}
int isfm_gp_reset_timers(isfm_gp* h, int32_t enable) {
  ISFM_TRY ISFM_GUARD(h); h->impl->timers.reset(enable != 0); ISFM_CATCH
}

}  // extern "C"
