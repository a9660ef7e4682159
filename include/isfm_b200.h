/* isfm_b200.h -- C ABI of the B200-native InstantSfM bundle-adjustment / global-positioning
 * Levenberg-Marquardt path.
 *
 * The reference (mrcabellom/InstantSfM) is pure Python and has no FFI of its own for this
 * path: TorchBA.Solve / TorchGP.Optimize build torch tensors and loop
 * `bae.optim.LM.step(input)` (instantsfm/processors/bundle_adjustment.py:115-142,
 * instantsfm/processors/global_positioning.py:154-184).  Each entry point below names the
 * reference interface it replaces.  Signatures carry plain pointers and sizes only; every
 * data pointer may be a HOST or a DEVICE pointer (copies use cudaMemcpyDefault), so the
 * reference-side binding can hand over numpy arrays or torch.Tensor.data_ptr() alike.
 *
 * Scalar type: `dtype` 0 = float32 (product), 1 = float64 (validation build of the same
 * kernels).  All `const void*` parameter/observation arrays are of that type.
 * Index arrays are int32, as in the reference (bundle_adjustment.py:99-100).
 *
 * Threading: one host thread per handle; a handle is bound to the CUDA device that is
 * current when it is created and issues all work on `stream` (0 = legacy default stream).
 * Errors: 0 on success, negative isfm_status otherwise; isfm_last_error() returns a
 * thread-local message.
 */
#ifndef ISFM_B200_H_
#define ISFM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

typedef enum isfm_status {
  ISFM_OK = 0,
  ISFM_EINVAL = -1,            /* bad argument / inconsistent sizes                        */
  ISFM_EUNSUPPORTED_MODEL = -2,/* FOV / THIN_PRISM_FISHEYE: NotImplementedError in the
                                  reference (utils/cost_function.py:128,182;
                                  bundle_adjustment.py:47-50)                              */
  ISFM_ECUDA = -3,             /* CUDA runtime failure (message has the call site)         */
  ISFM_ENCCL = -4,             /* NCCL failure or NCCL not loadable                        */
  ISFM_ENONFINITE = -5,        /* non-finite cost or PCG breakdown                         */
  ISFM_ESTATE = -6             /* call order violated (e.g. step before set_problem)       */
} isfm_status;

typedef struct isfm_comm isfm_comm; /* NCCL communicator wrapper, one per rank (process) */
typedef struct isfm_ba isfm_ba;     /* bundle-adjustment handle                          */
typedef struct isfm_gp isfm_gp;     /* global-positioning handle                         */

/* Per-step record: what bae.optim.LM.step leaves in optimizer state (loss, damping) plus
 * what the reference only prints. */
typedef struct isfm_step_stats {
  double loss_before;   /* `self.last`                                                   */
  double loss;          /* returned by LM.step (bundle_adjustment.py:132)                */
  double damping;       /* TrustRegion damping 1/radius AFTER the step                   */
  double quality;       /* trust-region gain ratio of the last trial                     */
  double model_term;    /* -(JD)^T (2R + JD) of the last trial                           */
  double step_norm_cam; /* ||D_c||_2 of the last trial                                   */
  int32_t trials;       /* damped solves in this step (1 + rejects)                      */
  int32_t rejects;
  int32_t pcg_iters;    /* total PCG iterations over all trials                          */
  int32_t accepted;     /* 1 if the final trial was kept                                 */
  int32_t pcg_status;   /* last trial's PCG: 1 converged (true residual b - S x verified
                           below pcg_tol), 0 hit pcg_max_iter, 2 breakdown (non-positive
                           curvature / non-finite: x holds the last good iterate; bae
                           prints and breaks, bundle_adjustment.py:132), 4 the true
                           residual stagnated above pcg_tol (floating-point floor)        */
  int32_t reserved;
} isfm_step_stats;

/* ------------------------------------------------------------------------------------ */
/* library                                                                                */
/* ------------------------------------------------------------------------------------ */
const char* isfm_version(void);
const char* isfm_last_error(void);
/* number of kernel launches issued by this library in this process (bench: gpu_launches) */
int64_t isfm_launch_count(void);
/* Device memory: the library allocates from a PRIVATE stream-ordered pool per device (never the
 * device's default pool, which torch's allocator may share) and keeps freed blocks >= 1 MB in a
 * process-wide cache for the next handle (the pipeline creates solver after solver of the same
 * size).  isfm_trim_cache() returns everything to the driver (call it at stage boundaries when
 * other GPU stages need the memory); isfm_set_cache_limit() bounds the cache (default 4 GB,
 * also ISFM_CACHE_LIMIT_MB; 0 disables caching).                                             */
void isfm_trim_cache(void);
void isfm_set_cache_limit(uint64_t bytes);

/* ------------------------------------------------------------------------------------ */
/* communicator (multi-GPU; one process per GPU).  No reference counterpart: the           */
/* reference BA/GP is single-GPU (device="cuda:0", bundle_adjustment.py:39).               */
/* ------------------------------------------------------------------------------------ */
/* Writes a 128-byte ncclUniqueId; rank 0 calls it and broadcasts the bytes.              */
int isfm_comm_unique_id(uint8_t id_out[128]);
int isfm_comm_create(const uint8_t id[128], int rank, int world, isfm_comm** out);
void isfm_comm_destroy(isfm_comm* comm);
/* Transport of the per-PCG-iteration sum: 1 = peer-memory exchange over NVLink (CUDA IPC,   */
/* fused into the producing / consuming kernels), 0 = ncclAllReduce (not set up yet, or IPC  */
/* unavailable, or ISFM_NO_PEER set).  Valid after the first isfm_*_set_problem on the comm. */
int isfm_comm_peer_enabled(const isfm_comm* comm);

/* ------------------------------------------------------------------------------------ */
/* integer prep on host arrays (bit-exact contract)                          */
/* ------------------------------------------------------------------------------------ */
/* Contiguous point ranges balanced by observation count.  point_offsets[n_pt+1] is the   */
/* CSR of observations per point; part_begin_out[world+1] receives point boundaries:      */
/* boundary g = first point whose starting observation offset >= g * n_obs / world.       */
int isfm_partition_points(const int64_t* point_offsets, int64_t n_pt, int world,
                          int64_t* part_begin_out);

/* ------------------------------------------------------------------------------------ */
/* bundle adjustment  (replaces the body of TorchBA.Solve below the tensor set-up)         */
/* ------------------------------------------------------------------------------------ */
typedef struct isfm_ba_desc {
  int32_t dtype;          /* 0 f32, 1 f64                                                 */
  int32_t model_id;       /* CameraModelId.value, scene/defs.py:101-113                   */
  int32_t optimize_poses; /* BUNDLE_ADJUSTER_OPTIONS['optimize_poses'], b_a.py:55         */
  int32_t reject;         /* LM(reject=30), bundle_adjustment.py:119                      */
  double huber_delta;     /* Huber(thres_loss_function), bundle_adjustment.py:118         */
  double tr_radius;       /* TrustRegion(radius=1e4, max=1e10, up=2, down=0.5**4) :116    */
  double tr_max;
  double tr_up;
  double tr_down;
  double pcg_tol;         /* PCG(tol=1e-5), bundle_adjustment.py:117.  Default here 1e-6:
                             the reference's 1e-5 is on the full Jacobi-scaled system; on
                             the reduced camera system 1e-6 keeps the first (largest) LM
                             step within 1e-4 of the reference's cost                     */
  int32_t pcg_max_iter;   /* 0 = default (10 * n_cam * d, as bae's PCG: 10 n)             */
  int32_t reserved;
  void* stream;           /* cudaStream_t                                                 */
  isfm_comm* comm;        /* NULL = single GPU                                            */
} isfm_ba_desc;

void isfm_ba_default_desc(isfm_ba_desc* desc);
int isfm_ba_create(const isfm_ba_desc* desc, isfm_ba** out);
void isfm_ba_destroy(isfm_ba* h);

/* Tensors of bundle_adjustment.py:111-126 in the compacted index space:
 *   cam  [n_cam, 7 + n_intr]  = [t, q_xyzw, intrinsics without pp]   (model.pose)
 *   pp   [n_cam, 2]           (input["camera_pps"])
 *   pts  [n_pt, 3]            (model.points_3d)
 *   obs  [n_obs, 2]           (input["points_2d"])
 *   cam_idx, pt_idx int32 [n_obs] (input["camera_indices"], input["point_indices"])
 * Observations may come in any order; the library stable-sorts them by point.  With a
 * communicator, cam/pp are the full replicated set and pts/obs this rank's shard.        */
int isfm_ba_set_problem(isfm_ba* h, int64_t n_cam, int64_t n_pt, int64_t n_obs,
                        const void* cam, const void* pp, const void* pts, const void* obs,
                        const int32_t* cam_idx, const int32_t* pt_idx);

/* One `optimizer.step(input)` (bundle_adjustment.py:132).  *loss_out = returned loss.     */
int isfm_ba_step(isfm_ba* h, double* loss_out, isfm_step_stats* stats /* may be NULL */);

/* The loop of bundle_adjustment.py:128-141 (window 4, function_tolerance, identical-loss
 * stop).  loss_history_out[max_iterations]; *n_iterations_out = steps actually taken.    */
int isfm_ba_solve(isfm_ba* h, int32_t max_iterations, double function_tolerance,
                  double* loss_history_out, int32_t* n_iterations_out);

/* Current parameters in the caller's original order (what update() reads back through the
 * aliased tensors, bundle_adjustment.py:18-36).  Either pointer may be NULL.             */
int isfm_ba_get_params(isfm_ba* h, void* cam_out, void* pts_out);
/* Overwrite current parameters (same layout as set_problem); resets the cached loss.     */
int isfm_ba_set_params(isfm_ba* h, const void* cam, const void* pts);

/* Robust cost sum rho(||r||^2) and plain sum ||r||^2 at the current parameters.           */
int isfm_ba_cost(isfm_ba* h, double* robust_cost_out, double* sq_cost_out);

/* Integer structure for bit-exact tests (any pointer may be NULL):
 *   obs_perm[n_obs]      stable argsort of pt_idx (position -> original observation)
 *   point_offsets[n_pt+1] CSR of sorted observations by point
 *   cam_perm[n_obs]      stable argsort of the point-sorted cam_idx
 *   cam_offsets[n_cam+1] CSR by camera                                                    */
int isfm_ba_get_structure(isfm_ba* h, int32_t* obs_perm, int64_t* point_offsets,
                          int32_t* cam_perm, int64_t* cam_offsets);
/* Reduced-camera-system block pattern (BSR, both triangles, row-major by block):
 * *nnzb_out blocks; row_ptr[n_cam+1], col_idx[nnzb] may be NULL to query the size first.
 * *n_pairs_out = number of (a, b) observation pairs feeding the off-diagonal blocks.     */
int isfm_ba_get_schur_pattern(isfm_ba* h, int64_t* nnzb_out, int64_t* n_pairs_out,
                              int64_t* row_ptr, int32_t* col_idx);

/* Mat-vec work units of the reduced camera system this rank multiplies per PCG iteration and
 * their total: equal unless the ranks share one block pattern and the summed matrix has been
 * split across them ("split mat-vec").                                */
int isfm_ba_get_matvec_units(isfm_ba* h, int64_t* owned_out, int64_t* total_out);

/* Kernel-level outputs for parity tests, in the ORIGINAL observation / point / camera
 * order.  `what` selects the buffer; `dst` must hold the documented element count of the
 * handle's dtype.  Runs the producing kernels at the current parameters and damping.      */
typedef enum isfm_ba_debug {
  ISFM_BA_RESIDUALS = 0,     /* [n_obs, 2]  unweighted r = proj - obs                     */
  ISFM_BA_JAC_CAM = 1,       /* [n_obs, 2, d] Triggs-weighted, d = 6 + n_intr            */
  ISFM_BA_JAC_POINT = 2,     /* [n_obs, 2, 3] Triggs-weighted                             */
  ISFM_BA_WEIGHTED_RES = 3,  /* [n_obs, 2]  Triggs-weighted residual                      */
  ISFM_BA_HPP = 4,           /* [n_pt, 6]   upper triangle xx xy xz yy yz zz, undamped    */
  ISFM_BA_GP = 5,            /* [n_pt, 3]   J_p^T R                                       */
  ISFM_BA_HCC = 6,           /* [n_cam, d, d] undamped (all ranks' sum)                   */
  ISFM_BA_GC = 7,            /* [n_cam, d]                                                */
  ISFM_BA_SCHUR_DENSE = 8,   /* [n_cam*d, n_cam*d] dense reduced system S at the current
                                damping (small problems only)                             */
  ISFM_BA_SCHUR_RHS = 9,     /* [n_cam*d]  b = -(g_c - Hcp Hpp^-1 g_p)                    */
  ISFM_BA_STEP_CAM = 10,     /* [n_cam, d]  D_c of the last trial                         */
  ISFM_BA_STEP_POINT = 11    /* [n_pt, 3]   D_p of the last trial                         */
} isfm_ba_debug;
int isfm_ba_debug_get(isfm_ba* h, int32_t what, void* dst);

/* CUDA-event timings (ms) accumulated since the last reset, per kernel family.
 * names_out receives a pointer to a static NUL-separated list; returns the count.        */
#define ISFM_N_TIMERS 16
int isfm_ba_get_timers(isfm_ba* h, double ms_out[ISFM_N_TIMERS], int64_t launches_out[ISFM_N_TIMERS]);
int isfm_ba_reset_timers(isfm_ba* h, int32_t enable);
const char* isfm_timer_name(int32_t i);
/* Persistent PCG kernel: time (ms, as seen by CTA 0) accumulated per phase since creation --
 * [0] mat-vec, [1] combine, [2] peer exchange (push + wait + q), [3] update, [4] coarse
 * correction, [5] direction; *solves_out = PCG solves run by that kernel; *two_level_out = 1 when
 * the two-level preconditioner (cluster similarity modes) is active.                          */
int isfm_ba_get_pcg_phases(isfm_ba* h, double ms_out[8], int64_t* solves_out, int32_t* two_level_out);

/* ------------------------------------------------------------------------------------ */
/* global positioning  (replaces the body of TorchGP.Optimize below the tensor set-up)     */
/* ------------------------------------------------------------------------------------ */
typedef struct isfm_gp_desc {
  int32_t dtype;
  int32_t reject;         /* LM(reject=30), global_positioning.py:161                     */
  double huber_delta;     /* Huber(thres_loss_function) :160                              */
  double tr_radius;       /* TrustRegion(radius=1e3, max=1e8, up=2, down=0.5**4) :158     */
  double tr_max;
  double tr_up;
  double tr_down;
  double pcg_tol;         /* PCG(tol=1e-5) :159                                           */
  int32_t pcg_max_iter;
  int32_t optimize_scales;/* 0 = PairwiseNonBatchedDepthOnly (scales are inputs) :73-83   */
  void* stream;
  isfm_comm* comm;
} isfm_gp_desc;

void isfm_gp_default_desc(isfm_gp_desc* desc);
int isfm_gp_create(const isfm_gp_desc* desc, isfm_gp** out);
void isfm_gp_destroy(isfm_gp* h);

/* Tensors of global_positioning.py:108-168:
 *   centres [n_cam,3] (model.translations), pts [n_pt,3], scales [n_obs,1] (model.scales),
 *   rays [n_obs,3] (input["translations"]), cam_idx / pt_idx int32 [n_obs],
 *   is_calibrated uint8 [n_cam], scale_fixed uint8 [n_obs] or NULL: 1 = row NOT in
 *   `scales.optimize_indices` (global_positioning.py:57-59).                             */
int isfm_gp_set_problem(isfm_gp* h, int64_t n_cam, int64_t n_pt, int64_t n_obs,
                        const void* centres, const void* pts, const void* scales,
                        const void* rays, const int32_t* cam_idx, const int32_t* pt_idx,
                        const uint8_t* is_calibrated, const uint8_t* scale_fixed);
int isfm_gp_step(isfm_gp* h, double* loss_out, isfm_step_stats* stats);
/* Loop of global_positioning.py:172-183 (no identical-loss test).                        */
int isfm_gp_solve(isfm_gp* h, int32_t max_iterations, double function_tolerance,
                  double* loss_history_out, int32_t* n_iterations_out);
int isfm_gp_get_params(isfm_gp* h, void* centres_out, void* pts_out, void* scales_out);
int isfm_gp_cost(isfm_gp* h, double* robust_cost_out, double* sq_cost_out);
int isfm_gp_get_timers(isfm_gp* h, double ms_out[ISFM_N_TIMERS], int64_t launches_out[ISFM_N_TIMERS]);
int isfm_gp_reset_timers(isfm_gp* h, int32_t enable);

/* ------------------------------------------------------------------------------------ */
/* inter-BA track filters (instantsfm/processors/track_filter.py).       */
/* Stateless; every pointer may be host or device memory (host arrays are staged); fp64   */
/* like the reference's numpy.  image_ids / track_idx are the flattened (image_id, index  */
/* of the track in dict order) of every observation, in the reference's loop order.       */
/* ------------------------------------------------------------------------------------ */
typedef enum isfm_filter_mode {
  ISFM_FILTER_ANGLE = 0,                   /* FilterTracksByAngle, track_filter.py:5-24:
                                              threshold = cos(deg2rad(max_angle_error))      */
  ISFM_FILTER_REPROJECTION_NORMALIZED = 1  /* FilterTracksByReprojectionNormalized :26-66:
                                              threshold = max_reprojection_error             */
} isfm_filter_mode;
/* world2cam [n_img,4,4] row-major, xyz [n_trk,3], features_undist [n_obs,3] (the bearing of
 * each observation, Image.features_undist[feature_id]); valid_out[n_obs] = 1 keeps it.    */
int isfm_filter_observations(int32_t mode, int64_t n_obs, int64_t n_img, int64_t n_trk,
                             const double* world2cam, const double* xyz,
                             const double* features_undist, const int32_t* image_ids,
                             const int32_t* track_idx, double threshold, uint8_t* valid_out,
                             void* stream);
/* FilterTracksTriangulationAngle, track_filter.py:116-137.  track_off[n_trk+1] is the CSR
 * of image_ids[n_obs] by track; centers [n_img,3] = Image.center(); cos_threshold =
 * cos(deg2rad(min_angle)); remove_out[n_trk] = 1 where the reference deletes the track.   */
int isfm_filter_triangulation_angle(int64_t n_trk, int64_t n_obs, int64_t n_img,
                                    const int64_t* track_off, const int32_t* image_ids,
                                    const double* centers, const double* xyz,
                                    double cos_threshold, uint8_t* remove_out, void* stream);

/* ------------------------------------------------------------------------------------ */
/* batched reprojection test: the arithmetic of complete_tracks,        */
/* instantsfm/processors/track_retriangulation.py:81-91, fp64 like the reference.          */
/* cam [n_cam, 7 + n_intr] = [t, q_xyzw, intrinsics without pp] per IMAGE, pp [n_cam, 2],  */
/* pts [n_pt, 3], obs [n_obs, 2] observed pixels, cam_idx / pt_idx int32 [n_obs]; host or  */
/* device pointers.  pass_out[a] = (||reproject(X, cam, pp) - obs|| <= max_error) &&       */
/* (rotate_quat(X, cam).z > min_depth); err_out (may be NULL) receives the error norm.     */
/* ------------------------------------------------------------------------------------ */
int isfm_reprojection_test(int32_t model_id, int64_t n_obs, int64_t n_cam, int64_t n_pt,
                           const double* cam, const double* pp, const double* pts,
                           const double* obs, const int32_t* cam_idx, const int32_t* pt_idx,
                           double max_error, double min_depth, uint8_t* pass_out,
                           double* err_out, void* stream);

/* ------------------------------------------------------------------------------------ */
/* pixel-space camera maps of the scene layer, fp64 like the   */
/* reference's numpy, ALL eleven camera models of scene/defs.py:101-113.                   */
/* cam_table [n_cam, ISFM_CAMERA_ROW] doubles per camera, the attributes Camera.set_params */
/* (scene/defs.py:177-237) derives from `params`:                                          */
/*   [0] model id, [1] fx, [2] fy, [3] cx, [4] cy, [5..10] k[0..5], [11] p0, [12] p1,      */
/*   [13] omega, [14] sx0, [15] sx1   (unused entries 0)                                   */
/* ------------------------------------------------------------------------------------ */
#define ISFM_CAMERA_ROW 16
/* FilterTracksByReprojection, instantsfm/processors/track_filter.py:68-114 (called by
 * filter_points, track_retriangulation.py:200-204): per observation
 *   p = world2cam[image] [X; 1];  e = || Camera.cam2img(p) - feature ||   (defs.py:371-412)
 *   valid_out = (p.z > 1e-10) && (e < max_error);   err_out (may be NULL) = e.
 * world2cam [n_img,4,4], image_cam int32 [n_img] (Image.cam_id), xyz [n_trk,3], features
 * [n_obs,2] = Image.features[feature_id] of every observation, image_ids / track_idx as in
 * isfm_filter_observations.                                                               */
int isfm_filter_reprojection(int64_t n_obs, int64_t n_img, int64_t n_trk, int64_t n_cam,
                             const double* world2cam, const int32_t* image_cam,
                             const double* cam_table, const double* xyz, const double* features,
                             const int32_t* image_ids, const int32_t* track_idx, double max_error,
                             uint8_t* valid_out, double* err_out, void* stream);
/* UndistortImages, instantsfm/processors/image_undistortion.py:3-9 (global_mapper.py:98,116,
 * 121,142): features [n_feat,2] pixels, cam_idx int32 [n_feat] = camera of the feature's image;
 * features_undist_out [n_feat,3] = [Camera.img2cam(xy), 1] / norm (defs.py:315-369; the
 * distorted models restate cv2.undistortPoints: 5 fixed-point iterations).                 */
int isfm_undistort_features(int64_t n_feat, int64_t n_cam, const double* cam_table,
                            const double* features, const int32_t* cam_idx,
                            double* features_undist_out, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* ISFM_B200_H_ */
