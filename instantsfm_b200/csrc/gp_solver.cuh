// gp_solver.cuh -- global positioning (TorchGP.Optimize, global_positioning.py:45-206) as
// one LM step = linearise, eliminate the per-observation scales (1x1), eliminate the points
// (3x3), block-Jacobi PCG on the 3x3-block camera system, back-substitute (K6).
//
// residual (utils/cost_function.py:22-29):  r = w (d - s (X - c)),  w = 1 (calibrated) / 0.5
// weighted by omega = sqrt(rho'(||r||^2)):  a = omega w s,  j = dr/ds = -omega w (X - c)
//   J_c = a I3   J_X = -a I3   J_s = j
// Scale elimination per observation (h = damp(j^T j)):
//   Q = a^2 (I - j j^T / h)     rt = a (r - j (j^T r) / h)          (fixed scale: Q = a^2 I)
//   H'cc = damp(sum a^2) I - sum a^2 j j^T / h ,  H'cX = -Q ,  g'c = sum rt ,  g'X = -sum rt
// Per-observation store (sorted by point): A [1], JV [3], RT [3]; per trial QW [15] = Q (6) |
// W = Q Hxx^-1 (9).
#pragma once
#include <algorithm>
#include <cmath>

#include "ba_solver.cuh"  // TrustRegionState
#include "comm.cuh"
#include "index_prep.cuh"
#include "pcg.cuh"

namespace isfm {

constexpr int GP_TPB = 256;

template <typename T>
__device__ __forceinline__ void gp_residual(const T* __restrict__ centres, const T* __restrict__ pts, const T* __restrict__ rays,
                                            const T* __restrict__ scales, const uint8_t* __restrict__ calib, int64_t a,
                                            int orig, int c, int p, T e[3], T r[3], T& w, T& s) {
  w = calib[c] ? T(1) : T(0.5);
  s = scales[orig];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    e[k] = pts[3 * (size_t)p + k] - centres[3 * (size_t)c + k];
    r[k] = w * (rays[3 * (size_t)a + k] - s * e[k]);
  }
}

// scales are indexed by ORIGINAL observation id (so get_params needs no un-permute); rays are
// stored in sorted order.
template <typename T>
__global__ void __launch_bounds__(GP_TPB)
gp_linearize_kernel(int64_t n_obs, const T* __restrict__ centres, const T* __restrict__ pts, const T* __restrict__ rays,
                    const T* __restrict__ scales, const uint8_t* __restrict__ calib, const int32_t* __restrict__ obs_perm,
                    const int32_t* __restrict__ cam_of, const int32_t* __restrict__ pt_of, T delta, T* __restrict__ A,
                    T* __restrict__ JV, T* __restrict__ RT, double* __restrict__ part_rho, double* __restrict__ part_sq) {
  double rho_sum = 0.0, sq_sum = 0.0;
  for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < n_obs; a += (int64_t)gridDim.x * blockDim.x) {
    T e[3], r[3], w, s;
    gp_residual(centres, pts, rays, scales, calib, a, obs_perm[a], cam_of[a], pt_of[a], e, r, w, s);
    T ss = r[0] * r[0] + r[1] * r[1] + r[2] * r[2], rho, om;
    huber(ss, delta, rho, om);
    rho_sum += (double)rho; sq_sum += (double)ss;
    if (A) {
      A[a] = om * w * s;
#pragma unroll
      for (int k = 0; k < 3; ++k) { JV[3 * a + k] = -om * w * e[k]; RT[3 * a + k] = om * r[k]; }
    }
  }
  rho_sum = block_sum(rho_sum);
  sq_sum = block_sum(sq_sum);
  if (threadIdx.x == 0) { part_rho[blockIdx.x] = rho_sum; part_sq[blockIdx.x] = sq_sum; }
}

// Q (xx xy xz yy yz zz) and rt for one observation at damping mu
template <typename T>
__device__ __forceinline__ void gp_eliminate_scale(T a, const T* j, const T* r, bool fixed, T mu, T Q[6], T rt[3], T& h, T& jr) {
  jr = j[0] * r[0] + j[1] * r[1] + j[2] * r[2];
  T a2 = a * a;
  if (fixed) {
    h = T(1);
    Q[0] = a2; Q[1] = 0; Q[2] = 0; Q[3] = a2; Q[4] = 0; Q[5] = a2;
    rt[0] = a * r[0]; rt[1] = a * r[1]; rt[2] = a * r[2];
    return;
  }
  h = damp_diag(j[0] * j[0] + j[1] * j[1] + j[2] * j[2], mu);
  T ih = T(1) / h, k = a2 * ih;
  Q[0] = a2 - k * j[0] * j[0]; Q[1] = -k * j[0] * j[1]; Q[2] = -k * j[0] * j[2];
  Q[3] = a2 - k * j[1] * j[1]; Q[4] = -k * j[1] * j[2]; Q[5] = a2 - k * j[2] * j[2];
  T f = jr * ih;
  rt[0] = a * (r[0] - j[0] * f); rt[1] = a * (r[1] - j[1] * f); rt[2] = a * (r[2] - j[2] * f);
}

// one thread per point: H'XX, its inverse, t_p = H'XX^-1 g'X, and per observation Q | W = Q H'XX^-1
template <typename T>
__global__ void __launch_bounds__(GP_TPB)
gp_point_solve_kernel(int64_t n_pt, const int32_t* __restrict__ pt_off, const int32_t* __restrict__ obs_perm,
                      const uint8_t* __restrict__ fixed, const T* __restrict__ A, const T* __restrict__ JV,
                      const T* __restrict__ RT, T mu, T* __restrict__ HINV, T* __restrict__ GX, T* __restrict__ TP,
                      T* __restrict__ QW) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n_pt) return;
  const int beg = pt_off[p], end = pt_off[p + 1];
  T hq[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0}, beta = 0;
  for (int a = beg; a < end; ++a) {
    T Q[6], rt[3], h, jr;
    gp_eliminate_scale(A[a], JV + 3 * (size_t)a, RT + 3 * (size_t)a, fixed && fixed[obs_perm[a]], mu, Q, rt, h, jr);
    beta += A[a] * A[a];
#pragma unroll
    for (int k = 0; k < 6; ++k) { hq[k] += Q[k]; QW[(size_t)a * 15 + k] = Q[k]; }
    g[0] -= rt[0]; g[1] -= rt[1]; g[2] -= rt[2];
  }
  T dd = damp_diag(beta, mu) - beta;
  hq[0] += dd; hq[3] += dd; hq[5] += dd;
  T iv[6];
  sym3_inverse(hq, iv);
#pragma unroll
  for (int k = 0; k < 6; ++k) HINV[(size_t)p * 6 + k] = iv[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) GX[(size_t)p * 3 + k] = g[k];
  TP[3 * p + 0] = iv[0] * g[0] + iv[1] * g[1] + iv[2] * g[2];
  TP[3 * p + 1] = iv[1] * g[0] + iv[3] * g[1] + iv[4] * g[2];
  TP[3 * p + 2] = iv[2] * g[0] + iv[4] * g[1] + iv[5] * g[2];
  const T I[9] = {iv[0], iv[1], iv[2], iv[1], iv[3], iv[4], iv[2], iv[4], iv[5]};
  for (int a = beg; a < end; ++a) {
    T* q = QW + (size_t)a * 15;
    const T Qm[9] = {q[0], q[1], q[2], q[1], q[3], q[4], q[2], q[4], q[5]};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) q[6 + 3 * r + c] = Qm[3 * r] * I[c] + Qm[3 * r + 1] * I[3 + c] + Qm[3 * r + 2] * I[6 + c];
  }
}

// one CTA per camera: [sum Q (6) | sum a^2 (1) | g'c (3) | E_ii upper (6) | e_i (3)] = 19 values
constexpr int GP_CAM_ACC = 19;
template <typename T>
__global__ void __launch_bounds__(128)
gp_camera_blocks_kernel(const int32_t* __restrict__ cam_off, const int32_t* __restrict__ cam_perm,
                        const int32_t* __restrict__ pt_of, const int32_t* __restrict__ obs_perm,
                        const uint8_t* __restrict__ fixed, const T* __restrict__ A, const T* __restrict__ JV,
                        const T* __restrict__ RT, const T* __restrict__ QW, const T* __restrict__ TP, T mu,
                        T* __restrict__ out /* [n_cam][19] */) {
  __shared__ T sh[4][GP_CAM_ACC];
  const int cam = blockIdx.x;
  T acc[GP_CAM_ACC];
#pragma unroll
  for (int i = 0; i < GP_CAM_ACC; ++i) acc[i] = T(0);
  for (int k = cam_off[cam] + threadIdx.x; k < cam_off[cam + 1]; k += 128) {
    const int a = cam_perm[k];
    T Q[6], rt[3], h, jr;
    gp_eliminate_scale(A[a], JV + 3 * (size_t)a, RT + 3 * (size_t)a, fixed && fixed[obs_perm[a]], mu, Q, rt, h, jr);
    const T* q = QW + (size_t)a * 15;
    const T* t = TP + 3 * (size_t)pt_of[a];
#pragma unroll
    for (int i = 0; i < 6; ++i) acc[i] += Q[i];
    acc[6] += A[a] * A[a];
    acc[7] += rt[0]; acc[8] += rt[1]; acc[9] += rt[2];
    // E_ii += W Q  (W = Q Hinv, 3x3 row-major at q+6), upper triangle
    const T Qm[9] = {Q[0], Q[1], Q[2], Q[1], Q[3], Q[4], Q[2], Q[4], Q[5]};
    int u = 10;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = r; c < 3; ++c) acc[u++] += q[6 + 3 * r] * Qm[c] + q[6 + 3 * r + 1] * Qm[3 + c] + q[6 + 3 * r + 2] * Qm[6 + c];
    acc[16] += Qm[0] * t[0] + Qm[1] * t[1] + Qm[2] * t[2];
    acc[17] += Qm[3] * t[0] + Qm[4] * t[1] + Qm[5] * t[2];
    acc[18] += Qm[6] * t[0] + Qm[7] * t[1] + Qm[8] * t[2];
  }
#pragma unroll
  for (int i = 0; i < GP_CAM_ACC; ++i) {
    T v = acc[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    acc[i] = v;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < GP_CAM_ACC; ++i) sh[w][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < GP_CAM_ACC)
    out[(size_t)cam * GP_CAM_ACC + threadIdx.x] = sh[0][threadIdx.x] + sh[1][threadIdx.x] + sh[2][threadIdx.x] + sh[3][threadIdx.x];
}

// E_ij = sum W_a Q_b (upper triangle only).  A view graph has many camera pairs with FEW common
// tracks (C4: 3.1 M lists of 2.4 pairs on average), so the lists are dealt to sub-warp groups:
// GP_LPL lanes per list -- 4 by default, 8 lists per warp -- instead of a warp per list (2-3 active
// lanes of 32: measured 1.13 ms at C4).  The lanes of a group take the list's pairs round-robin; the
// nine sums are folded inside the group by shuffles (fixed order: deterministic).
constexpr int GP_LPL = 4;
template <typename T>
__global__ void __launch_bounds__(128)
gp_schur_offdiag_kernel(int64_t n_lists, const int64_t* __restrict__ list_off, const uint64_t* __restrict__ pairs,
                        const int32_t* __restrict__ list_slot, const uint8_t* __restrict__ list_diag,
                        const T* __restrict__ QW, T* __restrict__ E) {
  const int sub = threadIdx.x % GP_LPL;
  const int64_t u = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / GP_LPL;
  const bool on = u < n_lists;
  T acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (on) {
    for (int64_t t = list_off[u] + sub; t < list_off[u + 1]; t += GP_LPL) {
      const uint64_t ab = pairs[t];
      const T* wa = QW + (size_t)(uint32_t)(ab >> 32) * 15 + 6;
      const T* qb = QW + (size_t)(uint32_t)ab * 15;
      const T Qm[9] = {qb[0], qb[1], qb[2], qb[1], qb[3], qb[4], qb[2], qb[4], qb[5]};
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[3 * r + c] += wa[3 * r] * Qm[c] + wa[3 * r + 1] * Qm[3 + c] + wa[3 * r + 2] * Qm[6 + c];
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int o = GP_LPL / 2; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  if (on) {
    // lane `sub` of the group stores entries sub, sub + GP_LPL, ...
    T* e = E + (size_t)list_slot[u] * 9;
    const bool diag = list_diag[u] != 0;
#pragma unroll
    for (int i = 0; i < 9; ++i)
      if (i % GP_LPL == sub) { if (diag) e[i] += acc[i]; else e[i] = acc[i]; }
  }
}

// after the all-reduce of the camera accumulators: E_ii into BSR is done before (local); here
// Hd = H'cc, Minv = (H'cc - sum E_ii)^-1, b = -(g'c + e)
template <typename T>
__global__ void gp_precond_kernel(int n_cam, const T* __restrict__ acc_local, const T* __restrict__ acc_sum,
                                  const int32_t* __restrict__ diag_slot, T mu, T* __restrict__ E, T* __restrict__ HD,
                                  T* __restrict__ MINV, T* __restrict__ bvec, int* __restrict__ fail) {
  int cam = blockIdx.x * blockDim.x + threadIdx.x;
  if (cam >= n_cam) return;
  const T* l = acc_local + (size_t)cam * GP_CAM_ACC;
  const T* g = acc_sum + (size_t)cam * GP_CAM_ACC;
  // local E_ii into this rank's BSR
  T* e = E + (size_t)diag_slot[cam] * 9;
  e[0] = l[10]; e[1] = l[11]; e[2] = l[12]; e[3] = l[11]; e[4] = l[13]; e[5] = l[14]; e[6] = l[12]; e[7] = l[14]; e[8] = l[15];
  T dd = damp_diag(g[6], mu) - g[6];
  T hd[9] = {g[0] + dd, g[1], g[2], g[1], g[3] + dd, g[4], g[2], g[4], g[5] + dd};
  double M[9] = {(double)hd[0] - g[10], (double)hd[1] - g[11], (double)hd[2] - g[12], (double)hd[3] - g[11], (double)hd[4] - g[13],
                 (double)hd[5] - g[14], (double)hd[6] - g[12], (double)hd[7] - g[14], (double)hd[8] - g[15]};
  if (!spd_inverse<3>(M)) {   // cancellation: fall back to the SPD block H'cc as preconditioner
    *fail = cam + 1;
#pragma unroll
    for (int k = 0; k < 9; ++k) M[k] = (double)hd[k];
    spd_inverse<3>(M);
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) { HD[(size_t)cam * 9 + k] = hd[k]; MINV[(size_t)cam * 9 + k] = (T)M[k]; }
#pragma unroll
  for (int k = 0; k < 3; ++k) bvec[(size_t)cam * 3 + k] = -(g[7 + k] + g[16 + k]);
}

// diagonal pair lists add onto E_ii AFTER gp_precond_kernel wrote it: run the off-diagonal
// kernel after the precond kernel (their contribution to the preconditioner is ignored; it
// only exists when a track sees the same image twice).
template <typename T>
__global__ void __launch_bounds__(GP_TPB)
gp_backsub_kernel(int64_t n_pt, const int32_t* __restrict__ pt_off, const int32_t* __restrict__ cam_of,
                  const int32_t* __restrict__ obs_perm, const uint8_t* __restrict__ fixed, const T* __restrict__ A,
                  const T* __restrict__ JV, const T* __restrict__ RT, const T* __restrict__ QW, const T* __restrict__ HINV,
                  const T* __restrict__ GX, const T* __restrict__ DC, T mu, const T* __restrict__ pts, const T* __restrict__ scales,
                  T* __restrict__ pts_trial, T* __restrict__ scales_trial, double* __restrict__ part_m) {
  double msum = 0.0;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_pt; p += (int64_t)gridDim.x * blockDim.x) {
    const int beg = pt_off[p], end = pt_off[p + 1];
    T u0 = GX[3 * p], u1 = GX[3 * p + 1], u2 = GX[3 * p + 2];
    for (int a = beg; a < end; ++a) {
      const T* q = QW + (size_t)a * 15;
      const T* dc = DC + 3 * (size_t)cam_of[a];
      u0 -= q[0] * dc[0] + q[1] * dc[1] + q[2] * dc[2];
      u1 -= q[1] * dc[0] + q[3] * dc[1] + q[4] * dc[2];
      u2 -= q[2] * dc[0] + q[4] * dc[1] + q[5] * dc[2];
    }
    const T* iv = HINV + (size_t)p * 6;
    T dx[3] = {-(iv[0] * u0 + iv[1] * u1 + iv[2] * u2), -(iv[1] * u0 + iv[3] * u1 + iv[4] * u2),
               -(iv[2] * u0 + iv[4] * u1 + iv[5] * u2)};
#pragma unroll
    for (int k = 0; k < 3; ++k) pts_trial[3 * p + k] = pts[3 * p + k] + dx[k];
    for (int a = beg; a < end; ++a) {
      const T* j = JV + 3 * (size_t)a;
      const T* r = RT + 3 * (size_t)a;
      const T* dc = DC + 3 * (size_t)cam_of[a];
      const int orig = obs_perm[a];
      const T av = A[a];
      T dv[3] = {dc[0] - dx[0], dc[1] - dx[1], dc[2] - dx[2]};
      T ds = T(0);
      if (!(fixed && fixed[orig])) {
        T h = damp_diag(j[0] * j[0] + j[1] * j[1] + j[2] * j[2], mu);
        T jr = j[0] * r[0] + j[1] * r[1] + j[2] * r[2];
        ds = -(jr + av * (j[0] * dv[0] + j[1] * dv[1] + j[2] * dv[2])) / h;
      }
      scales_trial[orig] = scales[orig] + ds;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        T jd = av * dv[k] + j[k] * ds;
        msum += (double)(jd * (2 * r[k] + jd));
      }
    }
  }
  msum = block_sum(msum);
  if (threadIdx.x == 0) part_m[blockIdx.x] = msum;
}

template <typename T>
__global__ void axpy_kernel(int64_t n, const T* __restrict__ x, const T* __restrict__ d, T* __restrict__ out,
                            double* __restrict__ part_norm) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double nrm = 0.0;
  if (i < n) { out[i] = x[i] + d[i]; nrm = (double)d[i] * (double)d[i]; }
  nrm = block_sum(nrm);
  if (threadIdx.x == 0) part_norm[blockIdx.x] = nrm;
}

struct GPSolverBase {
  virtual ~GPSolverBase() {}
  virtual void set_problem(int64_t n_cam, int64_t n_pt, int64_t n_obs, const void* centres, const void* pts, const void* scales,
                           const void* rays, const int32_t* cam_idx, const int32_t* pt_idx, const uint8_t* is_calibrated,
                           const uint8_t* scale_fixed) = 0;
  virtual void step(double* loss_out, isfm_step_stats* stats) = 0;
  virtual void get_params(void* centres_out, void* pts_out, void* scales_out) = 0;
  virtual void cost(double* robust, double* sq) = 0;
  KernelTimers timers;
  bool has_problem = false;
};

template <typename T>
struct GPSolver : GPSolverBase {
  isfm_gp_desc desc;
  cudaStream_t s;
  isfm_comm* comm;
  TrustRegionState tr;
  int64_t n_cam = 0, n_pt = 0, n_obs = 0;
  ObsIndex ix;
  SchurPattern sp;
  DeviceBuffer<T> centres[2], pts[2], scales[2], rays;
  DeviceBuffer<uint8_t> calib, fixed;
  bool all_fixed = false;
  DeviceBuffer<T> A, JV, RT, QW, HINV, GX, TP, ACC, ACCSUM, E, HD, MINV, bvec;
  DeviceBuffer<double> part_a, part_b, part_c, scalars;
  DeviceBuffer<int> fail;
  double* h_scalars = nullptr;
  BlockPCG<T, 3> pcg;
  int cur = 0;
  bool have_loss = false;
  double loss = 0.0;

  explicit GPSolver(const isfm_gp_desc& d) : desc(d) {
    s = static_cast<cudaStream_t>(d.stream);   // never the legacy default stream: isfm_gp_create substitutes an own stream
    comm = d.comm;
    timers.stream = s;
    tr.init(d.tr_radius, d.tr_max, d.tr_up, d.tr_down);
    ISFM_CUDA(cudaMallocHost(&h_scalars, 4 * sizeof(double)));
  }
  ~GPSolver() override {
    cudaStreamSynchronize(s);   // buffers go back to the stream-ordered pool after all work has finished
    if (h_scalars) cudaFreeHost(h_scalars);
  }

  template <typename U>
  void upload(DeviceBuffer<U>& dst, const void* src, size_t count) {
    dst.alloc(count);
    ISFM_CUDA(cudaMemcpyAsync(dst.get(), src, count * sizeof(U), cudaMemcpyDefault, s));
  }

  void set_problem(int64_t nc, int64_t np, int64_t no, const void* centres_in, const void* pts_in, const void* scales_in,
                   const void* rays_in, const int32_t* cam_idx, const int32_t* pt_idx, const uint8_t* is_cal,
                   const uint8_t* scale_fixed) override {
    ISFM_REQUIRE(centres_in && pts_in && scales_in && rays_in && cam_idx && pt_idx && is_cal, ISFM_EINVAL, "null input");
    n_cam = nc; n_pt = np; n_obs = no;
    upload(centres[0], centres_in, (size_t)nc * 3); centres[1].alloc((size_t)nc * 3);
    upload(pts[0], pts_in, (size_t)np * 3); pts[1].alloc((size_t)np * 3);
    upload(scales[0], scales_in, (size_t)no); scales[1].alloc((size_t)no);
    upload(calib, is_cal, (size_t)nc);
    all_fixed = !desc.optimize_scales;
    if (all_fixed) {
      fixed.alloc((size_t)no);
      ISFM_CUDA(cudaMemsetAsync(fixed.get(), 1, (size_t)no, s));
    } else if (scale_fixed) {
      upload(fixed, scale_fixed, (size_t)no);
    } else {
      fixed.release();
    }
    DeviceBuffer<T> rays_raw; DeviceBuffer<int32_t> ci, pi;
    upload(rays_raw, rays_in, (size_t)no * 3);
    upload(ci, cam_idx, (size_t)no); upload(pi, pt_idx, (size_t)no);
    build_obs_index(ix, nc, np, no, ci.get(), pi.get(), s, timers);
    rays.alloc((size_t)no * 3);
    { TimerScope ts(timers, T_INDEX_PREP);
      gather_rows_kernel<T><<<div_up(no * 3, GP_TPB), GP_TPB, 0, s>>>(no, 3, rays_raw.get(), ix.obs_perm.get(), rays.get()); }
    build_schur_pattern(sp, ix, s, timers, nullptr, SpmvCfg<T, 3>::WB, persistent_grid_warps<T, 3>());
    A.alloc((size_t)no); JV.alloc((size_t)no * 3); RT.alloc((size_t)no * 3); QW.alloc((size_t)no * 15);
    HINV.alloc((size_t)np * 6); GX.alloc((size_t)np * 3); TP.alloc((size_t)np * 3);
    ACC.alloc((size_t)nc * GP_CAM_ACC); ACCSUM.alloc((size_t)nc * GP_CAM_ACC);
    E.alloc((size_t)sp.nnzu * 9); E.zero(s);   // padding slots stay zero
    HD.alloc((size_t)nc * 9); MINV.alloc((size_t)nc * 9); bvec.alloc((size_t)nc * 3);
    part_a.alloc(std::max<int64_t>(RED_BLOCKS, nc)); part_b.alloc(std::max<int64_t>(RED_BLOCKS, nc));
    part_c.alloc(std::max<int64_t>(RED_BLOCKS, nc));
    scalars.alloc(4); fail.alloc(1); fail.zero(s);
    pcg.resize((int)nc, sp.n_off, sp.n_chunks, comm, s);
    ISFM_CUDA(cudaStreamSynchronize(s));
    cur = 0; have_loss = false; has_problem = true;
    tr.init(desc.tr_radius, desc.tr_max, desc.tr_up, desc.tr_down);
  }

  int red_grid(int64_t n) const { return (int)std::min<int64_t>(RED_BLOCKS, div_up(n, GP_TPB)); }
  const uint8_t* fixed_ptr() const { return fixed.get(); }

  void fetch_scalars(const double* p0, int n0, const double* p1, int n1, const double* p2, int n2) {
    { TimerScope ts(timers, T_REDUCE);
      reduce_scalars_kernel<<<3, 256, 0, s>>>(p0, p1, p2, n0, n1, n2, scalars.get()); }
    if (comm_world(comm) > 1) {
      TimerScope ts(timers, T_COMM);
      comm_allreduce_sum(comm, scalars.get(), 3, true, s);
    }
    ISFM_CUDA(cudaMemcpyAsync(h_scalars, scalars.get(), 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    ISFM_CUDA(cudaStreamSynchronize(s));
  }

  void run_residual_pass(int which, bool store) {
    const int g = red_grid(n_obs);
    TimerScope ts(timers, store ? T_LINEARIZE : T_COST);
    gp_linearize_kernel<T><<<g, GP_TPB, 0, s>>>(n_obs, centres[which].get(), pts[which].get(), rays.get(), scales[which].get(),
                                                calib.get(), ix.obs_perm.get(), ix.cam_of.get(), ix.pt_of.get(),
                                                (T)desc.huber_delta, store ? A.get() : nullptr, JV.get(), RT.get(),
                                                part_a.get(), part_b.get());
  }

  void step(double* loss_out, isfm_step_stats* st) override {
    ISFM_REQUIRE(has_problem, ISFM_ESTATE, "isfm_gp_step before isfm_gp_set_problem");
    run_residual_pass(cur, true);
    const int g = red_grid(n_obs);
    if (!have_loss) {
      fetch_scalars(part_a.get(), g, part_b.get(), g, nullptr, 0);
      loss = h_scalars[0];
      have_loss = true;
      ISFM_REQUIRE(std::isfinite(loss), ISFM_ENONFINITE, "initial cost is not finite");
    }
    const double last = loss;
    isfm_step_stats stats;
    memset(&stats, 0, sizeof stats);
    stats.loss_before = last;
    double mu = 1.0;
    int rejects = 0;
    const int trial = cur ^ 1;
    while (last <= loss) {
      // fp32: same damping floor as bundle adjustment (ba_solver.cuh) -- the reduced
      // centre system has gauge modes (translation, scale) too, left with nothing but the damping term
      const double floor_T = sizeof(T) == 4 ? 1e-6 : 0.0;
      mu = std::min(mu * (1.0 + std::max(tr.damping, floor_T)), sizeof(T) == 4 ? 1e24 : 1e100);
      const T m = (T)mu;
      { TimerScope ts(timers, T_POINT_SOLVE);
        gp_point_solve_kernel<T><<<div_up(n_pt, GP_TPB), GP_TPB, 0, s>>>(n_pt, ix.pt_off.get(), ix.obs_perm.get(), fixed_ptr(),
                                                                        A.get(), JV.get(), RT.get(), m, HINV.get(), GX.get(),
                                                                        TP.get(), QW.get()); }
      { TimerScope ts(timers, T_CAMERA_BLOCKS);
        gp_camera_blocks_kernel<T><<<(int)n_cam, 128, 0, s>>>(ix.cam_off.get(), ix.cam_perm.get(), ix.pt_of.get(),
                                                             ix.obs_perm.get(), fixed_ptr(), A.get(), JV.get(), RT.get(),
                                                             QW.get(), TP.get(), m, ACC.get()); }
      ISFM_CUDA(cudaMemcpyAsync(ACCSUM.get(), ACC.get(), (size_t)n_cam * GP_CAM_ACC * sizeof(T), cudaMemcpyDeviceToDevice, s));
      if (comm_world(comm) > 1) {
        TimerScope ts(timers, T_COMM);
        comm_allreduce_sum(comm, ACCSUM.get(), (size_t)n_cam * GP_CAM_ACC, sizeof(T) == 8, s);
      }
      { TimerScope ts(timers, T_PRECOND);
        gp_precond_kernel<T><<<div_up(n_cam, 128), 128, 0, s>>>((int)n_cam, ACC.get(), ACCSUM.get(), sp.diag_slot.get(), m,
                                                               E.get(), HD.get(), MINV.get(), bvec.get(), fail.get()); }
      if (sp.n_lists > 0) {
        TimerScope ts(timers, T_SCHUR_OFFDIAG);
        gp_schur_offdiag_kernel<T><<<div_up(sp.n_lists * GP_LPL, 128), 128, 0, s>>>(sp.n_lists, sp.list_off.get(), sp.pairs.get(),
                                                                        sp.list_slot.get(), sp.list_diag.get(), QW.get(),
                                                                        E.get());
      }
      int pcg_status = 0;
      int max_iter = desc.pcg_max_iter > 0 ? desc.pcg_max_iter : (int)std::min<int64_t>(10 * n_cam * 3, 5000);
      stats.pcg_iters += pcg.solve(sp, E.get(), HD.get(),
                                   MINV.get(), bvec.get(), desc.pcg_tol, max_iter, comm, s, timers, &pcg_status);
      const int mparts = red_grid(n_pt);
      { TimerScope ts(timers, T_BACKSUB);
        gp_backsub_kernel<T><<<mparts, GP_TPB, 0, s>>>(n_pt, ix.pt_off.get(), ix.cam_of.get(), ix.obs_perm.get(), fixed_ptr(),
                                                      A.get(), JV.get(), RT.get(), QW.get(), HINV.get(), GX.get(), pcg.x.get(),
                                                      m, pts[cur].get(), scales[cur].get(), pts[trial].get(),
                                                      scales[trial].get(), part_c.get()); }
      { TimerScope ts(timers, T_UPDATE);
        axpy_kernel<T><<<div_up(n_cam * 3, 128), 128, 0, s>>>(n_cam * 3, centres[cur].get(), pcg.x.get(), centres[trial].get(),
                                                             pcg.part_a.get()); }
      run_residual_pass(trial, false);
      fetch_scalars(part_a.get(), g, part_b.get(), g, part_c.get(), mparts);
      const double new_loss = h_scalars[0], mterm = h_scalars[2];
      stats.trials++;
      stats.model_term = -mterm;
      stats.quality = tr.update(last, new_loss, mterm);
      const bool worse = !(new_loss <= last);
      if (worse && rejects < desc.reject) { rejects++; loss = last; stats.accepted = 0; }
      else { cur = trial; loss = new_loss; stats.accepted = 1; break; }
    }
    stats.rejects = rejects; stats.loss = loss; stats.damping = tr.damping;
    ISFM_CUDA(cudaGetLastError());
    if (loss_out) *loss_out = loss;
    if (st) *st = stats;
  }

  void get_params(void* c_out, void* p_out, void* s_out) override {
    ISFM_REQUIRE(has_problem, ISFM_ESTATE, "no problem set");
    if (c_out) ISFM_CUDA(cudaMemcpyAsync(c_out, centres[cur].get(), (size_t)n_cam * 3 * sizeof(T), cudaMemcpyDefault, s));
    if (p_out) ISFM_CUDA(cudaMemcpyAsync(p_out, pts[cur].get(), (size_t)n_pt * 3 * sizeof(T), cudaMemcpyDefault, s));
    if (s_out) ISFM_CUDA(cudaMemcpyAsync(s_out, scales[cur].get(), (size_t)n_obs * sizeof(T), cudaMemcpyDefault, s));
    ISFM_CUDA(cudaStreamSynchronize(s));
  }

  void cost(double* robust, double* sq) override {
    ISFM_REQUIRE(has_problem, ISFM_ESTATE, "no problem set");
    run_residual_pass(cur, false);
    const int g = red_grid(n_obs);
    fetch_scalars(part_a.get(), g, part_b.get(), g, nullptr, 0);
    if (robust) *robust = h_scalars[0];
    if (sq) *sq = h_scalars[1];
  }
};

GPSolverBase* make_gp_solver_f32(const isfm_gp_desc& d);
GPSolverBase* make_gp_solver_f64(const isfm_gp_desc& d);

}  // namespace isfm
