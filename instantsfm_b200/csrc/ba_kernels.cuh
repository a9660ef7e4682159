// ba_kernels.cuh -- CUDA kernels of the bundle-adjustment LM step (K1, K2, K3, K5).
//
// Data layout in HBM (T = float in the product build), "sorted position" a = observation
// index after the stable sort by point:
//   cam   [n_cam][7+NI]   t, q_xyzw, intrinsics (pp removed)      pp  [n_cam][2]
//   pts   [n_pt][3]                                               obs [n_obs][2]   (sorted)
//   R     [n_obs][2]      Triggs-weighted residual
//   OBS   [n_obs][REC]    one 16-byte-aligned record per observation (128 B for D = 9, fp32):
//                           [0, 2D)        Jc  weighted camera block, row-major 2 x D (D = 6 + NI)
//                           [2D, 2D+6)     Jp  weighted point block, 2 x 3
//                           [2D+6, 2D+12)  V = Jp * Hpp^-1 (2 x 3), rewritten per trial
//                           [2D+12, 2D+14) rho = R - Jp t_p (t_p = Hpp^-1 g_p), rewritten per trial
//                         so that every gather of an observation is one or two full lines
//                         fetched with 128-bit loads
//   HPP   [n_pt][6]  GPT [n_pt][3]  HPPINV [n_pt][6]  TP [n_pt][3] = Hpp^-1 g_p
//   HCC   [n_cam][D*D]  GC [n_cam][D]   (undamped, all ranks' sum)
//   E     [nnzb][D*D]     BSR values of sum_p Hcp Hpp^-1 Hcp^T; S = damp(Hcc) - E
// The J-form is kept instead of Hcp = Jc^T Jp (27 floats / obs): every Schur product is
// Jc_a^T (V_a Jp_b^T) Jc_b with a 2x2 middle factor.
#pragma once
#include <cuda.h>   // CUtensorMap (the driver entry point is resolved at run time, no libcuda link)

#include "common.cuh"
#include "math.cuh"

namespace isfm {

constexpr int BA_TPB = 256;
constexpr int RED_BLOCKS = 148 * 4;  // grid of the grid-stride reduction kernels (4 CTAs / SM)

// per-observation record geometry
template <int D> struct ObsRec {
  static constexpr int JP = 2 * D;                       // offset of Jp
  static constexpr int V = 2 * D + 6;                    // offset of V
  static constexpr int RHO = 2 * D + 12;                 // offset of rho = R - Jp t_p
  static constexpr int REC = (2 * D + 14 + 3) / 4 * 4;   // stride in elements, multiple of 4
};

// 128-bit loads / stores of "quads" (4 consecutive elements; two double2 for T = double)
template <typename T> struct QuadIO;
template <> struct QuadIO<float> {
  static __device__ __forceinline__ void ld(const float* p, float* d) {
    float4 v = *reinterpret_cast<const float4*>(p); d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  static __device__ __forceinline__ void ldg(const float* p, float* d) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p)); d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  static __device__ __forceinline__ void st(float* p, const float* d) {
    *reinterpret_cast<float4*>(p) = make_float4(d[0], d[1], d[2], d[3]);
  }
};
template <> struct QuadIO<double> {
  static __device__ __forceinline__ void ld(const double* p, double* d) {
    double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y;
  }
  static __device__ __forceinline__ void ldg(const double* p, double* d) {
    double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p + 2));
    d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y;
  }
  static __device__ __forceinline__ void st(double* p, const double* d) {
    *reinterpret_cast<double2*>(p) = make_double2(d[0], d[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(d[2], d[3]);
  }
};
// asynchronous global -> shared copy of one quad (L2 only)
template <typename T> __device__ __forceinline__ void cp_async_quad(T* smem_dst, const T* gmem_src) {
  cp_async16(smem_dst, gmem_src);
  if (sizeof(T) == 8) cp_async16(smem_dst + 2, gmem_src + 2);
}
// loads quads [Q0, Q1) of a record into dst[0 .. 4 (Q1 - Q0)); element e of the record is
// dst[e - 4 Q0].  READONLY selects the non-coherent path.
template <typename T, int Q0, int Q1, bool READONLY>
__device__ __forceinline__ void load_quads(const T* __restrict__ rec, T* dst) {
#pragma unroll
  for (int q = Q0; q < Q1; ++q) {
    if (READONLY) QuadIO<T>::ldg(rec + 4 * q, dst + 4 * (q - Q0));
    else QuadIO<T>::ld(rec + 4 * q, dst + 4 * (q - Q0));
  }
}

// Packed camera table read by the per-observation kernels: row = [t, q, intrinsics | pp | pad],
// a multiple of four elements, so that a camera is gathered with ceil((CW + 2) / 4) 128-bit loads
// instead of CW + 2 scalar ones (the gathers are random over all cameras on wide-baseline scenes).
template <int CW> struct CamPack { static constexpr int CWP = (CW + 2 + 3) / 4 * 4; };

template <typename T, int CW>
__device__ __forceinline__ void load_camera(const T* __restrict__ camq, int c, T* cr, T* ppv) {
  constexpr int CWP = CamPack<CW>::CWP;
  T q[CWP];
#pragma unroll
  for (int i = 0; i < CWP / 4; ++i) QuadIO<T>::ldg(camq + (size_t)c * CWP + 4 * i, q + 4 * i);
#pragma unroll
  for (int i = 0; i < CW; ++i) cr[i] = q[i];
  ppv[0] = q[CW]; ppv[1] = q[CW + 1];
}

template <typename T, int CW>
__global__ void pack_cameras_kernel(int n_cam, const T* __restrict__ cam, const T* __restrict__ pp, T* __restrict__ camq) {
  constexpr int CWP = CamPack<CW>::CWP;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cam * CWP) return;
  const int c = i / CWP, k = i % CWP;
  camq[i] = k < CW ? cam[(size_t)c * CW + k] : (k < CW + 2 ? pp[2 * (size_t)c + (k - CW)] : T(0));
}

// ---------------------------------------------------------------------------------------
// K1: residual + Huber weight + Jacobian blocks per observation; cost partials per block.
// ---------------------------------------------------------------------------------------
template <typename T, int MODEL>
__global__ void __launch_bounds__(BA_TPB)
linearize_kernel(int64_t n_obs, const T* __restrict__ camq, const T* __restrict__ pts,
                 const T* __restrict__ obs, const int32_t* __restrict__ cam_of, const int32_t* __restrict__ pt_of,
                 T delta, T* __restrict__ R, T* __restrict__ OBS,
                 double* __restrict__ part_rho, double* __restrict__ part_sq) {
  constexpr int NI = ModelTraits<MODEL>::NI;
  constexpr int D = 6 + NI;
  constexpr int CW = 7 + NI;
  constexpr int REC = ObsRec<D>::REC;
  constexpr int NQ = (2 * D + 6 + 3) / 4;   // quads covering Jc | Jp (may spill into V, rewritten later)
  double rho_sum = 0.0, sq_sum = 0.0;
  for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < n_obs; a += (int64_t)gridDim.x * blockDim.x) {
    const int c = cam_of[a], p = pt_of[a];
    T cr[CW], ppv[2], X[3], o[2], r[2], rec[4 * NQ];
    load_camera<T, CW>(camq, c, cr, ppv);
#pragma unroll
    for (int i = 0; i < 3; ++i) X[i] = __ldg(pts + 3 * (size_t)p + i);
    o[0] = obs[2 * a]; o[1] = obs[2 * a + 1];
#pragma unroll
    for (int i = 2 * D + 6; i < 4 * NQ; ++i) rec[i] = T(0);
    ba_linearize<MODEL, T>(cr, ppv, X, o, r, rec, rec + 2 * D);
    T s = r[0] * r[0] + r[1] * r[1], rho, w;
    huber(s, delta, rho, w);
    rho_sum += (double)rho; sq_sum += (double)s;
    R[2 * a] = w * r[0]; R[2 * a + 1] = w * r[1];
#pragma unroll
    for (int i = 0; i < 2 * D + 6; ++i) rec[i] *= w;
    T* dst = OBS + (size_t)a * REC;
#pragma unroll
    for (int q = 0; q < NQ; ++q) QuadIO<T>::st(dst + 4 * q, rec + 4 * q);
  }
  rho_sum = block_sum(rho_sum);
  sq_sum = block_sum(sq_sum);
  if (threadIdx.x == 0) { part_rho[blockIdx.x] = rho_sum; part_sq[blockIdx.x] = sq_sum; }
}

// trial cost: residual only
template <typename T, int MODEL>
__global__ void __launch_bounds__(BA_TPB)
cost_kernel(int64_t n_obs, const T* __restrict__ camq, const T* __restrict__ pts,
            const T* __restrict__ obs, const int32_t* __restrict__ cam_of, const int32_t* __restrict__ pt_of,
            T delta, double* __restrict__ part_rho, double* __restrict__ part_sq) {
  constexpr int NI = ModelTraits<MODEL>::NI;
  constexpr int CW = 7 + NI;
  double rho_sum = 0.0, sq_sum = 0.0;
  for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < n_obs; a += (int64_t)gridDim.x * blockDim.x) {
    const int c = cam_of[a], p = pt_of[a];
    T cr[CW], ppv[2], X[3], o[2], r[2];
    load_camera<T, CW>(camq, c, cr, ppv);
#pragma unroll
    for (int i = 0; i < 3; ++i) X[i] = __ldg(pts + 3 * (size_t)p + i);
    o[0] = obs[2 * a]; o[1] = obs[2 * a + 1];
    ba_residual<MODEL, T>(cr, ppv, X, o, r);
    T s = r[0] * r[0] + r[1] * r[1], rho, w;
    huber(s, delta, rho, w);
    rho_sum += (double)rho; sq_sum += (double)s;
  }
  rho_sum = block_sum(rho_sum);
  sq_sum = block_sum(sq_sum);
  if (threadIdx.x == 0) { part_rho[blockIdx.x] = rho_sum; part_sq[blockIdx.x] = sq_sum; }
}

// unweighted residuals (debug / parity)
template <typename T, int MODEL>
__global__ void residual_kernel(int64_t n_obs, const T* __restrict__ camq,
                                const T* __restrict__ pts, const T* __restrict__ obs,
                                const int32_t* __restrict__ cam_of, const int32_t* __restrict__ pt_of, T* __restrict__ out) {
  constexpr int NI = ModelTraits<MODEL>::NI;
  constexpr int CW = 7 + NI;
  int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (a >= n_obs) return;
  T cr[CW], ppv[2];
  load_camera<T, CW>(camq, cam_of[a], cr, ppv);
  ba_residual<MODEL, T>(cr, ppv, pts + 3 * (size_t)pt_of[a], obs + 2 * a, out + 2 * a);
}

// out[k] = sum of partials k (one block per output scalar)
static __global__ void reduce_scalars_kernel(const double* __restrict__ p0, const double* __restrict__ p1,
                                      const double* __restrict__ p2, int n0, int n1, int n2, double* out) {
  const double* p = blockIdx.x == 0 ? p0 : (blockIdx.x == 1 ? p1 : p2);
  int n = blockIdx.x == 0 ? n0 : (blockIdx.x == 1 ? n1 : n2);
  if (!p) { if (threadIdx.x == 0) out[blockIdx.x] = 0.0; return; }
  double v = reduce_partials(p, n);
  if (threadIdx.x == 0) out[blockIdx.x] = v;
}

// first stage for long partial arrays (one per fused-kernel CTA: 234 k at 60 M observations):
// slice k of array a -> stage[a * RED_SLICES + k]; fixed slices, fixed order: deterministic
constexpr int RED_SLICES = 64;
static __global__ void reduce_stage_kernel(const double* __restrict__ p0, const double* __restrict__ p1,
                                           const double* __restrict__ p2, int n0, int n1, int n2, double* __restrict__ stage) {
  const int a = blockIdx.y, k = blockIdx.x;
  const double* p = a == 0 ? p0 : (a == 1 ? p1 : p2);
  const int n = a == 0 ? n0 : (a == 1 ? n1 : n2);
  double v = 0.0;
  if (p) {
    const int per = (n + RED_SLICES - 1) / RED_SLICES;
    const int lo = min(n, k * per), hi = min(n, lo + per);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) v += p[i];
  }
  v = block_sum(v);
  if (threadIdx.x == 0) stage[a * RED_SLICES + k] = v;
}

// ---------------------------------------------------------------------------------------
// K2 (point side) + K3 (3x3): one thread per point over its contiguous observation records.
// BUILD: accumulate Hpp = sum Jp^T Jp and g_p = sum Jp^T R.  Always: damp, invert, TP, V.
// D = 0 selects the points-only layout (record = Jp only is not used: the record keeps its
// full geometry, only V is skipped when WRITE_V is false).
// ---------------------------------------------------------------------------------------
template <typename T, int D, bool BUILD, bool WRITE_V>
__global__ void __launch_bounds__(BA_TPB)
point_solve_kernel(int64_t n_pt, const int32_t* __restrict__ pt_off, T* __restrict__ OBS, const T* __restrict__ R,
                   T mu, T* __restrict__ HPP, T* __restrict__ GPT, T* __restrict__ HPPINV, T* __restrict__ TP) {
  constexpr int REC = ObsRec<D>::REC, OJP = ObsRec<D>::JP, OV = ObsRec<D>::V;
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n_pt) return;
  const int beg = pt_off[p], end = pt_off[p + 1];
  T h[6], g[3];
  if (BUILD) {
#pragma unroll
    for (int i = 0; i < 6; ++i) h[i] = T(0);
    g[0] = g[1] = g[2] = T(0);
    for (int a = beg; a < end; ++a) {
      const T* j = OBS + (size_t)a * REC + OJP;
      T j0 = j[0], j1 = j[1], j2 = j[2], j3 = j[3], j4 = j[4], j5 = j[5];
      T r0 = R[2 * (size_t)a], r1 = R[2 * (size_t)a + 1];
      h[0] += j0 * j0 + j3 * j3; h[1] += j0 * j1 + j3 * j4; h[2] += j0 * j2 + j3 * j5;
      h[3] += j1 * j1 + j4 * j4; h[4] += j1 * j2 + j4 * j5; h[5] += j2 * j2 + j5 * j5;
      g[0] += j0 * r0 + j3 * r1; g[1] += j1 * r0 + j4 * r1; g[2] += j2 * r0 + j5 * r1;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) HPP[(size_t)p * 6 + i] = h[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) GPT[(size_t)p * 3 + i] = g[i];
  } else {
#pragma unroll
    for (int i = 0; i < 6; ++i) h[i] = HPP[(size_t)p * 6 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) g[i] = GPT[(size_t)p * 3 + i];
  }
  h[0] = damp_diag(h[0], mu); h[3] = damp_diag(h[3], mu); h[5] = damp_diag(h[5], mu);
  T iv[6];
  sym3_inverse(h, iv);
#pragma unroll
  for (int i = 0; i < 6; ++i) HPPINV[(size_t)p * 6 + i] = iv[i];
  const T t0 = iv[0] * g[0] + iv[1] * g[1] + iv[2] * g[2];
  const T t1 = iv[1] * g[0] + iv[3] * g[1] + iv[4] * g[2];
  const T t2 = iv[2] * g[0] + iv[4] * g[1] + iv[5] * g[2];
  TP[(size_t)p * 3 + 0] = t0; TP[(size_t)p * 3 + 1] = t1; TP[(size_t)p * 3 + 2] = t2;
  if (WRITE_V) {
    for (int a = beg; a < end; ++a) {
      T* rec = OBS + (size_t)a * REC;
      const T* j = rec + OJP;
      T* v = rec + OV;
      rec[ObsRec<D>::RHO] = R[2 * (size_t)a] - (j[0] * t0 + j[1] * t1 + j[2] * t2);
      rec[ObsRec<D>::RHO + 1] = R[2 * (size_t)a + 1] - (j[3] * t0 + j[4] * t1 + j[5] * t2);
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        T a0 = j[3 * row], a1 = j[3 * row + 1], a2 = j[3 * row + 2];
        v[3 * row + 0] = a0 * iv[0] + a1 * iv[1] + a2 * iv[2];
        v[3 * row + 1] = a0 * iv[1] + a1 * iv[3] + a2 * iv[4];
        v[3 * row + 2] = a0 * iv[2] + a1 * iv[4] + a2 * iv[5];
      }
    }
  }
}

// points-only BA (optimize_poses = False): D_p = -Hpp^-1 g_p, trial points, model term
template <typename T, int D>
__global__ void __launch_bounds__(BA_TPB)
point_only_step_kernel(int64_t n_pt, const int32_t* __restrict__ pt_off, const T* __restrict__ OBS,
                       const T* __restrict__ R, const T* __restrict__ TP, const T* __restrict__ pts,
                       T* __restrict__ pts_trial, T* __restrict__ DP, double* __restrict__ part_m) {
  constexpr int REC = ObsRec<D>::REC, OJP = ObsRec<D>::JP;
  double msum = 0.0;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_pt; p += (int64_t)gridDim.x * blockDim.x) {
    T d0 = -TP[3 * p], d1 = -TP[3 * p + 1], d2 = -TP[3 * p + 2];
    DP[3 * p] = d0; DP[3 * p + 1] = d1; DP[3 * p + 2] = d2;
    pts_trial[3 * p] = pts[3 * p] + d0; pts_trial[3 * p + 1] = pts[3 * p + 1] + d1; pts_trial[3 * p + 2] = pts[3 * p + 2] + d2;
    for (int a = pt_off[p]; a < pt_off[p + 1]; ++a) {
      const T* j = OBS + (size_t)a * REC + OJP;
      T jd0 = j[0] * d0 + j[1] * d1 + j[2] * d2, jd1 = j[3] * d0 + j[4] * d1 + j[5] * d2;
      msum += (double)(jd0 * (2 * R[2 * (size_t)a] + jd0) + jd1 * (2 * R[2 * (size_t)a + 1] + jd1));
    }
  }
  msum = block_sum(msum);
  if (threadIdx.x == 0) part_m[blockIdx.x] = msum;
}

// ---------------------------------------------------------------------------------------
// Undamped camera blocks Hcc_i = sum Jc^T Jc and g_c = sum Jc^T R, one CTA per camera over its
// camera-major record list.  Not on the step path any more (camera_schur_kernel forms
// Hcc - E_ii directly); kept for the parity / debug buffers ISFM_BA_HCC and ISFM_BA_GC.
// ---------------------------------------------------------------------------------------
constexpr int CAM_TPB = 128;

template <typename T, int D>
__global__ void __launch_bounds__(CAM_TPB)
camera_hessian_kernel(const int32_t* __restrict__ cam_off, const int32_t* __restrict__ cam_perm,
                      const T* __restrict__ OBS, const T* __restrict__ R, T* __restrict__ out_blocks, T* __restrict__ out_vec) {
  constexpr int REC = ObsRec<D>::REC;
  constexpr int NQ = (2 * D + 3) / 4;
  constexpr int NU = D * (D + 1) / 2;
  constexpr int NACC = NU + D;
  constexpr int NW = CAM_TPB / 32;
  __shared__ T sh[NW][NACC];
  const int cam = blockIdx.x;
  const int beg = cam_off[cam], end = cam_off[cam + 1];
  T acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = T(0);
  for (int k = beg + threadIdx.x; k < end; k += CAM_TPB) {
    const int a = cam_perm[k];
    T rec[4 * NQ];
    load_quads<T, 0, NQ, true>(OBS + (size_t)a * REC, rec);
    const T* jc = rec;
    const T s0 = R[2 * (size_t)a], s1 = R[2 * (size_t)a + 1];
    int u = 0;
#pragma unroll
    for (int r = 0; r < D; ++r) {
#pragma unroll
      for (int c = r; c < D; ++c) acc[u++] += jc[r] * jc[c] + jc[D + r] * jc[D + c];
      acc[NU + r] += jc[r] * s0 + jc[D + r] * s1;
    }
  }
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    T v = acc[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    acc[i] = v;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) sh[w][i] = acc[i];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < NACC; t += CAM_TPB) {
    T v = T(0);
#pragma unroll
    for (int k = 0; k < NW; ++k) v += sh[k][t];
    if (t >= NU) {
      out_vec[(size_t)cam * D + (t - NU)] = v;
    } else {
      int r = 0, rem = t;   // invert the packed upper-triangle index
      while (rem >= D - r) { rem -= D - r; ++r; }
      const int c = r + rem;
      T* blk = out_blocks + (size_t)cam * (D * D);
      blk[r * D + c] = v;
      blk[c * D + r] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// K2 (Schur off-diagonal): E_ij = sum over the pair list of block (i, j), i <= j, of
// Jc_a^T (V_a Jp_b^T) Jc_b.  A group of ceil(D/2) lanes owns one list; each lane owns two
// columns of the D x D block (2 D accumulators) and nothing is reduced across lanes.  The two 128-byte records of a pair are gathered with cp.async
// (16 B per lane, L2 only) into a per-group ring in shared memory, a few pairs ahead
// of the multiply: the dependent chain pair index -> record address -> record no longer
// serialises the list.  All groups of a warp run the same trip count (lists are visited in
// order of decreasing length) so the pipeline needs warp-level synchronisation only.  Only the upper triangle is
// stored; diagonal lists (same camera twice in a track) add onto E_ii written by the camera
// pass.  No atomics.
// ---------------------------------------------------------------------------------------
template <int D> struct SchurGroup {
  static constexpr int COLS = 2;                       // columns of the block owned by one lane
  static constexpr int LPL = (D + COLS - 1) / COLS;    // lanes per list
  static constexpr int PER_WARP = 32 / LPL;            // lists per warp
};
constexpr int SCHUR_TPB = 128;

template <typename T, int D>
__global__ void __launch_bounds__(SCHUR_TPB)
schur_offdiag_kernel(int64_t n_lists, const int32_t* __restrict__ list_order, const int64_t* __restrict__ list_off,
                     const uint64_t* __restrict__ pairs, const int32_t* __restrict__ list_slot,
                     const uint8_t* __restrict__ list_diag, const T* __restrict__ OBS, T* __restrict__ E) {
  constexpr int REC = ObsRec<D>::REC, OJP = ObsRec<D>::JP, OV = ObsRec<D>::V;
  constexpr int LPL = SchurGroup<D>::LPL, GPW = SchurGroup<D>::PER_WARP;
  constexpr int DEPTH = sizeof(T) == 4 ? 4 : 2;   // pairs in flight per list (ring in shared memory; 48 KB static limit)
  constexpr int VE = 16 / sizeof(T);              // elements per 16-byte vector
  constexpr int NVEC = REC / VE;                  // vectors per record
  // per-list ring [DEPTH][2][REC]; consecutive lists are offset by 4 extra words so that the
  // broadcast reads of the GPW lists of a warp fall into different banks
  constexpr int GS = DEPTH * 2 * REC + 4;
  __shared__ __align__(16) T ring_all[(SCHUR_TPB / 32) * GPW * GS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPL, gl = lane % LPL;
  const int c0 = 2 * gl, c1 = min(2 * gl + 1, D - 1);      // the last lane of an odd D repeats column D-1
  const int64_t warp = blockIdx.x * (int64_t)(SCHUR_TPB / 32) + w;
  const int64_t slot_in_order = warp * GPW + grp;
  const bool valid = grp < GPW && slot_in_order < n_lists;
  // lists are visited in row-major windows, by decreasing length inside a window: the GPW lists
  // of a warp have (almost) the same trip count and stay in the same one or two block rows
  const int64_t u = valid ? list_order[slot_in_order] : 0;
  const int64_t beg = valid ? list_off[u] : 0;
  const int len = valid ? (int)(list_off[u + 1] - beg) : 0;
  int maxlen = len;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
  T* ring = ring_all + (size_t)(w * GPW + (grp < GPW ? grp : 0)) * GS;

  auto issue = [&](int t) {
    if (t < len) {
      const uint64_t ab = __ldg(pairs + beg + t);
      const T* ra = OBS + (size_t)(uint32_t)(ab >> 32) * REC;
      const T* rb = OBS + (size_t)(uint32_t)ab * REC;
      T* da = ring + (t % DEPTH) * 2 * REC;
      T* db = da + REC;
      for (int v = gl; v < NVEC; v += LPL) { cp_async16(da + v * VE, ra + v * VE); cp_async16(db + v * VE, rb + v * VE); }
    }
    cp_async_commit();   // committed even when empty: every lane keeps the same group count
  };

  T acc0[D], acc1[D];
#pragma unroll
  for (int r = 0; r < D; ++r) { acc0[r] = T(0); acc1[r] = T(0); }
#pragma unroll
  for (int t = 0; t < DEPTH - 1; ++t) issue(t);
  for (int t = 0; t < maxlen; ++t) {
    issue(t + DEPTH - 1);
    cp_async_wait<DEPTH - 1>();
    __syncwarp();
    if (t < len) {
      const T* ra = ring + (t % DEPTH) * 2 * REC;
      const T* rb = ra + REC;
      const T* v = ra + OV;
      const T* j = rb + OJP;
      const T m00 = v[0] * j[0] + v[1] * j[1] + v[2] * j[2], m01 = v[0] * j[3] + v[1] * j[4] + v[2] * j[5];
      const T m10 = v[3] * j[0] + v[4] * j[1] + v[5] * j[2], m11 = v[3] * j[3] + v[4] * j[4] + v[5] * j[5];
      const T b00 = rb[c0], b10 = rb[D + c0], b01 = rb[c1], b11 = rb[D + c1];
      const T t00 = m00 * b00 + m01 * b10, t10 = m10 * b00 + m11 * b10;   // column c0 of M Jc_b
      const T t01 = m00 * b01 + m01 * b11, t11 = m10 * b01 + m11 * b11;   // column c1
#pragma unroll
      for (int r = 0; r < D; ++r) {
        const T a0 = ra[r], a1 = ra[D + r];
        acc0[r] += a0 * t00 + a1 * t10;
        acc1[r] += a0 * t01 + a1 * t11;
      }
    }
    __syncwarp();
  }
  if (!valid) return;
  T* blk = E + (size_t)list_slot[u] * (D * D);
  const bool two = 2 * gl + 1 < D;
  if (!list_diag[u]) {
#pragma unroll
    for (int r = 0; r < D; ++r) { blk[r * D + c0] = acc0[r]; if (two) blk[r * D + c1] = acc1[r]; }
  } else {
#pragma unroll
    for (int r = 0; r < D; ++r) { blk[r * D + c0] += acc0[r]; if (two) blk[r * D + c1] += acc1[r]; }
  }
}

// ---------------------------------------------------------------------------------------
// K2 (camera side, per trial): one CTA per camera over its observation records, one sweep for
// everything the reduced system needs from the diagonal:
//   HME_i = [ sum Jc^T (I - V Jp^T) Jc  |  diag(sum Jc^T Jc)  |  sum Jc^T rho ]
//         = [ Hcc_i - E_ii              |  diag(Hcc_i)        |  (g_c - e)_i  ]
// The difference Hcc - E_ii is formed per observation on the 2x2 factor I - M (M = Jp Hpp^-1 Jp^T
// has its eigenvalues in [0, 1)), not as the difference of two large sums: the cancellation
// that costs fp32 its digits on weakly observed cameras never happens.  diag(Hcc) is kept apart
// because the LM damping scales the clamped diagonal of Hcc only.  The record carries everything
// (Jc, Jp, V, rho): no gathers besides the record itself.  Also zeroes the diagonal slot of E
// (only duplicate-camera pair lists add to it).  No atomics.
// ---------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(CAM_TPB)
camera_schur_kernel(const int32_t* __restrict__ cam_off, const int32_t* __restrict__ cam_perm, const T* __restrict__ OBS,
                    T* __restrict__ HME, T* __restrict__ E, const int32_t* __restrict__ diag_slot) {
  constexpr int REC = ObsRec<D>::REC, OJP = ObsRec<D>::JP, OV = ObsRec<D>::V, ORHO = ObsRec<D>::RHO;
  constexpr int NQ = REC / 4;
  constexpr int NU = D * (D + 1) / 2;
  constexpr int NACC = NU + 2 * D;
  constexpr int NW = CAM_TPB / 32;
  constexpr int W = D * D + 2 * D;
  __shared__ T sh[NW][NACC];
  const int cam = blockIdx.x;
  const int beg = cam_off[cam], end = cam_off[cam + 1];
  T acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = T(0);
  for (int k = beg + threadIdx.x; k < end; k += CAM_TPB) {
    const int a = cam_perm[k];
    T rec[4 * NQ];
    load_quads<T, 0, NQ, true>(OBS + (size_t)a * REC, rec);
    const T* jc = rec;
    const T* v = rec + OV;
    const T* j = rec + OJP;
    const T x00 = T(1) - (v[0] * j[0] + v[1] * j[1] + v[2] * j[2]), x01 = -(v[0] * j[3] + v[1] * j[4] + v[2] * j[5]);
    const T x10 = -(v[3] * j[0] + v[4] * j[1] + v[5] * j[2]), x11 = T(1) - (v[3] * j[3] + v[4] * j[4] + v[5] * j[5]);
    const T s0 = rec[ORHO], s1 = rec[ORHO + 1];
    int u = 0;
#pragma unroll
    for (int r = 0; r < D; ++r) {
      const T l0 = jc[r] * x00 + jc[D + r] * x10, l1 = jc[r] * x01 + jc[D + r] * x11;
#pragma unroll
      for (int c = r; c < D; ++c) acc[u++] += l0 * jc[c] + l1 * jc[D + c];
      acc[NU + r] += jc[r] * jc[r] + jc[D + r] * jc[D + r];
      acc[NU + D + r] += jc[r] * s0 + jc[D + r] * s1;
    }
  }
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    T v = acc[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    acc[i] = v;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) sh[w][i] = acc[i];
  }
  __syncthreads();
  T* out = HME + (size_t)cam * W;
  for (int t = threadIdx.x; t < NACC; t += CAM_TPB) {
    T v = T(0);
#pragma unroll
    for (int k = 0; k < NW; ++k) v += sh[k][t];
    if (t >= NU) {
      out[D * D + (t - NU)] = v;
    } else {
      int r = 0, rem = t;
      while (rem >= D - r) { rem -= D - r; ++r; }
      const int c = r + rem;
      out[r * D + c] = v;
      out[c * D + r] = v;
    }
  }
  if (E) {
    T* blk = E + (size_t)diag_slot[cam] * (D * D);
    for (int t = threadIdx.x; t < D * D; t += CAM_TPB) blk[t] = T(0);
  }
}

// K3: S_ii = (Hcc - E_ii) + diag(clamp(diag Hcc) mu - diag Hcc); Minv = S_ii^-1; b = -(g_c - e).
// HD receives S_ii: the PCG operator is q = HD p - E p with the diagonal slot of E holding only
// the duplicate-camera contributions.
template <typename T, int D>
__global__ void precond_kernel(int n_cam, const T* __restrict__ HME, T mu, T* __restrict__ HD, T* __restrict__ MINV,
                               T* __restrict__ bvec, int* __restrict__ fail) {
  int cam = blockIdx.x * blockDim.x + threadIdx.x;
  if (cam >= n_cam) return;
  constexpr int W = D * D + 2 * D;
  double M[D * D];
  const T* h = HME + (size_t)cam * W;
#pragma unroll 1
  for (int r = 0; r < D; ++r)
    for (int c = 0; c < D; ++c) {
      double v = (double)h[r * D + c];
      if (r == c) { const T d = h[D * D + r]; v += (double)damp_diag(d, mu) - (double)d; }
      HD[(size_t)cam * (D * D) + r * D + c] = (T)v;
      M[r * D + c] = v;
    }
  if (!spd_inverse<D>(M)) {
    // rounding made S_ii indefinite (degenerate geometry): fall back to the damped diagonal of Hcc,
    // always positive -- a weaker preconditioner, the system itself is unchanged
    *fail = cam + 1;
#pragma unroll 1
    for (int k = 0; k < D * D; ++k) M[k] = 0.0;
#pragma unroll 1
    for (int r = 0; r < D; ++r) M[r * D + r] = 1.0 / (double)damp_diag(h[D * D + r], mu);
  }
#pragma unroll 1
  for (int k = 0; k < D * D; ++k) MINV[(size_t)cam * (D * D) + k] = (T)M[k];
#pragma unroll 1
  for (int k = 0; k < D; ++k) bvec[(size_t)cam * D + k] = -h[D * D + D + k];
}

// ---------------------------------------------------------------------------------------
// K5: back-substitution D_p = -Hpp^-1 (g_p + sum Jp^T (Jc D_c)), trial points, and the
// trust-region model term sum (JD)^T (2R + JD).  One thread per point, one sweep over its records.
// ---------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(BA_TPB)
backsub_kernel(int64_t n_pt, const int32_t* __restrict__ pt_off, const int32_t* __restrict__ cam_of,
               const T* __restrict__ OBS, const T* __restrict__ R, const T* __restrict__ GPT, const T* __restrict__ HPP,
               const T* __restrict__ HPPINV, const T* __restrict__ DC, const T* __restrict__ pts,
               T* __restrict__ pts_trial, T* __restrict__ DP, double* __restrict__ part_m) {
  constexpr int REC = ObsRec<D>::REC, OJP = ObsRec<D>::JP;
  constexpr int NQ = (2 * D + 6 + 3) / 4;
  double msum = 0.0;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_pt; p += (int64_t)gridDim.x * blockDim.x) {
    const int beg = pt_off[p], end = pt_off[p + 1];
    T u0 = GPT[3 * p], u1 = GPT[3 * p + 1], u2 = GPT[3 * p + 2];
    T A = T(0);   // sum_a w_a . (2 R_a + w_a),  w_a = Jc_a D_c
    for (int a = beg; a < end; ++a) {
      T rec[4 * NQ];
      load_quads<T, 0, NQ, true>(OBS + (size_t)a * REC, rec);
      const T* dc = DC + (size_t)cam_of[a] * D;
      T w0 = T(0), w1 = T(0);
#pragma unroll
      for (int c = 0; c < D; ++c) { T d = __ldg(dc + c); w0 += rec[c] * d; w1 += rec[D + c] * d; }
      const T* j = rec + OJP;
      u0 += j[0] * w0 + j[3] * w1; u1 += j[1] * w0 + j[4] * w1; u2 += j[2] * w0 + j[5] * w1;
      A += w0 * (2 * R[2 * (size_t)a] + w0) + w1 * (2 * R[2 * (size_t)a + 1] + w1);
    }
    const T* iv = HPPINV + (size_t)p * 6;
    T d0 = -(iv[0] * u0 + iv[1] * u1 + iv[2] * u2);
    T d1 = -(iv[1] * u0 + iv[3] * u1 + iv[4] * u2);
    T d2 = -(iv[2] * u0 + iv[4] * u1 + iv[5] * u2);
    DP[3 * p] = d0; DP[3 * p + 1] = d1; DP[3 * p + 2] = d2;
    pts_trial[3 * p] = pts[3 * p] + d0; pts_trial[3 * p + 1] = pts[3 * p + 1] + d1; pts_trial[3 * p + 2] = pts[3 * p + 2] + d2;
    // sum_a JD_a . (2 R_a + JD_a) with JD_a = w_a + Jp_a D_p
    //   = A + 2 D_p . (g_p + sum Jp_a^T w_a) + D_p^T (sum Jp_a^T Jp_a) D_p = A + 2 D_p . u + D_p^T Hpp D_p
    // (Hpp undamped) -- no second sweep over the observation records.
    const T* h = HPP + (size_t)p * 6;
    T hd0 = h[0] * d0 + h[1] * d1 + h[2] * d2, hd1 = h[1] * d0 + h[3] * d1 + h[4] * d2, hd2 = h[2] * d0 + h[4] * d1 + h[5] * d2;
    msum += (double)A + 2.0 * ((double)d0 * u0 + (double)d1 * u1 + (double)d2 * u2) + ((double)d0 * hd0 + (double)d1 * hd1 + (double)d2 * hd2);
  }
  msum = block_sum(msum);
  if (threadIdx.x == 0) part_m[blockIdx.x] = msum;
}

// x <- x (+) D for cameras: Exp(D[:6]) * pose, intrinsics += D[6:]   (bae update_parameter)
template <typename T, int NI>
__global__ void camera_update_kernel(int n_cam, const T* __restrict__ cam, const T* __restrict__ pp, const T* __restrict__ DC,
                                     T* __restrict__ cam_trial, T* __restrict__ camq_trial, double* __restrict__ part_norm) {
  constexpr int D = 6 + NI, CW = 7 + NI, CWP = CamPack<CW>::CWP;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double nrm = 0.0;
  if (i < n_cam) {
    T d[D], in[CW], out[CW];
#pragma unroll
    for (int k = 0; k < D; ++k) { d[k] = DC[(size_t)i * D + k]; nrm += (double)d[k] * (double)d[k]; }
#pragma unroll
    for (int k = 0; k < CW; ++k) in[k] = cam[(size_t)i * CW + k];
    se3_retract(in, d, out);
#pragma unroll
    for (int k = 0; k < NI; ++k) out[7 + k] = in[7 + k] + d[6 + k];
#pragma unroll
    for (int k = 0; k < CW; ++k) { cam_trial[(size_t)i * CW + k] = out[k]; camq_trial[(size_t)i * CWP + k] = out[k]; }
    camq_trial[(size_t)i * CWP + CW] = pp[2 * (size_t)i]; camq_trial[(size_t)i * CWP + CW + 1] = pp[2 * (size_t)i + 1];
#pragma unroll
    for (int k = CW + 2; k < CWP; ++k) camq_trial[(size_t)i * CWP + k] = T(0);
  }
  nrm = block_sum(nrm);
  if (threadIdx.x == 0) part_norm[blockIdx.x] = nrm;
}

// ---------------------------------------------------------------------------------------
// K1 + K2 (point side) + K3 (3x3) fused for the first trial of an LM step.
// A CTA owns a run of whole points with at most FUSED_TPB observations (greedy packing done
// at set-up: `cta_pt` holds the point boundaries), one thread per observation:
//   1. residual, Huber weight, weighted Jc / Jp in registers; the observation's contribution
//      to Hpp (6) and g_p (3) goes to shared memory;
//   2. one thread per point sums its (contiguous) contributions -- the segmented reduction
//      over observations sorted by point --, writes Hpp / g_p, damps, inverts, t = Hpp^-1 g_p;
//   3. V = Jp Hpp^-1 per observation; the complete 16-byte-aligned record (Jc | Jp | V) is
//      staged in shared memory (odd quad stride: conflict-free) and the CTA's contiguous slab
//      of OBS is written with fully coalesced 128-bit stores.
// Replaces linearize_kernel + point_solve_kernel<BUILD> when no track exceeds FUSED_TPB
// observations; rejected trials re-damp with point_solve_kernel<false>.
// ---------------------------------------------------------------------------------------
#ifndef ISFM_FUSED_TPB
#define ISFM_FUSED_TPB 256
#endif
constexpr int FUSED_TPB = ISFM_FUSED_TPB;
// phase-ablation switches of the timing harness (tools/kbench.cu); compiled out of the library
#ifdef ISFM_KBENCH
#define ISFM_ABL(bit) ((dbg & (bit)) != 0)
#else
#define ISFM_ABL(bit) false
#endif
template <typename T, int D> struct FusedCfg {
  static constexpr int REC = ObsRec<D>::REC;
  static constexpr int QR = REC / 4;            // quads per record
  static constexpr int QV = ObsRec<D>::V / 4;   // quads [0, QV): Jc | Jp only (staged); [QV, QR): end of Jp, V, padding
  // staged row: the record's QR quads + three quads of contributions to Hpp (6) and g_p (3); odd stride
  static constexpr int QC = QR;                 // first contribution quad
  static constexpr int SQ = (QR + 3) | 1;
  static constexpr size_t SMEM = (size_t)FUSED_TPB * SQ * 4 * sizeof(T);
};

template <typename T, int MODEL>
__global__ void __launch_bounds__(FUSED_TPB, sizeof(T) == 4 ? 4 * (256 / FUSED_TPB) : 1)
fused_linearize_kernel(const int4* __restrict__ tiles, const int32_t* __restrict__ pt_off,
                       const T* __restrict__ camq, const T* __restrict__ pts,
                       const T* __restrict__ obs, const int32_t* __restrict__ cam_of, const int32_t* __restrict__ pt_of,
                       T delta, T mu, T* __restrict__ R, T* __restrict__ OBS, T* __restrict__ HPP, T* __restrict__ GPT,
                       T* __restrict__ HPPINV, T* __restrict__ TP, double* __restrict__ part_rho,
                       double* __restrict__ part_sq, int want_cost, int dbg) {
  constexpr int NI = ModelTraits<MODEL>::NI;
  constexpr int D = 6 + NI;
  constexpr int CW = 7 + NI;
  typedef FusedCfg<T, D> Cfg;
  constexpr int REC = Cfg::REC, QR = Cfg::QR, SQ = Cfg::SQ, OJP = ObsRec<D>::JP, OV = ObsRec<D>::V;
  constexpr int QV = Cfg::QV, QC = Cfg::QC;
  extern __shared__ __align__(16) unsigned char fused_smem[];
  T* stage = reinterpret_cast<T*>(fused_smem);                 // [FUSED_TPB][SQ * 4]
  const int t = threadIdx.x;
  double rho_d = 0.0, sq_d = 0.0;
  const int4 tile = __ldg(tiles + blockIdx.x);   // {first point, end point, first observation, observations}
  const int o0 = tile.z, n = tile.w;
  const int64_t a = (int64_t)o0 + t;
  T* srow = stage + (size_t)t * SQ * 4;
  // a point without observations (possible in caller data) has no thread below: write its blocks here
  if (t < tile.y - tile.x) {
    const int pz = tile.x + t;
    if (pt_off[pz + 1] == pt_off[pz]) {
      const T dinv = T(1) / damp_diag(T(0), mu);
#pragma unroll
      for (int i = 0; i < 6; ++i) { HPP[(size_t)pz * 6 + i] = T(0); HPPINV[(size_t)pz * 6 + i] = (i == 0 || i == 3 || i == 5) ? dinv : T(0); }
#pragma unroll
      for (int i = 0; i < 3; ++i) { GPT[(size_t)pz * 3 + i] = T(0); TP[(size_t)pz * 3 + i] = T(0); }
    }
  }
  T tail[REC - 4 * QV];          // record elements [4 QV, REC): the end of Jp, V, rho, padding
  T rw0 = T(0), rw1 = T(0);
  int p = 0, kb = 0, ke = 0;
  if (t < n) {
    int c = cam_of[a];
    p = pt_of[a];
    kb = pt_off[p] - o0; ke = pt_off[p + 1] - o0;
    if (ISFM_ABL(8)) { c = t & 7; }
    const int pg = ISFM_ABL(8) ? (t & 15) : p;
    T cr[CW], ppv[2], X[3], o[2], r[2], rec[REC];
    load_camera<T, CW>(camq, c, cr, ppv);
#pragma unroll
    for (int i = 0; i < 3; ++i) X[i] = __ldg(pts + 3 * (size_t)pg + i);
    o[0] = obs[2 * a]; o[1] = obs[2 * a + 1];
    if (ISFM_ABL(4)) {
#pragma unroll
      for (int i = 0; i < REC; ++i) rec[i] = cr[i % CW] + X[i % 3] + o[i & 1] + ppv[i & 1];
      r[0] = rec[0]; r[1] = rec[1];
    } else
    ba_linearize<MODEL, T>(cr, ppv, X, o, r, rec, rec + OJP);
    T s = r[0] * r[0] + r[1] * r[1], rho, w;
    huber(s, delta, rho, w);
    rho_d += (double)rho; sq_d += (double)s;
    rw0 = w * r[0]; rw1 = w * r[1];
#pragma unroll
    for (int i = 0; i < 2 * D + 6; ++i) rec[i] *= w;
#pragma unroll
    for (int i = OV; i < REC; ++i) rec[i] = T(0);
    const T* j = rec + OJP;
    T cb[12];
    cb[0] = j[0] * j[0] + j[3] * j[3]; cb[1] = j[0] * j[1] + j[3] * j[4]; cb[2] = j[0] * j[2] + j[3] * j[5];
    cb[3] = j[1] * j[1] + j[4] * j[4]; cb[4] = j[1] * j[2] + j[4] * j[5]; cb[5] = j[2] * j[2] + j[5] * j[5];
    cb[6] = j[0] * rw0 + j[3] * rw1; cb[7] = j[1] * rw0 + j[4] * rw1; cb[8] = j[2] * rw0 + j[5] * rw1;
    cb[9] = cb[10] = cb[11] = T(0);
#pragma unroll
    for (int q = 0; q < 3; ++q) QuadIO<T>::st(srow + 4 * (QC + q), cb + 4 * q);
    if (!ISFM_ABL(32)) { R[2 * a] = rw0; R[2 * a + 1] = rw1; }
    // Jc and the quads fully covered by Jp leave the registers now; only the tail stays live
#pragma unroll
    for (int q = 0; q < QV; ++q) QuadIO<T>::st(srow + 4 * q, rec + 4 * q);
#pragma unroll
    for (int i = 0; i < REC - 4 * QV; ++i) tail[i] = rec[4 * QV + i];
  }
  __syncthreads();   // staged Jc | Jp quads and contributions are complete
  // Every observation thread sums the contributions of its own point (same order in every thread:
  // bit-identical, and no serial phase on a quarter of the warps), inverts the damped 3x3 block and
  // forms V = Jp Hpp^-1; the thread of the point's first observation stores the point arrays.
  if (t < n && !ISFM_ABL(2)) {
    T h[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
    for (int k = kb; k < ke; ++k) {
      T cb[12];
#pragma unroll
      for (int q = 0; q < 3; ++q) QuadIO<T>::ld(stage + ((size_t)k * SQ + QC + q) * 4, cb + 4 * q);
#pragma unroll
      for (int i = 0; i < 6; ++i) h[i] += cb[i];
      g[0] += cb[6]; g[1] += cb[7]; g[2] += cb[8];
    }
    const bool head = t == kb && !ISFM_ABL(32);
    if (head) {
#pragma unroll
      for (int i = 0; i < 6; ++i) HPP[(size_t)p * 6 + i] = h[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) GPT[(size_t)p * 3 + i] = g[i];
    }
    h[0] = damp_diag(h[0], mu); h[3] = damp_diag(h[3], mu); h[5] = damp_diag(h[5], mu);
    T iv[6];
    sym3_inverse(h, iv);
    const T t0 = iv[0] * g[0] + iv[1] * g[1] + iv[2] * g[2];
    const T t1 = iv[1] * g[0] + iv[3] * g[1] + iv[4] * g[2];
    const T t2 = iv[2] * g[0] + iv[4] * g[1] + iv[5] * g[2];
    if (head) {
#pragma unroll
      for (int i = 0; i < 6; ++i) HPPINV[(size_t)p * 6 + i] = iv[i];
      TP[(size_t)p * 3 + 0] = t0; TP[(size_t)p * 3 + 1] = t1; TP[(size_t)p * 3 + 2] = t2;
    }
    // Jp: elements [OJP, OJP + 6) of the record; those below 4 QV were staged, re-read them
    T jp[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) jp[i] = (OJP + i >= 4 * QV) ? tail[OJP + i - 4 * QV] : srow[OJP + i];
#pragma unroll
    for (int row = 0; row < 2; ++row) {
      T a0 = jp[3 * row], a1 = jp[3 * row + 1], a2 = jp[3 * row + 2];
      tail[OV - 4 * QV + 3 * row + 0] = a0 * iv[0] + a1 * iv[1] + a2 * iv[2];
      tail[OV - 4 * QV + 3 * row + 1] = a0 * iv[1] + a1 * iv[3] + a2 * iv[4];
      tail[OV - 4 * QV + 3 * row + 2] = a0 * iv[2] + a1 * iv[4] + a2 * iv[5];
    }
    tail[ObsRec<D>::RHO - 4 * QV] = rw0 - (jp[0] * t0 + jp[1] * t1 + jp[2] * t2);
    tail[ObsRec<D>::RHO - 4 * QV + 1] = rw1 - (jp[3] * t0 + jp[4] * t1 + jp[5] * t2);
#pragma unroll
    for (int q = QV; q < QR; ++q) QuadIO<T>::st(srow + 4 * q, tail + 4 * (q - QV));
  }
  __syncthreads();
  // coalesced write of the CTA's slab: slab quad i = record i / QR, quad i % QR (n <= FUSED_TPB
  // records = at most QR rounds; a thread's shared-memory loads all precede its stores)
  if (!ISFM_ABL(1)) {
    T* dst = OBS + (size_t)o0 * REC;
    T q4[QR][4];
#pragma unroll
    for (int u = 0; u < QR; ++u) {
      const int i = t + u * FUSED_TPB;
      if (i < n * QR) QuadIO<T>::ld(stage + ((size_t)(i / QR) * SQ + (i % QR)) * 4, q4[u]);
    }
#pragma unroll
    for (int u = 0; u < QR; ++u) {
      const int i = t + u * FUSED_TPB;
      if (i < n * QR) QuadIO<T>::st(dst + (size_t)i * 4, q4[u]);
    }
  }
  // cost partials: wanted on the first LM step only (later steps carry the accepted trial cost)
  if (!want_cost || ISFM_ABL(16)) return;
  rho_d = block_sum(rho_d);
  sq_d = block_sum(sq_d);
  if (t == 0) { part_rho[blockIdx.x] = rho_d; part_sq[blockIdx.x] = sq_d; }
}

// ---------------------------------------------------------------------------------------
// The same kernel with the record slab leaving shared memory through the TMA (fp32, 128-byte
// records).  In the kernel above every thread re-reads eight quads from the staging tile and
// stores them to HBM (8 LDS.128 + 8 STG.128 per observation) only to make the stores coalesced;
// that kernel is bound by the LSU pipe (78 % of the wavefront budget), not by HBM.  Here the tile
// is laid out as the TMA's SWIZZLE_128B pattern -- 16-byte chunk c of row t lives at chunk
// c ^ (t & 7): the thread-per-record STS.128 are conflict-free without padding -- and the slab is
// stored by bulk tensor copies (cp.async.bulk.tensor.2d, boxes of 8 records = 1 KB, issued by the
// first lanes, un-swizzled by the copy engine); only the < 8 records of a tile's tail are stored
// by threads.  `tmap`: 2-D tensor map over OBS ([n_obs][32] fp32, box {32, 8}, SWIZZLE_128B).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_2d(const void* tmap, int c0, int c1, const void* smem_src) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_src);
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(tmap), "r"(c0), "r"(c1), "r"(sa) : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct FusedTmaCfg {
  static constexpr int REC = 32;                 // floats per record (128 B)
  static constexpr int BOX_ROWS = 8;             // records per bulk store (1 KB)
  static constexpr size_t TILE_BYTES = (size_t)FUSED_TPB * REC * 4;
  static constexpr size_t PART_BYTES = (size_t)(FUSED_TPB / 32) * 2 * 12 * 4;   // per warp: parts of its first and last point (Hpp 6 | g_p 3 | pad)
  static constexpr size_t SMEM = TILE_BYTES + PART_BYTES + 1024;                // + slack to align the tile to 1 KB
};

#ifndef ISFM_FUSED_TMA_MINB
#define ISFM_FUSED_TMA_MINB 4
#endif
template <int MODEL>
__global__ void __launch_bounds__(FUSED_TPB, ISFM_FUSED_TMA_MINB * (256 / FUSED_TPB))
fused_linearize_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_big, int big_rows,
                           const int4* __restrict__ tiles,
                           const int32_t* __restrict__ pt_off, const float* __restrict__ camq, const float* __restrict__ pts,
                           const float* __restrict__ obs, const int32_t* __restrict__ cam_of, const int32_t* __restrict__ pt_of,
                           float delta, float mu, float* __restrict__ R, float* __restrict__ OBS, float* __restrict__ HPP,
                           float* __restrict__ GPT, float* __restrict__ HPPINV, float* __restrict__ TP,
                           double* __restrict__ part_rho, double* __restrict__ part_sq, int want_cost) {
  typedef float T;
  constexpr int NI = ModelTraits<MODEL>::NI;
  constexpr int D = 6 + NI;
  constexpr int CW = 7 + NI;
  constexpr int REC = ObsRec<D>::REC, QR = REC / 4, OJP = ObsRec<D>::JP, OV = ObsRec<D>::V, QV = OV / 4;
  static_assert(REC == FusedTmaCfg::REC, "TMA path: 128-byte records only");
  extern __shared__ __align__(16) unsigned char fused_smem[];
  // tile: [FUSED_TPB][32] floats, 1 KB aligned (the swizzle is a function of the shared-memory address)
  const unsigned base_sa = (unsigned)__cvta_generic_to_shared(fused_smem);
  T* tile = reinterpret_cast<T*>(fused_smem + (((base_sa + 1023u) & ~1023u) - base_sa));
  T* contrib = tile + FUSED_TPB * REC;                       // [FUSED_TPB / 32][2][12]: per-warp parts of the point sums
  const int t = threadIdx.x;
  const int sw = t & 7;
  T* row = tile + (size_t)t * REC;
  auto chunk = [&](int q) -> T* { return row + ((q ^ sw) << 2); };
  double rho_d = 0.0, sq_d = 0.0;
  const int4 tl = __ldg(tiles + blockIdx.x);   // {first point, end point, first observation, observations}
  const int o0 = tl.z, n = tl.w;
  const int64_t a = (int64_t)o0 + t;
  if (t < tl.y - tl.x) {   // a point without observations has no thread below: write its blocks here
    const int pz = tl.x + t;
    if (pt_off[pz + 1] == pt_off[pz]) {
      const T dinv = T(1) / damp_diag(T(0), mu);
#pragma unroll
      for (int i = 0; i < 6; ++i) { HPP[(size_t)pz * 6 + i] = T(0); HPPINV[(size_t)pz * 6 + i] = (i == 0 || i == 3 || i == 5) ? dinv : T(0); }
#pragma unroll
      for (int i = 0; i < 3; ++i) { GPT[(size_t)pz * 3 + i] = T(0); TP[(size_t)pz * 3 + i] = T(0); }
    }
  }
  T tail[REC - 4 * QV];          // record elements [4 QV, REC): the end of Jp, V, rho, padding
  T rw0 = T(0), rw1 = T(0);
  // this observation's contribution to Hpp (6) and g_p (3); a thread without observation is a
  // segment of its own with a zero contribution
  T cb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int p = 0, kb = t, ke = t + 1;
  if (t < n) {
    const int c = cam_of[a];
    p = pt_of[a];
    kb = pt_off[p] - o0; ke = pt_off[p + 1] - o0;
    T cr[CW], ppv[2], X[3], o[2], r[2], rec[REC];
    load_camera<T, CW>(camq, c, cr, ppv);
#pragma unroll
    for (int i = 0; i < 3; ++i) X[i] = __ldg(pts + 3 * (size_t)p + i);
    o[0] = obs[2 * a]; o[1] = obs[2 * a + 1];
    ba_linearize<MODEL, T>(cr, ppv, X, o, r, rec, rec + OJP);
    T s = r[0] * r[0] + r[1] * r[1], rho, w;
    huber(s, delta, rho, w);
    rho_d += (double)rho; sq_d += (double)s;
    rw0 = w * r[0]; rw1 = w * r[1];
#pragma unroll
    for (int i = 0; i < 2 * D + 6; ++i) rec[i] *= w;
#pragma unroll
    for (int i = OV; i < REC; ++i) rec[i] = T(0);
    const T* j = rec + OJP;
    cb[0] = j[0] * j[0] + j[3] * j[3]; cb[1] = j[0] * j[1] + j[3] * j[4]; cb[2] = j[0] * j[2] + j[3] * j[5];
    cb[3] = j[1] * j[1] + j[4] * j[4]; cb[4] = j[1] * j[2] + j[4] * j[5]; cb[5] = j[2] * j[2] + j[5] * j[5];
    cb[6] = j[0] * rw0 + j[3] * rw1; cb[7] = j[1] * rw0 + j[4] * rw1; cb[8] = j[2] * rw0 + j[5] * rw1;
    R[2 * a] = rw0; R[2 * a + 1] = rw1;
#pragma unroll
    for (int q = 0; q < QV; ++q) QuadIO<T>::st(chunk(q), rec + 4 * q);
#pragma unroll
    for (int i = 0; i < REC - 4 * QV; ++i) tail[i] = rec[4 * QV + i];
  }
  // Segmented sum over the observations of a point (they are consecutive threads).  Inside a warp:
  // inclusive scan by shuffles restricted to the segment, then the value of the segment's last lane
  // = the part of the point that lives in this warp.  Every warp publishes the part of its first
  // and of its last segment; a point that straddles warps adds the published parts in warp order
  // (the same sequence in every thread of the point: bit-identical Hpp in all of them).  The old
  // loop ran to the longest track of each warp (~15 trips of 3 LDS.128 + 9 FADD for a mean track
  // length of 5); this is 5 shuffle steps.
  {
    const int lane = t & 31, wbase = t & ~31, w = t >> 5;
    const int lo = max(kb, wbase) - wbase, hi = min(ke, wbase + 32) - 1 - wbase;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const bool take = lane - d >= lo;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const T v = __shfl_up_sync(0xffffffffu, cb[i], d);
        if (take) cb[i] += v;
      }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) cb[i] = __shfl_sync(0xffffffffu, cb[i], hi);
    if (lane == 0 || lane == 31) {
      T* dst = contrib + (size_t)(2 * w + (lane == 31 ? 1 : 0)) * 12;
#pragma unroll
      for (int i = 0; i < 9; ++i) dst[i] = cb[i];
    }
  }
  __syncthreads();   // per-warp parts are published
  if (t < n) {
    T h[6], g[3];
    const int wb = kb >> 5, we = (ke - 1) >> 5;
    if (wb == we) {
#pragma unroll
      for (int i = 0; i < 6; ++i) h[i] = cb[i];
      g[0] = cb[6]; g[1] = cb[7]; g[2] = cb[8];
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) h[i] = T(0);
      g[0] = g[1] = g[2] = T(0);
      for (int ww = wb; ww <= we; ++ww) {
        // first warp of the point: the part of that warp's LAST segment; later warps: of their FIRST
        const T* src = contrib + (size_t)(2 * ww + (ww == wb ? 1 : 0)) * 12;
        T c4[12];
#pragma unroll
        for (int q = 0; q < 3; ++q) QuadIO<T>::ld(src + 4 * q, c4 + 4 * q);
#pragma unroll
        for (int i = 0; i < 6; ++i) h[i] += c4[i];
        g[0] += c4[6]; g[1] += c4[7]; g[2] += c4[8];
      }
    }
    const bool head = t == kb;
    if (head) {
#pragma unroll
      for (int i = 0; i < 6; ++i) HPP[(size_t)p * 6 + i] = h[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) GPT[(size_t)p * 3 + i] = g[i];
    }
    h[0] = damp_diag(h[0], mu); h[3] = damp_diag(h[3], mu); h[5] = damp_diag(h[5], mu);
    T iv[6];
    sym3_inverse(h, iv);
    const T t0 = iv[0] * g[0] + iv[1] * g[1] + iv[2] * g[2];
    const T t1 = iv[1] * g[0] + iv[3] * g[1] + iv[4] * g[2];
    const T t2 = iv[2] * g[0] + iv[4] * g[1] + iv[5] * g[2];
    if (head) {
#pragma unroll
      for (int i = 0; i < 6; ++i) HPPINV[(size_t)p * 6 + i] = iv[i];
      TP[(size_t)p * 3 + 0] = t0; TP[(size_t)p * 3 + 1] = t1; TP[(size_t)p * 3 + 2] = t2;
    }
    // Jp: elements [OJP, OJP + 6) of the record; those below 4 QV were staged (swizzled), re-read them
    T jp[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int e = OJP + i;
      jp[i] = (e >= 4 * QV) ? tail[e - 4 * QV] : chunk(e >> 2)[e & 3];
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      T a0 = jp[3 * rr], a1 = jp[3 * rr + 1], a2 = jp[3 * rr + 2];
      tail[OV - 4 * QV + 3 * rr + 0] = a0 * iv[0] + a1 * iv[1] + a2 * iv[2];
      tail[OV - 4 * QV + 3 * rr + 1] = a0 * iv[1] + a1 * iv[3] + a2 * iv[4];
      tail[OV - 4 * QV + 3 * rr + 2] = a0 * iv[2] + a1 * iv[4] + a2 * iv[5];
    }
    tail[ObsRec<D>::RHO - 4 * QV] = rw0 - (jp[0] * t0 + jp[1] * t1 + jp[2] * t2);
    tail[ObsRec<D>::RHO - 4 * QV + 1] = rw1 - (jp[3] * t0 + jp[4] * t1 + jp[5] * t2);
#pragma unroll
    for (int q = QV; q < QR; ++q) QuadIO<T>::st(chunk(q), tail + 4 * (q - QV));
  }
  fence_proxy_async_smem();   // this thread's tile writes become visible to the copy engine ...
  __syncthreads();            // ... and every thread has made them
  // big boxes (big_rows records, a multiple of 8; 0 = none) first, then boxes of 8 records
  const int n_big = big_rows > 0 ? n / big_rows : 0;
  const int rb = n_big * big_rows;
  const int n_box = (n - rb) / FusedTmaCfg::BOX_ROWS;
  // one thread issues every store of the tile and waits once: the stores pipeline in the copy
  // engine (issued from several lanes they would be serialised together with their waits); the
  // thread sits in the LAST warp so that the hand-written tail below is not held up behind it
  if (t == FUSED_TPB - 1) {
    for (int i = 0; i < n_big; ++i) tma_store_2d(&tmap_big, 0, o0 + i * big_rows, tile + (size_t)i * big_rows * REC);
    for (int i = 0; i < n_box; ++i)
      tma_store_2d(&tmap, 0, o0 + rb + i * FusedTmaCfg::BOX_ROWS, tile + (size_t)(rb + i * FusedTmaCfg::BOX_ROWS) * REC);
    tma_store_commit_and_wait_read();   // shared memory must stay intact until the engine has read it
  }
  // tail of the tile (< 8 records): un-swizzle by hand, one quad per thread
  const int r0 = rb + n_box * FusedTmaCfg::BOX_ROWS;
  if (t < (n - r0) * QR) {
    const int rr = r0 + t / QR, q = t % QR;
    T v[4];
    QuadIO<T>::ld(tile + (size_t)rr * REC + ((q ^ (rr & 7)) << 2), v);
    QuadIO<T>::st(OBS + ((size_t)o0 + rr) * REC + 4 * q, v);
  }
  if (!want_cost) return;
  rho_d = block_sum(rho_d);
  sq_d = block_sum(sq_d);
  if (t == 0) { part_rho[blockIdx.x] = rho_d; part_sq[blockIdx.x] = sq_d; }
}

// ---------------------------------------------------------------------------------------
// K5 on the tiles of the fused K1 (whole points, <= FUSED_TPB observations per CTA), one thread
// per observation.  The thread-per-point sweep of backsub_kernel reads every record through 32
// different cache lines per warp instruction; here the Jc | Jp quads of the tile's contiguous
// record slab are copied to shared memory with coalesced 16-byte cp.async, the per-observation
// products Jp^T (Jc D_c) and w.(2R + w) go back to shared memory as one quad, and one thread
// per point sums its (contiguous) quads, solves for D_p and accumulates the model term.
// DCQ: the camera step padded to rows of DQ = 4 ceil(D / 4) elements (128-bit gathers).
// ---------------------------------------------------------------------------------------
template <typename T, int D> struct BacksubCfg {
  static constexpr int REC = ObsRec<D>::REC;
  static constexpr int QJ = (2 * D + 6 + 3) / 4;   // quads covering Jc | Jp
  static constexpr int SQ = (QJ + 1) | 1;          // + one quad of products; odd stride
  static constexpr int DQ = (D + 3) / 4 * 4;
  static constexpr size_t SMEM = (size_t)FUSED_TPB * SQ * 4 * sizeof(T);
};

template <typename T>
__global__ void pad_rows_kernel(int n, int W, int WP, const T* __restrict__ src, T* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * WP) return;
  const int r = i / WP, k = i % WP;
  dst[i] = k < W ? src[(size_t)r * W + k] : T(0);
}

template <typename T, int D>
__global__ void __launch_bounds__(FUSED_TPB)
backsub_tiles_kernel(const int4* __restrict__ tiles, const int32_t* __restrict__ pt_off, const int32_t* __restrict__ cam_of,
                     const T* __restrict__ OBS, const T* __restrict__ R, const T* __restrict__ GPT, const T* __restrict__ HPP,
                     const T* __restrict__ HPPINV, const T* __restrict__ DCQ, const T* __restrict__ pts,
                     T* __restrict__ pts_trial, T* __restrict__ DP, double* __restrict__ part_m) {
  typedef BacksubCfg<T, D> Cfg;
  constexpr int REC = Cfg::REC, QJ = Cfg::QJ, SQ = Cfg::SQ, DQ = Cfg::DQ, OJP = ObsRec<D>::JP;
  extern __shared__ __align__(16) unsigned char backsub_smem[];
  T* stage = reinterpret_cast<T*>(backsub_smem);   // [FUSED_TPB][SQ * 4]
  const int t = threadIdx.x;
  const int4 tile = __ldg(tiles + blockIdx.x);
  const int o0 = tile.z, n = tile.w, npts = tile.y - tile.x;
  {
    const T* src = OBS + (size_t)o0 * REC;
#pragma unroll
    for (int u = 0; u < QJ; ++u) {
      const int i = t + u * FUSED_TPB;
      if (i < n * QJ) cp_async_quad<T>(stage + ((size_t)(i / QJ) * SQ + (i % QJ)) * 4, src + (size_t)(i / QJ) * REC + (i % QJ) * 4);
    }
    cp_async_commit();
  }
  T dc[DQ], r0 = T(0), r1 = T(0);
  if (t < n) {
    const int64_t a = (int64_t)o0 + t;
    const int c = cam_of[a];
#pragma unroll
    for (int q = 0; q < DQ / 4; ++q) QuadIO<T>::ldg(DCQ + (size_t)c * DQ + 4 * q, dc + 4 * q);
    r0 = R[2 * a]; r1 = R[2 * a + 1];
  }
  int kb = 0, ke = 0;
  T u0 = T(0), u1 = T(0), u2 = T(0);
  const int p = tile.x + t;
  if (t < npts) {
    kb = pt_off[p] - o0; ke = pt_off[p + 1] - o0;
    u0 = GPT[3 * (size_t)p]; u1 = GPT[3 * (size_t)p + 1]; u2 = GPT[3 * (size_t)p + 2];
  }
  cp_async_wait<0>();
  __syncthreads();
  T* srow = stage + (size_t)t * SQ * 4;
  if (t < n) {
    T rec[4 * QJ];
#pragma unroll
    for (int q = 0; q < QJ; ++q) QuadIO<T>::ld(srow + 4 * q, rec + 4 * q);
    T w0 = T(0), w1 = T(0);
#pragma unroll
    for (int c = 0; c < D; ++c) { w0 += rec[c] * dc[c]; w1 += rec[D + c] * dc[c]; }
    const T* j = rec + OJP;
    T prod[4];
    prod[0] = j[0] * w0 + j[3] * w1; prod[1] = j[1] * w0 + j[4] * w1; prod[2] = j[2] * w0 + j[5] * w1;
    prod[3] = w0 * (2 * r0 + w0) + w1 * (2 * r1 + w1);
    QuadIO<T>::st(srow + 4 * QJ, prod);
  }
  __syncthreads();
  double msum = 0.0;
  if (t < npts) {
    T A = T(0);   // sum_a w_a . (2 R_a + w_a),  w_a = Jc_a D_c
    for (int k = kb; k < ke; ++k) {
      T prod[4];
      QuadIO<T>::ld(stage + ((size_t)k * SQ + QJ) * 4, prod);
      u0 += prod[0]; u1 += prod[1]; u2 += prod[2]; A += prod[3];
    }
    const T* iv = HPPINV + (size_t)p * 6;
    T d0 = -(iv[0] * u0 + iv[1] * u1 + iv[2] * u2);
    T d1 = -(iv[1] * u0 + iv[3] * u1 + iv[4] * u2);
    T d2 = -(iv[2] * u0 + iv[4] * u1 + iv[5] * u2);
    DP[3 * (size_t)p] = d0; DP[3 * (size_t)p + 1] = d1; DP[3 * (size_t)p + 2] = d2;
    pts_trial[3 * (size_t)p] = pts[3 * (size_t)p] + d0; pts_trial[3 * (size_t)p + 1] = pts[3 * (size_t)p + 1] + d1;
    pts_trial[3 * (size_t)p + 2] = pts[3 * (size_t)p + 2] + d2;
    // model term, closed form (see backsub_kernel): A + 2 D_p . u + D_p^T Hpp D_p
    const T* h = HPP + (size_t)p * 6;
    T hd0 = h[0] * d0 + h[1] * d1 + h[2] * d2, hd1 = h[1] * d0 + h[3] * d1 + h[4] * d2, hd2 = h[2] * d0 + h[4] * d1 + h[5] * d2;
    msum = (double)A + 2.0 * ((double)d0 * u0 + (double)d1 * u1 + (double)d2 * u2) + ((double)d0 * hd0 + (double)d1 * hd1 + (double)d2 * hd2);
  }
  msum = block_sum(msum);
  if (t == 0) part_m[blockIdx.x] = msum;
}

// rows of y = E p that can be non-zero on this rank: upper rows with an off-diagonal block and the
// columns of those blocks (their transposed products); diagonal lists (duplicate cameras in a track)
static __global__ void touched_rows_kernel(int n_cam, const int32_t* __restrict__ urow_ptr, const int32_t* __restrict__ ucol, uint8_t* touched) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n_cam) return;
  bool any = false;
  for (int e = urow_ptr[row] + lane; e < urow_ptr[row + 1]; e += 32) {
    const int j = ucol[e];
    if (j != row) { any = true; touched[j] = 1; }   // same value from every writer
  }
  if (__any_sync(0xffffffffu, any) && lane == 0) touched[row] = 1;
}
static __global__ void touched_diag_lists_kernel(int64_t n_lists, const uint8_t* __restrict__ list_diag, const int32_t* __restrict__ list_slot,
                                                 const int32_t* __restrict__ ucol, uint8_t* touched) {
  const int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (u < n_lists && list_diag[u]) touched[ucol[list_slot[u]]] = 1;   // a diagonal slot's column is its row
}

// gather rows: dst[i] = src[idx[i]] with row width W (set-up only)
template <typename T>
__global__ void gather_rows_kernel(int64_t n, int W, const T* __restrict__ src, const int32_t* __restrict__ idx, T* __restrict__ dst) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * W) return;
  dst[i] = src[(size_t)idx[i / W] * W + i % W];
}

}  // namespace isfm
