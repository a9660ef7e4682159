// coarse.cuh -- coarse level of the two-level PCG preconditioner for the reduced camera system
// (see pcg_persistent.cuh).
//
// Cameras are grouped into clusters of `cs` consecutive cameras (image ids follow the capture
// order on the sequences where this matters: street / video).  Every cluster carries the seven
// similarity modes of a rigid piece of the scene -- world translation v, rotation w about the
// cluster's centroid, scale s about the centroid -- expressed in each camera's left-perturbation
// tangent [dtau, dphi] (pose = world2cam (R, t), update T <- Exp(delta) T):
//     X_w' = c0 + s Exp(w) (X_w - c0) + v   =>   dphi = -R w,  dtau = -R v - tt x (R w) + s tt,
//     tt = t + R c0  (translation of the camera w.r.t. the shifted world origin c0).
// Moving a whole cluster rigidly (with its points, which the Schur complement has eliminated)
// costs almost no energy, so these are the slowest modes of S on long camera chains; block-Jacobi
// cannot see them.  Ac = P^T S P is formed from the stored upper blocks (deterministic segmented
// sums, no atomics), all-reduced across ranks, and inverted densely in fp64 by a blocked
// Gauss-Jordan sweep (SPD, no pivoting) -- the coarse dimension is 7 * clusters <= ~2 100.
#pragma once
#include <cub/cub.cuh>

#include "comm.cuh"
#include "common.cuh"
#include "index_prep.cuh"
#include "math.cuh"

namespace isfm {

constexpr int CM = 7;          // coarse modes per cluster (== PCG_MODES)
constexpr int GJ_B = 32;       // pivot block of the dense inverse

struct CoarseGeom {
  int n_cam = 0, cs = 0, ncl = 0, ncp = 0;
  __host__ __device__ int cluster_of(int cam) const { const int c = cam / cs; return c < ncl ? c : ncl - 1; }
  __host__ __device__ int begin(int cl) const { return cl * cs; }
  __host__ __device__ int end(int cl) const { return cl == ncl - 1 ? n_cam : (cl + 1) * cs; }
};

// ---- set-up: upper slots grouped by coarse block (I, J), I <= J ----
static __global__ void coarse_slot_keys_kernel(CoarseGeom g, const int32_t* __restrict__ urow_ptr, const int32_t* __restrict__ ucol,
                                               uint32_t* keys, int32_t* vals, int32_t* row_of) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= g.n_cam) return;
  const int ci = g.cluster_of(row);
  for (int e = urow_ptr[row] + lane; e < urow_ptr[row + 1]; e += 32) {
    keys[e] = (uint32_t)(ci * g.ncl + g.cluster_of(ucol[e]));
    vals[e] = e;
    row_of[e] = row;
  }
}
static __global__ void coarse_offsets_kernel(const uint32_t* __restrict__ sorted, int64_t n, int32_t* off, int64_t n_keys) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k > n_keys) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted[mid] < (uint32_t)k) lo = mid + 1; else hi = mid;
  }
  off[k] = (int32_t)lo;
}

// ---- per LM step: the prolongation blocks P_i (6 x 7) from the current poses ----
template <typename T>
__global__ void coarse_centroid_kernel(CoarseGeom g, int cw, const T* __restrict__ cam, double* __restrict__ c0) {
  // one warp per cluster; lanes stride over its cameras, fixed shuffle tree: deterministic
  const int cl = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (cl >= g.ncl) return;
  double s[3] = {0, 0, 0};
  for (int i = g.begin(cl) + lane; i < g.end(cl); i += 32) {
    const T* c = cam + (size_t)i * cw;
    double R[9];
    const double q[4] = {(double)c[3], (double)c[4], (double)c[5], (double)c[6]};
    quat_to_rot(q, R);
    // centre = -R^T t
    for (int k = 0; k < 3; ++k) s[k] -= R[0 * 3 + k] * (double)c[0] + R[1 * 3 + k] * (double)c[1] + R[2 * 3 + k] * (double)c[2];
  }
  for (int k = 0; k < 3; ++k)
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
  const double inv = 1.0 / (double)(g.end(cl) - g.begin(cl));
  if (lane < 3) c0[cl * 3 + lane] = (lane == 0 ? s[0] : (lane == 1 ? s[1] : s[2])) * inv;
}
template <typename T>
__global__ void coarse_modes_kernel(CoarseGeom g, int cw, const T* __restrict__ cam, const double* __restrict__ c0, T* __restrict__ Pm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n_cam) return;
  similarity_modes<T>(cam + (size_t)i * cw, c0 + g.cluster_of(i) * 3, Pm + (size_t)i * (6 * CM));
}

// ---- per trial: G = P^T E P (dense [ncp][ncp], fp64) ----
// CTA (k, c): chunk c of the member blocks of the k-th non-empty coarse block (I <= J).  A group of
// 8 lanes (7 working) owns one member block at a time; lane k forms column k of X = P_i^T B P_j
// (fp32 products of fp32 data) and keeps 7 fp64 accumulators.  The groups' sums are added in
// group order into the chunk's own partial matrix; coarse_reduce_parts_kernel adds the chunks in
// chunk order: deterministic, no atomics.
constexpr int GAL_TPB = 128;
constexpr int GAL_SPLIT = 8;
template <typename T, int D>
__global__ void __launch_bounds__(GAL_TPB)
coarse_galerkin_kernel(CoarseGeom g, const int2* __restrict__ cblocks, const int32_t* __restrict__ coff, const int32_t* __restrict__ cslot,
                       const int32_t* __restrict__ row_of, const int32_t* __restrict__ ucol, const T* __restrict__ E,
                       const T* __restrict__ Pm, double* __restrict__ Gparts) {
  const int2 ij = cblocks[blockIdx.x];
  const int I = ij.x, J = ij.y, chunk = blockIdx.y;
  constexpr int NG = GAL_TPB / 8;
  __shared__ double part[NG][2][CM * CM];   // [group][off-diagonal members | diagonal members][r * CM + k]
  const int grp = threadIdx.x >> 3, k = threadIdx.x & 7;
  double acc[CM], accd[CM];
#pragma unroll
  for (int r = 0; r < CM; ++r) { acc[r] = 0.0; accd[r] = 0.0; }
  const int b0 = coff[I * g.ncl + J], b1 = coff[I * g.ncl + J + 1];
  const int beg = b0 + (int)(((long long)(b1 - b0) * chunk) / GAL_SPLIT), end = b0 + (int)(((long long)(b1 - b0) * (chunk + 1)) / GAL_SPLIT);
  if (k < CM) {
    for (int m = beg + grp; m < end; m += NG) {
      const int e = cslot[m], i = row_of[e], j = ucol[e];
      const T* __restrict__ B = E + (size_t)e * (D * D);
      const T* __restrict__ Pi = Pm + (size_t)i * (6 * CM);
      const T* __restrict__ Pj = Pm + (size_t)j * (6 * CM);
      float t[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 6; ++c) s += (float)B[r * D + c] * (float)Pj[c * CM + k];
        t[r] = s;
      }
      const bool dg = i == j;
#pragma unroll
      for (int r = 0; r < CM; ++r) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 6; ++c) s += (float)Pi[c * CM + r] * t[c];
        if (dg) accd[r] += (double)s; else acc[r] += (double)s;
      }
    }
#pragma unroll
    for (int r = 0; r < CM; ++r) { part[grp][0][r * CM + k] = acc[r]; part[grp][1][r * CM + k] = accd[r]; }
  }
  __syncthreads();
  if (threadIdx.x < CM * CM) {
    const int r = threadIdx.x / CM, c = threadIdx.x % CM;
    double x = 0.0, xt = 0.0, xd = 0.0;
    for (int q = 0; q < NG; ++q) { x += part[q][0][r * CM + c]; xt += part[q][0][c * CM + r]; xd += part[q][1][r * CM + c]; }
    double* G = Gparts + (size_t)chunk * g.ncp * g.ncp;
    if (I == J) {
      G[(size_t)(I * CM + r) * g.ncp + J * CM + c] = x + xt + xd;   // pairs (i, j) and (j, i) of one cluster, diagonal blocks once
    } else {
      G[(size_t)(I * CM + r) * g.ncp + J * CM + c] = x + xd;
      G[(size_t)(J * CM + c) * g.ncp + I * CM + r] = x + xd;        // mirrored coarse block
    }
  }
}
static __global__ void coarse_reduce_parts_kernel(size_t n, const double* __restrict__ parts, double* __restrict__ G) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = 0.0;
#pragma unroll
  for (int c = 0; c < GAL_SPLIT; ++c) v += parts[(size_t)c * n + i];
  G[i] = v;
}

// A = P^T Hd P - G  (Hd: the damped diagonal blocks S_ii; added once, after the all-reduce of G);
// padding rows / columns: identity.  One CTA per coarse row block.
// Ridge: along the seven gauge directions of the whole scene P^T (Hcc - E) P vanishes and only the
// damping term is left, while G carries the rounding noise of the stored E (eps_T relative): with a
// late-LM damping of 1e-8 the computed A would be indefinite there.  `ridge` * diag(P^T Hd P) is
// added to the diagonal (ridge_eps * eps_T): the coarse level then treats those modes as slightly more damped
// than they are, the preconditioner stays positive definite.  Default 4 eps_T: with the fp32 damping floor of
// 16 eps_T (ba_solver.cuh) the damping term itself dominates the noise; the ridge is a second line of defence.
constexpr int ASM_PARTS = 8;
template <typename T, int D>
__global__ void __launch_bounds__(512)
coarse_assemble_kernel(CoarseGeom g, const T* __restrict__ Hd, const T* __restrict__ Pm, const double* __restrict__ G, double* __restrict__ A,
                       double ridge) {
  __shared__ double Hp[ASM_PARTS][CM * CM];
  __shared__ double H[CM * CM];
  const int I = blockIdx.x;
  if (I < g.ncl) {
    // H_I = sum_{i in I} P_i^T Hd_i P_i: part p takes the cameras begin + p, begin + p + 8, ...; parts added in order
    const int part = threadIdx.x / (CM * CM), e = threadIdx.x % (CM * CM);
    if (part < ASM_PARTS) {
      const int r = e / CM, c = e % CM;
      double s = 0.0;
      for (int i = g.begin(I) + part; i < g.end(I); i += ASM_PARTS) {
        const T* __restrict__ h = Hd + (size_t)i * (D * D);
        const T* __restrict__ P = Pm + (size_t)i * (6 * CM);
        for (int a = 0; a < 6; ++a) {
          double t = 0.0;
          for (int b = 0; b < 6; ++b) t += (double)h[a * D + b] * (double)P[b * CM + c];
          s += (double)P[a * CM + r] * t;
        }
      }
      Hp[part][e] = s;
    }
    __syncthreads();
    if (threadIdx.x < CM * CM) {
      double s = 0.0;
      for (int q = 0; q < ASM_PARTS; ++q) s += Hp[q][threadIdx.x];
      H[threadIdx.x] = s;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < CM * g.ncp; idx += blockDim.x) {
      const int r = idx / g.ncp, col = idx % g.ncp;
      const size_t o = (size_t)(I * CM + r) * g.ncp + col;
      double v = col < g.ncl * CM ? -G[o] : 0.0;
      if (col / CM == I && col < g.ncl * CM) v += H[r * CM + col % CM] * (col % CM == r ? 1.0 + ridge : 1.0);
      A[o] = v;
    }
  } else {
    // padding rows [ncl * CM, ncp): identity
    const int r0 = g.ncl * CM;
    for (size_t idx = threadIdx.x; idx < (size_t)(g.ncp - r0) * g.ncp; idx += blockDim.x) {
      const int r = r0 + (int)(idx / g.ncp), col = (int)(idx % g.ncp);
      A[(size_t)r * g.ncp + col] = r == col ? 1.0 : 0.0;
    }
  }
}

// ---- dense SPD inverse, in place, blocked Gauss-Jordan without pivoting (fp64) ----
// step k, kernel A (CTA per 32-column tile J): Pinv = A[K,K]^-1 (every CTA, in shared memory);
//   ROW[:, J] = Pinv A[K, J];  COL[J, :] = A[J, K] (copy);  CTA K stores Pinv.
// step k, kernel B (CTA per 32 x 32 tile (I, J)):
//   I != K, J != K: A[I,J] -= COL[I] ROW[J];  I == K: A[K,J] = ROW[J];  J == K: A[I,K] = -COL[I] Pinv;  (K,K): Pinv.
static __global__ void __launch_bounds__(GJ_B * GJ_B)
gj_panel_kernel(int n, int kb, const double* __restrict__ A, double* __restrict__ ROW, double* __restrict__ COL,
                double* __restrict__ PINV, int* __restrict__ fail) {
  __shared__ double M[GJ_B][GJ_B + 1], Inv[GJ_B][GJ_B + 1], Tl[GJ_B][GJ_B + 1];
  const int tx = threadIdx.x % GJ_B, ty = threadIdx.x / GJ_B, J = blockIdx.x, K = kb;
  M[ty][tx] = A[(size_t)(K * GJ_B + ty) * n + K * GJ_B + tx];
  Inv[ty][tx] = ty == tx ? 1.0 : 0.0;
  Tl[ty][tx] = A[(size_t)(K * GJ_B + ty) * n + J * GJ_B + tx];           // A[K, J] tile
  COL[(size_t)(J * GJ_B + ty) * GJ_B + tx] = A[(size_t)(J * GJ_B + ty) * n + K * GJ_B + tx];   // A[J, K] tile
  __syncthreads();
  // Gauss-Jordan on [M | Inv]
  for (int p = 0; p < GJ_B; ++p) {
    // read phase (old values) / write phase: two barriers per pivot
    const double piv = M[p][p];
    if (!(piv > 0.0) && threadIdx.x == 0) *fail = 1;
    const double ip = 1.0 / piv;
    const double mp = M[p][tx] * ip, vp = Inv[p][tx] * ip, f = M[ty][p];
    __syncthreads();
    if (ty == p) { M[p][tx] = mp; Inv[p][tx] = vp; }
    else { M[ty][tx] -= f * mp; Inv[ty][tx] -= f * vp; }
    __syncthreads();
  }
  double s = 0.0;
#pragma unroll 8
  for (int c = 0; c < GJ_B; ++c) s += Inv[ty][c] * Tl[c][tx];
  ROW[(size_t)ty * n + J * GJ_B + tx] = s;
  if (J == K) PINV[ty * GJ_B + tx] = Inv[ty][tx];
}
// 64 x 64 tile per CTA (2 x 2 pivot-sized blocks), 256 threads, a 4 x 4 micro-tile each; n is a multiple of 32, so
// edge tiles are guarded per 32-block.
constexpr int GJ_T = 64;
static __global__ void __launch_bounds__(256)
gj_update_kernel(int n, int kb, double* __restrict__ A, const double* __restrict__ ROW, const double* __restrict__ COL,
                 const double* __restrict__ PINV) {
  __shared__ double Cs[GJ_T][GJ_B + 1], Rs[GJ_B][GJ_T + 1];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int i0 = blockIdx.y * GJ_T, j0 = blockIdx.x * GJ_T, K0 = kb * GJ_B;
  // stage the column panel rows [i0, i0 + 64) and the row panel columns [j0, j0 + 64); inside the pivot
  // block column the "row panel" is Pinv (A[I,K] <- -COL[I] Pinv), inside the pivot block row nothing is multiplied
  for (int e = threadIdx.x; e < GJ_T * GJ_B; e += 256) {
    const int r = e / GJ_B, c = e % GJ_B;
    Cs[r][c] = i0 + r < n ? COL[(size_t)(i0 + r) * GJ_B + c] : 0.0;
  }
  for (int e = threadIdx.x; e < GJ_B * GJ_T; e += 256) {
    const int r = e / GJ_T, c = e % GJ_T, col = j0 + c;
    double v = 0.0;
    if (col < n) v = (col >= K0 && col < K0 + GJ_B) ? PINV[r * GJ_B + (col - K0)] : ROW[(size_t)r * n + col];
    Rs[r][c] = v;
  }
  __syncthreads();
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 4
  for (int c = 0; c < GJ_B; ++c) {
    double cv[4], rv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) cv[a] = Cs[ty * 4 + a][c];
#pragma unroll
    for (int b = 0; b < 4; ++b) rv[b] = Rs[c][tx * 4 + b];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] += cv[a] * rv[b];
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int row = i0 + ty * 4 + a;
    if (row >= n) continue;
    const bool prow = row >= K0 && row < K0 + GJ_B;   // pivot block row
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int col = j0 + tx * 4 + b;
      if (col >= n) continue;
      const bool pcol = col >= K0 && col < K0 + GJ_B;
      double* dst = A + (size_t)row * n + col;
      if (prow) *dst = pcol ? PINV[(row - K0) * GJ_B + (col - K0)] : ROW[(size_t)(row - K0) * n + col];
      else *dst = pcol ? -acc[a][b] : *dst - acc[a][b];
    }
  }
}
template <typename T, int D>
struct CoarseLevel {
  bool enabled = false;
  CoarseGeom g;
  DeviceBuffer<int32_t> coff, cslot, row_of;
  DeviceBuffer<int2> cblocks;   // non-empty coarse blocks (I <= J)
  int n_cblocks = 0;
  DeviceBuffer<double> c0, G, Gparts, A, ROW, COL, PINV, rc;
  DeviceBuffer<T> Pm;
  DeviceBuffer<double> Ainv;   // fp64: the inverse of a coarse matrix of condition 1e8 rounded to fp32 is no longer positive definite
  DeviceBuffer<int> fail;
  double ridge_eps = 4.0;   // ridge on the coarse diagonal in units of eps_T (see coarse_assemble_kernel)

  // decides whether the coarse level pays (sparse, chain-like camera graph) and builds the slot lists
  // n_off_global: strictly-upper blocks of the whole reduced system (all ranks)
  void setup(const SchurPattern& sp, int64_t n_off_global, int64_t n_cam, cudaStream_t s, KernelTimers& kt) {
    enabled = false;
    if (D < CM) return;
    if (const char* e = getenv("ISFM_COARSE_RIDGE")) ridge_eps = atof(e);
    const char* env = getenv("ISFM_TWO_LEVEL");   // "0": never, "1": always, unset: by sparsity
    if (env && atoi(env) == 0) return;
    // Default: every system of at least 32 cameras.  On chain-like camera graphs (street scenes) the
    // coarse level removes the bending modes block-Jacobi needs thousands of iterations for; on dense
    // BAL-like systems it removes the weakly damped global modes (measured at C3: 31 -> 11 PCG
    // iterations per LM step for 0.3 ms of coarse set-up per trial).
    if (!(env ? atoi(env) != 0 : n_cam >= 32) || n_cam < 8) return;
    // cluster size: a quarter of the mean upper row length (~ the co-visibility window), at most
    // ISFM_COARSE_MAX_CLUSTERS clusters (default 148: the dense inverse of the 7 * clusters coarse
    // matrix is replicated on every rank and costs O(clusters^3) per trial)
    int cs = (int)std::max<int64_t>(4, std::min<int64_t>(n_off_global / std::max<int64_t>(n_cam, 1) / 4, n_cam / 16));
    if (const char* e = getenv("ISFM_COARSE_CS")) cs = std::max(2, atoi(e));
    int max_cl = 148;   // <= SMs of a B200: one cluster per CTA in the update / coarse phases of the persistent kernel
    if (const char* e = getenv("ISFM_COARSE_MAX_CLUSTERS")) max_cl = std::max(1, atoi(e));
    cs = std::max<int>(cs, (int)((n_cam + max_cl - 1) / max_cl));
    cs = (int)std::min<int64_t>(cs, std::max<int64_t>(n_cam / 2, 2));
    g.n_cam = (int)n_cam; g.cs = cs; g.ncl = std::max(1, (int)(n_cam / cs));
    g.ncp = (g.ncl * CM + GJ_B - 1) / GJ_B * GJ_B;
    TimerScope ts(kt, T_INDEX_PREP);
    const int64_t nnzu = sp.nnzu;
    DeviceBuffer<uint32_t> keys, keys_s; DeviceBuffer<int32_t> vals;
    keys.alloc(nnzu); keys_s.alloc(nnzu); vals.alloc(nnzu); cslot.alloc(nnzu); row_of.alloc(nnzu);
    coarse_slot_keys_kernel<<<div_up(n_cam, 8), 256, 0, s>>>(g, sp.urow_ptr.get(), sp.ucol.get(), keys.get(), vals.get(), row_of.get());
    size_t bytes = 0;
    int end_bit = 1;
    while ((1ll << end_bit) < (int64_t)g.ncl * g.ncl) ++end_bit;
    ISFM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.get(), keys_s.get(), vals.get(), cslot.get(), nnzu, 0, end_bit, s));
    DeviceBuffer<uint8_t> tmp; tmp.alloc(bytes);
    ISFM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), bytes, keys.get(), keys_s.get(), vals.get(), cslot.get(), nnzu, 0, end_bit, s));
    const int64_t nk = (int64_t)g.ncl * g.ncl;
    coff.alloc(nk + 1);
    coarse_offsets_kernel<<<div_up(nk + 1, 256), 256, 0, s>>>(keys_s.get(), nnzu, coff.get(), nk);
    ISFM_CUDA(cudaGetLastError());
    {
      std::vector<int32_t> h_off((size_t)nk + 1);
      ISFM_CUDA(cudaMemcpyAsync(h_off.data(), coff.get(), h_off.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
      ISFM_CUDA(cudaStreamSynchronize(s));   // temporaries go out of scope
      std::vector<int2> nz;
      for (int I = 0; I < g.ncl; ++I)
        for (int J = I; J < g.ncl; ++J)
          if (h_off[(size_t)I * g.ncl + J + 1] > h_off[(size_t)I * g.ncl + J]) nz.push_back(make_int2(I, J));
      n_cblocks = (int)nz.size();
      cblocks.alloc(std::max<size_t>(nz.size(), 1));
      if (!nz.empty()) ISFM_CUDA(cudaMemcpyAsync(cblocks.get(), nz.data(), nz.size() * sizeof(int2), cudaMemcpyHostToDevice, s));
      ISFM_CUDA(cudaStreamSynchronize(s));
    }
    const size_t nn = (size_t)g.ncp * g.ncp;
    Gparts.alloc(nn * GAL_SPLIT);
    c0.alloc((size_t)g.ncl * 3); G.alloc(nn); A.alloc(nn); ROW.alloc((size_t)GJ_B * g.ncp); COL.alloc((size_t)g.ncp * GJ_B);
    PINV.alloc(GJ_B * GJ_B); rc.alloc(g.ncp); Pm.alloc((size_t)n_cam * 6 * CM); Ainv.alloc(nn); fail.alloc(1);
    enabled = true;
  }

  // prolongation from the current camera rows ([n_cam][cw], t | q_xyzw | ...): once per LM step
  void update_modes(const T* cam, int cw, cudaStream_t s, KernelTimers& kt) {
    if (!enabled) return;
    TimerScope ts(kt, T_MISC);
    coarse_centroid_kernel<T><<<div_up(g.ncl, 4), 128, 0, s>>>(g, cw, cam, c0.get());
    coarse_modes_kernel<T><<<div_up(g.n_cam, 128), 128, 0, s>>>(g, cw, cam, c0.get(), Pm.get());
  }

  // Ac^-1 for the current damped system: once per trial.  E: this rank's (partial) upper blocks.
  void factor(const T* E, const T* Hd, const SchurPattern& sp, isfm_comm* comm, cudaStream_t s, KernelTimers& kt) {
    if (!enabled) return;
    { TimerScope ts(kt, T_COARSE);
      const size_t nn = (size_t)g.ncp * g.ncp;
      ISFM_CUDA(cudaMemsetAsync(Gparts.get(), 0, nn * GAL_SPLIT * sizeof(double), s));
      if (n_cblocks > 0)
        coarse_galerkin_kernel<T, D><<<dim3(n_cblocks, GAL_SPLIT), GAL_TPB, 0, s>>>(g, cblocks.get(), coff.get(), cslot.get(), row_of.get(),
                                                                                  sp.ucol.get(), E, Pm.get(), Gparts.get());
      coarse_reduce_parts_kernel<<<div_up((int64_t)nn, 256), 256, 0, s>>>(nn, Gparts.get(), G.get()); }
    if (comm_world(comm) > 1) {
      TimerScope ts(kt, T_COMM);
      comm_allreduce_sum(comm, G.get(), (size_t)g.ncp * g.ncp, true, s);
    }
    TimerScope ts(kt, T_COARSE);
    coarse_assemble_kernel<T, D><<<g.ncl + 1, 512, 0, s>>>(g, Hd, Pm.get(), G.get(), A.get(),
                                                          ridge_eps * (sizeof(T) == 4 ? 5.96e-8 : 1.11e-16));
    ISFM_CUDA(cudaMemsetAsync(fail.get(), 0, sizeof(int), s));
    const int nb = g.ncp / GJ_B;
    for (int k = 0; k < nb; ++k) {
      gj_panel_kernel<<<nb, GJ_B * GJ_B, 0, s>>>(g.ncp, k, A.get(), ROW.get(), COL.get(), PINV.get(), fail.get());
      gj_update_kernel<<<dim3(div_up(g.ncp, GJ_T), div_up(g.ncp, GJ_T)), 256, 0, s>>>(g.ncp, k, A.get(), ROW.get(), COL.get(), PINV.get());
    }
    ISFM_CUDA(cudaMemcpyAsync(Ainv.get(), A.get(), (size_t)g.ncp * g.ncp * sizeof(double), cudaMemcpyDeviceToDevice, s));
    ISFM_CUDA(cudaGetLastError());
  }
};

}  // namespace isfm
