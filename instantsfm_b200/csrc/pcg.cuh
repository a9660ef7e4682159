// pcg.cuh -- block-Jacobi preconditioned CG on the reduced camera system (K4).
//
//   S p = Hd p - E p
//
// Hd: damped block diagonal Hcc (replicated on every rank), E: this rank's Schur
// contribution sum_p Hcp Hpp^-1 Hcp^T, upper triangle as BSR.  With several ranks the only
// communication per iteration is the all-reduce of y = E p; every rank holds the full
// vectors and computes the dot products redundantly (bit-identical across ranks).
// Replaces bae.utils.pysolvers.PCG (bundle_adjustment.py:117): x0 = 0, stop when
// ||r|| < tol ||b||.
#pragma once
#include "comm.cuh"
#include "common.cuh"
#include "index_prep.cuh"

namespace isfm {

struct PcgState {
  double rho;      // r.z of the current iteration
  double rho_next; // r.z computed by the direction kernel, promoted by the next mat-vec tail
  int has_next;
  double bb;       // ||b||^2
  double rr;       // ||r||^2 after the last completed iteration
  int done;        // 1: converged, 2: breakdown (non-finite / non-positive curvature)
  int iters;
  double pq;       // p.q of the current iteration (published by the last CTA of the mat-vec tail)
  unsigned int ticket;
  unsigned int ticket2;   // last-CTA election of the merged update / direction kernel
  int max_iter;
  unsigned int bar;       // persistent kernel: grid-barrier arrivals (monotonic within a solve)
  int abort;              // persistent kernel: a barrier or a peer wait timed out
  unsigned long long phase_ns[8];   // persistent kernel: time per phase of CTA 0 (PcgPhase), this solve
};

// The CTA that finishes last sums the per-CTA partials in index order (deterministic) and
// publishes the scalar: the next kernel reads one double instead of re-reducing n partials in
// every CTA.  Integer ticket only; no floating-point atomics.
__device__ __forceinline__ void publish_pq_last_cta(double cta_value, double* partial, PcgState* st) {
  __shared__ bool is_last__;
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = cta_value;
    __threadfence();
    is_last__ = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last__) {
    double v = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) v += __ldcg(partial + i);
    v = block_sum(v);
    if (threadIdx.x == 0) {
      st->pq = v;
      st->ticket = 0u;
      // nobody reads rho inside this kernel: promote the value the previous direction kernel left
      if (st->has_next) { st->rho = st->rho_next; st->has_next = 0; }
    }
  }
}

constexpr int PCG_TPB = 128;

// Symmetric mat-vec y = E p with only the upper triangle of E stored (BSR, sorted by (i, j)).
// One CTA per block row i.  Each warp streams its share of the row's blocks through a private
// shared-memory stage with fully coalesced loads (a 128-byte line per warp instruction; the
// D x D blocks are 36-byte-row AoS, unfriendly to direct per-lane row loads), then
//   lane (b, r): yup_i[r] += sum_c B[r][c] p_j[c]                 (row r of block b)
//   lane (b, c): C[tpos(e)][c] = sum_r B[r][c] p_i[r]             (B^T p_i, deposited at the
//                position of block (j, i) in the row-major order of the lower triangle)
// A second kernel adds row j's deposits: no atomics, deterministic.
template <typename T, int D> struct SpmvCfg {
  static constexpr int NW = PCG_TPB / 32;                            // warps = work units per CTA
  static constexpr int GPW = 32 / D;                                 // blocks per warp pass
  // blocks per stage: a multiple of 4 (16-byte alignment of every stage source)
  static constexpr int WB = sizeof(T) == 4 ? GPW * 4 : (GPW * 2 + 3) / 4 * 4;   // (8-block stages measured slower)
  static constexpr int PASSES = (WB + GPW - 1) / GPW;
  static constexpr int VE = 16 / sizeof(T);                          // elements per 16-byte vector
  static constexpr int NV = WB * D * D / VE;                         // vectors per stage
  static constexpr int NLD = (NV + 31) / 32;                         // cp.async per lane per stage
  static constexpr int STG = NLD * 32 * VE + (WB * D + VE - 1) / VE * VE;   // elements per stage buffer: blocks | p_j of each block
  static constexpr int POFF = NLD * 32 * VE;                         // offset of the p_j area inside a stage buffer
  static constexpr size_t SMEM = (size_t)NW * 2 * STG * sizeof(T);   // double-buffered, per warp
};

// One WARP per work unit (<= SPMV_CHUNK consecutive slots of one upper row), no CTA-wide
// synchronisation.  The unit's blocks are streamed with cp.async (LDGSTS, 16 B, L2-only) into
// a per-warp double buffer: stage k+1 is in flight while stage k is multiplied.
template <typename T, int D>
__global__ void __launch_bounds__(PCG_TPB)
pcg_spmv_upper_kernel(int n_units, int unit_slots, const int32_t* __restrict__ unit_row, const int32_t* __restrict__ unit_beg,
                      const int32_t* __restrict__ urow_ptr, const int32_t* __restrict__ ucol,
                      const int32_t* __restrict__ tpos, const T* __restrict__ EU, const T* __restrict__ p,
                      T* __restrict__ yup_part, T* __restrict__ C, const PcgState* __restrict__ st) {
  if (st->done) return;
  typedef SpmvCfg<T, D> Cfg;
  constexpr int GPW = Cfg::GPW, WB = Cfg::WB, DD = D * D, VE = Cfg::VE, NLD = Cfg::NLD, STG = Cfg::STG;
  extern __shared__ __align__(16) unsigned char spmv_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int unit = blockIdx.x * Cfg::NW + w;
  if (unit >= n_units) return;
  T* buf = reinterpret_cast<T*>(spmv_smem) + (size_t)w * 2 * STG;
  const int bl = lane / D, r = lane % D;
  const int row = unit_row[unit];
  const int beg = unit_beg[unit], end = min(beg + unit_slots, urow_ptr[row + 1]);
  const int ns = (end - beg + WB - 1) / WB;
  T pi[D];
#pragma unroll
  for (int c = 0; c < D; ++c) pi[c] = p[(size_t)row * D + c];
  // E is streamed once per mat-vec (evict_first); the deposits are read back by the combine
  // kernel right after and fit the L2 (evict_last): they need not travel to HBM and back
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();

  // column / deposit indices of stage k for this lane's blocks (-1 where the lane has no block)
  auto load_idx = [&](int k, int* jj, int* tp) {
    const int base = beg + k * WB;
    const int nb = min(WB, end - base);
#pragma unroll
    for (int pass = 0; pass < Cfg::PASSES; ++pass) {
      const int b = pass * GPW + bl;
      const bool on = k < ns && bl < GPW && b < nb;
      jj[pass] = on ? __ldg(ucol + base + b) : -1;
      tp[pass] = on ? __ldg(tpos + base + b) : -1;
    }
  };
  // stage k: the blocks (16-byte cp.async, L2 only) and, per block, its p_j (element-wise
  // cp.async) so that the multiply phase reads everything from shared memory
  auto issue = [&](int k, const int* jj) {
    const int base = beg + k * WB;
    const int nb = min(WB, end - base);
    const int last = nb * DD / VE - 1;   // indices past the end re-copy the last vector
    const T* src = EU + (size_t)base * DD;
    T* dst = buf + (size_t)(k & 1) * STG;
#pragma unroll
    for (int q = 0; q < NLD; ++q) cp_async16_hint(dst + (size_t)(lane + 32 * q) * VE, src + (size_t)min(lane + 32 * q, last) * VE, pol_stream);
#pragma unroll
    for (int pass = 0; pass < Cfg::PASSES; ++pass)
      if (jj[pass] >= 0) cp_async_small<sizeof(T)>(dst + Cfg::POFF + (pass * GPW + bl) * D + r, p + (size_t)jj[pass] * D + r);
    cp_async_commit();
  };

  // software pipeline: indices two stages ahead, data one stage ahead of the multiply
  int j0[Cfg::PASSES], t0[Cfg::PASSES], j1[Cfg::PASSES], t1[Cfg::PASSES], j2[Cfg::PASSES], t2[Cfg::PASSES];
  load_idx(0, j0, t0);
  load_idx(1, j1, t1);
  issue(0, j0);
  T acc = T(0);
  for (int k = 0; k < ns; ++k) {
    load_idx(k + 2, j2, t2);
    if (k + 1 < ns) { issue(k + 1, j1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncwarp();
    const T* S = buf + (size_t)(k & 1) * STG;
#pragma unroll
    for (int pass = 0; pass < Cfg::PASSES; ++pass) {
      if (j0[pass] >= 0) {
        const T* B = S + (pass * GPW + bl) * DD;
        const T* pj = S + Cfg::POFF + (pass * GPW + bl) * D;
        T t = T(0);
#pragma unroll
        for (int c = 0; c < D; ++c) { acc += B[r * D + c] * pj[c]; t += B[c * D + r] * pi[c]; }
        if (t0[pass] >= 0) st_global_hint(C + (size_t)t0[pass] * D + r, t, pol_keep);
      }
    }
    __syncwarp();
#pragma unroll
    for (int pass = 0; pass < Cfg::PASSES; ++pass) { j0[pass] = j1[pass]; t0[pass] = t1[pass]; j1[pass] = j2[pass]; t1[pass] = t2[pass]; }
  }
  // fold the GPW block lanes onto lanes 0..D-1
#pragma unroll
  for (int k = 1; k < GPW; ++k) {
    T o = __shfl_sync(0xffffffffu, acc, (lane + k * D) & 31);
    if (lane < D) acc += o;
  }
  if (lane < D) st_global_hint(yup_part + (size_t)unit * D + lane, acc, pol_keep);
}

// y_i = sum of row i's chunk partials + sum of the deposits of row i (contiguous in C).
// COMB_FUSED (single rank): q_i = Hd_i p_i - y_i and the per-CTA partial of p.q.
// COMB_PLAIN: y is written for an NCCL all-reduce.
// COMB_PUSH: this rank's y rows are stored straight into the exchange slot of EVERY rank
// (remote stores over NVLink); the CTA that finishes last publishes the sequence number in all
// peers' flags behind a system-scope fence -- the first half of the all-reduce is this
// kernel's epilogue (see comm.cuh), pcg_apply_diag_kernel<.., true> is the second half.  Row i of the lower triangle holds up to i deposits, so a CTA takes
// the two rows b and n-1-b: every CTA of a dense system sums the same number of deposits (no
// tail of long rows), with eight independent loads in flight per thread.
constexpr int COMB_TPB = 256;
constexpr int COMB_PLAIN = 0, COMB_FUSED = 1, COMB_PUSH = 2;

template <typename T, int D, int MODE>
__global__ void __launch_bounds__(COMB_TPB)
pcg_combine_kernel(int n_cam, const int32_t* __restrict__ dep_beg, const int32_t* __restrict__ dep_end,
                   const int32_t* __restrict__ chunk_ptr, int unit_lo, int unit_hi,
                   const T* __restrict__ yup_part, const T* __restrict__ C, const T* __restrict__ Hd,
                   const T* __restrict__ p, T* __restrict__ out, double* __restrict__ partial,
                   PcgState* __restrict__ st, const PeerExchange px) {
  if (st->done) return;
  constexpr bool FUSED = MODE == COMB_FUSED;
  // sequence number of THIS exchange; the counter is advanced by the last CTA only, after every
  // CTA has read it (each CTA reads before it takes its ticket)
  uint32_t seq = 0;
  if (MODE == COMB_PUSH) seq = *reinterpret_cast<volatile uint32_t*>(peer_seq(px)) + 1u;
  const int parity = (int)(seq & 1u);
  constexpr int G = COMB_TPB / D;   // entry groups per CTA; threads >= G * D idle
  constexpr int MLP = 8;
  __shared__ T sh[G][D];
  __shared__ T qs[2][D];
  const int g = threadIdx.x / D, c = threadIdx.x % D;
  int halves = 1;
  for (int half = 0; half < 2; ++half) {
    const int row = half == 0 ? (int)blockIdx.x : n_cam - 1 - (int)blockIdx.x;
    if (half == 1 && row <= (int)blockIdx.x) break;   // middle row of an odd system: once
    halves = half + 1;
    if (g < G) {
      T acc = T(0);
      // only the units / deposits this rank produced (all of them unless the mat-vec is split)
      const int ke = min(chunk_ptr[row + 1], unit_hi);
      for (int k = max(chunk_ptr[row], unit_lo) + g; k < ke; k += G) acc += yup_part[(size_t)k * D + c];
      const int end = dep_end[row];
      int k = dep_beg[row] + g;
      T a[MLP];
#pragma unroll
      for (int u = 0; u < MLP; ++u) a[u] = T(0);
      for (; k + (MLP - 1) * G < end; k += MLP * G) {
#pragma unroll
        for (int u = 0; u < MLP; ++u) a[u] += __ldcs(C + (size_t)(k + u * G) * D + c);
      }
      for (; k < end; k += G) a[0] += __ldcs(C + (size_t)k * D + c);
#pragma unroll
      for (int u = 1; u < MLP; ++u) a[0] += a[u];
      sh[g][c] = acc + a[0];
    }
    __syncthreads();
    if (threadIdx.x < D) {
      T y = T(0);
      for (int k = 0; k < G; ++k) y += sh[k][threadIdx.x];
      if (FUSED) {
        const T* __restrict__ h = Hd + (size_t)row * (D * D) + threadIdx.x * D;
        const T* __restrict__ pi = p + (size_t)row * D;
        T q = T(0);
#pragma unroll
        for (int k = 0; k < D; ++k) q += h[k] * pi[k];
        q -= y;
        out[(size_t)row * D + threadIdx.x] = q;
        qs[half][threadIdx.x] = q * pi[threadIdx.x];
      } else if (MODE == COMB_PUSH) {
        for (int dst = 0; dst < px.world; ++dst) peer_slot<T>(px, dst, parity, px.rank)[(size_t)row * D + threadIdx.x] = y;
      } else {
        out[(size_t)row * D + threadIdx.x] = y;
      }
    }
    __syncthreads();   // sh is reused by the second row; qs complete
  }
  if (MODE == COMB_PUSH) {
    // the stores of threads 0..D-1 are ordered before thread 0's fence by the barrier above
    __shared__ bool last_push__;
    if (threadIdx.x == 0) {
      __threadfence_system();
      last_push__ = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last_push__) {
      __threadfence_system();
      if ((int)threadIdx.x < px.world) st_release_sys(peer_flags(px, threadIdx.x, parity) + px.rank, seq);
      if (threadIdx.x == 0) { *peer_seq(px) = seq; st->ticket = 0u; }
    }
  }
  if (FUSED) {
    double s = 0.0;
    if (threadIdx.x == 0) {
      for (int h = 0; h < halves; ++h)
#pragma unroll
        for (int k = 0; k < D; ++k) s += (double)qs[h][k];
    }
    publish_pq_last_cta(s, partial, st);
  }
}

// Large camera systems: the combine kernel writes y locally (COMB_PLAIN) and this kernel pushes
// it to every rank's exchange slot with coalesced 16-byte remote stores from the whole grid (the
// 36-byte row pieces the combine CTAs would store themselves are too small for NVLink once the
// vector is hundreds of KB); the CTA that finishes last publishes the sequence number.
constexpr int PUSH_TPB = 256;
template <typename T>
__global__ void __launch_bounds__(PUSH_TPB)
pcg_push_kernel(size_t n, const T* __restrict__ y, PcgState* __restrict__ st, const PeerExchange px) {
  if (st->done) return;
  const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(peer_seq(px)) + 1u;
  const int parity = (int)(seq & 1u);
  constexpr int VE = 16 / sizeof(T);
  const size_t nv = n / VE;
  for (size_t i = blockIdx.x * (size_t)PUSH_TPB + threadIdx.x; i < nv; i += (size_t)gridDim.x * PUSH_TPB) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(y) + i);
    for (int dst = 0; dst < px.world; ++dst) reinterpret_cast<float4*>(peer_slot<T>(px, dst, parity, px.rank))[i] = v;
  }
  if (blockIdx.x == 0)
    for (size_t i = nv * VE + threadIdx.x; i < n; i += PUSH_TPB) {
      const T v = y[i];
      for (int dst = 0; dst < px.world; ++dst) peer_slot<T>(px, dst, parity, px.rank)[i] = v;
    }
  __syncthreads();
  __shared__ bool last_push__;
  if (threadIdx.x == 0) {
    __threadfence_system();
    last_push__ = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last_push__) {
    __threadfence_system();
    if ((int)threadIdx.x < px.world) st_release_sys(peer_flags(px, threadIdx.x, parity) + px.rank, seq);
    if (threadIdx.x == 0) { *peer_seq(px) = seq; st->ticket = 0u; }
  }
}

// multi-rank second half: q = Hd p - y, per-row partial of p.q.  y is either the NCCL
// all-reduced vector or (PEER) the sum, in rank order, of the `world` partial vectors the ranks
// pushed into this rank's exchange slots: every rank adds the same numbers in the same order, so
// q, and with it the whole PCG state, stays bit-identical across ranks.
template <typename T, int D, bool PEER>
__global__ void __launch_bounds__(PCG_TPB)
pcg_apply_diag_kernel(int n_cam, const T* __restrict__ Hd, const T* __restrict__ p, const T* __restrict__ y,
                      T* __restrict__ q, double* __restrict__ partial, PcgState* __restrict__ st, const PeerExchange px) {
  if (st->done) return;
  int parity = 0;
  if (PEER) {
    const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(peer_seq(px));   // advanced by the push kernel before us
    parity = (int)(seq & 1u);
    if (!peer_wait_all(px, parity, seq)) {   // a peer never arrived: stop the solve, the host reports it
      if (threadIdx.x == 0) st->done = 3;
      return;
    }
  }
  // D threads per camera: thread (cam, k) owns component k -- row k of Hd_cam and entry k of
  // every partial vector (consecutive threads read consecutive words of the exchange slots)
  constexpr int CPB = PCG_TPB / D;
  const int cam = blockIdx.x * CPB + threadIdx.x / D, k = threadIdx.x % D;
  double s = 0.0;
  if (threadIdx.x < CPB * D && cam < n_cam) {
    const T* __restrict__ h = Hd + (size_t)cam * (D * D) + k * D;
    const T* __restrict__ pi = p + (size_t)cam * D;
    const size_t o = (size_t)cam * D + k;
    T v = T(0);
#pragma unroll
    for (int c = 0; c < D; ++c) v += h[c] * pi[c];
    if (PEER) {
      T ysum = T(0);
      for (int src = 0; src < px.world; ++src) ysum += __ldcg(peer_slot<T>(px, px.rank, parity, src) + o);
      v -= ysum;
    } else {
      v -= y[o];
    }
    q[o] = v;
    s = (double)v * (double)pi[k];
  }
  s = block_sum(s);
  publish_pq_last_cta(s, partial, st);
}

// r0 = b, x0 = 0, z0 = Minv r0, p0 = z0; partials of (r.z, b.b) per camera
template <typename T, int D>
__global__ void pcg_init_kernel(int n_cam, const T* __restrict__ b, const T* __restrict__ Minv, T* __restrict__ x,
                                T* __restrict__ r, T* __restrict__ p, double* __restrict__ part_rz,
                                double* __restrict__ part_bb) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_cam) return;
  const T* m = Minv + (size_t)row * (D * D);
  T rv[D];
  double bb = 0.0, rz = 0.0;
#pragma unroll
  for (int c = 0; c < D; ++c) { rv[c] = b[(size_t)row * D + c]; bb += (double)rv[c] * (double)rv[c]; }
#pragma unroll
  for (int k = 0; k < D; ++k) {
    T z = T(0);
#pragma unroll
    for (int c = 0; c < D; ++c) z += m[k * D + c] * rv[c];
    x[(size_t)row * D + k] = T(0);
    r[(size_t)row * D + k] = rv[k];
    p[(size_t)row * D + k] = z;
    rz += (double)z * (double)rv[k];
  }
  part_rz[row] = rz;
  part_bb[row] = bb;
}

template <int D>
__global__ void pcg_init_state_kernel(int n, int max_iter, const double* __restrict__ part_rz,
                                      const double* __restrict__ part_bb, PcgState* st) {
  double rz = reduce_partials(part_rz, n);
  double bb = reduce_partials(part_bb, n);
  if (threadIdx.x == 0) {
    st->rho = rz; st->rho_next = 0.0; st->has_next = 0; st->bb = bb; st->rr = bb; st->iters = 0; st->pq = 0.0;
    st->ticket = 0u; st->ticket2 = 0u; st->max_iter = max_iter; st->bar = 0u; st->abort = 0;
    for (int i = 0; i < 8; ++i) st->phase_ns[i] = 0ull;
    st->done = (bb == 0.0) ? 1 : ((isfinite(rz) && isfinite(bb)) ? 0 : 2);
  }
}

// End-of-iteration bookkeeping (one thread): new state, convergence test, and -- inside the
// device-side WHILE graph -- the loop condition of the conditional node.
__device__ __forceinline__ void pcg_finish_iteration(PcgState* st, double rho_new, double rr, double tol2, bool promote_now,
                                                     cudaGraphConditionalHandle cond, int use_cond) {
  const double pq = st->pq;
  if (promote_now) { st->rho = rho_new; st->has_next = 0; }
  else { st->rho_next = rho_new; st->has_next = 1; }   // promoted by the next mat-vec tail (other CTAs still read rho)
  st->rr = rr;
  const int it = st->iters + 1;
  st->iters = it;
  int done = 0;
  // `done` is read at kernel entry by every block of the NEXT launch only
  if (!(isfinite(rho_new) && isfinite(rr)) || !(pq > 0.0)) done = 2;
  else if (rr < tol2 * st->bb) done = 1;
  if (done) st->done = done;
  if (use_cond) cudaGraphSetConditional(cond, (!done && it < st->max_iter) ? 1u : 0u);
}

// alpha = rho / (p.q); x += alpha p; r -= alpha q; z = Minv r; partials of r.z and r.r.
// D threads per camera (thread k owns component k and row k of Minv).
// MERGED (small systems, where launches dominate): the CTA that finishes last also computes
// beta = rho_new / rho, the new direction p = z + beta p for the whole vector and the iteration
// state -- one launch less per iteration.  Otherwise pcg_direction_kernel follows.
constexpr int UPD_TPB = 128;

template <typename T, int D, bool MERGED>
__global__ void __launch_bounds__(UPD_TPB)
pcg_update_kernel(int n_cam, double tol2, const T* __restrict__ Minv, T* __restrict__ p, const T* __restrict__ q,
                  T* __restrict__ x, T* __restrict__ r, T* __restrict__ z, double* __restrict__ part_rz,
                  double* __restrict__ part_rr, PcgState* __restrict__ st, cudaGraphConditionalHandle cond, int use_cond) {
  if (st->done) {
    if (MERGED && use_cond && blockIdx.x == 0 && threadIdx.x == 0) cudaGraphSetConditional(cond, 0u);
    return;
  }
  constexpr int CPB = UPD_TPB / D;
  const double rho = st->rho;
  const double pq_now = st->pq;
  if (!(pq_now > 0.0) || !isfinite(pq_now)) {
    // breakdown (non-positive curvature / non-finite): x keeps the last good iterate; the solve
    // ends with status 2 (uniform: every CTA reads the same scalar)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      st->done = 2;
      if (use_cond) cudaGraphSetConditional(cond, 0u);
    }
    return;
  }
  const T alpha = (T)(rho / pq_now);
  const int cam = blockIdx.x * CPB + threadIdx.x / D, k = threadIdx.x % D;
  const bool on = threadIdx.x < CPB * D && cam < n_cam;
  T rv[D];
  if (on) {
#pragma unroll
    for (int c = 0; c < D; ++c) { const size_t o = (size_t)cam * D + c; rv[c] = r[o] - alpha * q[o]; }
  }
  __syncthreads();   // every thread of a camera has read the old r before anyone overwrites it
  double rz = 0.0, rr = 0.0;
  if (on) {
    const size_t o = (size_t)cam * D + k;
    x[o] += alpha * p[o];
    r[o] = rv[k];
    const T* __restrict__ m = Minv + (size_t)cam * (D * D) + k * D;
    T zz = T(0);
#pragma unroll
    for (int c = 0; c < D; ++c) zz += m[c] * rv[c];
    z[o] = zz;
    rz = (double)zz * (double)rv[k];
    rr = (double)rv[k] * (double)rv[k];
  }
  rz = block_sum(rz);
  rr = block_sum(rr);
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    part_rz[blockIdx.x] = rz; part_rr[blockIdx.x] = rr;
    if (MERGED) {
      __threadfence();
      is_last = atomicAdd(&st->ticket2, 1u) == gridDim.x - 1;
    }
  }
  if (!MERGED) return;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += UPD_TPB) { a += __ldcg(part_rz + i); b += __ldcg(part_rr + i); }
  a = block_sum(a);
  b = block_sum(b);
  __shared__ double rho_new_sh;
  if (threadIdx.x == 0) rho_new_sh = a;
  __syncthreads();
  const T beta = (T)(rho_new_sh / rho);
  const int n = n_cam * D;
#pragma unroll 4
  for (int i = threadIdx.x; i < n; i += UPD_TPB) p[i] = __ldcg(z + i) + beta * p[i];
  if (threadIdx.x == 0) {
    st->ticket2 = 0u;
    pcg_finish_iteration(st, a, b, tol2, true, cond, use_cond);
  }
}

// beta = rho_new / rho; p = z + beta p (one thread per vector entry); block 0 publishes the new state
template <typename T, int D>
__global__ void __launch_bounds__(PCG_TPB)
pcg_direction_kernel(int n_cam, int n_part, double tol2, const double* __restrict__ part_rz,
                     const double* __restrict__ part_rr, const T* __restrict__ z, T* __restrict__ p, PcgState* st,
                     cudaGraphConditionalHandle cond, int use_cond) {
  if (st->done) {
    if (use_cond && blockIdx.x == 0 && threadIdx.x == 0) cudaGraphSetConditional(cond, 0u);
    return;
  }
  const double rho_new = reduce_partials(part_rz, n_part);
  const double rr = reduce_partials(part_rr, n_part);
  const double rho = st->rho;
  const T beta = (T)(rho_new / rho);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_cam * D) p[i] = z[i] + beta * p[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) pcg_finish_iteration(st, rho_new, rr, tol2, false, cond, use_cond);
}

}  // namespace isfm
#include "pcg_persistent.cuh"
namespace isfm {

// warps of the persistent PCG grid on the current device (0: that kernel will not be used)
template <typename T, int D>
inline int persistent_grid_warps() {
  if (getenv("ISFM_NO_PERSISTENT")) return 0;
  int dev = 0, coop = 0, n_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  return coop ? n_sm * PersistCfg<T, D>::NW : 0;
}

// deposit positions written by the upper slots [slot_lo, slot_hi): min / max per destination row
static __global__ void own_deposit_range_kernel(int slot_lo, int slot_hi, const int32_t* __restrict__ ucol,
                                                const int32_t* __restrict__ tpos, int32_t* beg, int32_t* end) {
  const int e = slot_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= slot_hi) return;
  const int k = tpos[e];
  if (k < 0) return;   // diagonal or padding slot: no deposit
  atomicMin(beg + ucol[e], k);
  atomicMax(end + ucol[e], k + 1);
}
static __global__ void own_deposit_fix_kernel(int n, int32_t* beg, int32_t* end) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n && end[j] == 0) beg[j] = 0;   // row without own deposits: empty range
}

template <typename T, int D>
struct BlockPCG {
  int n_cam = 0;
  DeviceBuffer<T> x, r, z, p, pp, q, y, yup, C;
  DeviceBuffer<double> part_pq, part_a, part_b;
  // persistent kernel (pcg_persistent.cuh): grid size (0 = unavailable), accumulated phase times
  int persist_grid = 0;
  // Stage stream of the persistent kernel: the stage table of the pattern re-ordered so that warp w
  // of the grid owns the CONTIGUOUS range [warp_stage_ptr[w], warp_stage_ptr[w + 1]) holding the
  // units w, w + W, w + 2 W, ... (W warps in the grid).  All warps therefore advance through the
  // matrix together, as a window of W units (~55 MB) that slides over it -- the DRAM / TLB
  // locality of a grid of short-lived CTAs launched in order -- instead of W far-apart streams.
  DeviceBuffer<int4> stream;
  DeviceBuffer<int32_t> warp_stage_ptr;
  int64_t stream_lo = -1, stream_hi = -1;
  int stream_grid = 0;
  const void* stream_src = nullptr;
  double phase_ms[8] = {0};
  int64_t persist_solves = 0;
  // two-level preconditioner and sparse exchange ranges, set by the owner before solve()
  struct CoarseRef { int enabled = 0, cs = 0, ncl = 0, ncp = 0; const T* Pm = nullptr; const double* Ainv = nullptr; double* rc = nullptr; int* fail = nullptr; } coarse;
  int row_lo[ISFM_MAX_PEERS] = {0}, row_len[ISFM_MAX_PEERS] = {0};
  bool ring_valid = false;
  DeviceBuffer<PcgState> state;
  DeviceBuffer<double> rc_parts;   // persistent kernel, two-level: [grid][maxov][8]
  bool env_push_set = false, env_push_grid = false;   // ISFM_PEER_PUSH given / starts with 'g' (values copied: the environment may change)
  bool env_verify = false, env_no_l2_keep = false, env_no_graph = false;
  PcgState* h_state = nullptr;  // pinned
  // The whole iteration loop of a single-rank solve is ONE graph launch: a conditional WHILE node
  // whose body holds the kernels of one iteration; the last kernel sets the loop condition on the
  // device (converged / breakdown / max_iter), so the host neither polls nor re-launches.
  cudaGraphExec_t graph_exec = nullptr;
  bool graph_disabled = false;
  const T* g_E = nullptr; const T* g_Hd = nullptr; const T* g_Minv = nullptr; double g_tol2 = 0.0; int64_t g_units = -1;
  const void* g_peer_base = nullptr;   // the graph bakes the exchange pointers in
  // split mat-vec: [own_dep_beg[j], own_dep_end[j]) = the deposits of lower row j that come from
  // this rank's units (contiguous: deposits of a row are ordered by source row)
  DeviceBuffer<int32_t> own_dep_beg, own_dep_end;
  bool own_valid = false;
  int64_t g_unit_lo = 0, g_unit_hi = -1;

  ~BlockPCG() {
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (h_state) cudaFreeHost(h_state);
  }

  // `comm` (may be NULL): every rank calls resize with the same n, so the peer exchange can be
  // (re)sized collectively here
  void resize(int n, int64_t n_off, int64_t n_chunks, isfm_comm* comm = nullptr, cudaStream_t s = 0) {
    n_cam = n;
    if (comm_world(comm) > 1) comm_peer_ensure(comm, (size_t)n * D * sizeof(T), s);
    size_t len = (size_t)n * D;
    x.alloc(len); r.alloc(len); z.alloc(len); p.alloc(len); q.alloc(len); y.alloc(len);
    pp.alloc((size_t)n * PersistCfg<T, D>::DP); pp.zero(s);
    // + 16 bytes: the combine phase copies whole 16-byte vectors and may read past a run's last entry
    yup.alloc((size_t)std::max<int64_t>(n_chunks, 1) * D + 16 / sizeof(T));
    C.alloc((size_t)std::max<int64_t>(n_off, 1) * D + 16 / sizeof(T));
    const size_t n_part = (size_t)std::max(n, 1024);   // >= any grid of the persistent kernel
    part_pq.alloc(n_part); part_a.alloc(n_part); part_b.alloc(n_part);
    state.alloc(1);
    own_valid = false;
    coarse = CoarseRef{};
    ring_valid = false;
    stream_lo = stream_hi = -1;
    // environment switches: read once per problem, not per solve
    { const char* e = getenv("ISFM_PEER_PUSH"); env_push_set = e != nullptr; env_push_grid = e && e[0] == 'g'; }
    env_verify = getenv("ISFM_PCG_VERIFY") != nullptr;
    env_no_l2_keep = getenv("ISFM_NO_L2_KEEP") != nullptr; env_no_graph = getenv("ISFM_NO_GRAPH") != nullptr;
    // persistent solve kernel: one CTA per SM, if the device can co-schedule them
    persist_grid = 0;
    if (!getenv("ISFM_NO_PERSISTENT")) {
      int dev = 0, coop = 0, n_sm = 0, per_sm = 0;
      ISFM_CUDA(cudaGetDevice(&dev));
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
      if (coop && cudaFuncSetAttribute(pcg_persistent_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)PersistCfg<T, D>::SMEM) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_persistent_kernel<T, D>, PersistCfg<T, D>::NT,
                                                        PersistCfg<T, D>::SMEM) == cudaSuccess && per_sm >= 1)
        persist_grid = n_sm;
      cudaGetLastError();
    }
    ISFM_CUDA(cudaFuncSetAttribute(pcg_spmv_upper_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)SpmvCfg<T, D>::SMEM));
    if (!h_state) ISFM_CUDA(cudaMallocHost(&h_state, sizeof(PcgState)));
    if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
  }

  // (re)builds the stage stream for the units [unit_lo, unit_hi) and the current grid; no-op when unchanged
  void build_stream(const SchurPattern& sp, int64_t unit_lo, int64_t unit_hi, cudaStream_t s) {
    if (stream_lo == unit_lo && stream_hi == unit_hi && stream_grid == persist_grid && stream_src == (const void*)sp.stages.get()) return;
    const int W = persist_grid * PersistCfg<T, D>::NW;
    const bool interleave = !getenv("ISFM_PCG_BLOCKED_ORDER");   // A/B: contiguous unit ranges per warp instead
    std::vector<int4> h_all((size_t)std::max<int64_t>(sp.n_stages, 1));
    ISFM_CUDA(cudaMemcpyAsync(h_all.data(), sp.stages.get(), (size_t)sp.n_stages * sizeof(int4), cudaMemcpyDeviceToHost, s));
    ISFM_CUDA(cudaStreamSynchronize(s));
    std::vector<int4> out;
    std::vector<int32_t> ptr((size_t)W + 1, 0);
    const int64_t n_units = unit_hi - unit_lo;
    out.reserve((size_t)(sp.h_unit_stage_ptr[(size_t)unit_hi] - sp.h_unit_stage_ptr[(size_t)unit_lo]) + 1);
    for (int w = 0; w < W; ++w) {
      ptr[(size_t)w] = (int32_t)out.size();
      if (interleave) {
        for (int64_t u = unit_lo + w; u < unit_hi; u += W)
          for (int32_t g = sp.h_unit_stage_ptr[(size_t)u]; g < sp.h_unit_stage_ptr[(size_t)u + 1]; ++g) out.push_back(h_all[(size_t)g]);
      } else {
        const int64_t u0 = unit_lo + (int64_t)w * n_units / W, u1 = unit_lo + (int64_t)(w + 1) * n_units / W;
        for (int64_t u = u0; u < u1; ++u)
          for (int32_t g = sp.h_unit_stage_ptr[(size_t)u]; g < sp.h_unit_stage_ptr[(size_t)u + 1]; ++g) out.push_back(h_all[(size_t)g]);
      }
    }
    ptr[(size_t)W] = (int32_t)out.size();
    stream.alloc(std::max<size_t>(out.size(), 1)); warp_stage_ptr.alloc(ptr.size());
    if (!out.empty()) ISFM_CUDA(cudaMemcpyAsync(stream.get(), out.data(), out.size() * sizeof(int4), cudaMemcpyHostToDevice, s));
    ISFM_CUDA(cudaMemcpyAsync(warp_stage_ptr.get(), ptr.data(), ptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    ISFM_CUDA(cudaStreamSynchronize(s));   // the host vectors go out of scope
    stream_lo = unit_lo; stream_hi = unit_hi; stream_grid = persist_grid; stream_src = (const void*)sp.stages.get();
  }

  // split mat-vec set-up: restrict the combine kernel to the deposits of the upper slots [slot_lo, slot_hi)
  void set_owned_slots(const SchurPattern& sp, int64_t slot_lo, int64_t slot_hi, cudaStream_t s) {
    own_dep_beg.alloc(n_cam); own_dep_end.alloc(n_cam);
    ISFM_CUDA(cudaMemsetAsync(own_dep_beg.get(), 0x7f, (size_t)n_cam * sizeof(int32_t), s));   // 0x7f7f7f7f: larger than any position
    ISFM_CUDA(cudaMemsetAsync(own_dep_end.get(), 0, (size_t)n_cam * sizeof(int32_t), s));
    if (slot_hi > slot_lo)
      own_deposit_range_kernel<<<div_up(slot_hi - slot_lo, 256), 256, 0, s>>>((int)slot_lo, (int)slot_hi, sp.ucol.get(), sp.tpos.get(),
                                                                             own_dep_beg.get(), own_dep_end.get());
    own_deposit_fix_kernel<<<div_up(n_cam, 256), 256, 0, s>>>(n_cam, own_dep_beg.get(), own_dep_end.get());
    ISFM_CUDA(cudaGetLastError());
    own_valid = true;
    if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
  }

  // Solves S x = b.  Returns iterations; result in x.  status: 1 converged (persistent kernel: the
  // TRUE residual b - S x, recomputed after the recursive one converged, is below the tolerance),
  // 0 hit max_iter, 2 breakdown, 4 the true residual stagnated above the tolerance (fp32 floor).
  // unit_lo / unit_hi: the mat-vec work units this rank multiplies (all of them unless the ranks
  // share one block pattern and the summed E has been reduce-scattered by unit ranges; the
  // partials and deposits of the other units are zero, see BASolver::setup_matvec_split)
  int solve(const SchurPattern& sp, const T* E,
            const T* Hd, const T* Minv, const T* b, double tol, int max_iter, isfm_comm* comm, cudaStream_t s,
            KernelTimers& kt, int* status_out, int64_t unit_lo = 0, int64_t unit_hi = -1) {
    if (unit_hi < 0) unit_hi = sp.n_chunks;
    const int n_units = (int)(unit_hi - unit_lo);
    // deposit range of every row: the whole lower row, or (split mat-vec) the part this rank writes
    const bool own_ranges = own_dep_beg.get() != nullptr && own_valid;
    const int32_t* dep_beg = own_ranges ? own_dep_beg.get() : sp.lrow_ptr.get();
    const int32_t* dep_end = own_ranges ? own_dep_end.get() : sp.lrow_ptr.get() + 1;
    const int nb = div_up(n_cam, PCG_TPB);
    const int nb_upd = div_up(n_cam, UPD_TPB / D), nb_dir = div_up((int64_t)n_cam * D, PCG_TPB);
    const int nb_comb = (n_cam + 1) / 2, nb_diag = div_up(n_cam, PCG_TPB / D);
    const bool multi = comm_world(comm) > 1;
    const bool peer = multi && comm->peer_ready && (size_t)n_cam * D * sizeof(T) <= comm->px.slot_bytes;
    const PeerExchange px = peer ? comm->px : PeerExchange{};
    // ISFM_PEER_PUSH = "grid" / "fused" forces the push variant (tests); default: by vector size
    const bool push_env = env_push_set;
    const bool big_push = peer && (push_env ? env_push_grid : (size_t)n_cam * D * sizeof(T) > ((size_t)128 << 10));
    const bool merged = (int64_t)n_cam * D <= 4096;   // the last CTA of the update kernel also builds p
    { TimerScope ts(kt, T_PCG_VEC);
      pcg_init_kernel<T, D><<<nb, PCG_TPB, 0, s>>>(n_cam, b, Minv, x.get(), r.get(), p.get(), part_a.get(), part_b.get()); }
    { TimerScope ts(kt, T_PCG_VEC);
      pcg_init_state_kernel<D><<<1, 256, 0, s>>>(n_cam, max_iter, part_a.get(), part_b.get(), state.get()); }
    const double tol2 = tol * tol;
    // ---- persistent path: the whole solve in one cooperative kernel (single rank or peer exchange) ----
    if (persist_grid > 0 && (!multi || peer) && !kt.enabled_fine() && sp.stage_blocks == SpmvCfg<T, D>::WB) {
      PcgArgs<T> a;
      memset(&a, 0, sizeof a);
      a.n_cam = n_cam; a.unit_lo = (int)unit_lo; a.unit_hi = (int)unit_hi; a.max_iter = max_iter;
      a.unit_row = sp.chunk_row.get(); a.unit_beg = sp.chunk_beg.get(); a.urow_ptr = sp.urow_ptr.get(); a.ucol = sp.ucol.get();
      a.tpos = sp.tpos.get(); a.dep_beg = dep_beg; a.dep_end = dep_end; a.chunk_ptr = sp.chunk_ptr.get();
      build_stream(sp, unit_lo, unit_hi, s);
      a.stages = stream.get();
      a.warp_stage_ptr = warp_stage_ptr.get();
      a.E = E; a.Hd = Hd; a.Minv = Minv; a.b = b;
      // ISFM_PCG_VERIFY=1: recompute the true residual when the recursive one has converged and go on
      // from it if it is above the tolerance.  Off by default: measured +60 % PCG iterations at C3 /
      // C5 for an LM trajectory that agrees to 7 digits either way (the fp32 recursive residual under-
      // states the true one, but LM only needs an inexact Newton step).
      a.verify = env_verify ? 1 : 0;
      a.x = x.get(); a.r = r.get(); a.z = z.get(); a.p = p.get(); a.pp = pp.get(); a.q = q.get(); a.y = y.get(); a.yup = yup.get(); a.C = C.get();
      a.part_pq = part_pq.get(); a.part_a = part_a.get(); a.part_b = part_b.get();
      a.st = state.get(); a.tol2 = tol2;
      a.cams_per_cta = div_up(n_cam, persist_grid);
      const int64_t my_slots = n_units > 0 ? (int64_t)std::min<int64_t>((int64_t)n_units * sp.unit_slots, sp.nnzu) : 0;
      a.keep_in_l2 = (size_t)my_slots * D * D * sizeof(T) <= ((size_t)72 << 20) && !env_no_l2_keep;
      a.peer = peer ? 1 : 0;
      a.push_grid = big_push ? 1 : 0;
      a.px = px;
      for (int rk = 0; rk < ISFM_MAX_PEERS; ++rk) { a.row_lo[rk] = ring_valid ? row_lo[rk] : 0; a.row_len[rk] = ring_valid ? row_len[rk] : n_cam; }
      if (peer && ring_valid) {
        // the push variant follows the bytes this rank actually sends
        if (!push_env) a.push_grid = (size_t)row_len[comm_rank(comm)] * D * sizeof(T) > ((size_t)128 << 10);
      }
      a.coarse = coarse.enabled; a.cs = coarse.cs; a.ncl = coarse.ncl; a.ncp = coarse.ncp;
      a.maxov = 0;
      if (coarse.enabled) {   // most clusters a CTA's camera share can overlap; its partial coarse residuals live in rc_parts
        a.maxov = (a.cams_per_cta + coarse.cs - 2) / coarse.cs + 1;
        rc_parts.alloc((size_t)persist_grid * a.maxov * 8);
      }
      a.Pm = coarse.Pm; a.Ainv = coarse.Ainv; a.rc = rc_parts.get(); a.coarse_fail = coarse.fail;
      const size_t upd_smem = (size_t)a.cams_per_cta * (2 * D * sizeof(T) + 64) + 16 + ((size_t)a.ncl * PCG_MODES + (size_t)a.maxov * 8) * 8;
      const size_t persist_smem = PersistCfg<T, D>::SMEM;
      ISFM_REQUIRE(upd_smem <= persist_smem, ISFM_EINVAL, "persistent PCG kernel: camera share per CTA exceeds shared memory");
      a.phase_ns = &state.get()->phase_ns[0];
      if (coarse.enabled) ISFM_CUDA(cudaMemsetAsync(q.get(), 0, (size_t)n_cam * D * sizeof(T), s));   // alpha = 0 pass reads q
      { TimerScope ts(kt, T_PCG_SPMV);
        void* params[] = {&a};
        ISFM_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&pcg_persistent_kernel<T, D>), dim3(persist_grid),
                                              dim3(PersistCfg<T, D>::NT), params, PersistCfg<T, D>::SMEM, s)); }
      ISFM_CUDA(cudaMemcpyAsync(h_state, state.get(), sizeof(PcgState), cudaMemcpyDeviceToHost, s));
      ISFM_CUDA(cudaStreamSynchronize(s));
      if (h_state->abort || h_state->done == 3)
        throw IsfmError(ISFM_ENCCL, h_state->done == 3 ? "peer-memory exchange timed out waiting for a rank (PCG)"
                                                       : "grid barrier timed out inside the persistent PCG kernel");
      for (int i = 0; i < 8; ++i) phase_ms[i] += (double)h_state->phase_ns[i] * 1e-6;
      persist_solves++;
      if (status_out) *status_out = h_state->done;
      return h_state->iters;
    }
    auto launch_iteration = [&](cudaGraphConditionalHandle cond, int use_cond) {
      if (n_units > 0) {
        TimerScope ts(kt, T_PCG_SPMV);
        pcg_spmv_upper_kernel<T, D><<<div_up(n_units, SpmvCfg<T, D>::NW), PCG_TPB, SpmvCfg<T, D>::SMEM, s>>>(
            n_units, sp.unit_slots, sp.chunk_row.get() + unit_lo, sp.chunk_beg.get() + unit_lo, sp.urow_ptr.get(), sp.ucol.get(), sp.tpos.get(), E,
            p.get(), yup.get() + (size_t)unit_lo * D, C.get(), state.get()); }
      if (!multi) {
        TimerScope ts(kt, T_PCG_VEC);
        pcg_combine_kernel<T, D, COMB_FUSED><<<nb_comb, COMB_TPB, 0, s>>>(n_cam, dep_beg, dep_end, sp.chunk_ptr.get(), (int)unit_lo, (int)unit_hi, yup.get(),
                                                                         C.get(), Hd, p.get(), q.get(), part_pq.get(), state.get(), px);
      } else if (peer) {
        // all-reduce of y over peer memory: pushed by the combine epilogue (small systems) or by a
        // grid-wide copy kernel (large ones), summed by the next kernel
        if (big_push) {
          { TimerScope ts(kt, T_PCG_VEC);
            pcg_combine_kernel<T, D, COMB_PLAIN><<<nb_comb, COMB_TPB, 0, s>>>(n_cam, dep_beg, dep_end, sp.chunk_ptr.get(), (int)unit_lo, (int)unit_hi, yup.get(),
                                                                             C.get(), Hd, p.get(), y.get(), part_pq.get(), state.get(), px); }
          { TimerScope ts(kt, T_COMM);
            const size_t len = (size_t)n_cam * D;
            pcg_push_kernel<T><<<(int)std::min<size_t>(div_up(len, PUSH_TPB * (16 / sizeof(T))), 148), PUSH_TPB, 0, s>>>(len, y.get(), state.get(), px); }
        } else {
          TimerScope ts(kt, T_PCG_VEC);
          pcg_combine_kernel<T, D, COMB_PUSH><<<nb_comb, COMB_TPB, 0, s>>>(n_cam, dep_beg, dep_end, sp.chunk_ptr.get(), (int)unit_lo, (int)unit_hi, yup.get(),
                                                                          C.get(), Hd, p.get(), y.get(), part_pq.get(), state.get(), px);
        }
        { TimerScope ts(kt, T_COMM);
          pcg_apply_diag_kernel<T, D, true><<<nb_diag, PCG_TPB, 0, s>>>(n_cam, Hd, p.get(), y.get(), q.get(), part_pq.get(), state.get(), px); }
      } else {
        { TimerScope ts(kt, T_PCG_VEC);
          pcg_combine_kernel<T, D, COMB_PLAIN><<<nb_comb, COMB_TPB, 0, s>>>(n_cam, dep_beg, dep_end, sp.chunk_ptr.get(), (int)unit_lo, (int)unit_hi, yup.get(),
                                                                           C.get(), Hd, p.get(), y.get(), part_pq.get(), state.get(), px); }
        { TimerScope ts(kt, T_COMM);
          comm_allreduce_sum(comm, y.get(), (size_t)n_cam * D, sizeof(T) == 8, s); }
        { TimerScope ts(kt, T_PCG_VEC);
          pcg_apply_diag_kernel<T, D, false><<<nb_diag, PCG_TPB, 0, s>>>(n_cam, Hd, p.get(), y.get(), q.get(), part_pq.get(), state.get(), px); }
      }
      if (merged) {
        TimerScope ts(kt, T_PCG_VEC);
        pcg_update_kernel<T, D, true><<<nb_upd, UPD_TPB, 0, s>>>(n_cam, tol2, Minv, p.get(), q.get(), x.get(), r.get(), z.get(),
                                                                part_a.get(), part_b.get(), state.get(), cond, use_cond);
      } else {
        { TimerScope ts(kt, T_PCG_VEC);
          pcg_update_kernel<T, D, false><<<nb_upd, UPD_TPB, 0, s>>>(n_cam, tol2, Minv, p.get(), q.get(), x.get(), r.get(), z.get(),
                                                                   part_a.get(), part_b.get(), state.get(), cond, 0); }
        { TimerScope ts(kt, T_PCG_VEC);
          pcg_direction_kernel<T, D><<<nb_dir, PCG_TPB, 0, s>>>(n_cam, nb_upd, tol2, part_a.get(), part_b.get(), z.get(), p.get(),
                                                               state.get(), cond, use_cond); }
      }
    };
    const int vec_per_iter = (merged ? 2 : 3) + (multi ? 1 : 0) + (big_push ? 1 : 0);
    // No per-kernel timing and no NCCL call inside the loop (single rank, or the peer-memory
    // exchange): device-side WHILE graph.
    bool use_graph = (!multi || peer) && !kt.enabled && !graph_disabled && !env_no_graph;
    if (use_graph && !(graph_exec && g_E == E && g_Hd == Hd && g_Minv == Minv && g_tol2 == tol2 && g_units == sp.n_chunks && g_unit_lo == unit_lo && g_unit_hi == unit_hi &&
                       g_peer_base == (peer ? (const void*)comm->px.base[0] : nullptr))) {
      if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
      tl_count_launches() = false;                 // captured, not executed
      int64_t saved[ISFM_N_TIMERS];
      for (int i = 0; i < ISFM_N_TIMERS; ++i) saved[i] = kt.launches[i];
      cudaGraph_t graph = nullptr;
      cudaGraphConditionalHandle cond = 0;
      bool ok = cudaGraphCreate(&graph, 0) == cudaSuccess;
      if (ok) ok = cudaGraphConditionalHandleCreate(&cond, graph, 1u, cudaGraphCondAssignDefault) == cudaSuccess;
      cudaGraphNode_t node = nullptr;
      cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
      np.type = cudaGraphNodeTypeConditional;
      np.conditional.handle = cond;
      np.conditional.type = cudaGraphCondTypeWhile;
      np.conditional.size = 1;
      if (ok) ok = cudaGraphAddNode(&node, graph, nullptr, 0, &np) == cudaSuccess;
      if (ok) {
        cudaGraph_t body = np.conditional.phGraph_out[0];
        ok = cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
          launch_iteration(cond, 1);
          cudaGraph_t same = nullptr;
          ok = cudaStreamEndCapture(s, &same) == cudaSuccess;
        }
      }
      if (ok) ok = cudaGraphInstantiate(&graph_exec, graph, 0) == cudaSuccess;
      if (graph) cudaGraphDestroy(graph);
      if (!ok) { graph_exec = nullptr; graph_disabled = true; cudaGetLastError(); }   // plain launches + host polling instead
      tl_count_launches() = true;
      for (int i = 0; i < ISFM_N_TIMERS; ++i) kt.launches[i] = saved[i];
      g_E = E; g_Hd = Hd; g_Minv = Minv; g_tol2 = tol2; g_units = sp.n_chunks; g_unit_lo = unit_lo; g_unit_hi = unit_hi;
      g_peer_base = peer ? (const void*)comm->px.base[0] : nullptr;
      use_graph = graph_exec != nullptr;
    }
    h_state->done = 0; h_state->iters = 0;
    if (use_graph) {
      ISFM_CUDA(cudaGraphLaunch(graph_exec, s));
      ISFM_CUDA(cudaMemcpyAsync(h_state, state.get(), sizeof(PcgState), cudaMemcpyDeviceToHost, s));
      ISFM_CUDA(cudaStreamSynchronize(s));
      const int it = h_state->iters;
      g_launch_count += (int64_t)(1 + vec_per_iter) * it;
      kt.launches[T_PCG_SPMV] += it; kt.launches[T_PCG_VEC] += (int64_t)vec_per_iter * it;
    } else {
      const int check_every = 8;
      int it = 0;
      while (it < max_iter) {
        const int chunk = std::min(check_every, max_iter - it);
        for (int k = 0; k < chunk; ++k) launch_iteration(0, 0);
        it += chunk;
        ISFM_CUDA(cudaMemcpyAsync(h_state, state.get(), sizeof(PcgState), cudaMemcpyDeviceToHost, s));
        ISFM_CUDA(cudaStreamSynchronize(s));
        if (h_state->done) break;
      }
    }
    ISFM_CUDA(cudaGetLastError());
    if (h_state->done == 3) throw IsfmError(ISFM_ENCCL, "peer-memory exchange timed out waiting for a rank (PCG)");
    if (status_out) *status_out = h_state->done;
    return h_state->iters;
  }
};

}  // namespace isfm
