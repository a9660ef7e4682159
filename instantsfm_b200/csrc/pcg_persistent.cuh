// pcg_persistent.cuh -- the whole PCG solve of the reduced camera system as ONE persistent
// cooperative kernel (one CTA per SM, grid-wide barriers between the phases of an iteration).
//
//   per iteration          phase                                                   barrier after
//   P1  mat-vec            y_up, deposits <- E p   (warp per work unit, cp.async ring)      yes
//   P2  combine            q = Hd p - (unit partials + deposits)        [+ push y to peers] yes
//       (multi-rank)       publish flags, wait for every peer, q = Hd p - sum_ranks y       yes
//   P3  update             alpha = rho / p.q ; x, r ; z = Minv r ; [rc = P^T r]             yes
//       (two-level)        zc = Ac^-1 rc (own clusters' rows) ; z += P zc                    yes
//   P4  direction          beta ; p = z + beta p ; convergence test                         yes
//
// Every dot product is reduced the same way on every CTA (per-CTA partials in fp64, summed in CTA
// order after the barrier), so all CTAs -- and, with several ranks, all GPUs -- take the same
// decisions without a broadcast; no floating-point atomics anywhere.  Replaces the WHILE-graph of
// four kernels per iteration (pcg.cuh): the ~10 us kernel boundaries become ~1.5 us barriers.
//
// Two-level preconditioner (optional): M^-1 = blockdiag(S_ii)^-1 + P Ac^-1 P^T, Ac = P^T S P, with
// P the seven similarity modes (3 translations, 3 rotations, scale) of every cluster of
// consecutive cameras expressed in the cameras' left-perturbation tangents (coarse.cuh).  It
// removes the rigid "bending" modes of long camera chains that make block-Jacobi PCG need
// thousands of iterations on city-scale street scenes.
#pragma once
#include <type_traits>

namespace isfm {

constexpr int PCG_MODES = 7;

template <typename T, int D> struct PersistCfg {
  typedef SpmvCfg<T, D> S;
  // The direction vector is gathered from a copy padded to rows of DP elements (16-byte multiples):
  // inside one long-running kernel p changes every iteration, so its gathers must bypass the
  // (non-coherent) L1 -- cp.async.cg exists for 16-byte copies only.
  static constexpr int DP = (D + S::VE - 1) / S::VE * S::VE;
  static constexpr int NCH = DP / S::VE;                              // 16-byte chunks per padded row
  static constexpr int POFF = S::NLD * 32 * S::VE;                    // offset of the p_j area inside a stage buffer
  static constexpr int PIOFF = POFF + S::WB * DP;                     // offset of p_i (the stage's own row)
  static constexpr int STG = PIOFF + DP;                              // elements per stage buffer
  // Ring depth and warps per CTA: 16 warps with a 3-stage ring where that fits (two stages in flight
  // per warp while one is multiplied, and 512 threads leave ptxas 128 registers -- at 24 warps the
  // 80-register cap made it re-materialise every address inside the stream loop), else a 2-stage ring.
  static constexpr int NW3 = (int)((size_t)222 * 1024 / (3 * (size_t)STG * sizeof(T)));
  static constexpr int NST = NW3 >= 16 ? 3 : 2;
  static constexpr size_t PER_WARP = NST * (size_t)STG * sizeof(T);
  static constexpr int NW_RAW = (int)((size_t)222 * 1024 / PER_WARP);
  static constexpr int NW = NW_RAW >= 16 ? 16 : (NW_RAW >= 12 ? 12 : 8);
  static constexpr int NT = NW * 32;
  static constexpr size_t SMEM = (size_t)NW * PER_WARP;
};

enum PcgPhase { PH_SPMV = 0, PH_COMBINE, PH_EXCHANGE, PH_UPDATE, PH_COARSE, PH_DIRECTION, PH_N };

template <typename T>
struct PcgArgs {
  int n_cam, unit_lo, unit_hi, max_iter;
  const int32_t *unit_row, *unit_beg, *urow_ptr, *ucol, *tpos, *dep_beg, *dep_end, *chunk_ptr;
  const int4* stages;          // this solve's stage stream (BlockPCG::stream): the stages of the units
                               // [unit_lo, unit_hi) dealt to the warps of the grid
  const int32_t* warp_stage_ptr;   // [grid * NW + 1] stage range of every warp
  const T *E, *Hd, *Minv, *b;
  int verify;                  // re-compute the TRUE residual b - S x when the recursive one converged (see kernel)
  T *x, *r, *z, *p, *pp, *q, *y, *yup, *C;   // pp: p padded to rows of PersistCfg::DP elements
  double *part_pq, *part_a, *part_b;   // [gridDim.x] per-CTA partials of p.q, r.z, r.r
  PcgState* st;
  double tol2;
  int cams_per_cta;   // cameras of the update / direction phases owned by one CTA (whole clusters when coarse)
  int keep_in_l2;     // this rank's slice of E fits the L2: stream it without the evict_first hint
  // multi-rank exchange of y over peer memory (comm.cuh); row_lo / row_len: circular range of the
  // camera rows rank r contributes to (multiples of 4 cameras), the only rows it pushes
  int peer, push_grid;
  PeerExchange px;
  int row_lo[ISFM_MAX_PEERS], row_len[ISFM_MAX_PEERS];
  // two-level preconditioner (coarse.cuh)
  int coarse, cs, ncl, ncp, maxov;   // cluster size (cameras; the last cluster also takes the remainder), clusters,
                                     // padded coarse dimension, most clusters one CTA's camera share overlaps
  const T* Pm;                // [n_cam][6][PCG_MODES]
  const double* Ainv;         // [ncp][ncp]
  double* rc;                 // [gridDim.x][maxov][8]: per-CTA partial coarse residuals P^T r of the clusters its cameras overlap
  const int* coarse_fail;     // set by the dense inverse when Ac was not positive definite: block-Jacobi only
  unsigned long long* phase_ns;   // [PH_N] accumulated by CTA 0 (may be NULL)
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Grid-wide barrier on a monotonic counter (all CTAs are co-resident: cooperative launch).
// Bounded: a CTA that waits longer than a few seconds raises `abort` and every CTA leaves the
// kernel -- a lost peer or a bug ends the solve with an error instead of hanging the GPU.
// SYS: the fence covers stores to peer GPUs (the flag published after the barrier orders them).
template <bool SYS>
__device__ __forceinline__ bool grid_barrier(PcgState* st, unsigned& epoch) {
  __shared__ int ok__;
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch++;
    if (SYS) __threadfence_system(); else __threadfence();
    atomicAdd(&st->bar, 1u);
    const unsigned target = epoch * gridDim.x;
    int good = 1;
    long long spins = 0;
    while ((int)(ld_acquire_gpu(&st->bar) - target) < 0) {
      if (++spins > (1ll << 25) || *reinterpret_cast<volatile int*>(&st->abort)) {
        *reinterpret_cast<volatile int*>(&st->abort) = 1;
        good = 0;
        break;
      }
      if (spins > 4096) __nanosleep(100);
    }
    ok__ = good;
  }
  __syncthreads();
  return ok__ != 0;
}

// deterministic sum of the per-CTA partials (same order on every CTA): warp 0 strides, shuffle tree
__device__ __forceinline__ void sum_partials2(const double* __restrict__ pa, const double* __restrict__ pb, int n, double& a, double& b) {
  __shared__ double res__[2];
  if (threadIdx.x < 32) {
    double va = 0.0, vb = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) { va += __ldcg(pa + i); if (pb) vb += __ldcg(pb + i); }
    for (int o = 16; o > 0; o >>= 1) { va += __shfl_xor_sync(0xffffffffu, va, o); vb += __shfl_xor_sync(0xffffffffu, vb, o); }
    if (threadIdx.x == 0) { res__[0] = va; res__[1] = vb; }
  }
  __syncthreads();
  a = res__[0]; b = res__[1];
  __syncthreads();
}

// 16-byte shared-memory load of VE consecutive elements (LDS.128)
__device__ __forceinline__ void lds16(const float* p, float* out) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
__device__ __forceinline__ void lds16(const double* p, double* out) {
  const double2 v = *reinterpret_cast<const double2*>(p);
  out[0] = v.x; out[1] = v.y;
}

// Mat-vec phase of one warp: a contiguous range [g0, g1) of the stage table, walked as ONE
// continuous cp.async stream (stage g + 1 is in flight while stage g is multiplied; the per-stage
// descriptors are fetched three, the column / deposit indices two stages ahead).  Unit and row
// changes happen inside the stream -- the stage's own p_i travels with it into shared memory, the
// row partial is folded and stored when a stage is flagged last-of-unit -- so a warp never drains
// its pipeline between work units (the stand-alone kernel, one unit per warp, does).
//
// Instruction diet (the loop is issue / latency bound, not bandwidth bound: ncu, profiles/):
//  * the column and deposit indices of a stage are ONE coalesced load each (lane k holds block k's)
//    and reach the lane that needs them by shuffle -- no per-pass predicated index loads;
//  * lanes beyond GPW * D shadow the last block lane (same loads, same arithmetic, no stores), full
//    stages run a branch-free body, only a row's last stage takes the predicated one;
//  * every shared-memory address is one per-lane base + an immediate; p_j rows are read with 16-byte
//    loads from their padded rows.
template <typename T, int D, bool HINT>
__device__ __forceinline__ void spmv_stream(T* buf, int lane, int g0, int g1, const PcgArgs<T>& a, uint64_t pol_stream,
                                            uint64_t pol_keep) {
  typedef SpmvCfg<T, D> Cfg;
  typedef PersistCfg<T, D> PC;
  constexpr int GPW = Cfg::GPW, WB = Cfg::WB, PASSES = Cfg::PASSES, DD = D * D, VE = Cfg::VE, NLD = Cfg::NLD, STG = PC::STG, DP = PC::DP,
                NCH = PC::NCH;
  constexpr int NIDX = (WB + 31) / 32;             // index registers per lane (WB > 32 only for the 3x3 blocks of GP)
  constexpr int NPC = (WB * NCH + 31) / 32;        // p_j chunk copies per lane and stage
  constexpr unsigned FULL = 0xffffffffu;
  if (g0 >= g1) return;
  const int bl_raw = lane / D, r = lane % D;
  const bool lane_on = bl_raw < GPW;
  const int bl = lane_on ? bl_raw : GPW - 1;
  const int o_row = bl * DD + r * D, o_col = bl * DD + r, o_pj = PC::POFF + bl * DP;
  auto meta = [&](int g) -> int4 { return g < g1 ? __ldg(a.stages + g) : make_int4(0, 0, 0, 0); };
  // lane k (+ 32 i) holds the column / deposit position of the stage's block k (+ 32 i)
  auto load_idx = [&](const int4& m, int* col, int* tp) {
    const int nb = m.z & 0xff;
#pragma unroll
    for (int i = 0; i < NIDX; ++i) {
      const int k = lane + 32 * i;
      const bool on = k < nb;
      col[i] = on ? __ldg(a.ucol + m.y + k) : 0;
      tp[i] = on ? __ldg(a.tpos + m.y + k) : -1;
    }
  };
  auto pick = [&](const int* v, int blk) -> int {   // v of block blk, from the lane that holds it
    int x = __shfl_sync(FULL, v[0], blk & 31);
#pragma unroll
    for (int i = 1; i < NIDX; ++i) { const int y = __shfl_sync(FULL, v[i], blk & 31); if (blk >= 32 * i) x = y; }
    return x;
  };
  auto issue = [&](T* dst, const int4& m, const int* col) {
    const int nb = m.z & 0xff;
    const int nv = nb * DD / VE;                   // rows are padded to multiples of 4 slots: exact
    const T* src = a.E + (size_t)m.y * DD + (size_t)lane * VE;
    T* d = dst + lane * VE;
#pragma unroll
    for (int q = 0; q < NLD; ++q) {
      if (lane + 32 * q < nv) {
        if (HINT) cp_async16_hint(d + q * 32 * VE, src + q * 32 * VE, pol_stream);
        else cp_async16(d + q * 32 * VE, src + q * 32 * VE);
      }
    }
    // p_j: chunk c of the stage = 16-byte piece (c % NCH) of the padded row of block c / NCH's column;
    // consecutive columns (dense rows, banded street rows) make these a few full 128-byte lines
#pragma unroll
    for (int k = 0; k < NPC; ++k) {
      const int c = lane + 32 * k, blk = c / NCH, ch = c - blk * NCH;
      const int j = pick(col, blk);
      if (c < nb * NCH) cp_async16(dst + PC::POFF + c * VE, a.pp + (size_t)j * DP + ch * VE);
    }
    if (lane < NCH) cp_async16(dst + PC::PIOFF + lane * VE, a.pp + (size_t)m.x * DP + lane * VE);
    cp_async_commit();
  };
  // ring of NST stage buffers: stages g + 1 .. g + NST - 1 are in flight while stage g is multiplied;
  // descriptors are fetched NST + 1, indices NST stages ahead
  constexpr int NST = PC::NST;
  int4 m0 = meta(g0), m1 = meta(g0 + 1), m2 = meta(g0 + 2), m3 = NST == 3 ? meta(g0 + 3) : make_int4(0, 0, 0, 0);
  int c0[NIDX], t0[NIDX], c1[NIDX], t1[NIDX], c2[NIDX], t2[NIDX], c3[NIDX], t3[NIDX];
  load_idx(m0, c0, t0);
  load_idx(m1, c1, t1);
  issue(buf, m0, c0);
  if (NST == 3) {
    load_idx(m2, c2, t2);
    if (g0 + 1 < g1) issue(buf + STG, m1, c1);
  }
  T acc = T(0);
  T pi[D];
  int par = 0, nxt = NST - 1;   // ring slots of the stage being multiplied / being issued
  for (int g = g0; g < g1; ++g) {
    const int4 mn = meta(g + NST + 1);   // consumed (load_idx) in the NEXT iteration
    if (NST == 3) {
      load_idx(m3, c3, t3);
      if (g + 2 < g1) { issue(buf + nxt * STG, m2, c2); cp_async_wait<2>(); }
      else if (g + 1 < g1) cp_async_wait<1>();
      else cp_async_wait<0>();
    } else {
      load_idx(m2, c2, t2);
      if (g + 1 < g1) { issue(buf + nxt * STG, m1, c1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    }
    __syncwarp();
    const T* S = buf + par * STG;
    if (m0.z & (1 << 8)) {   // first stage of a unit: a new row, its p_i moves from the stage buffer into registers
      T v[NCH * VE];
#pragma unroll
      for (int k = 0; k < NCH; ++k) lds16(S + PC::PIOFF + k * VE, v + k * VE);
#pragma unroll
      for (int c = 0; c < D; ++c) pi[c] = v[c];
    }
    const int nb0 = m0.z & 0xff;
    auto body = [&](auto full_tag) {
      constexpr bool FULLSTAGE = decltype(full_tag)::value;
#pragma unroll
      for (int pass = 0; pass < PASSES; ++pass) {
        const int blk = pass * GPW + bl;
        const int tp = pick(t0, blk);
        const bool on = FULLSTAGE ? (PASSES * GPW == WB || blk < WB) : blk < nb0;
        if (on) {
          const T* Br = S + o_row + pass * GPW * DD;
          const T* Bc = S + o_col + pass * GPW * DD;
          T pj[NCH * VE];
#pragma unroll
          for (int k = 0; k < NCH; ++k) lds16(S + o_pj + pass * GPW * DP + k * VE, pj + k * VE);
          T t = T(0);
#pragma unroll
          for (int c = 0; c < D; ++c) { acc += Br[c] * pj[c]; t += Bc[c * D] * pi[c]; }
          if (lane_on && tp >= 0) st_global_hint(a.C + (size_t)tp * D + r, t, pol_keep);   // diagonal / padding slots: no deposit
        }
      }
    };
    if (nb0 == WB) body(std::true_type{}); else body(std::false_type{});
    if (m0.z & (1 << 9)) {   // last stage of its unit: fold the GPW block lanes onto lanes 0..D-1, store the unit's row partial
      if (!lane_on) acc = T(0);
#pragma unroll
      for (int k = 1; k < GPW; ++k) {
        T o = __shfl_sync(FULL, acc, (lane + k * D) & 31);
        if (lane < D) acc += o;
      }
      if (lane < D) st_global_hint(a.yup + (size_t)m0.w * D + lane, acc, pol_keep);
      acc = T(0);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NIDX; ++i) {
      c0[i] = c1[i]; t0[i] = t1[i]; c1[i] = c2[i]; t1[i] = t2[i];
      if (NST == 3) { c2[i] = c3[i]; t2[i] = t3[i]; }
    }
    m0 = m1; m1 = m2;
    if (NST == 3) { m2 = m3; m3 = mn; } else { m2 = mn; }
    par = par == NST - 1 ? 0 : par + 1;
    nxt = nxt == NST - 1 ? 0 : nxt + 1;
  }
}

// true when camera row `row` lies in the circular range [lo, lo + len) of an n-row system
__device__ __forceinline__ bool in_ring(int row, int lo, int len, int n) {
  int d = row - lo;
  if (d < 0) d += n;
  return d < len;
}

template <typename T, int D>
__global__ void __launch_bounds__(PersistCfg<T, D>::NT, 1)
pcg_persistent_kernel(const PcgArgs<T> a) {
  typedef PersistCfg<T, D> PC;
  constexpr int NT = PC::NT, NW = PC::NW;
  extern __shared__ __align__(16) unsigned char pcg_smem[];
  T* smem = reinterpret_cast<T*>(pcg_smem);
  PcgState* st = a.st;
  if (st->done) return;   // b == 0 or not finite (pcg_init_state_kernel); uniform over the grid
  const int n = a.n_cam, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int nblk = gridDim.x;
  unsigned epoch = 0;
  double rho = st->rho;
  const double bb = st->bb;
  uint32_t seq = 0;
  const int me = a.px.rank, world = a.peer ? a.px.world : 1;
  if (a.peer) seq = *reinterpret_cast<volatile uint32_t*>(peer_seq(a.px));
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  const bool prof = a.phase_ns != nullptr && blockIdx.x == 0 && tid == 0;
  unsigned long long t_prev = prof ? global_ns() : 0ull;
  auto lap = [&](int ph) {
    if (prof) { const unsigned long long t = global_ns(); a.phase_ns[ph] += t - t_prev; t_prev = t; }
  };

  // mat-vec stages of this warp (whole units: a unit's row partial is produced by exactly one warp)
  const int wg = blockIdx.x * NW + w;
  const int wg0 = __ldg(a.warp_stage_ptr + wg), wg1 = __ldg(a.warp_stage_ptr + wg + 1);
  T* wbuf = smem + (size_t)w * PC::NST * PC::STG;

  // combine phase geometry.  Virtual row sequence of NR rows: single rank / shared pattern -- row(i) =
  // i / 2 for even i, n - 1 - i / 2 for odd i (a short and a long lower-triangle row side by side:
  // equal work per pair on a dense system); multi-rank with a local pattern -- only the rows of this
  // rank's ring carry anything: row(i) = (row_lo + i) mod n.  CTAs take contiguous shares of whole
  // pairs; inside a CTA one warp streams one row at a time (see P2).
  const bool ring_items = a.peer && a.row_len[me] < n;
  const int NR = ring_items ? a.row_len[me] : n;
  const int n_pairs = (NR + 1) / 2;
  const int ri0 = min(NR, 2 * (int)(((long long)blockIdx.x * n_pairs) / nblk)), ri1 = min(NR, 2 * (int)(((long long)(blockIdx.x + 1) * n_pairs) / nblk));
  auto row_at = [&](int i) -> int {
    if (ring_items) { const int rr2 = a.row_lo[me] + i; return rr2 >= n ? rr2 - n : rr2; }
    return (i & 1) ? n - 1 - (i >> 1) : (i >> 1);
  };
  // shared memory of the combine phase: row table and row sums of a batch of RBC rows, then NCB chunk
  // buffers per warp
  constexpr int RBC = 256, VEc = 16 / (int)sizeof(T), G = 32 / D;
  constexpr size_t CHEAD = ((size_t)RBC * 5 * sizeof(int) + (size_t)RBC * D * sizeof(T) + 15) / 16 * 16;
  constexpr int NCB = 2;                                                              // chunk buffers per warp (measured: 2 large beat 4 small)
  constexpr int BUFE = (int)((PC::SMEM - CHEAD) / NW / NCB / sizeof(T)) / VEc * VEc;  // elements per chunk buffer
  constexpr int CHD = (BUFE - VEc) / D;                                               // D-element entries per chunk (room for the alignment skew)
  int* rowtab = reinterpret_cast<int*>(pcg_smem);
  T* ysm = reinterpret_cast<T*>(pcg_smem + (size_t)RBC * 5 * sizeof(int));
  T* cbuf = reinterpret_cast<T*>(pcg_smem + CHEAD) + (size_t)w * NCB * BUFE;

  // cameras of the update / direction phases: an even share of the cameras per CTA.  Two-level: the
  // share overlaps the clusters [cl_lo, cl_lo + nov); the CTA publishes ITS part of their coarse
  // residuals, every CTA sums the parts in CTA order (same numbers everywhere, no atomics).
  constexpr int CPB = NT / D;
  const bool coarse = a.coarse && *a.coarse_fail == 0;
  const int cam0 = min(n, blockIdx.x * a.cams_per_cta), cam1 = min(n, cam0 + a.cams_per_cta);
  auto cluster_of = [&](int cam) { return min(cam / a.cs, a.ncl - 1); };
  int cl_lo = 0, nov = 0;
  if (a.coarse && cam0 < cam1) { cl_lo = cluster_of(cam0); nov = cluster_of(cam1 - 1) - cl_lo + 1; }
  const int ucam = tid / D, uk = tid % D;

  int done = 0, it = 0;
  double rr = bb;
  // two-level: the initial z = M^-1 r needs the coarse correction too: one pass of the update /
  // direction phases with alpha = beta = 0 (q is zeroed by the host) before the first mat-vec
  // Pass kinds.  ITER: a PCG iteration.  INIT (two-level only): update / direction phases alone,
  // alpha = beta = 0.  VERIFY: the recursive residual says "converged" -- in fp32 it keeps
  // shrinking after the true residual has stagnated -- so the true residual r = b - S x is
  // recomputed with one more mat-vec (p := x); if it is above the tolerance the solve goes on from
  // it (restart: p = z, at most two times), else it ends.  The reported residual is a true one.
  enum { ITER = 0, INIT = 1, VERIFY = 2 };
  int mode = coarse ? INIT : ITER;
  int restarts = 0;
  bool first = mode == INIT;
  if (!first) {   // padded copy of the initial direction (pcg_init_kernel wrote p)
    for (int c0 = cam0; c0 < cam1; c0 += CPB) {
      const int cam = c0 + ucam;
      if (ucam < CPB && cam < cam1) a.pp[(size_t)cam * PC::DP + uk] = a.p[(size_t)cam * D + uk];
    }
    if (!grid_barrier<false>(st, epoch)) return;
  }
  while (true) {
    double pq_acc = 0.0;
    const int parity = (int)((seq + 1u) & 1u);
    if (mode != INIT) {
    // ---------------- P1: mat-vec ----------------
    if (a.keep_in_l2) spmv_stream<T, D, false>(wbuf, lane, wg0, wg1, a, pol_stream, pol_keep);
    else spmv_stream<T, D, true>(wbuf, lane, wg0, wg1, a, pol_stream, pol_keep);
    if (!grid_barrier<false>(st, epoch)) return;
    lap(PH_SPMV);

    // ---------------- P2: combine ----------------
    // y_row = sum of the row's deposits (one contiguous run of C) + the unit partials of the row (one
    // contiguous run of yup).  Both runs are D-element entries and stream through the warp's two
    // chunk buffers with 16-byte cp.async (the next chunk -- of this row or the warp's next row -- is
    // in flight while the current one is summed): one L2 round trip per 6 KB instead of one per batch
    // of 4-byte loads.  Row sums land in shared memory; q = Hd p - y (or the push to the peers)
    // follows for the whole batch with every thread busy.
    for (int rb0 = ri0; rb0 < ri1; rb0 += RBC) {
      const int nbr = min(RBC, ri1 - rb0);
      __syncthreads();   // previous batch finished with the row table / sums
      for (int j = tid; j < nbr; j += NT) {
        const int row = row_at(rb0 + j);
        rowtab[5 * j] = row;
        rowtab[5 * j + 1] = __ldg(a.dep_beg + row);
        rowtab[5 * j + 2] = __ldg(a.dep_end + row);
        rowtab[5 * j + 3] = max(__ldg(a.chunk_ptr + row), a.unit_lo);
        rowtab[5 * j + 4] = min(__ldg(a.chunk_ptr + row + 1), a.unit_hi);
      }
      for (int idx = tid; idx < nbr * D; idx += NT) ysm[idx] = T(0);
      __syncthreads();
      {
        // cursor over the chunks of this warp's rows j = w, w + NW, ...: (row slot, run, entry range)
        int j = w, run = 0, k = 0, ke = 0;
        bool open = false;
        auto next = [&](int& cj, int& crun, int& ck0, int& ck1) -> bool {
          while (true) {
            if (!open) {
              if (j >= nbr) return false;
              k = rowtab[5 * j + 1 + 2 * run]; ke = rowtab[5 * j + 2 + 2 * run]; open = true;
            }
            if (k < ke) { cj = j; crun = run; ck0 = k; ck1 = min(k + CHD, ke); k = ck1; return true; }
            open = false;
            if (run == 0) run = 1; else { run = 0; j += NW; }
          }
        };
        // one commit group per call, empty when there is no chunk: the wait count below stays constant
        auto issue = [&](T* dst, bool valid, int crun, int ck0, int ck1) {
          if (valid) {
            const T* base = crun ? a.yup : a.C;
            const size_t e0 = (size_t)ck0 * D, a0 = e0 / VEc * VEc;
            const int nvec = (int)(((size_t)ck1 * D + VEc - 1) / VEc - a0 / VEc);   // may read < 16 bytes past the run: the arrays are padded
            const T* src = base + a0;
            for (int v = lane; v < nvec; v += 32) cp_async16(dst + v * VEc, src + (size_t)v * VEc);
          }
          cp_async_commit();
        };
        const int lg = lane / D, lc = lane % D;
        // descriptors of the chunks in flight: [0] is consumed next
        int qj[NCB - 1], qrun[NCB - 1], qk0[NCB - 1], qk1[NCB - 1];
        bool qv[NCB - 1];
#pragma unroll
        for (int i = 0; i < NCB - 1; ++i) {
          qj[i] = qrun[i] = qk0[i] = qk1[i] = 0;
          qv[i] = next(qj[i], qrun[i], qk0[i], qk1[i]);
          issue(cbuf + i * BUFE, qv[i], qrun[i], qk0[i], qk1[i]);
        }
        int pb = 0, acc_j = -1;
        T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
        auto flush = [&]() {
          T v = (acc0 + acc1) + (acc2 + acc3);
          if (lg >= G) v = T(0);
#pragma unroll
          for (int g2 = 1; g2 < G; ++g2) { const T o = __shfl_sync(0xffffffffu, v, (lane + g2 * D) & 31); if (lane < D) v += o; }
          if (lane < D && acc_j >= 0) ysm[acc_j * D + lane] = v;
          acc0 = acc1 = acc2 = acc3 = T(0);
        };
        while (qv[0]) {
          int nj = 0, nrun = 0, nk0 = 0, nk1 = 0;
          const bool hn = next(nj, nrun, nk0, nk1);
          const int pin = pb == 0 ? NCB - 1 : pb - 1;   // the buffer consumed in the previous round
          issue(cbuf + pin * BUFE, hn, nrun, nk0, nk1);
          cp_async_wait<NCB - 1>();
          __syncwarp();
          if (qj[0] != acc_j) { flush(); acc_j = qj[0]; }
          {
            const T* src = cbuf + pb * BUFE + (int)(((size_t)qk0[0] * D) % VEc) + lc;
            const int m = qk1[0] - qk0[0];
            if (lg < G) {
              int d = lg;
              for (; d + 3 * G < m; d += 4 * G) {
                acc0 += src[d * D]; acc1 += src[(d + G) * D]; acc2 += src[(d + 2 * G) * D]; acc3 += src[(d + 3 * G) * D];
              }
              for (; d < m; d += G) acc0 += src[d * D];
            }
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i + 1 < NCB - 1; ++i) { qj[i] = qj[i + 1]; qrun[i] = qrun[i + 1]; qk0[i] = qk0[i + 1]; qk1[i] = qk1[i + 1]; qv[i] = qv[i + 1]; }
          qj[NCB - 2] = nj; qrun[NCB - 2] = nrun; qk0[NCB - 2] = nk0; qk1[NCB - 2] = nk1; qv[NCB - 2] = hn;
          pb = pb == NCB - 1 ? 0 : pb + 1;
        }
        cp_async_wait<0>();
        flush();
      }
      __syncthreads();
      for (int idx = tid; idx < nbr * D; idx += NT) {
        const int jj = idx / D, c = idx - jj * D, row = rowtab[5 * jj];
        const T y = ysm[idx];
        if (!a.peer) {
          const T* __restrict__ hd = a.Hd + (size_t)row * (D * D) + c * D;
          const T* __restrict__ pi = a.p + (size_t)row * D;
          T qv = T(0);
#pragma unroll
          for (int k2 = 0; k2 < D; ++k2) qv += hd[k2] * __ldcg(pi + k2);
          qv -= y;
          a.q[(size_t)row * D + c] = qv;
          pq_acc += (double)qv * (double)__ldcg(pi + c);
        } else if (a.push_grid) {
          a.y[(size_t)row * D + c] = y;
        } else {
          for (int dst = 0; dst < world; ++dst) peer_slot<T>(a.px, dst, parity, me)[(size_t)row * D + c] = y;
        }
      }
    }
    if (a.peer) {
      if (a.push_grid) {
        // the rows this rank touches leave with coalesced 16-byte remote stores from the whole grid
        if (!grid_barrier<false>(st, epoch)) return;
        constexpr int VE = 16 / sizeof(T);
        const int lo = a.row_lo[me], len = a.row_len[me];
        const int len0 = min(len, n - lo), len1 = len - len0;   // [lo, lo + len0) and the wrapped [0, len1)
        for (int piece = 0; piece < 2; ++piece) {
          const size_t off = piece == 0 ? (size_t)lo * D : 0, cnt = (size_t)(piece == 0 ? len0 : len1) * D;
          const size_t nv = cnt / VE;   // row_lo / row_len are multiples of 4 cameras: 16-byte aligned, no tail
          for (size_t i = (size_t)blockIdx.x * NT + tid; i < nv; i += (size_t)nblk * NT) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(a.y + off) + i);
            for (int dst = 0; dst < world; ++dst) reinterpret_cast<float4*>(peer_slot<T>(a.px, dst, parity, me) + off)[i] = v;
          }
          for (size_t i = nv * VE + (size_t)blockIdx.x * NT + tid; i < cnt; i += (size_t)nblk * NT) {
            const T v = a.y[off + i];
            for (int dst = 0; dst < world; ++dst) peer_slot<T>(a.px, dst, parity, me)[off + i] = v;
          }
        }
      }
      if (!grid_barrier<true>(st, epoch)) return;
      lap(PH_COMBINE);
      seq += 1u;
      if (blockIdx.x == 0 && tid < world) {
        __threadfence_system();
        st_release_sys(peer_flags(a.px, tid, parity) + me, seq);
      }
      if (!peer_wait_all(a.px, parity, seq)) {   // a peer never arrived: every CTA of every rank times out alike
        if (tid == 0) { *reinterpret_cast<volatile int*>(&st->abort) = 1; st->done = 3; }
        return;
      }
      // q = Hd p - sum over ranks (in rank order: bit-identical on every rank) of their y
      for (int c0 = cam0; c0 < cam1; c0 += CPB) {
        const int cam = c0 + ucam;
        if (ucam < CPB && cam < cam1) {
          const T* __restrict__ hd = a.Hd + (size_t)cam * (D * D) + uk * D;
          const T* __restrict__ pi = a.p + (size_t)cam * D;
          const size_t o = (size_t)cam * D + uk;
          T v = T(0);
#pragma unroll
          for (int c = 0; c < D; ++c) v += hd[c] * __ldcg(pi + c);
          T ysum = T(0);
          for (int src = 0; src < world; ++src)
            if (in_ring(cam, a.row_lo[src], a.row_len[src], n)) ysum += __ldcg(peer_slot<T>(a.px, me, parity, src) + o);
          v -= ysum;
          a.q[o] = v;
          pq_acc += (double)v * (double)__ldcg(pi + uk);
        }
      }
    }
    pq_acc = block_sum(pq_acc);
    if (tid == 0) a.part_pq[blockIdx.x] = pq_acc;
    if (!grid_barrier<false>(st, epoch)) return;
    lap(a.peer ? PH_EXCHANGE : PH_COMBINE);
    }   // mode != INIT

    // ---------------- P3: update ----------------
    T alpha = T(0);
    if (mode == ITER) {
      double pq, dummy;
      sum_partials2(a.part_pq, nullptr, nblk, pq, dummy);
      if (!(pq > 0.0) || !isfinite(pq)) { done = 2; break; }   // breakdown: x keeps the last good iterate
      alpha = (T)(rho / pq);
    }
    const bool verify_pass = mode == VERIFY;
    double rz_acc = 0.0, rr_acc = 0.0;
    // Shared memory of the update / coarse / direction phases (the stage buffers are idle): the new
    // residual and z of the CTA's cameras, their coarse contributions, the coarse residual and the
    // coarse solution of the overlapped clusters.  Every global value is touched once, by one thread.
    const int ncam_cta = cam1 - cam0;
    T* rs = reinterpret_cast<T*>(pcg_smem);                                   // [cams_per_cta][D]
    T* zs = rs + (size_t)a.cams_per_cta * D;                                  // [cams_per_cta][D]
    double* cw = reinterpret_cast<double*>(pcg_smem + ((size_t)a.cams_per_cta * D * sizeof(T) * 2 + 15) / 16 * 16);   // [cams_per_cta][8]
    double* rcs = cw + (size_t)a.cams_per_cta * 8;                            // [ncl * PCG_MODES]
    double* zc = rcs + (size_t)a.ncl * PCG_MODES;                             // [nov][8]
    {
      const size_t e0 = (size_t)cam0 * D;
      const int ne = ncam_cta * D;
      for (int e = tid; e < ne; e += NT) {
        const size_t o = e0 + e;
        T rn;
        if (verify_pass) rn = __ldg(a.b + o) - __ldcg(a.q + o);   // true residual: q holds S x
        else { rn = __ldcg(a.r + o) - alpha * __ldcg(a.q + o); a.x[o] += alpha * __ldcg(a.p + o); }
        a.r[o] = rn;
        rs[e] = rn;
        rr_acc += (double)rn * (double)rn;
      }
    }
    __syncthreads();
    for (int c0 = 0; c0 < ncam_cta; c0 += CPB) {
      const int lcam = c0 + ucam;
      if (ucam < CPB && lcam < ncam_cta) {
        const int cam = cam0 + lcam;
        const T* __restrict__ m = a.Minv + (size_t)cam * (D * D) + uk * D;
        const T* rv = rs + lcam * D;
        T zz = T(0);
#pragma unroll
        for (int c = 0; c < D; ++c) zz += __ldg(m + c) * rv[c];
        zs[lcam * D + uk] = zz;
        if (!coarse) a.z[(size_t)cam * D + uk] = zz;
        rz_acc += (double)zz * (double)rv[uk];
        if (coarse && uk < PCG_MODES) {
          // (P_cam^T r_cam)[uk]: the pose part of the residual against mode uk
          const T* __restrict__ pm = a.Pm + (size_t)cam * (6 * PCG_MODES) + uk;
          double sum = 0.0;
#pragma unroll
          for (int mm = 0; mm < 6; ++mm) sum += (double)__ldg(pm + mm * PCG_MODES) * (double)rv[mm];
          cw[lcam * 8 + uk] = sum;
        }
      }
    }
    if (coarse) {
      __syncthreads();
      // this CTA's part of the coarse residual of every cluster it overlaps: a warp per (cluster, mode)
      for (int pr = w; pr < nov * PCG_MODES; pr += NW) {
        const int slot = pr / PCG_MODES, k = pr - slot * PCG_MODES, cl = cl_lo + slot;
        const int b0 = max(cl * a.cs, cam0) - cam0, b1 = min(cl == a.ncl - 1 ? n : (cl + 1) * a.cs, cam1) - cam0;
        double sum = 0.0;
        for (int cc = b0 + lane; cc < b1; cc += 32) sum += cw[cc * 8 + k];
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) a.rc[((size_t)blockIdx.x * a.maxov + slot) * 8 + k] = sum;
      }
      rr_acc = block_sum(rr_acc);
      if (tid == 0) a.part_b[blockIdx.x] = rr_acc;
      if (!grid_barrier<false>(st, epoch)) return;
      lap(PH_UPDATE);
      // the whole coarse residual, summed from the CTAs' parts in CTA order; then zc = rows of Ac^-1
      // of the clusters this CTA's cameras overlap times rc; then z += P zc
      for (int idx = tid; idx < a.ncl * PCG_MODES; idx += NT) {
        const int cl = idx / PCG_MODES, k = idx - cl * PCG_MODES;
        const int c_lo = cl * a.cs, c_hi = cl == a.ncl - 1 ? n : (cl + 1) * a.cs;
        const int b_lo = c_lo / a.cams_per_cta, b_hi = (c_hi - 1) / a.cams_per_cta;
        double sum = 0.0;
        for (int bb2 = b_lo; bb2 <= b_hi; ++bb2) {
          const int slot = cl - cluster_of(bb2 * a.cams_per_cta);
          sum += __ldcg(a.rc + ((size_t)bb2 * a.maxov + slot) * 8 + k);
        }
        rcs[idx] = sum;
      }
      __syncthreads();
      for (int rowi = w; rowi < nov * PCG_MODES; rowi += NW) {
        const int slot = rowi / PCG_MODES, k = rowi % PCG_MODES;
        double sum = 0.0;
        {
          const double* __restrict__ arow = a.Ainv + (size_t)((cl_lo + slot) * PCG_MODES + k) * a.ncp;
          for (int j = lane; j < a.ncl * PCG_MODES; j += 32) sum += __ldg(arow + j) * rcs[j];
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) zc[slot * 8 + k] = sum;
      }
      __syncthreads();
      rz_acc = 0.0;
      for (int c0 = 0; c0 < ncam_cta; c0 += CPB) {
        const int lcam = c0 + ucam;
        if (ucam < CPB && lcam < ncam_cta) {
          const int cam = cam0 + lcam;
          T zz = zs[lcam * D + uk];
          if (uk < 6) {
            const double* zcl = zc + (cluster_of(cam) - cl_lo) * 8;
            const T* __restrict__ pm = a.Pm + (size_t)cam * (6 * PCG_MODES) + uk * PCG_MODES;
            double sum = 0.0;
#pragma unroll
            for (int mm = 0; mm < PCG_MODES; ++mm) sum += (double)__ldg(pm + mm) * zcl[mm];
            zz += (T)sum;
            zs[lcam * D + uk] = zz;
          }
          a.z[(size_t)cam * D + uk] = zz;
          rz_acc += (double)zz * (double)rs[lcam * D + uk];
        }
      }
      rz_acc = block_sum(rz_acc);
      if (tid == 0) a.part_a[blockIdx.x] = rz_acc;
      if (!grid_barrier<false>(st, epoch)) return;
      lap(PH_COARSE);
    } else {
      rz_acc = block_sum(rz_acc);
      rr_acc = block_sum(rr_acc);
      if (tid == 0) { a.part_a[blockIdx.x] = rz_acc; a.part_b[blockIdx.x] = rr_acc; }
      if (!grid_barrier<false>(st, epoch)) return;
      lap(PH_UPDATE);
    }

    // ---------------- P4: direction ----------------
    double rho_new;
    sum_partials2(a.part_a, a.part_b, nblk, rho_new, rr);
    T beta = T(0);
    bool to_verify = false;
    if (mode == ITER) {
      ++it;
      if (!(isfinite(rho_new) && isfinite(rr)) || (!(rho_new > 0.0) && rr > 0.0)) { done = 2; break; }   // r.z <= 0: preconditioner not SPD
      if (rr < a.tol2 * bb) {
        if (!a.verify) { done = 1; break; }
        to_verify = true;               // p := x below, then one mat-vec for the true residual
      } else if (it >= a.max_iter) break;
      beta = (T)(rho_new / rho);
    } else if (mode == VERIFY) {
      ++it;   // a verification pass costs one mat-vec: counted as an iteration
      if (!(isfinite(rho_new) && isfinite(rr))) { done = 2; break; }
      // rr is the TRUE squared residual now (a small slack: the two residuals differ by rounding)
      if (rr < 2.25 * a.tol2 * bb || restarts >= 2 || it >= a.max_iter) { done = rr < 2.25 * a.tol2 * bb ? 1 : 4; break; }
      ++restarts;                        // go on from the true residual: p = z
    } else if (!isfinite(rho_new)) { done = 2; break; }
    first = false;
    mode = to_verify ? VERIFY : ITER;
    rho = rho_new;
    if (to_verify) {
      for (int c0 = cam0; c0 < cam1; c0 += CPB) {
        const int cam = c0 + ucam;
        if (ucam < CPB && cam < cam1) {
          const size_t o = (size_t)cam * D + uk;
          const T xv = a.x[o];
          a.p[o] = xv;
          a.pp[(size_t)cam * PC::DP + uk] = xv;
        }
      }
      if (!grid_barrier<false>(st, epoch)) return;
      lap(PH_DIRECTION);
      continue;
    }
    {
      // p = z + beta p: z of the CTA's cameras is still in shared memory
      const int ne = ncam_cta * D;
      for (int e = tid; e < ne; e += NT) {
        const int lcam = e / D, k2 = e - lcam * D;
        const size_t o = (size_t)cam0 * D + e;
        const T pn = zs[e] + beta * __ldcg(a.p + o);
        a.p[o] = pn;
        a.pp[(size_t)(cam0 + lcam) * PC::DP + k2] = pn;
      }
    }
    if (!grid_barrier<false>(st, epoch)) return;
    lap(PH_DIRECTION);
  }
  if (blockIdx.x == 0 && tid == 0) {
    st->iters = it; st->done = done; st->rr = rr; st->rho = rho;
    if (a.peer) *peer_seq(a.px) = seq;
  }
}

}  // namespace isfm
