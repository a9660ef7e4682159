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
  static constexpr size_t PER_WARP = 2 * (size_t)STG * sizeof(T);
  static constexpr int NW_RAW = (int)((size_t)222 * 1024 / PER_WARP);
  static constexpr int NW = NW_RAW >= 24 ? 24 : (NW_RAW >= 16 ? 16 : (NW_RAW >= 12 ? 12 : 8));
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
  int coarse, cs, ncl, ncp, kcl;   // cluster size (cameras; the last cluster also takes the remainder), clusters,
                                   // padded coarse dimension, clusters per CTA
  const T* Pm;                // [n_cam][6][PCG_MODES]
  const T* Ainv;              // [ncp][ncp]
  double* rc;                 // [ncp]
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

// Mat-vec phase of one warp: a contiguous range [g0, g1) of the stage table, walked as ONE
// continuous cp.async stream (stage g + 1 is in flight while stage g is multiplied; the per-stage
// descriptors are fetched three, the column / deposit indices two stages ahead).  Unit and row
// changes happen inside the stream -- the stage's own p_i travels with it into shared memory, the
// row partial is folded and stored when a stage is flagged last-of-unit -- so a warp never drains
// its pipeline between work units (the stand-alone kernel, one unit per warp, does).
template <typename T, int D>
__device__ __forceinline__ void spmv_stream(T* buf, int lane, int g0, int g1, const PcgArgs<T>& a, uint64_t pol_stream,
                                            uint64_t pol_keep, bool hint) {
  typedef SpmvCfg<T, D> Cfg;
  typedef PersistCfg<T, D> PC;
  constexpr int GPW = Cfg::GPW, DD = D * D, VE = Cfg::VE, NLD = Cfg::NLD, STG = PC::STG, DP = PC::DP;
  if (g0 >= g1) return;
  const int bl = lane / D, r = lane % D;
  auto meta = [&](int g) -> int4 { return g < g1 ? __ldg(a.stages + g) : make_int4(0, 0, 0, 0); };
  // m.z bit 10: the stage's columns are consecutive (dense rows, banded street rows): no column
  // loads, and the p_j rows form ONE contiguous piece of pp.  Loaded values are only STORED here
  // (consumed one iteration later): no load-to-use stall in the stream.
  auto load_idx = [&](const int4& m, int* jj, int* tp) {
    const int nb = m.z & 0xff;
    const bool contig = (m.z & (1 << 10)) != 0;
#pragma unroll
    for (int pass = 0; pass < Cfg::PASSES; ++pass) {
      const int b = pass * GPW + bl;
      const bool on = bl < GPW && b < nb;
      // contiguous stage: every lane keeps the stage's FIRST column in jj[0] (jj[1..] unused)
      if (pass == 0) jj[0] = contig ? (nb > 0 ? __ldg(a.ucol + m.y) : -1) : (on ? __ldg(a.ucol + m.y + b) : -1);
      else jj[pass] = (!contig && on) ? __ldg(a.ucol + m.y + b) : -1;
      tp[pass] = on ? __ldg(a.tpos + m.y + b) : -1;
    }
  };
  auto issue = [&](int g, const int4& m, const int* jj) {
    const int nb = m.z & 0xff;
    const int last = nb * DD / VE - 1;   // indices past the end re-copy the last vector
    const T* src = a.E + (size_t)m.y * DD;
    T* dst = buf + (size_t)((g - g0) & 1) * STG;
    if (hint) {
#pragma unroll
      for (int q = 0; q < NLD; ++q) cp_async16_hint(dst + (size_t)(lane + 32 * q) * VE, src + (size_t)min(lane + 32 * q, last) * VE, pol_stream);
    } else {
#pragma unroll
      for (int q = 0; q < NLD; ++q) cp_async16(dst + (size_t)(lane + 32 * q) * VE, src + (size_t)min(lane + 32 * q, last) * VE);
    }
    if (m.z & (1 << 10)) {
      // consecutive columns: nb padded rows of pp in one run -- coalesced 16-byte copies (a few
      // 128-byte lines) instead of nb scattered ones (one L2 transaction each)
      const T* psrc = a.pp + (size_t)jj[0] * DP;
      for (int c = lane; c < nb * PC::NCH; c += 32) cp_async16(dst + PC::POFF + c * VE, psrc + c * VE);
    } else {
#pragma unroll
      for (int pass = 0; pass < Cfg::PASSES; ++pass)
        if (jj[pass] >= 0 && r < PC::NCH) cp_async16(dst + PC::POFF + (pass * GPW + bl) * DP + r * VE, a.pp + (size_t)jj[pass] * DP + r * VE);
    }
    if (lane < PC::NCH) cp_async16(dst + PC::PIOFF + lane * VE, a.pp + (size_t)m.x * DP + lane * VE);
    cp_async_commit();
  };
  int4 m0 = meta(g0), m1 = meta(g0 + 1), m2 = meta(g0 + 2);
  int j0[Cfg::PASSES], t0[Cfg::PASSES], j1[Cfg::PASSES], t1[Cfg::PASSES], j2[Cfg::PASSES], t2[Cfg::PASSES];
  load_idx(m0, j0, t0);
  load_idx(m1, j1, t1);
  issue(g0, m0, j0);
  T acc = T(0);
  T pi[D];
  for (int g = g0; g < g1; ++g) {
    const int4 m3 = meta(g + 3);   // descriptor three stages ahead: consumed (load_idx) in the NEXT iteration
    load_idx(m2, j2, t2);
    if (g + 1 < g1) { issue(g + 1, m1, j1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncwarp();
    const T* S = buf + (size_t)((g - g0) & 1) * STG;
    if (m0.z & (1 << 8)) {   // first stage of a unit: a new row, its p_i moves from the stage buffer into registers
#pragma unroll
      for (int c = 0; c < D; ++c) pi[c] = S[PC::PIOFF + c];
    }
    const int nb0 = m0.z & 0xff;
#pragma unroll
    for (int pass = 0; pass < Cfg::PASSES; ++pass) {
      if (bl < GPW && pass * GPW + bl < nb0) {
        const T* B = S + (pass * GPW + bl) * DD;
        const T* pj = S + PC::POFF + (pass * GPW + bl) * DP;
        T t = T(0);
#pragma unroll
        for (int c = 0; c < D; ++c) { acc += B[r * D + c] * pj[c]; t += B[c * D + r] * pi[c]; }
        if (t0[pass] >= 0) st_global_hint(a.C + (size_t)t0[pass] * D + r, t, pol_keep);
      }
    }
    if (m0.z & (1 << 9)) {   // last stage of its unit: fold the GPW block lanes onto lanes 0..D-1, store the unit's row partial
#pragma unroll
      for (int k = 1; k < GPW; ++k) {
        T o = __shfl_sync(0xffffffffu, acc, (lane + k * D) & 31);
        if (lane < D) acc += o;
      }
      if (lane < D) st_global_hint(a.yup + (size_t)m0.w * D + lane, acc, pol_keep);
      acc = T(0);
    }
    __syncwarp();
#pragma unroll
    for (int pass = 0; pass < Cfg::PASSES; ++pass) { j0[pass] = j1[pass]; t0[pass] = t1[pass]; j1[pass] = j2[pass]; t1[pass] = t2[pass]; }
    m0 = m1; m1 = m2; m2 = m3;
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
  T* wbuf = smem + (size_t)w * 2 * PC::STG;

  // combine phase geometry: the row pairs (b, n - 1 - b) -- equal work on a dense system -- are dealt
  // to the CTAs in contiguous shares; a CTA with m pairs forms groups of NW / m warps (all its
  // threads work whether it owns 1 pair or 100), one pair per group and round
  // (multi-rank, local pattern: only the rows of this rank's ring carry anything -- the items are then
  // those rows, one per group, so that ALL CTAs share them)
  const bool ring_items = a.peer && a.row_len[me] < n;
  const int n_pairs = ring_items ? (a.row_len[me] + 1) / 2 : (n + 1) / 2;
  const int pair0 = (int)(((long long)blockIdx.x * n_pairs) / nblk), pair1 = (int)(((long long)(blockIdx.x + 1) * n_pairs) / nblk);
  const int m_pairs = pair1 - pair0;
  const int wpr = m_pairs >= NW ? 1 : NW / max(m_pairs, 1);
  const int gthreads = 32 * wpr, G = gthreads / D;
  const int grp = w / wpr, gpc = NW / wpr, gt = tid - grp * gthreads;   // group in CTA, groups per CTA, thread in group
  const int gg = gt / D, gc = gt % D;
  constexpr int MLP = 16;

  // cameras of the update / direction phases: an even share, or -- two-level -- whole clusters
  constexpr int CPB = NT / D;
  const bool coarse = a.coarse && *a.coarse_fail == 0;
  int cam0, cam1, ncl_cta = 0, first_cl = 0;
  if (a.coarse) {
    first_cl = blockIdx.x * a.kcl;
    ncl_cta = max(0, min(a.kcl, a.ncl - first_cl));
    cam0 = ncl_cta ? first_cl * a.cs : n;
    cam1 = ncl_cta ? (first_cl + ncl_cta == a.ncl ? n : (first_cl + ncl_cta) * a.cs) : n;
  } else {
    cam0 = min(n, blockIdx.x * a.cams_per_cta);
    cam1 = min(n, cam0 + a.cams_per_cta);
  }
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
    spmv_stream<T, D>(wbuf, lane, wg0, wg1, a, pol_stream, pol_keep, !a.keep_in_l2);
    if (!grid_barrier<false>(st, epoch)) return;
    lap(PH_SPMV);

    // ---------------- P2: combine ----------------
    {
      int round = 0;
      for (int item0 = pair0; item0 < pair1; item0 += gpc, ++round) {
        const int item = item0 + grp;
        T* sh = smem + (size_t)(round & 1) * (NW * 32 * 2) + (size_t)grp * (gthreads * 2);   // [half][G][D] per group
        int rows[2] = {-1, -1};
        if (item < pair1 && grp < gpc) {
          if (ring_items) {   // two consecutive rows of the ring
            rows[0] = a.row_lo[me] + 2 * item;
            rows[1] = 2 * item + 1 < a.row_len[me] ? rows[0] + 1 : -1;
            if (rows[0] >= n) rows[0] -= n;
            if (rows[1] >= n) rows[1] -= n;
          } else {
            rows[0] = (int)item;
            rows[1] = n - 1 - (int)item;
            if (rows[1] <= rows[0]) rows[1] = -1;   // middle row of an odd system: once
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = rows[h];
          if (row >= 0 && gg < G) {
            T acc = T(0);
            const int ke = min(__ldg(a.chunk_ptr + row + 1), a.unit_hi);
            for (int k = max(__ldg(a.chunk_ptr + row), a.unit_lo) + gg; k < ke; k += G) acc += __ldcg(a.yup + (size_t)k * D + gc);
            const int end = __ldg(a.dep_end + row);
            int k = __ldg(a.dep_beg + row) + gg;
            T v[MLP];
#pragma unroll
            for (int u = 0; u < MLP; ++u) v[u] = T(0);
            for (; k + (MLP - 1) * G < end; k += MLP * G) {
#pragma unroll
              for (int u = 0; u < MLP; ++u) v[u] += __ldcg(a.C + (size_t)(k + u * G) * D + gc);
            }
            for (; k < end; k += G) v[0] += __ldcg(a.C + (size_t)k * D + gc);
#pragma unroll
            for (int u = 1; u < MLP; ++u) v[0] += v[u];
            sh[(h * G + gg) * D + gc] = acc + v[0];
          }
        }
        __syncthreads();
        if (gt < 2 * D) {
          const int h = gt / D, c = gt % D, row = rows[h];
          if (row >= 0) {
            T y = T(0);
            for (int k = 0; k < G; ++k) y += sh[(h * G + k) * D + c];
            if (!a.peer) {
              const T* __restrict__ hd = a.Hd + (size_t)row * (D * D) + c * D;
              const T* __restrict__ pi = a.p + (size_t)row * D;
              T qv = T(0);
#pragma unroll
              for (int k = 0; k < D; ++k) qv += hd[k] * __ldcg(pi + k);
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
    double* cw = reinterpret_cast<double*>(pcg_smem);                 // [CPB][8] coarse contributions of one pass
    double* rc_loc = cw + (size_t)CPB * 8;                            // [ncl_cta][8]
    if (coarse) {
      for (int i = tid; i < ncl_cta * 8; i += NT) rc_loc[i] = 0.0;
      __syncthreads();
    }
    for (int c0 = cam0; c0 < cam1; c0 += CPB) {
      const int cam = c0 + ucam;
      const bool on = ucam < CPB && cam < cam1;
      T rv[D];
      if (on) {
        if (verify_pass) {   // true residual: q holds S x
#pragma unroll
          for (int c = 0; c < D; ++c) { const size_t o = (size_t)cam * D + c; rv[c] = __ldg(a.b + o) - __ldcg(a.q + o); }
        } else {
#pragma unroll
          for (int c = 0; c < D; ++c) { const size_t o = (size_t)cam * D + c; rv[c] = __ldcg(a.r + o) - alpha * __ldcg(a.q + o); }
        }
      }
      __syncthreads();   // every thread of a camera has read the old r before anyone overwrites it
      if (on) {
        const size_t o = (size_t)cam * D + uk;
        if (!verify_pass) a.x[o] += alpha * __ldcg(a.p + o);
        a.r[o] = rv[uk];
        const T* __restrict__ m = a.Minv + (size_t)cam * (D * D) + uk * D;
        T zz = T(0);
#pragma unroll
        for (int c = 0; c < D; ++c) zz += m[c] * rv[c];
        a.z[o] = zz;
        rz_acc += (double)zz * (double)rv[uk];
        rr_acc += (double)rv[uk] * (double)rv[uk];
        if (coarse && uk < PCG_MODES) {
          // (P_cam^T r_cam)[uk]: the pose part of the residual against mode uk
          const T* __restrict__ pm = a.Pm + (size_t)cam * (6 * PCG_MODES) + uk;
          double s = 0.0;
#pragma unroll
          for (int mm = 0; mm < 6; ++mm) s += (double)pm[mm * PCG_MODES] * (double)rv[mm];
          cw[ucam * 8 + uk] = s;
        }
      }
      if (coarse) {
        __syncthreads();
        if (tid < ncl_cta * PCG_MODES) {
          const int cl = tid / PCG_MODES, k = tid % PCG_MODES;
          const int b0 = max(cam0 + cl * a.cs, c0), b1 = min(cl == ncl_cta - 1 ? cam1 : cam0 + (cl + 1) * a.cs, c0 + CPB);
          double s = rc_loc[cl * 8 + k];
          for (int cc = b0; cc < b1; ++cc) s += cw[(cc - c0) * 8 + k];
          rc_loc[cl * 8 + k] = s;
        }
        __syncthreads();
      }
    }
    if (coarse) {
      if (tid < ncl_cta * PCG_MODES) {
        const int cl = tid / PCG_MODES, k = tid % PCG_MODES;
        a.rc[(first_cl + cl) * PCG_MODES + k] = rc_loc[cl * 8 + k];
      }
      rr_acc = block_sum(rr_acc);
      if (tid == 0) a.part_b[blockIdx.x] = rr_acc;
      if (!grid_barrier<false>(st, epoch)) return;
      lap(PH_UPDATE);
      // zc = rows of Ac^-1 of the own clusters times rc; then z += P zc
      double* zc = reinterpret_cast<double*>(pcg_smem);   // [ncl_cta][8]
      for (int rowi = w; rowi < ncl_cta * PCG_MODES; rowi += NW) {
        const int cl = rowi / PCG_MODES, k = rowi % PCG_MODES;
        double s = 0.0;
        {
          const T* __restrict__ arow = a.Ainv + (size_t)((first_cl + cl) * PCG_MODES + k) * a.ncp;
          for (int j = lane; j < a.ncl * PCG_MODES; j += 32) s += (double)__ldg(arow + j) * __ldcg(a.rc + j);
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) zc[cl * 8 + k] = s;
      }
      __syncthreads();
      rz_acc = 0.0;
      for (int c0 = cam0; c0 < cam1; c0 += CPB) {
        const int cam = c0 + ucam;
        if (ucam < CPB && cam < cam1) {
          const size_t o = (size_t)cam * D + uk;
          T zz = a.z[o];
          if (uk < 6) {
            const double* zcl = zc + min((cam - cam0) / a.cs, ncl_cta - 1) * 8;
            const T* __restrict__ pm = a.Pm + (size_t)cam * (6 * PCG_MODES) + uk * PCG_MODES;
            double s = 0.0;
#pragma unroll
            for (int mm = 0; mm < PCG_MODES; ++mm) s += (double)pm[mm] * zcl[mm];
            zz += (T)s;
            a.z[o] = zz;
          }
          rz_acc += (double)zz * (double)a.r[o];
        }
      }
      __syncthreads();   // zc (shared memory) is dead before block_sum / the next phase reuse it
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
      if (!(isfinite(rho_new) && isfinite(rr))) { done = 2; break; }
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
    for (int c0 = cam0; c0 < cam1; c0 += CPB) {
      const int cam = c0 + ucam;
      if (ucam < CPB && cam < cam1) {
        const size_t o = (size_t)cam * D + uk;
        const T pn = a.z[o] + beta * a.p[o];
        a.p[o] = pn;
        a.pp[(size_t)cam * PC::DP + uk] = pn;
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
