// ba_solver.cuh -- host driver of one bae.optim.LM.step for bundle adjustment, solved by
// Schur complement + block-Jacobi PCG instead of the reference's full-system Jacobi PCG
// (SURVEY.md 9.3).  One instance per handle; T in {float, double}, MODEL = CameraModelId.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "ba_kernels.cuh"
#include <chrono>

#include "coarse.cuh"
#include "comm.cuh"
#include "index_prep.cuh"
#include "pcg.cuh"

namespace isfm {

// pp.optim.strategy.TrustRegion; scalars live on the host in double.
struct TrustRegionState {
  double radius, high = 0.5, low = 1e-3, up, down0, down, factor = 0.5, max, min = 1e-6, damping;
  void init(double radius_, double max_, double up_, double down_) {
    radius = radius_; max = max_; up = up_; down0 = down = down_; damping = 1.0 / radius_;
  }
  double update(double last, double loss, double model_term /* (JD)^T (2R + JD) */) {
    double denom = -model_term;
    double quality = denom != 0.0 ? (last - loss) / denom : 0.0;
    double r = 1.0 / damping;
    if (quality > high) { r = up * r; down = down0; }
    else if (quality > low) { down = down0; }
    else { r = r * down; down = down * factor; }
    down = std::max(min, std::min(down, max));
    r = std::max(min, std::min(r, max));
    radius = r; damping = 1.0 / r;
    return quality;
  }
};

struct BASolverBase {
  virtual ~BASolverBase() {}
  virtual void set_problem(int64_t n_cam, int64_t n_pt, int64_t n_obs, const void* cam, const void* pp, const void* pts,
                           const void* obs, const int32_t* cam_idx, const int32_t* pt_idx) = 0;
  virtual void step(double* loss_out, isfm_step_stats* stats) = 0;
  virtual void get_params(void* cam_out, void* pts_out) = 0;
  virtual void set_params(const void* cam, const void* pts) = 0;
  virtual void cost(double* robust, double* sq) = 0;
  virtual void get_structure(int32_t* obs_perm, int64_t* pt_off, int32_t* cam_perm, int64_t* cam_off) = 0;
  virtual void get_schur_pattern(int64_t* nnzb, int64_t* n_pairs, int64_t* row_ptr, int32_t* col_idx) = 0;
  virtual void debug_get(int what, void* dst) = 0;
  virtual void get_pcg_phases(double* ms_out, int64_t* solves_out, int32_t* two_level_out) = 0;
  KernelTimers timers;
  bool has_problem = false;
  int64_t matvec_units_owned = 0, matvec_units_total = 0;   // isfm_ba_get_matvec_units
};

template <typename T, int MODEL>
struct BASolver : BASolverBase {
  static constexpr int NI = ModelTraits<MODEL>::NI;
  static constexpr int D = 6 + NI;
  static constexpr int CW = 7 + NI;

  isfm_ba_desc desc;
  cudaStream_t s;
  isfm_comm* comm;
  TrustRegionState tr;
  int64_t n_cam = 0, n_pt = 0, n_obs = 0;
  ObsIndex ix;
  SchurPattern sp;
  DeviceBuffer<T> cam[2], camq[2], pts[2], pp, obs;   // camq: packed [cam row | pp | pad] rows read by the kernels
  static constexpr int CWP = CamPack<CW>::CWP;
  static constexpr int REC = ObsRec<D>::REC;
  DeviceBuffer<T> R, OBS, HPP, GPT, HPPINV, TP, DP, DCQ;
  DeviceBuffer<T> HCC_GC, HME, HD, E, MINV, bvec;  // HME = [Hcc - E_ii | diag Hcc | g_c - e] per camera: one all-reduce per trial
  DeviceBuffer<double> part_a, part_b, part_c, scalars, red_stage;
  DeviceBuffer<int> fail;
  double* h_scalars = nullptr;  // pinned [4]
  BlockPCG<T, D> pcg;
  CoarseLevel<T, D> coarse;   // two-level preconditioner (chain-like camera graphs), see coarse.cuh
  int coarse_fallbacks = 0;   // solves repeated with block-Jacobi alone (see run_schur_and_pcg)
  int coarse_age = -1, coarse_period = 3, iters_at_factor = -1;   // LM steps since the coarse inverse was built; rebuild period
  bool coarse_refresh = false;
  int cur = 0;
  bool have_loss = false;
  double loss = 0.0;
  double mu_last = 1.0;
  CUtensorMap obs_tmap;       // OBS as a 2-D tensor (TMA store of the record slabs), valid when tma_ok
  CUtensorMap obs_tmap_big;   // the same tensor with boxes of tma_big_rows records
  int tma_big_rows = 0;
  bool tma_ok = false;
  bool fused_ok = false;      // every track fits one CTA of the fused K1 (<= FUSED_TPB observations)
  int n_fused_cta = 0;        // tiles of whole points with <= FUSED_TPB observations
  DeviceBuffer<int4> fused_tiles;   // {first point, end point, first observation, observations} per CTA
  // Multi-rank, identical block pattern on every rank (dense co-visibility): the summed E is
  // reduce-scattered by ranges of mat-vec units once per trial and each rank multiplies only its
  // own range in every PCG iteration (see setup_matvec_split).
  bool split_matvec = false;
  int64_t unit_lo = 0, unit_hi = -1;
  std::vector<size_t> split_off, split_cnt;   // element ranges of E per owner rank
  DeviceBuffer<T> E_own;                      // sum over ranks of this rank's range of E
  bool union_pattern = false;                 // the block pattern is the union over ranks (zero blocks where no local pairs)
  bool debug = false, no_tile_backsub = false;   // environment switches, read once at creation
  // Floor on the LM damping used to build the systems (DESIGN.md, fp32 conditioning).  The reduced
  // system S = damp(Hcc) - E is stored in T: along the seven gauge directions of the scene only the
  // damping term is left of it, and below ~16 eps_T that term is smaller than the rounding noise of
  // the stored blocks -- S turns numerically indefinite (PCG breakdown, rejected trials; measured at
  // C3: 67 rejected trials and a diverged cost within 25 LM steps without the floor, none with it).
  // The trust region itself (radius, reported damping) is not touched.
  double min_damping = sizeof(T) == 4 ? 1e-6 : 0.0;

  explicit BASolver(const isfm_ba_desc& d) : desc(d) {
    s = static_cast<cudaStream_t>(d.stream);   // never the legacy default stream: isfm_ba_create substitutes an own stream
    comm = d.comm;
    timers.stream = s;
    tr.init(d.tr_radius, d.tr_max, d.tr_up, d.tr_down);
    ISFM_CUDA(cudaMallocHost(&h_scalars, 4 * sizeof(double)));
    if (const char* e = getenv("ISFM_MIN_DAMPING")) min_damping = atof(e);
    debug = getenv("ISFM_DEBUG") != nullptr;
    if (const char* e = getenv("ISFM_COARSE_PERIOD")) coarse_period = std::max(1, atoi(e));
    no_tile_backsub = getenv("ISFM_NO_TILE_BACKSUB") != nullptr;
  }
  ~BASolver() override {
    cudaStreamSynchronize(s);   // buffers go back to the stream-ordered pool after all work has finished
    if (h_scalars) cudaFreeHost(h_scalars);
  }

  T* HCC() { return HCC_GC.get(); }
  T* GC() { return HCC_GC.get() + (size_t)n_cam * D * D; }

  template <typename U>
  void upload(DeviceBuffer<U>& dst, const void* src, size_t count) {
    dst.alloc(count);
    ISFM_CUDA(cudaMemcpyAsync(dst.get(), src, count * sizeof(U), cudaMemcpyDefault, s));
  }

  void set_problem(int64_t nc, int64_t np, int64_t no, const void* cam_in, const void* pp_in, const void* pts_in,
                   const void* obs_in, const int32_t* cam_idx, const int32_t* pt_idx) override {
    ISFM_REQUIRE(cam_in && pp_in && pts_in && obs_in && cam_idx && pt_idx, ISFM_EINVAL, "null input");
    n_cam = nc; n_pt = np; n_obs = no;
    // ISFM_DEBUG_SETUP=1: wall-clock per set-up phase (each mark synchronises the stream)
    const bool dbg_setup = getenv("ISFM_DEBUG_SETUP") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
      if (!dbg_setup) return;
      cudaStreamSynchronize(s);
      auto now = std::chrono::steady_clock::now();
      fprintf(stderr, "[isfm setup] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
      t_last = now;
    };
    upload(cam[0], cam_in, (size_t)nc * CW); cam[1].alloc((size_t)nc * CW);
    upload(pts[0], pts_in, (size_t)np * 3); pts[1].alloc((size_t)np * 3);
    upload(pp, pp_in, (size_t)nc * 2);
    camq[0].alloc((size_t)nc * CWP); camq[1].alloc((size_t)nc * CWP);
    pack_cameras(0);
    DeviceBuffer<T> obs_raw; DeviceBuffer<int32_t> ci, pi;
    upload(obs_raw, obs_in, (size_t)no * 2);
    upload(ci, cam_idx, (size_t)no); upload(pi, pt_idx, (size_t)no);
    mark("alloc + H2D");
    build_obs_index(ix, nc, np, no, ci.get(), pi.get(), s, timers);
    mark("build_obs_index (sorts, CSR)");
    obs.alloc((size_t)no * 2);
    { TimerScope ts(timers, T_INDEX_PREP);
      gather_rows_kernel<T><<<div_up(no * 2, BA_TPB), BA_TPB, 0, s>>>(no, 2, obs_raw.get(), ix.obs_perm.get(), obs.get()); }
    R.alloc((size_t)no * 2); OBS.alloc((size_t)no * REC);
    HPP.alloc((size_t)np * 6); GPT.alloc((size_t)np * 3); HPPINV.alloc((size_t)np * 6); TP.alloc((size_t)np * 3);
    DP.alloc((size_t)np * 3);
    mark("obs gather + allocs");
    build_fused_partition();
    if (fused_ok) build_obs_tensor_map();
    mark("fused tile partition");
    if (getenv("ISFM_NO_FUSED")) fused_ok = false;
    const int64_t n_part = std::max<int64_t>(std::max<int64_t>(RED_BLOCKS, nc), n_fused_cta);
    part_a.alloc(n_part); part_b.alloc(n_part); part_c.alloc(n_part);
    scalars.alloc(4); fail.alloc(1); fail.zero(s); red_stage.alloc(3 * RED_SLICES);
    if (desc.optimize_poses) {
      UnionKeys hook(this);
      build_schur_pattern(sp, ix, s, timers, comm_world(comm) > 1 ? &hook : nullptr, SpmvCfg<T, D>::WB, persistent_grid_warps<T, D>());
      mark("build_schur_pattern");
      HCC_GC.alloc((size_t)nc * (D * D + D)); HD.alloc((size_t)nc * D * D);
      E.alloc((size_t)sp.nnzu * D * D); E.zero(s);   // padding slots stay zero
      HME.alloc((size_t)nc * (D * D + 2 * D));
      MINV.alloc((size_t)nc * D * D); bvec.alloc((size_t)nc * D); DCQ.alloc((size_t)nc * BacksubCfg<T, D>::DQ);
      pcg.resize((int)nc, sp.n_off, sp.n_chunks, comm, s);
      setup_matvec_split();
      setup_exchange_ring();
      setup_coarse();
      matvec_units_total = sp.n_chunks;
      matvec_units_owned = split_matvec ? unit_hi - unit_lo : sp.n_chunks;
      mark("Schur / PCG allocs + E.zero");
    }
    ISFM_CUDA(cudaStreamSynchronize(s));
    cur = 0; have_loss = false; has_problem = true;
    coarse_age = -1; coarse_refresh = false; iters_at_factor = -1;
    tr.init(desc.tr_radius, desc.tr_max, desc.tr_up, desc.tr_down);
  }

  // Multi-rank pattern hook: all-gathers every rank's block keys; when the ranks' patterns
  // overlap heavily (sum of the local block counts >= 1.5 x the size of their union: dense
  // co-visibility) every rank builds the UNION pattern so that the mat-vec can be split
  // (setup_matvec_split); otherwise each rank keeps its own local pattern (Variant N).
  // COLLECTIVE; ISFM_SPLIT_MATVEC=0 disables, =1 forces the union.
  struct UnionKeys : PatternKeyHook {
    BASolver* self;
    explicit UnionKeys(BASolver* s_) : self(s_) {}
    int64_t extra_keys(const uint64_t* list_key, int64_t n_lists, int64_t n_cam, DeviceBuffer<uint64_t>& out,
                       cudaStream_t s) override {
      isfm_comm* comm = self->comm;
      self->union_pattern = false;
      const int world = comm_world(comm), rank = comm_rank(comm);
      const char* env = getenv("ISFM_SPLIT_MATVEC");
      if (world <= 1 || (env && atoi(env) == 0)) return 0;
      // counts of every rank (exact in fp64)
      std::vector<double> cnt((size_t)world, 0.0);
      cnt[rank] = (double)n_lists;
      DeviceBuffer<double> d_cnt; d_cnt.alloc(world);
      ISFM_CUDA(cudaMemcpyAsync(d_cnt.get(), cnt.data(), world * sizeof(double), cudaMemcpyHostToDevice, s));
      comm_allreduce_sum(comm, d_cnt.get(), world, true, s);
      ISFM_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt.get(), world * sizeof(double), cudaMemcpyDeviceToHost, s));
      ISFM_CUDA(cudaStreamSynchronize(s));
      int64_t max_cnt = 0, sum_cnt = 0;
      for (double c : cnt) { max_cnt = std::max<int64_t>(max_cnt, (int64_t)c); sum_cnt += (int64_t)c; }
      if (max_cnt == 0) return 0;
      DeviceBuffer<uint64_t> mine;
      mine.alloc(max_cnt); out.alloc((size_t)max_cnt * world);
      ISFM_CUDA(cudaMemsetAsync(mine.get(), 0xff, (size_t)max_cnt * sizeof(uint64_t), s));   // padding: ~0 = ignored
      if (n_lists > 0) ISFM_CUDA(cudaMemcpyAsync(mine.get(), list_key, (size_t)n_lists * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
      comm_allgather_bytes(comm, mine.get(), out.get(), (size_t)max_cnt * sizeof(uint64_t), s);
      const int64_t n_union = count_unique_upper_keys(out.get(), max_cnt * world, n_cam, s);
      const bool overlap = (double)sum_cnt >= 1.5 * (double)std::max<int64_t>(n_union, 1);
      if (!(env ? atoi(env) != 0 : overlap)) { out.release(); return 0; }
      self->union_pattern = true;
      matvec_share = world;   // the summed matrix will be split across the ranks
      return max_cnt * world;
    }
  };

  // Variant R of SURVEY 8(e), for the case where it pays: when every rank has the SAME block
  // pattern (BAL-like scenes: every camera pair co-observes points of every shard), each rank's
  // partial E_g is a full-size matrix and Variant N makes every rank stream all of it in every
  // PCG iteration -- the mat-vec, 70 % of a step, would not scale at all.  Instead the mat-vec
  // units are cut into `world` contiguous ranges of (almost) equal slot counts; once per trial
  // the ranks sum E range by range into the owner (comm_reduce_ranges) and every PCG iteration
  // multiplies 1 / world of the matrix per rank.  The per-iteration exchange of y is unchanged:
  // partials and deposits of the units a rank does not own stay zero.
  // COLLECTIVE: every rank takes the same decision (signatures are compared through an all-reduce).
  void setup_matvec_split() {
    split_matvec = false; unit_lo = 0; unit_hi = -1;
    const int world = comm_world(comm), rank = comm_rank(comm);
    if (world <= 1) return;
    const char* env = getenv("ISFM_SPLIT_MATVEC");   // "0": never, "1": whenever the patterns agree
    uint64_t sig[2];
    schur_pattern_signature(sp, n_cam, s, sig);
    // every rank writes its (nnzu, n_chunks, hash, hash) into its own row of a zero table; after
    // the sum every rank holds every row (all values < 2^48: exact in fp64)
    std::vector<double> table((size_t)world * 4, 0.0);
    table[(size_t)rank * 4 + 0] = (double)sp.nnzu; table[(size_t)rank * 4 + 1] = (double)sp.n_chunks;
    table[(size_t)rank * 4 + 2] = (double)sig[0]; table[(size_t)rank * 4 + 3] = (double)sig[1];
    DeviceBuffer<double> d_table;
    d_table.alloc(table.size());
    ISFM_CUDA(cudaMemcpyAsync(d_table.get(), table.data(), table.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    comm_allreduce_sum(comm, d_table.get(), table.size(), true, s);
    ISFM_CUDA(cudaMemcpyAsync(table.data(), d_table.get(), table.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    ISFM_CUDA(cudaStreamSynchronize(s));
    bool same = true;
    for (int r = 1; r < world; ++r)
      for (int k = 0; k < 4; ++k) same = same && table[(size_t)r * 4 + k] == table[k];
    const bool worth = (size_t)sp.nnzu * D * D * sizeof(T) >= ((size_t)8 << 20);   // below that the reduce costs more than it saves
    if (!same || (env ? atoi(env) == 0 : !(worth || union_pattern)) || sp.n_chunks < world) return;
    // unit boundaries: first unit whose first slot >= r * nnzu / world
    std::vector<int64_t> ub((size_t)world + 1);
    int64_t u = 0;
    for (int r = 0; r <= world; ++r) {
      const int64_t target = (int64_t)((__int128)r * sp.nnzu / world);
      while (u < sp.n_chunks && sp.h_chunk_beg[(size_t)u] < target) ++u;
      ub[r] = (r == world) ? sp.n_chunks : u;
    }
    split_off.assign(world, 0); split_cnt.assign(world, 0);
    for (int r = 0; r < world; ++r) {
      const int64_t lo = ub[r] < sp.n_chunks ? sp.h_chunk_beg[(size_t)ub[r]] : sp.nnzu;
      const int64_t hi = ub[r + 1] < sp.n_chunks ? sp.h_chunk_beg[(size_t)ub[r + 1]] : sp.nnzu;
      split_off[r] = (size_t)lo * D * D; split_cnt[r] = (size_t)(hi - lo) * D * D;
    }
    unit_lo = ub[rank]; unit_hi = ub[rank + 1];
    pcg.yup.zero(s); pcg.C.zero(s);   // never written for the units of other ranks
    E_own.alloc(std::max<size_t>(split_cnt[rank], 1));
    pcg.set_owned_slots(sp, (int64_t)(split_off[rank] / (D * D)), (int64_t)((split_off[rank] + split_cnt[rank]) / (D * D)), s);
    split_matvec = true;
  }

  // Multi-rank, local patterns (street scenes): the rows of y = E_g p this rank can be non-zero in
  // form (almost) one arc of the camera ring.  Every rank publishes the smallest circular range
  // [lo, lo + len) covering its rows; the PCG exchange pushes and sums only those rows.
  // COLLECTIVE.
  void setup_exchange_ring() {
    pcg.ring_valid = false;
    const int world = comm_world(comm), rank = comm_rank(comm);
    if (world <= 1 || world > ISFM_MAX_PEERS) return;
    int lo = 0, len = (int)n_cam;
    if (!split_matvec && !union_pattern && n_cam >= 64 && !getenv("ISFM_FULL_EXCHANGE")) {
      DeviceBuffer<uint8_t> touched;
      touched.alloc((size_t)n_cam);
      ISFM_CUDA(cudaMemsetAsync(touched.get(), 0, (size_t)n_cam, s));
      touched_rows_kernel<<<div_up(n_cam, 8), 256, 0, s>>>((int)n_cam, sp.urow_ptr.get(), sp.ucol.get(), touched.get());
      if (sp.n_lists > 0)
        touched_diag_lists_kernel<<<div_up(sp.n_lists, 256), 256, 0, s>>>(sp.n_lists, sp.list_diag.get(), sp.list_slot.get(), sp.ucol.get(), touched.get());
      std::vector<uint8_t> h((size_t)n_cam);
      ISFM_CUDA(cudaMemcpyAsync(h.data(), touched.get(), (size_t)n_cam, cudaMemcpyDeviceToHost, s));
      ISFM_CUDA(cudaStreamSynchronize(s));
      // largest circular gap of untouched rows; the ring is its complement
      const int n = (int)n_cam;
      int best_len = 0, best_end = 0, run = 0;
      for (int k = 0; k < 2 * n; ++k) {
        if (!h[(size_t)(k % n)]) { if (++run > best_len && run <= n) { best_len = run; best_end = k; } }
        else run = 0;
      }
      if (best_len >= n) { lo = 0; len = 0; }   // nothing touched at all
      else if (best_len > 0) {
        lo = (best_end + 1) % n;
        len = n - best_len;
        const int lo4 = lo / 4 * 4;
        len = (len + (lo - lo4) + 3) / 4 * 4;
        lo = lo4;
      }
      if (len >= n * 6 / 10) { lo = 0; len = n; }   // not worth the bookkeeping
    }
    int32_t mine[2] = {lo, len};
    DeviceBuffer<int32_t> d_mine, d_all;
    d_mine.alloc(2); d_all.alloc((size_t)2 * world);
    ISFM_CUDA(cudaMemcpyAsync(d_mine.get(), mine, sizeof mine, cudaMemcpyHostToDevice, s));
    comm_allgather_bytes(comm, d_mine.get(), d_all.get(), sizeof mine, s);
    std::vector<int32_t> all((size_t)2 * world);
    ISFM_CUDA(cudaMemcpyAsync(all.data(), d_all.get(), all.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    ISFM_CUDA(cudaStreamSynchronize(s));
    for (int r = 0; r < world; ++r) { pcg.row_lo[r] = all[(size_t)2 * r]; pcg.row_len[r] = all[(size_t)2 * r + 1]; }
    pcg.ring_valid = true;
    (void)rank;
  }

  // Two-level preconditioner: decided from the GLOBAL sparsity (sum of the ranks' block counts) so
  // that every rank builds the same clusters.  COLLECTIVE.
  void setup_coarse() {
    double n_off_sum = (double)sp.n_off;
    if (comm_world(comm) > 1) {
      DeviceBuffer<double> d; d.alloc(1);
      ISFM_CUDA(cudaMemcpyAsync(d.get(), &n_off_sum, sizeof(double), cudaMemcpyHostToDevice, s));
      comm_allreduce_sum(comm, d.get(), 1, true, s);
      ISFM_CUDA(cudaMemcpyAsync(&n_off_sum, d.get(), sizeof(double), cudaMemcpyDeviceToHost, s));
      ISFM_CUDA(cudaStreamSynchronize(s));
      if (union_pattern) n_off_sum /= comm_world(comm);   // every rank holds the whole pattern
    }
    coarse.setup(sp, (int64_t)n_off_sum, n_cam, s, timers);
    pcg.coarse = typename BlockPCG<T, D>::CoarseRef{};
    if (coarse.enabled) {
      pcg.coarse.enabled = 1; pcg.coarse.cs = coarse.g.cs; pcg.coarse.ncl = coarse.g.ncl; pcg.coarse.ncp = coarse.g.ncp;
      pcg.coarse.Pm = coarse.Pm.get(); pcg.coarse.Ainv = coarse.Ainv.get(); pcg.coarse.rc = coarse.rc.get();
      pcg.coarse.fail = coarse.fail.get();
    }
  }

  void pack_cameras(int which) {
    TimerScope ts(timers, T_MISC);
    pack_cameras_kernel<T, CW><<<div_up(n_cam * CWP, 256), 256, 0, s>>>((int)n_cam, cam[which].get(), pp.get(), camq[which].get());
  }

  int red_grid(int64_t n) const { return (int)std::min<int64_t>(RED_BLOCKS, div_up(n, BA_TPB)); }

  // reduce up to three partial arrays to scalars, all-reduce them, copy to the host
  void fetch_scalars(const double* p0, int n0, const double* p1, int n1, const double* p2, int n2, bool allreduce) {
    if (std::max(n0, std::max(n1, n2)) > 8192) {
      // long partial arrays: 64 slices per array first (a single CTA per array took 0.2 ms at 60 M observations)
      TimerScope ts(timers, T_REDUCE);
      reduce_stage_kernel<<<dim3(RED_SLICES, 3), 256, 0, s>>>(p0, p1, p2, n0, n1, n2, red_stage.get());
      const double* st = red_stage.get();
      reduce_scalars_kernel<<<3, 64, 0, s>>>(p0 ? st : nullptr, p1 ? st + RED_SLICES : nullptr, p2 ? st + 2 * RED_SLICES : nullptr,
                                             RED_SLICES, RED_SLICES, RED_SLICES, scalars.get());
    } else {
      TimerScope ts(timers, T_REDUCE);
      reduce_scalars_kernel<<<3, 256, 0, s>>>(p0, p1, p2, n0, n1, n2, scalars.get());
    }
    if (allreduce && comm_world(comm) > 1) {
      TimerScope ts(timers, T_COMM);
      comm_allreduce_sum(comm, scalars.get(), 3, true, s);
    }
    ISFM_CUDA(cudaMemcpyAsync(h_scalars, scalars.get(), 4 * sizeof(double), cudaMemcpyDeviceToHost, s));   // [3]: ||D_c||^2
    ISFM_CUDA(cudaStreamSynchronize(s));
  }

  void run_linearize() {
    const int g = red_grid(n_obs);
    TimerScope ts(timers, T_LINEARIZE);
    linearize_kernel<T, MODEL><<<g, BA_TPB, 0, s>>>(n_obs, camq[cur].get(), pts[cur].get(), obs.get(),
                                                    ix.cam_of.get(), ix.pt_of.get(), (T)desc.huber_delta, R.get(), OBS.get(),
                                                    part_a.get(), part_b.get());
  }

  // fused K1 + point solve of the first trial; returns the number of cost partials
  // 2-D tensor map over OBS ([n_obs][32] fp32 = 128-byte records), boxes of 8 records, 128-byte
  // swizzle: the fused K1 stores its record slab with bulk tensor copies (fused_linearize_tma_kernel).
  // cuTensorMapEncodeTiled is resolved through the runtime (no link against libcuda).
  bool build_obs_tensor_map() {
    tma_ok = false;
    if constexpr (sizeof(T) != 4 || REC != FusedTmaCfg::REC) {
      return false;
    } else {
    if (getenv("ISFM_NO_TMA") || n_obs <= 0) return false;
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) { cudaGetLastError(); return false; }
    const cuuint64_t dims[2] = {(cuuint64_t)FusedTmaCfg::REC, (cuuint64_t)n_obs};
    const cuuint64_t strides[1] = {(cuuint64_t)FusedTmaCfg::REC * 4};
    const cuuint32_t box[2] = {(cuuint32_t)FusedTmaCfg::REC, (cuuint32_t)FusedTmaCfg::BOX_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = reinterpret_cast<EncodeFn>(fn)(&obs_tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, OBS.get(), dims, strides, box, estr,
                                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    tma_big_rows = 64;   // measured: 8 / 32 / 64-record boxes 268 / 254 / 254 us on C3
    if (const char* e = getenv("ISFM_TMA_BOX")) tma_big_rows = atoi(e);
    if (tma_big_rows % 8 != 0 || tma_big_rows < 0 || tma_big_rows > 256) tma_big_rows = 0;
    obs_tmap_big = obs_tmap;
    if (tma_big_rows > 0) {
      const cuuint32_t box_big[2] = {(cuuint32_t)FusedTmaCfg::REC, (cuuint32_t)tma_big_rows};
      if (reinterpret_cast<EncodeFn>(fn)(&obs_tmap_big, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, OBS.get(), dims, strides, box_big, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { tma_big_rows = 0; obs_tmap_big = obs_tmap; }
    }
    if (cudaFuncSetAttribute(fused_linearize_tma_kernel<MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)FusedTmaCfg::SMEM) != cudaSuccess) { cudaGetLastError(); return false; }
    tma_ok = true;
    return true;
    }
  }

  int run_fused_linearize(T mu, bool want_cost) {
    TimerScope ts(timers, T_LINEARIZE);
    if (tma_ok) {
      launch_fused_tma(mu, want_cost);
      return n_fused_cta;
    }
    fused_linearize_kernel<T, MODEL><<<n_fused_cta, FUSED_TPB, FusedCfg<T, D>::SMEM, s>>>(
        fused_tiles.get(), ix.pt_off.get(), camq[cur].get(), pts[cur].get(), obs.get(), ix.cam_of.get(),
        ix.pt_of.get(), (T)desc.huber_delta, mu, R.get(), OBS.get(), HPP.get(), GPT.get(), HPPINV.get(), TP.get(),
        part_a.get(), part_b.get(), want_cost ? 1 : 0, 0);
    return n_fused_cta;
  }

  // (the kernel exists for float records only; the double build never takes this path)
  void launch_fused_tma(T mu, bool want_cost) {
    if constexpr (sizeof(T) == 4 && REC == FusedTmaCfg::REC) {
      fused_linearize_tma_kernel<MODEL><<<n_fused_cta, FUSED_TPB, FusedTmaCfg::SMEM, s>>>(
          obs_tmap, obs_tmap_big, tma_big_rows, fused_tiles.get(), ix.pt_off.get(), camq[cur].get(), pts[cur].get(), obs.get(), ix.cam_of.get(),
          ix.pt_of.get(), (float)desc.huber_delta, mu, R.get(), OBS.get(), HPP.get(), GPT.get(), HPPINV.get(), TP.get(),
          part_a.get(), part_b.get(), want_cost ? 1 : 0);
    }
  }

  // greedy packing of whole points into CTAs of <= FUSED_TPB observations (host, one pass)
  void build_fused_partition() {
    std::vector<int32_t> off((size_t)n_pt + 1), cta;
    ISFM_CUDA(cudaMemcpyAsync(off.data(), ix.pt_off.get(), off.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    ISFM_CUDA(cudaStreamSynchronize(s));
    fused_ok = true;
    cta.push_back(0);
    int32_t start = 0;
    for (int64_t p = 0; p < n_pt; ++p) {
      if (off[p + 1] - off[p] > FUSED_TPB) { fused_ok = false; break; }
      if (off[p + 1] - off[start] > FUSED_TPB) { cta.push_back((int32_t)p); start = (int32_t)p; }
    }
    cta.push_back((int32_t)n_pt);
    if (!fused_ok) return;
    n_fused_cta = (int)cta.size() - 1;
    std::vector<int4> tiles((size_t)n_fused_cta);
    for (int i = 0; i < n_fused_cta; ++i) tiles[i] = make_int4(cta[i], cta[i + 1], off[cta[i]], off[cta[i + 1]] - off[cta[i]]);
    fused_tiles.alloc(tiles.size());
    ISFM_CUDA(cudaMemcpyAsync(fused_tiles.get(), tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice, s));
    ISFM_CUDA(cudaStreamSynchronize(s));
    ISFM_CUDA(cudaFuncSetAttribute(fused_linearize_kernel<T, MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)FusedCfg<T, D>::SMEM));
    ISFM_CUDA(cudaFuncSetAttribute(backsub_tiles_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)BacksubCfg<T, D>::SMEM));
  }

  void run_cost(int which, double* robust, double* sq) {
    const int g = red_grid(n_obs);
    { TimerScope ts(timers, T_COST);
      cost_kernel<T, MODEL><<<g, BA_TPB, 0, s>>>(n_obs, camq[which].get(), pts[which].get(), obs.get(),
                                                 ix.cam_of.get(), ix.pt_of.get(), (T)desc.huber_delta, part_a.get(),
                                                 part_b.get()); }
    fetch_scalars(part_a.get(), g, part_b.get(), g, nullptr, 0, true);
    *robust = h_scalars[0];
    if (sq) *sq = h_scalars[1];
  }

  void run_point_solve(bool build, T mu) {
    TimerScope ts(timers, build ? T_POINT_BLOCKS : T_POINT_SOLVE);
    const int g = div_up(n_pt, BA_TPB);
#define ISFM_PS(B, W)                                                                                              \
  point_solve_kernel<T, D, B, W><<<g, BA_TPB, 0, s>>>(n_pt, ix.pt_off.get(), OBS.get(), R.get(), mu, HPP.get(), GPT.get(), \
                                                      HPPINV.get(), TP.get())
    if (desc.optimize_poses) { if (build) ISFM_PS(true, true); else ISFM_PS(false, true); }
    else { if (build) ISFM_PS(true, false); else ISFM_PS(false, false); }
#undef ISFM_PS
  }

  void run_camera_hessian() {
    { TimerScope ts(timers, T_CAMERA_BLOCKS);
      camera_hessian_kernel<T, D><<<(int)n_cam, CAM_TPB, 0, s>>>(ix.cam_off.get(), ix.cam_perm.get(), OBS.get(), R.get(), HCC(),
                                                                GC()); }
    if (comm_world(comm) > 1) {
      TimerScope ts(timers, T_COMM);
      comm_allreduce_sum(comm, HCC_GC.get(), (size_t)n_cam * (D * D + D), sizeof(T) == 8, s);
    }
  }

  // builds E, the preconditioner and the right-hand side for damping mu; solves for D_c
  int run_schur_and_pcg(T mu, int* pcg_status) {
    { TimerScope ts(timers, T_CAMERA_BLOCKS);
      camera_schur_kernel<T, D><<<(int)n_cam, CAM_TPB, 0, s>>>(ix.cam_off.get(), ix.cam_perm.get(), OBS.get(), HME.get(), E.get(),
                                                              sp.diag_slot.get()); }
    if (sp.n_lists > 0) {
      TimerScope ts(timers, T_SCHUR_OFFDIAG);
      const int lists_per_cta = SchurGroup<D>::PER_WARP * (SCHUR_TPB / 32);
      schur_offdiag_kernel<T, D><<<div_up(sp.n_lists, lists_per_cta), SCHUR_TPB, 0, s>>>(
          sp.n_lists, sp.list_order.get(), sp.list_off.get(), sp.pairs.get(), sp.list_slot.get(), sp.list_diag.get(), OBS.get(),
          E.get());
    }
    if (comm_world(comm) > 1) {
      TimerScope ts(timers, T_COMM);
      comm_allreduce_sum(comm, HME.get(), (size_t)n_cam * (D * D + 2 * D), sizeof(T) == 8, s);
    }
    if (split_matvec) {
      TimerScope ts(timers, T_COMM);
      comm_reduce_ranges(comm, E.get(), E_own.get(), split_off.data(), split_cnt.data(), sizeof(T) == 8, s);
    }
    { TimerScope ts(timers, T_PRECOND);
      precond_kernel<T, D><<<div_up(n_cam, 64), 64, 0, s>>>((int)n_cam, HME.get(), mu, HD.get(), MINV.get(), bvec.get(),
                                                            fail.get()); }
    // Two-level preconditioner: the coarse inverse is LAGGED -- rebuilt on the first trial of every
    // `coarse_period`-th LM step, or as soon as a solve needed 30 % more iterations than the one right
    // after the last rebuild.  A stale coarse level only slows PCG down, it never changes the system.
    // Same decisions on every rank (iteration counts are identical across ranks).
    if (coarse.enabled && (coarse_age < 0 || coarse_age >= coarse_period || coarse_refresh)) {
      coarse.factor(E.get(), HD.get(), sp, comm, s, timers);
      coarse_age = 0; coarse_refresh = false; iters_at_factor = -1;
    }
    int max_iter = desc.pcg_max_iter > 0 ? desc.pcg_max_iter : (int)std::min<int64_t>(10 * n_cam * D, 5000);
    // split mat-vec: the kernel indexes blocks by their global slot, E_own starts at this rank's first slot
    const T* Emat = split_matvec ? E_own.get() - split_off[comm_rank(comm)] : E.get();
    int status = 0;
    int iters = pcg.solve(sp, Emat, HD.get(), MINV.get(), bvec.get(), desc.pcg_tol, max_iter, comm, s, timers, &status, unit_lo, unit_hi);
    if (coarse.enabled) {
      if (status == 2 || status == 0) {
        // Safety net: breakdown (r.z <= 0: the coarse inverse lost positive definiteness) or no
        // convergence with the two-level preconditioner -- solve this system again with block-Jacobi
        // alone (the kernel ignores the coarse level while its `fail` flag is set) and rebuild the
        // coarse inverse before the next solve.  Same decision on every rank.
        const int one = 1;
        ISFM_CUDA(cudaMemcpyAsync(coarse.fail.get(), &one, sizeof(int), cudaMemcpyHostToDevice, s));
        ISFM_CUDA(cudaStreamSynchronize(s));
        coarse_fallbacks++;
        iters += pcg.solve(sp, Emat, HD.get(), MINV.get(), bvec.get(), desc.pcg_tol, max_iter, comm, s, timers, &status, unit_lo, unit_hi);
        coarse_refresh = true;
      } else if (iters_at_factor < 0) iters_at_factor = iters;
      else if (iters > iters_at_factor + iters_at_factor * 3 / 10 + 4) coarse_refresh = true;
    }
    if (pcg_status) *pcg_status = status;
    return iters;
  }

  void step(double* loss_out, isfm_step_stats* st) override {
    ISFM_REQUIRE(has_problem, ISFM_ESTATE, "isfm_ba_step before isfm_ba_set_problem");
    // R, J at the current parameters (+ the initial loss on the first call: `self.loss`).
    // The first trial's damping is known here, so the fused kernel also does the point solve.
    const double mu_cap = sizeof(T) == 4 ? 1e24 : 1e100;
    const double mu_first = std::min(1.0 + std::max(tr.damping, min_damping), mu_cap);
    int lin_parts;
    if (fused_ok) lin_parts = run_fused_linearize((T)mu_first, !have_loss);
    else { run_linearize(); lin_parts = red_grid(n_obs); }
    if (!have_loss) {
      const int g = lin_parts;
      fetch_scalars(part_a.get(), g, part_b.get(), g, nullptr, 0, true);
      loss = h_scalars[0];
      have_loss = true;
      ISFM_REQUIRE(std::isfinite(loss), ISFM_ENONFINITE, "initial cost is not finite");
    }
    const double last = loss;
    isfm_step_stats stats;
    memset(&stats, 0, sizeof stats);
    stats.loss_before = last;
    if (desc.optimize_poses && coarse.enabled) {
      if (coarse_age >= 0) ++coarse_age;
      if (coarse_age < 0 || coarse_age >= coarse_period || coarse_refresh)
        coarse.update_modes(cam[cur].get(), CW, s, timers);   // cluster modes at the current poses (only when the inverse is rebuilt)
    }
    double mu = 1.0;
    bool built = false;
    int rejects = 0;
    const int trial = cur ^ 1;
    while (last <= loss) {
      // cumulative across rejected trials (pypose LM); capped so that damped blocks stay finite in T
      mu = std::min(mu * (1.0 + std::max(tr.damping, min_damping)), mu_cap);
      if (!(fused_ok && !built)) run_point_solve(!built, (T)mu);   // first trial: done by the fused kernel (mu == mu_first)
      built = true;
      int mterm_parts;
      if (desc.optimize_poses) {
        int pcg_status = 0;
        stats.pcg_iters += run_schur_and_pcg((T)mu, &pcg_status);
        stats.pcg_status = pcg_status;
        if (debug) {
          int h_fail = 0;
          ISFM_CUDA(cudaMemcpyAsync(&h_fail, fail.get(), sizeof(int), cudaMemcpyDeviceToHost, s));
          ISFM_CUDA(cudaStreamSynchronize(s));
          fprintf(stderr, "[isfm] pcg status %d (1 converged, 0 max_iter, 2 breakdown) precond_fail_cam %d bb %.6e rho0 %.6e\n",
                  pcg_status, h_fail - 1, pcg.h_state->bb, pcg.h_state->rho);
          if (h_fail) {
            const int c = h_fail - 1;
            std::vector<T> hh((size_t)D * D), ee((size_t)D * D + D);
            ISFM_CUDA(cudaMemcpy(hh.data(), HD.get() + (size_t)c * D * D, hh.size() * sizeof(T), cudaMemcpyDeviceToHost));
            ISFM_CUDA(cudaMemcpy(ee.data(), HME.get() + (size_t)c * (D * D + 2 * D), ee.size() * sizeof(T), cudaMemcpyDeviceToHost));
            for (int r = 0; r < D; ++r) fprintf(stderr, "[isfm]   cam %d diag %d: S_ii %.9e  Hcc - E_ii %.9e\n", c, r, (double)hh[r * D + r],
                                               (double)ee[r * D + r]);
            ISFM_CUDA(cudaMemsetAsync(fail.get(), 0, sizeof(int), s));
          }
        }
        if (fused_ok && !no_tile_backsub) {
          constexpr int DQ = BacksubCfg<T, D>::DQ;
          { TimerScope ts(timers, T_MISC);
            pad_rows_kernel<T><<<div_up(n_cam * DQ, 256), 256, 0, s>>>((int)n_cam, D, DQ, pcg.x.get(), DCQ.get()); }
          TimerScope ts(timers, T_BACKSUB);
          mterm_parts = n_fused_cta;
          backsub_tiles_kernel<T, D><<<n_fused_cta, FUSED_TPB, BacksubCfg<T, D>::SMEM, s>>>(
              fused_tiles.get(), ix.pt_off.get(), ix.cam_of.get(), OBS.get(), R.get(), GPT.get(), HPP.get(), HPPINV.get(),
              DCQ.get(), pts[cur].get(), pts[trial].get(), DP.get(), part_c.get());
        } else {
          TimerScope ts(timers, T_BACKSUB);
          mterm_parts = red_grid(n_pt);
          backsub_kernel<T, D><<<mterm_parts, BA_TPB, 0, s>>>(n_pt, ix.pt_off.get(), ix.cam_of.get(), OBS.get(), R.get(),
                                                             GPT.get(), HPP.get(), HPPINV.get(), pcg.x.get(), pts[cur].get(),
                                                             pts[trial].get(), DP.get(), part_c.get()); }
        { TimerScope ts(timers, T_UPDATE);
          camera_update_kernel<T, NI><<<div_up(n_cam, 128), 128, 0, s>>>((int)n_cam, cam[cur].get(), pp.get(), pcg.x.get(),
                                                                        cam[trial].get(), camq[trial].get(), pcg.part_a.get()); }
        // ||D_c||^2 of this trial (identical on every rank: not all-reduced), fetched with the trial scalars
        { TimerScope ts(timers, T_REDUCE);
          reduce_scalars_kernel<<<1, 256, 0, s>>>(pcg.part_a.get(), nullptr, nullptr, div_up(n_cam, 128), 0, 0, scalars.get() + 3); }
      } else {
        TimerScope ts(timers, T_BACKSUB);
        mterm_parts = red_grid(n_pt);
        point_only_step_kernel<T, D><<<mterm_parts, BA_TPB, 0, s>>>(n_pt, ix.pt_off.get(), OBS.get(), R.get(), TP.get(),
                                                                 pts[cur].get(), pts[trial].get(), DP.get(), part_c.get());
        ISFM_CUDA(cudaMemcpyAsync(cam[trial].get(), cam[cur].get(), (size_t)n_cam * CW * sizeof(T), cudaMemcpyDeviceToDevice, s));
        ISFM_CUDA(cudaMemcpyAsync(camq[trial].get(), camq[cur].get(), (size_t)n_cam * CWP * sizeof(T), cudaMemcpyDeviceToDevice, s));
      }
      const int g = red_grid(n_obs);
      { TimerScope ts(timers, T_COST);
        cost_kernel<T, MODEL><<<g, BA_TPB, 0, s>>>(n_obs, camq[trial].get(), pts[trial].get(), obs.get(),
                                                   ix.cam_of.get(), ix.pt_of.get(), (T)desc.huber_delta, part_a.get(),
                                                   part_b.get()); }
      // scalars: [0] new robust cost, [1] plain squared cost, [2] model term (summed over ranks)
      fetch_scalars(part_a.get(), g, part_b.get(), g, part_c.get(), mterm_parts, true);
      const double new_loss = h_scalars[0];
      const double mterm = h_scalars[2];
      stats.trials++;
      stats.model_term = -mterm;
      stats.quality = tr.update(last, new_loss, mterm);
      if (debug)
        fprintf(stderr, "[isfm] trial %d mu %.3e last %.17g new %.17g mterm %.6e quality %.4f pcg_iters %d damping->%.3e\n",
                stats.trials, mu, last, new_loss, mterm, stats.quality, stats.pcg_iters, tr.damping);
      const bool worse = !(new_loss <= last);  // NaN counts as worse (the reference would keep it)
      // never accept a non-finite state (the reference would; its write-back would then poison the scene)
      if (!std::isfinite(new_loss) && rejects >= desc.reject)
        throw IsfmError(ISFM_ENONFINITE, "trial cost is not finite after the last allowed rejection");
      if (worse && rejects < desc.reject) {
        rejects++;
        loss = last;
        stats.accepted = 0;
      } else {
        cur = trial;
        loss = new_loss;
        stats.accepted = 1;
        break;
      }
    }
    mu_last = mu;
    stats.rejects = rejects;
    stats.loss = loss;
    stats.damping = tr.damping;
    if (desc.optimize_poses) stats.step_norm_cam = std::sqrt(h_scalars[3]);   // last trial
    ISFM_CUDA(cudaGetLastError());
    if (loss_out) *loss_out = loss;
    if (st) *st = stats;
  }

  void get_pcg_phases(double* ms_out, int64_t* solves_out, int32_t* two_level_out) override {
    if (ms_out) for (int i = 0; i < 8; ++i) ms_out[i] = pcg.phase_ms[i];
    if (solves_out) *solves_out = pcg.persist_solves;
    if (two_level_out) *two_level_out = coarse.enabled ? 1 : 0;
  }

  void get_params(void* cam_out, void* pts_out) override {
    ISFM_REQUIRE(has_problem, ISFM_ESTATE, "no problem set");
    if (cam_out) ISFM_CUDA(cudaMemcpyAsync(cam_out, cam[cur].get(), (size_t)n_cam * CW * sizeof(T), cudaMemcpyDefault, s));
    if (pts_out) ISFM_CUDA(cudaMemcpyAsync(pts_out, pts[cur].get(), (size_t)n_pt * 3 * sizeof(T), cudaMemcpyDefault, s));
    ISFM_CUDA(cudaStreamSynchronize(s));
  }

  void set_params(const void* cam_in, const void* pts_in) override {
    ISFM_REQUIRE(has_problem, ISFM_ESTATE, "no problem set");
    if (cam_in) {
      ISFM_CUDA(cudaMemcpyAsync(cam[cur].get(), cam_in, (size_t)n_cam * CW * sizeof(T), cudaMemcpyDefault, s));
      pack_cameras(cur);
    }
    if (pts_in) ISFM_CUDA(cudaMemcpyAsync(pts[cur].get(), pts_in, (size_t)n_pt * 3 * sizeof(T), cudaMemcpyDefault, s));
    ISFM_CUDA(cudaStreamSynchronize(s));
    have_loss = false;
  }

  void cost(double* robust, double* sq) override {
    ISFM_REQUIRE(has_problem, ISFM_ESTATE, "no problem set");
    double r, q;
    run_cost(cur, &r, &q);
    if (robust) *robust = r;
    if (sq) *sq = q;
  }

  template <typename U>
  static void d2h(std::vector<U>& h, const U* d, size_t n, cudaStream_t s) {
    h.resize(n);
    ISFM_CUDA(cudaMemcpyAsync(h.data(), d, n * sizeof(U), cudaMemcpyDeviceToHost, s));
    ISFM_CUDA(cudaStreamSynchronize(s));
  }

  void get_structure(int32_t* obs_perm, int64_t* pt_off, int32_t* cam_perm, int64_t* cam_off) override {
    ISFM_REQUIRE(has_problem, ISFM_ESTATE, "no problem set");
    std::vector<int32_t> h;
    if (obs_perm) { d2h(h, ix.obs_perm.get(), (size_t)n_obs, s); std::copy(h.begin(), h.end(), obs_perm); }
    if (cam_perm) { d2h(h, ix.cam_perm.get(), (size_t)n_obs, s); std::copy(h.begin(), h.end(), cam_perm); }
    if (pt_off) { d2h(h, ix.pt_off.get(), (size_t)n_pt + 1, s); std::copy(h.begin(), h.end(), pt_off); }
    if (cam_off) { d2h(h, ix.cam_off.get(), (size_t)n_cam + 1, s); std::copy(h.begin(), h.end(), cam_off); }
  }

  // full (both triangles) BSR pattern, rebuilt on the host from the stored upper triangle
  void get_schur_pattern(int64_t* nnzb, int64_t* n_pairs, int64_t* row_ptr, int32_t* col_idx) override {
    ISFM_REQUIRE(has_problem && desc.optimize_poses, ISFM_ESTATE, "no reduced camera system (optimize_poses = 0?)");
    if (nnzb) *nnzb = sp.n_blocks + sp.n_off;
    if (n_pairs) *n_pairs = sp.n_pairs;
    if (!row_ptr && !col_idx) return;
    std::vector<int32_t> up, uc;
    d2h(up, sp.urow_ptr.get(), (size_t)n_cam + 1, s); d2h(uc, sp.ucol.get(), (size_t)sp.nnzu, s);
    std::vector<std::vector<int32_t>> rows((size_t)n_cam);
    for (int64_t i = 0; i < n_cam; ++i)
      for (int e = up[i]; e < up[i + 1]; ++e) {
        rows[i].push_back(uc[e]);
        if (uc[e] != i) rows[uc[e]].push_back((int32_t)i);
      }
    int64_t pos = 0;
    for (int64_t i = 0; i < n_cam; ++i) {
      std::sort(rows[i].begin(), rows[i].end());
      rows[i].erase(std::unique(rows[i].begin(), rows[i].end()), rows[i].end());   // row-padding slots repeat (i, i)
      if (row_ptr) row_ptr[i] = pos;
      for (int32_t c : rows[i]) { if (col_idx) col_idx[pos] = c; ++pos; }
    }
    if (row_ptr) row_ptr[n_cam] = pos;
  }

  // copies columns [off, off + W) of a [n_obs, stride] sorted-order buffer back in the caller's
  // observation order
  void unsort_rows(const T* dev, int stride, int off, int W, void* dst) {
    std::vector<T> h; std::vector<int32_t> perm;
    d2h(h, dev, (size_t)n_obs * stride, s);
    d2h(perm, ix.obs_perm.get(), (size_t)n_obs, s);
    T* out = static_cast<T*>(dst);
    for (int64_t a = 0; a < n_obs; ++a)
      for (int k = 0; k < W; ++k) out[(size_t)perm[a] * W + k] = h[(size_t)a * stride + off + k];
  }

  void debug_get(int what, void* dst) override {
    ISFM_REQUIRE(has_problem && dst, ISFM_ESTATE, "no problem set");
    if (what == ISFM_BA_STEP_CAM || what == ISFM_BA_STEP_POINT) {
      if (what == ISFM_BA_STEP_CAM) {
        ISFM_REQUIRE(desc.optimize_poses, ISFM_ESTATE, "no camera step in points-only mode");
        ISFM_CUDA(cudaMemcpyAsync(dst, pcg.x.get(), (size_t)n_cam * D * sizeof(T), cudaMemcpyDeviceToHost, s));
      } else {
        ISFM_CUDA(cudaMemcpyAsync(dst, DP.get(), (size_t)n_pt * 3 * sizeof(T), cudaMemcpyDeviceToHost, s));
      }
      ISFM_CUDA(cudaStreamSynchronize(s));
      return;
    }
    if (what == ISFM_BA_RESIDUALS) {
      DeviceBuffer<T> tmp; tmp.alloc((size_t)n_obs * 2);
      residual_kernel<T, MODEL><<<div_up(n_obs, BA_TPB), BA_TPB, 0, s>>>(n_obs, camq[cur].get(), pts[cur].get(),
                                                                       obs.get(), ix.cam_of.get(), ix.pt_of.get(), tmp.get());
      unsort_rows(tmp.get(), 2, 0, 2, dst);
      return;
    }
    const T mu = (T)(1.0 + tr.damping);
    run_linearize();
    run_point_solve(true, mu);
    switch (what) {
      case ISFM_BA_JAC_CAM: unsort_rows(OBS.get(), REC, 0, 2 * D, dst); return;
      case ISFM_BA_JAC_POINT: unsort_rows(OBS.get(), REC, ObsRec<D>::JP, 6, dst); return;
      case ISFM_BA_WEIGHTED_RES: unsort_rows(R.get(), 2, 0, 2, dst); return;
      case ISFM_BA_HPP: ISFM_CUDA(cudaMemcpyAsync(dst, HPP.get(), (size_t)n_pt * 6 * sizeof(T), cudaMemcpyDeviceToHost, s)); break;
      case ISFM_BA_GP: ISFM_CUDA(cudaMemcpyAsync(dst, GPT.get(), (size_t)n_pt * 3 * sizeof(T), cudaMemcpyDeviceToHost, s)); break;
      default: {
        ISFM_REQUIRE(desc.optimize_poses, ISFM_ESTATE, "camera blocks need optimize_poses = 1");
        run_camera_hessian();
        if (what == ISFM_BA_HCC) { ISFM_CUDA(cudaMemcpyAsync(dst, HCC(), (size_t)n_cam * D * D * sizeof(T), cudaMemcpyDeviceToHost, s)); break; }
        if (what == ISFM_BA_GC) { ISFM_CUDA(cudaMemcpyAsync(dst, GC(), (size_t)n_cam * D * sizeof(T), cudaMemcpyDeviceToHost, s)); break; }
        ISFM_REQUIRE(what == ISFM_BA_SCHUR_DENSE || what == ISFM_BA_SCHUR_RHS, ISFM_EINVAL, "unknown debug buffer");
        int status = 0;
        run_schur_and_pcg(mu, &status);
        if (what == ISFM_BA_SCHUR_RHS) { ISFM_CUDA(cudaMemcpyAsync(dst, bvec.get(), (size_t)n_cam * D * sizeof(T), cudaMemcpyDeviceToHost, s)); break; }
        std::vector<T> hE, hHD; std::vector<int32_t> rp, ci;
        d2h(hE, E.get(), (size_t)sp.nnzu * D * D, s); d2h(hHD, HD.get(), (size_t)n_cam * D * D, s);
        d2h(rp, sp.urow_ptr.get(), (size_t)n_cam + 1, s); d2h(ci, sp.ucol.get(), (size_t)sp.nnzu, s);
        const size_t n = (size_t)n_cam * D;
        T* out = static_cast<T*>(dst);
        std::fill(out, out + n * n, T(0));
        for (int64_t i = 0; i < n_cam; ++i) {
          for (int r = 0; r < D; ++r)
            for (int c = 0; c < D; ++c) out[(i * D + r) * n + i * D + c] = hHD[(size_t)i * D * D + r * D + c];
          for (int e = rp[i]; e < rp[i + 1]; ++e)
            for (int r = 0; r < D; ++r)
              for (int c = 0; c < D; ++c) {
                const T v = hE[(size_t)e * D * D + r * D + c];
                out[(i * D + r) * n + (size_t)ci[e] * D + c] -= v;
                if (ci[e] != i) out[((size_t)ci[e] * D + c) * n + i * D + r] -= v;   // mirrored block
              }
        }
        return;
      }
    }
    ISFM_CUDA(cudaStreamSynchronize(s));
  }
};

BASolverBase* make_ba_solver_f32(const isfm_ba_desc& d);
BASolverBase* make_ba_solver_f64(const isfm_ba_desc& d);

template <typename T>
BASolverBase* make_ba_solver_t(const isfm_ba_desc& d) {
  switch (d.model_id) {
    case 0: return new BASolver<T, 0>(d);
    case 1: return new BASolver<T, 1>(d);
    case 2: return new BASolver<T, 2>(d);
    case 3: return new BASolver<T, 3>(d);
    case 4: return new BASolver<T, 4>(d);
    case 5: return new BASolver<T, 5>(d);
    case 6: return new BASolver<T, 6>(d);
    case 8: return new BASolver<T, 8>(d);
    case 9: return new BASolver<T, 9>(d);
    default: throw IsfmError(ISFM_EUNSUPPORTED_MODEL, "Unsupported camera model");
  }
}

}  // namespace isfm
