#!/usr/bin/env python
"""Benchmark of the BA / GP Levenberg-Marquardt hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3|C1|C2|C5|C4]

A "step" is one LM iteration (`isfm_ba_step` == `optimizer.step(input)`,
bundle_adjustment.py:132) on a synthetic BAL-shaped problem.  Every config is ONE seeded global
instance (instantsfm_b200/synthetic.py); with N > 1 ranks (torchrun, one rank per GPU) each rank
holds a contiguous range of its points, so N ranks solve the same problem as one -- STRONG
scaling, same costs.

N = 1 (default): C3 (Venice-shaped, 1 778 cameras / 993 k points / 5.0 M observations).
N > 1: the same C3 instance split over the ranks (`value`, comparable with the N = 1 line), plus
       -- object `c5` -- the north-star multi-GPU problem C5 (20 k cameras / 10 M points / 60 M
       observations, street geometry) at N ranks with its cost compared against the committed
       1-GPU run (profiles/r2_c5_n1.json).
--config C4: global positioning (2.5 k cameras / 500 k tracks / 3.0 M observations).
Prints ONE JSON line (rank 0).

value        observations/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
e2e          same metric through the C ABI from pinned HOST buffers: create + set_problem
             (H2D of every tensor + integer prep) + K steps + get_params (D2H), cold and warm
roofline     dominant kernel: algorithmic bytes / CUDA-event time, and the
             same on the DRAM bytes ncu measured (profiles/)
cpu_baseline the oracle (restated reference algorithm, "port") on the host cores, bounded sample;
             `parity` = the GPU (fp32) on the SAME sample, per-step cost difference
c1           BASELINE.json config 1 like for like: GPU and CPU port on the full C1 problem
reference_gpu the reference's torch loop restated (oracle/torch_ba.py, fp64) on THIS GPU, C3
--impl reference  times the CPU port alone (bae / pypose are not installable)
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Libraries (NCCL's version banner, torchrun) write to fd 1; the contract is ONE JSON line on
# stdout, so everything else is diverted to stderr and the line goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


METRIC = "ba_lm_observations_per_sec"
UNIT = "obs/s"
CPU_SAMPLE_SCALE = 0.02   # cpu_baseline: config with points/observations scaled by this factor
# C5 window, fixed so that every rank count repeats the run of profiles/r2_c5_n1.json: LM steps 4-6.
# The first six LM steps take the cost from 1.18e8 to 3.887e7 and are the same at every rank count (PCG
# iterations 56 / 64 / 68 / 72 / ...).  From step 7 on the cost moves by < 1e-3 per step and the trust
# region has grown to the damping floor: accept / reject decisions then depend on the rounding of the
# summation order, and runs of the SAME configuration differ by rejected trials and 13 - 238 PCG
# iterations per step (profiles/README.md) -- timing those steps compares trajectories, not machines.
C5_STEPS, C5_WARMUP = 3, 3
# dram__bytes_read.sum + dram__bytes_write.sum per launch on config C3 (1 GPU, full size), from the
# `ncu --set full` captures summarised under profiles/ (ncu cannot run inside the bench).  pcg_solve: per PCG iteration.
NCU_TRAFFIC_C3 = {"pcg_solve": 549.4e6, "schur_offdiag": 2882.6e6, "linearize": 797.6e6, "camera_blocks": 665.8e6,
                  "backsub": 802.0e6}
NCU_TRAFFIC_SOURCE = ("profiles/r2_ncu_full_summary.txt (dram__bytes_read + write per launch; pcg_solve: the captured solve of 33 "
                      "PCG iterations moved 18.13 GB = 549 MB per iteration)")
PHASES = ["matvec", "combine", "exchange", "update", "coarse", "direction"]


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.002)   # the timed region is ~50 ms: sample every ~2 ms

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def algorithmic_bytes(name, n_obs, n_pt, n_cam, d, n_pairs, n_lists, nnzb, fsize=4):
    """Algorithmic HBM bytes of ONE launch of a kernel family; for
    `pcg_solve` of one PCG iteration."""
    dd = d * d
    n_up = (nnzb + n_cam) // 2           # stored upper blocks incl. the diagonal
    table = {
        # fused K1 + point solve: obs + indices read; R and the full record (Jc | Jp | V | rho, padded) written;
        # per point X read, Hpp | g_p | Hpp^-1 | t_p written; camera table read
        "linearize": n_obs * (8 + 8 + fsize * (2 + (2 * d + 14 + 3) // 4 * 4)) + n_pt * (8 + fsize * (3 + 18)) + n_cam * (9 + d) * fsize,
        "point_blocks": n_obs * fsize * (6 + 2 + 6) + n_pt * (4 + fsize * (6 + 3 + 6 + 3)),
        "point_solve": n_obs * fsize * (6 + 6) + n_pt * (4 + fsize * (6 + 3 + 6 + 3)),
        # one sweep per trial: camera-major index + Jc | Jp | V | rho of every record; Hcc - E_ii, diag Hcc, g_c - e written
        "camera_blocks": n_obs * (4 + fsize * (2 * d + 14)) + n_cam * fsize * (2 * dd + 2 * d),
        "schur_offdiag": n_pairs * (8 + fsize * (4 * d + 12)) + n_lists * (16 + dd * fsize),
        # symmetric storage: the upper blocks + their column / deposit indices are streamed once per
        # iteration; the vectors (x r z p q, Hd, Minv) are read / written once.  The transposed
        # products (deposits) stay in L2 and are NOT charged.
        "pcg_solve": n_up * (dd * fsize + 8) + n_cam * fsize * (2 * dd + 12 * d),
        "backsub": n_obs * (4 + fsize * (2 * d + 6 + 2)) + n_pt * fsize * (3 + 6 + 3 + 3 + 3),
        "cost": n_obs * (8 + 8) + n_pt * 3 * fsize + n_cam * (9 + d) * fsize,
    }
    return float(table.get(name, 0.0))


def _sync_barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(x, world):
    import torch
    t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _pin(a, dtype):
    import torch
    host = [np.ascontiguousarray(x, dtype=dtype) for x in (a.camera_params, a.camera_pps, a.points_3d, a.points_2d)]
    host += [np.ascontiguousarray(a.camera_indices, np.int32), np.ascontiguousarray(a.point_indices, np.int32)]
    pinned = [torch.from_numpy(h).pin_memory() for h in host]
    return host, pinned, [p.numpy() for p in pinned]


def run_resident(a, comm, world, local_rank, steps, warmup, n_total):
    """Device-resident throughput with the per-kernel CUDA-event timers ON in the timed pass (a
    handful of kernels per trial -- the PCG solve is one launch -- so the events cost < 1 %)."""
    import torch
    from instantsfm_b200.engine import BAEngine
    _, pinned, pinned_np = _pin(a, np.float32)
    eng = BAEngine(a.model_id, dtype=np.float32, comm=comm)
    eng.set_problem(*pinned_np)
    pat = eng.schur_pattern()
    mv_owned, mv_total = eng.matvec_units()
    rob_initial = eng.cost()[0]   # same instance at every rank count <=> same initial cost
    wl = [eng.step()[0] for _ in range(warmup)]
    ph0, solves0, two_level = eng.pcg_phases()
    sampler = ClockSampler(local_rank)
    eng.reset_timers(True)
    _sync_barrier(world)
    sampler.start()
    launches0 = eng.lib.isfm_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    losses, stats = [], []
    for _ in range(steps):
        loss, st = eng.step()
        losses.append(loss); stats.append(st)
    ev1.record()
    _sync_barrier(world)
    clocks = sampler.stop()
    launches = eng.lib.isfm_launch_count() - launches0
    ms_total = _max_over_ranks(ev0.elapsed_time(ev1), world)
    timers = eng.timers()
    eng.reset_timers(False)
    ph1, solves1, _ = eng.pcg_phases()
    rob, sq = eng.cost()
    d = a.camera_params.shape[1] - 1
    iters = int(sum(s["pcg_iters"] for s in stats))
    trials = int(sum(s["trials"] for s in stats))
    peak, peak_src = _peaks()
    kernels = {}
    for name, t in timers.items():
        if name in ("index_prep",) or t["launches"] == 0:
            continue
        fam = "pcg_solve" if name == "pcg_spmv" else name
        per_launch_ms = t["ms"] / max(t["launches"], 1)
        ab = algorithmic_bytes(fam, a.n_obs, a.n_pt, a.n_cam, d, pat["n_pairs"], (pat["nnzb"] - a.n_cam) // 2, pat["nnzb"])
        units = 1.0
        if fam == "pcg_solve":
            ab *= mv_owned / max(mv_total, 1)          # split mat-vec: this rank streams only its own unit range
            units = iters / max(t["launches"], 1)      # one launch = one whole solve
        dram = NCU_TRAFFIC_C3.get(fam) if (a.n_obs_total == 5_000_000 and world == 1) else None
        sec = per_launch_ms * 1e-3
        kernels[fam] = {"ms_per_step": t["ms"] / steps, "launches_per_step": t["launches"] / steps, "us_per_launch": per_launch_ms * 1e3,
                        "achieved_gbs": (ab * units / sec / 1e9) if ab and sec > 0 else None,
                        "frac_algorithmic": (ab * units / sec / 1e9 / peak) if ab and sec > 0 else None,
                        "frac_dram": (dram * units / sec / 1e9 / peak) if dram and sec > 0 else None}
    pcg_ms = kernels.get("pcg_solve", {}).get("ms_per_step", 0.0) * steps
    phases = {PHASES[i]: (ph1[i] - ph0[i]) * 1e3 / max(iters, 1) for i in range(len(PHASES))}   # us per PCG iteration
    work = {"pcg_iters_per_step": iters / steps, "trials_per_step": trials / steps,
            "us_per_pcg_iter": pcg_ms * 1e3 / max(iters, 1),
            "ms_per_trial_excl_pcg": (ms_total - pcg_ms) / max(trials, 1),
            "pcg_phase_us_per_iter": phases,
            "comm_ms_per_step": timers.get("comm", {"ms": 0.0})["ms"] / steps,
            "two_level_preconditioner": bool(two_level)}
    top = max((k for k in kernels if kernels[k]["achieved_gbs"] is not None), key=lambda k: kernels[k]["ms_per_step"])
    units_top = iters / max(timers["pcg_spmv"]["launches"], 1) if top == "pcg_solve" else 1.0
    roofline = {"kernel": top, "bound": "hbm", "achieved": kernels[top]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[top]["frac_algorithmic"], "frac_dram": kernels[top]["frac_dram"],
                "traffic": (NCU_TRAFFIC_C3.get(top) * units_top) if kernels[top]["frac_dram"] is not None else None,
                "traffic_source": NCU_TRAFFIC_SOURCE, "peak_source": peak_src,
                "note": "pcg_solve = one persistent kernel per PCG solve: bytes = per-iteration bytes x iterations of the launch",
                "share_of_step": kernels[top]["ms_per_step"] / sum(k["ms_per_step"] for k in kernels.values())}
    if top == "pcg_solve" and phases.get("matvec", 0.0) > 0.0:
        # the kernel's time includes the grid barriers and the small phases of every iteration; the
        # mat-vec phase is the part that streams the matrix (timed by CTA 0 with %globaltimer)
        try:
            ab_it = algorithmic_bytes("pcg_solve", a.n_obs, a.n_pt, a.n_cam, d, pat["n_pairs"], (pat["nnzb"] - a.n_cam) // 2, pat["nnzb"]) * mv_owned / max(mv_total, 1)
            gbs = ab_it / (phases["matvec"] * 1e-6) / 1e9
            roofline["matvec_phase"] = {"us_per_iteration": phases["matvec"], "achieved": gbs, "frac": gbs / peak,
                                        "note": "same algorithmic bytes per iteration over the mat-vec phase alone"}
        except Exception as e:   # informational only
            roofline["matvec_phase"] = {"error": repr(e)[:200]}
    out = {"ms_total": ms_total, "value": n_total * steps / (ms_total * 1e-3), "losses": [float(x) for x in wl + losses],
           "pcg_iters": [int(s["pcg_iters"]) for s in stats], "rejects": int(sum(s["rejects"] for s in stats)),
           "final_robust_cost": rob, "initial_robust_cost": rob_initial, "final_rmse_px": float(np.sqrt(sq / n_total)), "kernels": kernels, "work": work,
           "roofline": roofline, "clocks": clocks, "launches": int(launches), "pattern": pat, "matvec": (mv_owned, mv_total)}
    eng.close()
    return out


def run_e2e(a, comm, world, steps, runs=3):
    """Complete solves through the C ABI from pinned host buffers: create, set_problem (H2D of every
    tensor + sorts + Schur pattern), K steps, get_params (D2H), destroy.  The first run of a
    process is the COLD one (module load, memory pool growth); both are reported."""
    import torch
    from instantsfm_b200.engine import BAEngine
    host, pinned, pinned_np = _pin(a, np.float32)
    out = []
    for _ in range(runs):
        _sync_barrier(world)
        t0 = time.perf_counter()
        eng = BAEngine(a.model_id, dtype=np.float32, comm=comm)
        eng.set_problem(*pinned_np)
        t_setup = time.perf_counter() - t0
        for _ in range(steps):
            eng.step()
        cam_out, pts_out = eng.get_params()
        torch.cuda.synchronize()
        dt = _max_over_ranks(time.perf_counter() - t0, world)
        out.append((dt, t_setup))
        eng.close()
    h2d = sum(h.nbytes for h in host)
    d2h = cam_out.nbytes + pts_out.nbytes + 8 * steps
    return out, h2d, d2h


def run_cpu_port(config, n_steps, threads, scale, compare_gpu=False):
    """The reference's loop restated (oracle/torch_ba.py: full system, Jacobi PCG tol 1e-5, torch sparse
    CSR products, fp64) on the host cores; optionally the GPU (fp32, product settings) on the SAME
    problem for the same steps -> parity of the costs.  (The scipy oracle of the tests, oracle/lm.py,
    runs the same algorithm ~3x slower; the faster port is the baseline.)"""
    import torch
    from instantsfm_b200.synthetic import make_config
    from oracle.torch_ba import TorchRefBA
    torch.set_num_threads(threads)
    a = make_config(config, scale=scale)
    pb = TorchRefBA(a.model_id, a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices, device="cpu")
    costs = [pb.step()]  # warm-up (first call pays import / allocator costs)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        costs.append(pb.step())
    dt = time.perf_counter() - t0
    r = pb.residuals().numpy()
    rmse_cpu = float(np.sqrt((r * r).sum(-1).mean()))
    cam_cpu, pts_cpu = pb.cam.numpy(), pb.pts.numpy()
    sample = (f"{config}" + (f" scaled x{scale}" if scale != 1.0 else " (full)") + f": {a.n_cam} cameras / {a.n_pt} points / "
              f"{a.n_obs} observations, {n_steps} LM steps of the fp64 torch port of the reference loop (full normal equations, "
              "Jacobi PCG 1e-5, sparse CSR J / J^T)")
    res = {"value": a.n_obs * n_steps / dt, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
           "ms_per_lm_step": dt / n_steps * 1e3, "costs": [float(c) for c in costs],
           "pcg_iters_per_step": pb.pcg_iters / (n_steps + 1), "final_rmse_px": rmse_cpu}
    if compare_gpu:
        from instantsfm_b200.engine import BAEngine
        eng = BAEngine(a.model_id, dtype=np.float32)
        eng.set_problem(a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
        eng.step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = [eng.step()[0] for _ in range(n_steps)]
        torch.cuda.synchronize()
        gdt = time.perf_counter() - t0
        _, sq = eng.cost()
        cam, pts = eng.get_params()
        rel = [abs(x - y) / y for x, y in zip(g, costs[1:])]
        res["gpu_same_problem"] = {"ms_per_lm_step": gdt / n_steps * 1e3, "costs": [float(x) for x in g],
                                   "speedup_vs_cpu": dt / gdt, "same_config": True, "same_steps": True}
        res["parity"] = {"cost_rel_diff_max": float(max(rel)), "cost_rel_diff_last": float(rel[-1]),
                         "rmse_rel_diff": float(abs(np.sqrt(sq / a.n_obs) - rmse_cpu) / rmse_cpu),
                         "pose_rel_diff": float(np.linalg.norm(cam - cam_cpu) / np.linalg.norm(cam_cpu)),
                         "point_rel_diff": float(np.linalg.norm(pts - pts_cpu) / np.linalg.norm(pts_cpu)),
                         "note": "GPU fp32 (Schur + block-Jacobi PCG 1e-6) vs CPU fp64 port (full-system Jacobi PCG 1e-5), same inputs, same LM steps"}
        eng.close()
    return res


def run_reference_gpu(a, steps):
    """The reference's eager-torch LM loop restated (oracle/torch_ba.py), fp64, on this GPU."""
    import torch
    from oracle.torch_ba import TorchRefBA
    t = TorchRefBA(a.model_id, a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices, device="cuda")
    costs = [t.step()]   # warm-up step (allocator, autograd graph construction)
    torch.cuda.synchronize()
    it0, t0 = t.pcg_iters, time.perf_counter()
    for _ in range(steps):
        costs.append(t.step())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"ms_per_lm_step": dt / steps * 1e3, "steps": steps, "pcg_iters_per_step": (t.pcg_iters - it0) / steps, "dtype": "f64",
            "costs": [float(c) for c in costs],
            "what": "oracle/torch_ba.py: bundle_adjustment.py:115-142 restated with torch ops on device='cuda' (full camera+point "
                    "system, Jacobi PCG 1e-5, matrix-free J^T(Jp) on the per-observation blocks; bae itself is not installable)"}


def run_dropin(a, steps):
    """TorchBA.Solve (the drop-in processor) on scene objects of this size: flatten + solve + write-back."""
    import torch
    from instantsfm_b200.processors.bundle_adjustment import TorchBA
    from instantsfm_b200.synthetic import ba_arrays_to_scene
    t0 = time.perf_counter()
    cameras, images, tracks = ba_arrays_to_scene(a)
    t_scene = time.perf_counter() - t0
    opts = {"min_num_view_per_track": 2, "optimize_poses": True, "thres_loss_function": 1.0, "max_num_iterations": steps,
            "function_tolerance": 0.0}
    ba = TorchBA()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ba.Solve(cameras, images, tracks, opts)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"seconds": dt, "value": a.n_obs * len(ba.loss_history) / dt, "unit": UNIT, "steps": len(ba.loss_history),
            "host_phases_seconds": getattr(ba, "last_timing", None), "scene_build_seconds_not_timed": t_scene,
            "what": "TorchBA.Solve(cameras, images, tracks, options) on python scene objects (one Camera/Image per camera, one Track per "
                    "point): vectorised flatten, C-ABI solve, vectorised write-back"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, min(args.steps, 3))
    config = args.config if args.config != "C4" else "C3"
    res = run_cpu_port(config, steps, threads, CPU_SAMPLE_SCALE)
    c1 = run_cpu_port("C1", steps, threads, 1.0)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 1, "ms_per_step": res["ms_per_lm_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{config} BAL-shaped synthetic BA (reference arm: bounded CPU sample)", "sample": res["sample"]},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "c1": {"what": "BASELINE.json config 1 at full size (same config as the `c1` object of the GPU arm)",
                   "value": c1["value"], "ms_per_lm_step": c1["ms_per_lm_step"], "costs": c1["costs"], "sample": c1["sample"]},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "bae/pypose (where the reference's LM arithmetic lives) cannot be installed offline; "
                    "this is the restated reference algorithm (oracle/) on the host cores"}
    emit(line)


def main_gp(args):
    """Global positioning, BASELINE.json config 4 (C4)."""
    import torch
    from instantsfm_b200.engine import GPEngine
    from instantsfm_b200.synthetic import make_gp_config
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(0)
    g = make_gp_config("C4", scale=args.scale)
    n_obs, n_pt, n_cam = g.translations.shape[0], g.points_3d.shape[0], g.camera_translations.shape[0]
    arrs = (g.camera_translations, g.points_3d, g.scales, g.translations, g.camera_indices, g.point_indices, g.is_calibrated)
    f32 = [np.ascontiguousarray(x, np.float32) for x in arrs[:4]] + [np.ascontiguousarray(arrs[4], np.int32),
                                                                     np.ascontiguousarray(arrs[5], np.int32), np.ascontiguousarray(arrs[6], np.uint8)]
    pinned = [torch.from_numpy(x).pin_memory().numpy() for x in f32]
    eng = GPEngine(dtype=np.float32)
    eng.set_problem(*pinned)
    for _ in range(args.warmup):
        eng.step()
    eng.reset_timers(True)
    sampler = ClockSampler(0)
    torch.cuda.synchronize()
    sampler.start()
    l0 = eng.lib.isfm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    stats, losses = [], []
    for _ in range(args.steps):
        loss, st = eng.step()
        losses.append(loss); stats.append(st)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = eng.lib.isfm_launch_count() - l0
    ms = e0.elapsed_time(e1)
    timers = eng.timers()
    eng.reset_timers(False)
    peak, peak_src = _peaks()
    # algorithmic bytes per launch: linearise 28 B/obs read + 28 written; point solve 88 B/obs
    gp_bytes = {"linearize": n_obs * (12 + 4 + 8 + 4 * 7) + n_pt * 12, "point_solve": n_obs * 88 + n_pt * 60,
                "camera_blocks": n_obs * (4 + 4 * 22), "backsub": n_obs * (4 * 24) + n_pt * 48, "cost": n_obs * (12 + 4 + 8 + 4) + n_pt * 12}
    kernels = {}
    for name, t in timers.items():
        if name == "index_prep" or t["launches"] == 0:
            continue
        sec = t["ms"] / t["launches"] * 1e-3
        ab = gp_bytes.get(name)
        kernels[name] = {"ms_per_step": t["ms"] / args.steps, "launches_per_step": t["launches"] / args.steps, "us_per_launch": sec * 1e6,
                         "achieved_gbs": ab / sec / 1e9 if ab else None, "frac_algorithmic": ab / sec / 1e9 / peak if ab else None}
    top = max((k for k in kernels if kernels[k]["achieved_gbs"]), key=lambda k: kernels[k]["ms_per_step"])
    rob, sq = eng.cost()
    eng.close()
    runs = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e = GPEngine(dtype=np.float32)
        e.set_problem(*pinned)
        for _ in range(args.steps):
            e.step()
        out = e.get_params()
        torch.cuda.synchronize()
        runs.append(time.perf_counter() - t0)
        e.close()
    cpu = None
    if not args.no_cpu:
        from oracle.gp import GPProblem, make_optimizer as gp_opt
        import torch as _t
        threads = os.cpu_count() or 1
        _t.set_num_threads(threads)
        gs = make_gp_config("C4", scale=0.01 * args.scale)
        pb = GPProblem(gs.camera_translations, gs.points_3d, gs.scales, gs.translations, gs.camera_indices, gs.point_indices, gs.is_calibrated)
        opt = gp_opt(pb, 0.1, solver="pcg", pcg_tol=1e-5)
        opt.step()
        t0 = time.perf_counter()
        for _ in range(2):
            opt.step()
        dt = time.perf_counter() - t0
        ns = gs.translations.shape[0]
        cpu = {"value": ns * 2 / dt, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_lm_step": dt / 2 * 1e3,
               "sample": f"C4 scaled x{0.01 * args.scale}: {gs.camera_translations.shape[0]} cameras / {gs.points_3d.shape[0]} points / {ns} observations, "
                         "2 LM steps of the fp64 oracle (full system incl. one scale unknown per observation, Jacobi PCG 1e-5)"}
    emit({"metric": "gp_lm_observations_per_sec", "value": n_obs * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
          "dtype": "f32", "data": "synthetic",
          "config": {"workload": f"C4 global positioning, 1DSfM-scale synthetic view graph: {n_cam} cameras / {n_pt} tracks / {n_obs} observations",
                     "l2_policy": "inputs larger than L2"},
          "lm_iters_per_sec": args.steps / (ms * 1e-3), "final_robust_cost": rob, "final_rmse": float(np.sqrt(sq / n_obs)),
          "pcg_iters": [int(s["pcg_iters"]) for s in stats], "losses": [float(x) for x in losses], "clocks": clocks,
          "gpu_launches": int(launches), "kernels": kernels,
          "roofline": {"kernel": top, "bound": "hbm", "achieved": kernels[top]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                       "frac": kernels[top]["frac_algorithmic"], "traffic": None, "peak_source": peak_src},
          "e2e": {"value": n_obs * args.steps / sorted(runs)[1], "unit": UNIT, "cold_seconds": runs[0], "warm_seconds": sorted(runs[1:])[0],
                  "h2d_bytes_per_step": sum(x.nbytes for x in f32) / args.steps, "d2h_bytes_per_step": sum(x.nbytes for x in out) / args.steps},
          "cpu_baseline": cpu})


def main_ours(args):
    import torch
    import torch.distributed as dist
    from instantsfm_b200.engine import Communicator
    from instantsfm_b200.synthetic import make_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = Communicator()

    # one global instance; this rank's contiguous point range of it
    a = make_config(args.config, scale=args.scale, shard=(rank, world))
    n_total = a.n_obs_total
    res = run_resident(a, comm, world, local_rank, args.steps, args.warmup, n_total)
    runs, h2d, d2h = run_e2e(a, comm, world, args.steps, runs=1 if args.quick else 3)
    warm = sorted(r[0] for r in runs[1:])[0] if len(runs) > 1 else runs[0][0]
    e2e = {"value": n_total * args.steps / warm, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
           "includes": f"isfm_ba_create + set_problem (H2D, sort, Schur pattern) + {args.steps} steps + get_params (D2H) + destroy, per rank",
           "seconds": warm, "cold_seconds": runs[0][0], "cold_value": n_total * args.steps / runs[0][0],
           "setup_seconds": [r[1] for r in runs], "runs_seconds": [r[0] for r in runs],
           "note": "`value` = best warm run (device-block cache filled by the first solve); `cold_*` = first solve of the process"}

    extra = {}
    if world > 1 and args.config == "C3" and not args.no_c5:
        # the north-star multi-GPU problem, same instance at every rank count
        a5 = make_config("C5", shard=(rank, world))
        r5 = run_resident(a5, comm, world, local_rank, C5_STEPS, C5_WARMUP, a5.n_obs_total)
        c5 = {"workload": "C5 city-scale synthetic BA (street geometry): 20000 cameras / 10000000 points / 60000000 observations, strong scaling",
              "n_gpus": world, "steps": C5_STEPS, "warmup": C5_WARMUP, "ms_per_step": r5["ms_total"] / C5_STEPS, "value": r5["value"], "unit": UNIT,
              "losses": r5["losses"], "final_robust_cost": r5["final_robust_cost"], "initial_robust_cost": r5["initial_robust_cost"],
              "final_rmse_px": r5["final_rmse_px"],
              "pcg_iters": r5["pcg_iters"], "rejects": r5["rejects"], "work": r5["work"],
              "kernels": {k: {"ms_per_step": v["ms_per_step"], "us_per_launch": v["us_per_launch"]} for k, v in r5["kernels"].items()}}
        fx = os.path.join(ROOT, "profiles", "r2_c5_n1.json")
        if os.path.exists(fx):
            with open(fx) as f:
                ref = json.load(f)
            ref = ref.get("c5", ref)
            c5["n1_reference"] = {"file": "profiles/r2_c5_n1.json", "ms_per_step": ref["ms_per_step"], "final_robust_cost": ref["final_robust_cost"],
                                  "steps": ref.get("steps"), "warmup": ref.get("warmup"), "pcg_iters": ref.get("pcg_iters")}
            c5["same_window_as_n1"] = bool(ref.get("steps") == C5_STEPS and ref.get("warmup") == C5_WARMUP)
            c5["speedup_vs_n1"] = ref["ms_per_step"] / c5["ms_per_step"]
            c5["cost_rel_diff_vs_n1"] = abs(c5["final_robust_cost"] - ref["final_robust_cost"]) / ref["final_robust_cost"]
            c5["loss_rel_diff_vs_n1_max"] = float(max(abs(x - y) / y for x, y in zip(c5["losses"], ref["losses"])))
            c5["loss_rel_diff_note"] = ("max over the LM steps run; the first step is a long nonlinear step from the perturbed start whose outcome "
                                        "depends on the inexact (1e-6) PCG solve at the 1e-2 level; the rank counts meet again: see cost_rel_diff_vs_n1")
            if ref.get("initial_robust_cost"):
                c5["initial_cost_rel_diff_vs_n1"] = abs(c5["initial_robust_cost"] - ref["initial_robust_cost"]) / ref["initial_robust_cost"]
        extra["c5"] = c5
    if rank == 0 and world == 1 and args.config == "C5":
        extra["c5"] = {"ms_per_step": res["ms_total"] / args.steps, "final_robust_cost": res["final_robust_cost"], "losses": res["losses"],
                       "initial_robust_cost": res["initial_robust_cost"],
                       "steps": args.steps, "warmup": args.warmup, "pcg_iters": res["pcg_iters"], "rejects": res["rejects"], "work": res["work"]}

    cpu = c1 = ref_gpu = dropin = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        try:   # a baseline leg must never take the bench line down
            cpu = run_cpu_port(args.config if args.config != "C5" else "C3", 2, threads, CPU_SAMPLE_SCALE, compare_gpu=True)
        except Exception as e:
            cpu = {"error": repr(e)[:300], "kind": "port", "cores": threads}
        try:
            c1 = run_cpu_port("C1", 3, threads, 1.0, compare_gpu=True)
            c1["what"] = "BASELINE.json config 1 (64 cameras / 10 k points / 60 k observations) at FULL size on both arms"
        except Exception as e:
            c1 = {"error": repr(e)[:300]}
    if rank == 0 and world == 1 and not args.quick and args.config in ("C1", "C2", "C3"):
        try:
            ref_gpu = run_reference_gpu(a, 2)
            ref_gpu["ours_ms_per_lm_step"] = res["ms_total"] / args.steps
            ref_gpu["speedup_per_lm_step"] = ref_gpu["ms_per_lm_step"] / (res["ms_total"] / args.steps)
        except Exception as e:   # a baseline leg must never take the bench line down
            ref_gpu = {"error": repr(e)[:300]}
        try:
            dropin = run_dropin(a, args.steps)
        except Exception as e:
            dropin = {"error": repr(e)[:300]}

    if rank == 0:
        pat = res["pattern"]
        mv_owned, mv_total = res["matvec"]
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": res["ms_total"] / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{args.config} BAL-shaped synthetic BA, RADIAL cameras: {a.n_cam} cameras / {a.n_pt_total} points / "
                           f"{a.n_obs_total} observations" + (f", ONE instance split over {world} ranks by contiguous point ranges "
                                                              f"(rank 0: {a.n_pt} points / {a.n_obs} observations)" if world > 1 else ""),
                           "total_observations": n_total, "l2_policy": "inputs larger than L2 (J blocks alone exceed 126 MB)",
                           "pcg_tol": 1e-6, "schur_blocks": pat["nnzb"], "schur_pairs": pat["n_pairs"]},
                "lm_iters_per_sec": args.steps / (res["ms_total"] * 1e-3), "final_rmse_px": res["final_rmse_px"],
                "final_robust_cost": res["final_robust_cost"], "initial_robust_cost": res["initial_robust_cost"],
                "pcg_iters": res["pcg_iters"], "losses": res["losses"], "rejects": res["rejects"],
                "work": res["work"],
                "matvec_split": (None if world == 1 else {"units_owned_rank0": mv_owned, "units_total": mv_total}),
                "pcg_exchange": (None if world == 1 else ("peer-memory push over NVLink inside the persistent PCG kernel"
                                                          if comm.peer_enabled else "ncclAllReduce per iteration")),
                "clocks": res["clocks"], "e2e": e2e, "e2e_dropin": dropin, "gpu_launches": res["launches"], "roofline": res["roofline"],
                "kernels": res["kernels"], "cpu_baseline": cpu, "c1": c1, "reference_gpu": ref_gpu}
        line.update(extra)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink/grow points and observations (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / c1 legs")
    ap.add_argument("--no-c5", action="store_true", help="N > 1: skip the C5 (60 M observations) leg")
    ap.add_argument("--quick", action="store_true", help="one e2e run, no reference_gpu / drop-in legs")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    elif args.config == "C4":
        main_gp(args)
    else:
        main_ours(args)
