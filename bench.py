#!/usr/bin/env python
"""Benchmark of the BA Levenberg-Marquardt hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3]

A "step" is one LM iteration (`isfm_ba_step` == `optimizer.step(input)`,
bundle_adjustment.py:132) on a synthetic BAL-shaped problem.  N = 1: config C3
(Venice-shaped, 1 778 cameras / 993 k points / 5.0 M observations).  N > 1 (torchrun, one
rank per GPU): weak scaling -- every rank owns a C3-sized shard of points over the same
1 778 cameras, the ranks exchange the camera-system partial sums and the PCG mat-vecs over
NCCL.  Prints ONE JSON line (rank 0).

value       observations/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
e2e         same metric through the C ABI from pinned HOST buffers: create + set_problem
            (H2D of every tensor + integer prep) + K steps + get_params (D2H), per rank
roofline    dominant kernel family: algorithmic bytes / CUDA-event time
cpu_baseline the oracle (restated reference algorithm, "port") on the host cores, bounded sample
--impl reference  times that CPU port alone (bae / pypose are not installable)
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
# dram__bytes_read.sum + dram__bytes_write.sum per launch on config C3 (1 GPU, full size), from the
# `ncu --set full` captures summarised in profiles/r1_ncu_full_summary.txt (ncu cannot run inside the bench)
NCU_TRAFFIC_C3 = {"pcg_spmv": 535.1e6, "schur_offdiag": 2887.5e6, "linearize": 788.6e6, "camera_blocks": 664.7e6,
                  "backsub": 800.7e6}


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
    """Algorithmic HBM bytes of ONE launch of a kernel family."""
    dd = d * d
    table = {
        # fused K1 + point solve: obs + indices read; R and the full record (Jc | Jp | V | rho, padded) written;
        # per point X read, Hpp | g_p | Hpp^-1 | t_p written; camera table read
        "linearize": n_obs * (8 + 8 + fsize * (2 + (2 * d + 14 + 3) // 4 * 4)) + n_pt * (8 + fsize * (3 + 18)) + n_cam * (9 + d) * fsize,
        "point_blocks": n_obs * fsize * (6 + 2 + 6) + n_pt * (4 + fsize * (6 + 3 + 6 + 3)),
        "point_solve": n_obs * fsize * (6 + 6) + n_pt * (4 + fsize * (6 + 3 + 6 + 3)),
        # one sweep per trial: camera-major index + Jc | Jp | V | rho of every record; Hcc - E_ii, diag Hcc, g_c - e written
        "camera_blocks": n_obs * (4 + fsize * (2 * d + 14)) + n_cam * fsize * (2 * dd + 2 * d),
        "schur_offdiag": n_pairs * (8 + fsize * (4 * d + 12)) + n_lists * (16 + dd * fsize),
        # upper triangle only: blocks + col/tpos indices, B^T p_i deposits written, p gathered
        "pcg_spmv": ((nnzb + n_cam) // 2) * (dd * fsize + 8) + ((nnzb - n_cam) // 2) * d * fsize + n_cam * d * fsize,
        "backsub": n_obs * (4 + fsize * (2 * d + 6 + 2)) + n_pt * fsize * (3 + 6 + 3 + 3 + 3),
        "cost": n_obs * (8 + 8) + n_pt * 3 * fsize + n_cam * (9 + d) * fsize,
    }
    return float(table.get(name, 0.0))


def run_cpu_port(config, n_steps, threads):
    """The oracle run the reference's way (full system, Jacobi PCG tol 1e-5) on a bounded sample."""
    import torch
    from instantsfm_b200.synthetic import make_config
    from oracle.ba import BAProblem, make_optimizer
    torch.set_num_threads(threads)
    a = make_config(config, scale=CPU_SAMPLE_SCALE)
    pb = BAProblem(a.model_id, a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
    opt = make_optimizer(pb, 1.0, solver="pcg", pcg_tol=1e-5)
    opt.step()  # warm-up (first call pays torch / scipy import costs)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        opt.step()
    dt = time.perf_counter() - t0
    sample = (f"{config} scaled x{CPU_SAMPLE_SCALE}: {a.n_cam} cameras / {a.n_pt} points / {a.n_obs} observations, "
              f"{n_steps} LM steps of the fp64 torch/scipy oracle (full normal equations, Jacobi PCG 1e-5)")
    return a.n_obs * n_steps / dt, dt / n_steps * 1e3, sample, a


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, min(args.steps, 3))
    val, ms, sample, a = run_cpu_port(args.config, steps, threads)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.config} BAL-shaped synthetic BA (reference arm: bounded CPU sample)", "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "bae/pypose (where the reference's LM arithmetic lives) cannot be installed offline; "
                    "this is the restated reference algorithm (oracle/) on the host cores"}
    emit(line)


def main_ours(args):
    import torch
    import torch.distributed as dist
    from instantsfm_b200.engine import BAEngine, Communicator
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
    strong = args.scaling == "strong"
    # weak: every rank owns one full config worth of points; strong: 1/world of them
    a = make_config(args.config, scale=(1.0 if strong else float(world)) * args.scale, shard=(rank, world))
    n_total = a.n_obs * world
    d = a.camera_params.shape[1] - 1
    dtype = np.float32

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host = [np.ascontiguousarray(x, dtype=dtype) for x in (a.camera_params, a.camera_pps, a.points_3d, a.points_2d)]
    host += [np.ascontiguousarray(a.camera_indices, np.int32), np.ascontiguousarray(a.point_indices, np.int32)]
    pinned = [torch.from_numpy(h).pin_memory() for h in host]
    pinned_np = [p.numpy() for p in pinned]

    # ---- device-resident throughput ------------------------------------------------------
    eng = BAEngine(a.model_id, dtype=dtype, comm=comm)
    eng.set_problem(*pinned_np)
    pat = eng.schur_pattern()
    mv_owned, mv_total = eng.matvec_units()
    for _ in range(args.warmup):
        eng.step()
    saved_params = eng.get_params()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = eng.lib.isfm_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    losses, stats = [], []
    for _ in range(args.steps):
        loss, st = eng.step()
        losses.append(loss); stats.append(st)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    launches = eng.lib.isfm_launch_count() - launches0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = n_total * args.steps / (ms_total * 1e-3)
    rob, sq = eng.cost()
    rmse = float(np.sqrt(sq / n_total))

    # ---- per-kernel CUDA-event timing for the roofline (separate pass, timers on) ---------
    # replay of the timed steps from the same parameters (trust-region radius continues)
    eng.set_params(*saved_params)
    eng.reset_timers(True)
    prof_steps = 1 if args.quick else args.steps
    pcg_iters_prof = 0
    torch.cuda.synchronize()
    t_prof = time.perf_counter()
    for _ in range(prof_steps):
        _, st = eng.step()
        pcg_iters_prof += st["pcg_iters"]
    torch.cuda.synchronize()
    prof_wall_ms = (time.perf_counter() - t_prof) * 1e3 / prof_steps
    timers = eng.timers()
    eng.reset_timers(False)
    peak, peak_src = _peaks()
    kernels = {}
    for name, t in timers.items():
        if name in ("index_prep",) or t["launches"] == 0:
            continue
        # PCG kernels are launched in chunks of 8 iterations and exit early once converged:
        # charge their time to the iterations that did work
        eff = pcg_iters_prof if name == "pcg_spmv" else t["launches"]
        per_launch_ms = t["ms"] / max(eff, 1)
        ab = algorithmic_bytes(name, a.n_obs, a.n_pt, a.n_cam, d, pat["n_pairs"], (pat["nnzb"] - a.n_cam) // 2, pat["nnzb"])
        if name == "pcg_spmv":
            ab *= mv_owned / max(mv_total, 1)   # split mat-vec: this rank streams only its own unit range
        kernels[name] = {"ms_per_step": t["ms"] / prof_steps, "launches_per_step": t["launches"] / prof_steps,
                         "us_per_launch": per_launch_ms * 1e3,
                         "achieved_gbs": (ab / (per_launch_ms * 1e-3) / 1e9) if ab and per_launch_ms > 0 else None}
    top = max((k for k in kernels if kernels[k]["achieved_gbs"] is not None), key=lambda k: kernels[k]["ms_per_step"])
    roofline = {"kernel": top, "bound": "hbm", "achieved": kernels[top]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[top]["achieved_gbs"] / peak,
                "traffic": NCU_TRAFFIC_C3.get(top) if (args.config == "C3" and world == 1 and args.scale == 1.0) else None,
                "traffic_source": "profiles/r1_ncu_full_summary.txt (bytes per launch)", "peak_source": peak_src,
                "share_of_step": kernels[top]["ms_per_step"] / sum(k["ms_per_step"] for k in kernels.values())}
    # ---- end to end through the C ABI from pinned host buffers ---------------------------
    # three complete solves (create, set_problem = H2D of every tensor + sort + Schur pattern,
    # K steps, get_params = D2H, destroy); the median run is reported
    runs = []
    for _ in range(1 if args.quick else 3):
        barrier()
        t0 = time.perf_counter()
        eng2 = BAEngine(a.model_id, dtype=dtype, comm=comm)
        eng2.set_problem(*pinned_np)
        t_setup = time.perf_counter() - t0
        for _ in range(args.steps):
            eng2.step()
        cam_out, pts_out = eng2.get_params()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        runs.append((float(dt.item()), t_setup))
        eng2.close()
    e2e_s, t_setup = sorted(runs)[len(runs) // 2]
    h2d = sum(h.nbytes for h in host)
    d2h = cam_out.nbytes + pts_out.nbytes + 8 * args.steps
    e2e = {"value": n_total * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps,
           "d2h_bytes_per_step": d2h / args.steps, "includes": "isfm_ba_create + set_problem (H2D, sort, Schur pattern) + "
           f"{args.steps} steps + get_params (D2H) + destroy, per rank; median of 3 runs", "seconds": e2e_s,
           "setup_seconds": t_setup, "runs_seconds": [r[0] for r in runs]}
    eng.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        val, ms_cpu, sample, _ = run_cpu_port(args.config, 2, threads)
        cpu = {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "ms_per_lm_step": ms_cpu}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{args.config} BAL-shaped synthetic BA, RADIAL cameras, per rank {a.n_cam} cameras / "
                           f"{a.n_pt} points / {a.n_obs} observations" + (f" (x{world} ranks, cameras shared)" if world > 1 else ""),
                           "total_observations": n_total, "l2_policy": "inputs larger than L2 (J blocks alone exceed 126 MB)",
                           "pcg_tol": 1e-6, "schur_blocks": pat["nnzb"], "schur_pairs": pat["n_pairs"]},
                "lm_iters_per_sec": args.steps / (ms_total * 1e-3), "final_rmse_px": rmse, "final_robust_cost": rob,
                "pcg_iters_per_step": float(np.mean([s["pcg_iters"] for s in stats])),
                "pcg_iters": [int(s["pcg_iters"]) for s in stats], "losses": [float(x) for x in losses],
                "rejects": int(sum(s["rejects"] for s in stats)),
                "matvec_split": (None if world == 1 else {"units_owned_rank0": mv_owned, "units_total": mv_total,
                                                          "note": "identical block pattern on every rank: summed E reduce-scattered by unit "
                                                                  "ranges once per trial, each rank multiplies its own range"
                                                                  if mv_owned < mv_total else "every rank multiplies its own partial E_g"}),
                "pcg_exchange": (None if world == 1 else ("peer-memory push over NVLink fused into the PCG kernels (device-side WHILE graph)"
                                                          if comm.peer_enabled else "ncclAllReduce per iteration")),
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels,
                "profile_pass": {"pcg_iters_per_step": pcg_iters_prof / prof_steps, "wall_ms_per_step": prof_wall_ms,
                                 "kernel_ms_per_step": sum(k["ms_per_step"] for k in kernels.values())},
                "cpu_baseline": cpu}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink/grow points and observations (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="timed pass only (no per-kernel pass, one e2e run)")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)
