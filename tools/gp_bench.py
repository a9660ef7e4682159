"""Global-positioning benchmark on the C4 config (2.5 k cameras / 500 k tracks / 3.0 M observations):
LM iterations/s and observations/s with per-kernel CUDA-event timings.  python tools/gp_bench.py [--steps 10]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from instantsfm_b200.engine import GPEngine  # noqa: E402
from instantsfm_b200.synthetic import make_gp_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--scale", type=float, default=1.0)
args = ap.parse_args()
g = make_gp_config("C4", scale=args.scale)
eng = GPEngine(dtype=np.float32)
t0 = time.perf_counter()
eng.set_problem(g.camera_translations, g.points_3d, g.scales, g.translations, g.camera_indices, g.point_indices, g.is_calibrated)
setup = time.perf_counter() - t0
hist = [eng.step()[0] for _ in range(args.warmup)]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
stats = []
for _ in range(args.steps):
    loss, st = eng.step()
    hist.append(loss); stats.append(st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
eng.reset_timers(True)
for _ in range(args.steps):
    hist.append(eng.step()[0])
timers = eng.timers()
n_obs = g.translations.shape[0]
print(json.dumps({"config": "C4 GP", "n_cam": int(g.camera_translations.shape[0]), "n_pt": int(g.points_3d.shape[0]), "n_obs": int(n_obs),
                  "ms_per_step": ms / args.steps, "lm_iters_per_sec": args.steps / (ms * 1e-3), "obs_per_sec": n_obs * args.steps / (ms * 1e-3),
                  "pcg_iters_per_step": float(np.mean([s["pcg_iters"] for s in stats])), "rejects": int(sum(s["rejects"] for s in stats)),
                  "setup_seconds": setup, "loss_first": hist[0], "loss_last": hist[-1],
                  "kernels_ms_per_step": {k: round(v["ms"] / args.steps, 4) for k, v in timers.items()}}))
