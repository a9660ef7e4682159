"""Minimal driver for ncu captures: C3 (or --config) problem, a few LM steps, nothing else.
Usage: python tools/prof_run.py [--config C3] [--steps 2]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instantsfm_b200.engine import BAEngine  # noqa: E402
from instantsfm_b200.synthetic import make_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C3")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--scale", type=float, default=1.0)
args = ap.parse_args()
a = make_config(args.config, scale=args.scale, shard=(0, 1))
eng = BAEngine(a.model_id, dtype=np.float32)
eng.set_problem(a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
for _ in range(args.steps):
    loss, st = eng.step()
    print("loss %.6e pcg %d trials %d" % (loss, st["pcg_iters"], st["trials"]))
