"""Generates tests/golden/ba_trajectory_<config>.npz: the fp64 oracle's LM trajectory on a
BASELINE.json config (exact solves through the point Schur complement, oracle/lm.py::schur_direct).

    python tests/golden/make_ba_trajectory_golden.py C2 12

The GPU parity test (tests/test_baseline_configs_gpu.py) replays the same seeded instance
through the C ABI and compares per-iteration cost, final RMSE, every camera row and a fixed
sample of the points.  The oracle needs minutes per config at this size, which is why its
output is committed instead of being recomputed on the GPU box.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from instantsfm_b200.synthetic import make_config, CONFIGS  # noqa: E402
from oracle.ba import BAProblem, make_optimizer  # noqa: E402

POINT_SAMPLE = 4096


def main():
    config = sys.argv[1] if len(sys.argv) > 1 else "C2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    a = make_config(config)
    pb = BAProblem(a.model_id, a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
    opt = make_optimizer(pb, 1.0, solver="schur")
    costs, rmse = [], [pb.rmse()]
    for it in range(steps):
        t0 = time.time()
        costs.append(opt.step())
        rmse.append(pb.rmse())
        print(f"{config} step {it}: cost {costs[-1]:.9e} rmse {rmse[-1]:.6f} trials {len(opt.trace[-1]['trials'])} ({time.time() - t0:.1f} s)", flush=True)
    sample = np.linspace(0, a.n_pt - 1, POINT_SAMPLE).astype(np.int64)
    out = os.path.join(ROOT, "tests", "golden", f"ba_trajectory_{config}.npz")
    np.savez_compressed(out, config=config, sizes=np.array(CONFIGS[config][:3]), costs=np.array(costs), rmse=np.array(rmse),
                        trials=np.array([len(t["trials"]) for t in opt.trace]), cam=pb.cam, point_sample=sample,
                        points=pb.pts[sample], initial_cost_check=np.array([opt.trace[0]["loss_before"]]),
                        obs_checksum=np.array([a.points_2d.sum(), a.camera_indices.astype(np.int64).sum()]))
    print("wrote", out)


if __name__ == "__main__":
    main()
