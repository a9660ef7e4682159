"""Generates tests/golden/gp_trajectory_C4x0.1.npz: the fp64 oracle's LM trajectory of global
positioning on BASELINE.json config 4 scaled by 0.1 (2 500 cameras / 50 k tracks / 300 k observations;
full [centres | points | scales] system, Jacobi PCG to 1e-10 -- the exact sparse solve needs > 15
minutes per step at this size).

    python tests/golden/make_gp_trajectory_golden.py [steps]

tests/test_gp_gpu.py::test_c4_tenth_trajectory_fp32 replays the same seeded instance through the C ABI
in fp32 and compares per-iteration cost, trial counts and the final centres.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from instantsfm_b200.synthetic import make_gp_config  # noqa: E402
from oracle.gp import GPProblem, make_optimizer  # noqa: E402

SCALE = 0.1


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    g = make_gp_config("C4", scale=SCALE)
    pb = GPProblem(g.camera_translations, g.points_3d, g.scales, g.translations, g.camera_indices, g.point_indices, g.is_calibrated)
    opt = make_optimizer(pb, 0.1, solver="pcg", pcg_tol=1e-10)
    costs = []
    for it in range(steps):
        t0 = time.time()
        costs.append(opt.step())
        print(f"C4x{SCALE} step {it}: cost {costs[-1]:.9e} trials {len(opt.trace[-1]['trials'])} ({time.time() - t0:.1f} s)", flush=True)
    out = os.path.join(ROOT, "tests", "golden", f"gp_trajectory_C4x{SCALE}.npz")
    np.savez_compressed(out, costs=np.array(costs), trials=np.array([len(t["trials"]) for t in opt.trace]), centres=pb.c,
                        sizes=np.array([g.camera_translations.shape[0], g.points_3d.shape[0], g.translations.shape[0]]),
                        checksum=np.array([g.translations.sum(), g.camera_indices.astype(np.int64).sum()]))
    print("wrote", out)


if __name__ == "__main__":
    main()
