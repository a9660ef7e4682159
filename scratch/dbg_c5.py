import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from instantsfm_b200.engine import BAEngine
from instantsfm_b200.synthetic import make_config
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
a = make_config("C5", scale=scale, shard=(0, 1))
print("problem", a.n_cam, a.n_pt, a.n_obs, "cam obs min", np.bincount(a.camera_indices, minlength=a.n_cam).min())
eng = BAEngine(a.model_id, dtype=np.float32)
eng.set_problem(a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
pat = eng.schur_pattern()
print("nnzb", pat["nnzb"], "pairs", pat["n_pairs"])
rob, sq = eng.cost(); print("cost0", rob, "rmse0", np.sqrt(sq / a.n_obs))
for it in range(2):
    loss, st = eng.step()
    print(it, loss, st)
for what in ["hcc", "gc", "schur_rhs"]:
    v = eng.debug(what); print(what, np.isfinite(v).all(), np.abs(v).max())
