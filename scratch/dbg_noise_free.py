import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from instantsfm_b200.engine import BAEngine
from instantsfm_b200.synthetic import make_ba_problem
a = make_ba_problem(12, 400, 2400, seed=29, noise_px=0.0, outlier_frac=0.0, perturb=0.3)
eng = BAEngine(a.model_id, dtype=np.float64, pcg_tol=1e-12)
eng.set_problem(a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
for it in range(50):
    loss, st = eng.step()
    print(it, loss, st['trials'], st['rejects'], st['pcg_iters'], st['quality'], st['damping'], st['model_term'], st['step_norm_cam'])
    if not np.isfinite(loss): break
print(eng.cost())
