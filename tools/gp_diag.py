"""GP C4 x 0.1 fp32 trajectory vs the committed oracle trajectory for several PCG tolerances (GPU)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instantsfm_b200.engine import GPEngine
from instantsfm_b200.synthetic import make_gp_config
gold = np.load(os.path.join(ROOT, "tests", "golden", "gp_trajectory_C4x0.1.npz"))
g = make_gp_config("C4", scale=0.1)
for dtype in (np.float32, np.float64):
    for tol in (1e-5, 1e-6, 1e-7, 1e-8):
        eng = GPEngine(dtype=dtype, pcg_tol=tol)
        eng.set_problem(g.camera_translations, g.points_3d, g.scales, g.translations, g.camera_indices, g.point_indices, g.is_calibrated, None)
        out = []
        for it, ref in enumerate(gold["costs"]):
            loss, st = eng.step()
            out.append("%.1e%s" % (abs(loss - ref) / ref, "" if st["trials"] == int(gold["trials"][it]) else "!t%d" % st["trials"]))
        print(np.dtype(dtype).name, tol, out, flush=True)
        eng.close()
