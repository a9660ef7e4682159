"""Parity diagnostic (GPU): C1 (live oracle) and C2 (golden) errors of the fp32 build for several PCG tolerances."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instantsfm_b200.engine import BAEngine
from instantsfm_b200.synthetic import make_config
from oracle.ba import BAProblem, make_optimizer

def nrel(x, ref): return float(np.linalg.norm(np.asarray(x, np.float64) - ref) / np.linalg.norm(ref))

a = make_config("C1")
pb = BAProblem(a.model_id, a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
opt = make_optimizer(pb, 1.0, solver="schur")
ref = [opt.step() for _ in range(12)]
for dtype in (np.float32, np.float64):
    for tol in (1e-6, 3e-7, 1e-7, 1e-8):
        eng = BAEngine(a.model_id, dtype=dtype, pcg_tol=tol)
        eng.set_problem(a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
        costs, its = [], 0
        for _ in range(12):
            l, st = eng.step(); costs.append(l); its += st["pcg_iters"]
        cam, pts = eng.get_params()
        print(f"C1 {np.dtype(dtype).name} tol {tol:g}: cost rel max {max(abs(x-y)/y for x,y in zip(costs, ref)):.2e} cam {nrel(cam, pb.cam):.2e} "
              f"t {nrel(cam[:,:3], pb.cam[:,:3]):.2e} q {nrel(cam[:,3:7], pb.cam[:,3:7]):.2e} intr {nrel(cam[:,7:], pb.cam[:,7:]):.2e} pts {nrel(pts, pb.pts):.2e} pcg its {its}", flush=True)
        eng.close()
g = np.load(os.path.join(ROOT, "tests", "golden", "ba_trajectory_C2.npz"))
a = make_config("C2")
for tol in (1e-6, 3e-7, 1e-7):
    eng = BAEngine(a.model_id, dtype=np.float32, pcg_tol=tol)
    eng.set_problem(a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
    costs, its = [], 0
    t0 = time.perf_counter()
    for _ in range(12):
        l, st = eng.step(); costs.append(l); its += st["pcg_iters"]
    dt = time.perf_counter() - t0
    cam, pts = eng.get_params()
    print(f"C2 f32 tol {tol:g}: cost rel max {max(abs(x-y)/y for x,y in zip(costs, g['costs'])):.2e} cam {nrel(cam, g['cam']):.2e} pts {nrel(pts[g['point_sample']], g['points']):.2e} pcg its {its} {dt*1e3/12:.2f} ms/step", flush=True)
    eng.close()
