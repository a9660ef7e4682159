"""Parity diagnostic (GPU): the two fp32 parity tests that fail, under solver variants selected by env."""
import copy, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instantsfm_b200.engine import BAEngine
from instantsfm_b200.synthetic import make_config
from oracle.ba import BAProblem, make_optimizer

def nrel(x, ref): return float(np.linalg.norm(np.asarray(x, np.float64) - ref) / np.linalg.norm(ref))

which = sys.argv[1] if len(sys.argv) > 1 else "c1"
tols = [float(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1e-6]
if which == "c1":
    a = make_config("C1")
    pb = BAProblem(a.model_id, a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
    opt = make_optimizer(pb, 1.0, solver="schur")
    ref = [opt.step() for _ in range(12)]
    for dtype in (np.float32,):
        for tol in tols:
            eng = BAEngine(a.model_id, dtype=dtype, pcg_tol=tol)
            eng.set_problem(a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
            costs, its, stt = [], 0, []
            for _ in range(12):
                l, st = eng.step(); costs.append(l); its += st["pcg_iters"]; stt.append(st["pcg_status"])
            cam, pts = eng.get_params()
            d = np.linalg.norm(pts.astype(np.float64) - pb.pts, axis=1)
            print(f"C1 {np.dtype(dtype).name} tol {tol:g}: cost rel {['%.1e' % (abs(x-y)/y) for x,y in zip(costs, ref)]} cam {nrel(cam, pb.cam):.2e} "
                  f"pts {nrel(pts, pb.pts):.2e} worst pts {np.sort(d)[-5:]} median {np.median(d):.2e} pcg its {its} status {stt}", flush=True)
            eng.close()
else:
    from tests.test_processors_gpu import _ba_scene, BA_OPTS
    from instantsfm_b200.processors import TorchBA
    from oracle import ba as oba
    cameras, images, tracks = _ba_scene()
    c2, i2, t2 = copy.deepcopy((cameras, images, tracks))
    opts = dict(BA_OPTS, max_num_iterations=25)
    hist, _, pb = oba.solve(c2, i2, t2, opts, solver="direct")
    for tol in tols:
        c, i, t = copy.deepcopy((cameras, images, tracks))
        ba = TorchBA(dtype=np.float32, pcg_tol=tol)
        ba.Solve(c, i, t, opts)
        n = min(len(hist), len(ba.loss_history))
        print(f"proc f32 tol {tol:g}: len {len(ba.loss_history)}/{len(hist)} rel {['%.1e' % (abs(x-y)/y) for x,y in zip(ba.loss_history[:n], hist[:n])]}", flush=True)
