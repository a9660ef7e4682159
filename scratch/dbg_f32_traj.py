import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from instantsfm_b200.processors import TorchBA
from oracle import ba as oba
import tests.test_processors_gpu as T
cameras, images, tracks = T._ba_scene()
c2, i2, t2 = copy.deepcopy((cameras, images, tracks))
opts = dict(T.BA_OPTS, max_num_iterations=25)
hist, opt, pb = oba.solve(c2, i2, t2, opts, solver="direct")
for dtype, tol, floor in [(np.float32, 1e-6, 0), (np.float32, 1e-6, 1e-7), (np.float32, 1e-6, 1e-6), (np.float32, 1e-6, 1e-5), (np.float64, 1e-6, 1e-6), (np.float64, 1e-6, 1e-5)]:
    os.environ["ISFM_MIN_DAMPING"] = str(floor)
    c1, i1, t1 = copy.deepcopy(T._ba_scene())
    ba = TorchBA(dtype=dtype, pcg_tol=tol)
    ba.Solve(c1, i1, t1, opts)
    rel = (np.array(ba.loss_history) - np.array(hist)) / np.array(hist)
    print(dtype.__name__, tol, "floor", floor, "max rel", np.abs(rel).max(), "at", np.abs(rel).argmax())
    print(" rel:", " ".join("%.1e" % r for r in rel))
    print(" pcg:", [s["pcg_iters"] for s in ba.last_stats], "trials", [s["trials"] for s in ba.last_stats])
print("oracle trials", [len(t["trials"]) for t in opt.trace])
print("hist", [round(h, 2) for h in hist])
