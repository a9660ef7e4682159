"""Generate tests/golden/reference_cost_functions.npz FROM THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_reference_golden.py

It imports /root/reference/instantsfm/utils/cost_function.py unmodified.  Three of its
imports are not installable here (no network): ``pyceres`` (only used by the Fetzer cost
classes, out of scope), ``bae.autograd.function`` (``map_transform`` is a tracing
decorator, ``TrackingTensor`` a tensor subclass -- neither changes values) and
``bae.utils.ba.rotate_quat``.  They are stubbed; ``rotate_quat`` is stubbed with the
textbook R(q) p + t written directly from the storage convention
[t, q = (x, y, z, w)] that the reference's callers establish (bundle_adjustment.py:71,
track_retriangulation.py:65-67) via scipy's Rotation, i.e. independently of the oracle.

The file pins: the nine ``reproject_*`` functions, ``pairwise_cost`` and the
``get_camera_model_info`` tables of scene/defs.py.
"""
import os
import sys
import types

import numpy as np
import torch
from scipy.spatial.transform import Rotation

REF = "/root/reference"


def _install_stubs():
    pyceres = types.ModuleType("pyceres")

    class CostFunction:  # noqa: D401 - stub
        def __init__(self, *a, **k):
            pass
    pyceres.CostFunction = CostFunction
    sys.modules["pyceres"] = pyceres

    bae = types.ModuleType("bae")
    bae_utils = types.ModuleType("bae.utils")
    bae_ba = types.ModuleType("bae.utils.ba")
    bae_autograd = types.ModuleType("bae.autograd")
    bae_fn = types.ModuleType("bae.autograd.function")

    def rotate_quat(points, pose7):
        p = points.detach().numpy()
        pose = pose7.detach().numpy()
        R = Rotation.from_quat(pose[..., 3:7]).as_matrix()  # scipy: xyzw
        y = np.einsum("...ij,...j->...i", R, p) + pose[..., :3]
        return torch.from_numpy(y)

    bae_ba.rotate_quat = rotate_quat
    bae_fn.map_transform = lambda f: f
    bae_fn.TrackingTensor = lambda t: t
    for name, mod in [("bae", bae), ("bae.utils", bae_utils), ("bae.utils.ba", bae_ba),
                      ("bae.autograd", bae_autograd), ("bae.autograd.function", bae_fn)]:
        sys.modules[name] = mod


def main():
    _install_stubs()
    sys.path.insert(0, REF)
    from instantsfm.utils import cost_function as cf
    from instantsfm.scene.defs import CameraModelId, get_camera_model_info

    rng = np.random.default_rng(20261018)
    n = 64
    out = {}
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    t = rng.normal(size=(n, 3))
    # points in front of the camera: choose camera-frame y then map back to world
    y = np.concatenate([rng.uniform(-1.5, 1.5, size=(n, 2)), rng.uniform(2.0, 6.0, size=(n, 1))], 1)
    R = Rotation.from_quat(q).as_matrix()
    X = np.einsum("nji,nj->ni", R, y - t)
    pp = rng.uniform(300, 700, size=(n, 2))
    out["pose"] = np.concatenate([t, q], 1)
    out["points"] = X
    out["pp"] = pp

    n_intr = {0: 1, 1: 2, 2: 2, 3: 3, 4: 6, 5: 6, 6: 10, 8: 2, 9: 3}
    for mid, ni in n_intr.items():
        intr = np.empty((n, ni))
        nf = 1 if mid in (0, 2, 3, 8, 9) else 2
        intr[:, :nf] = rng.uniform(500, 1500, size=(n, nf))
        intr[:, nf:] = rng.normal(scale=0.02, size=(n, ni - nf))
        cam = np.concatenate([out["pose"], intr], 1)
        proj = cf.reproject_funcs[mid](torch.from_numpy(X), torch.from_numpy(cam), torch.from_numpy(pp))
        out[f"intr_{mid}"] = intr
        out[f"proj_{mid}"] = proj.numpy()
    for mid in (7, 10):
        try:
            cf.reproject_funcs[mid](torch.from_numpy(X), torch.from_numpy(out["pose"]), torch.from_numpy(pp))
            raised = False
        except NotImplementedError:
            raised = True
        out[f"raises_{mid}"] = np.array(raised)

    # pairwise_cost (GP)
    c = rng.normal(size=(n, 3)) * 10
    Xg = rng.normal(size=(n, 3)) * 10
    s = rng.uniform(0.1, 2.0, size=(n, 1))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    cal = rng.uniform(size=n) < 0.7
    out["gp_c"], out["gp_X"], out["gp_s"], out["gp_d"], out["gp_cal"] = c, Xg, s, d, cal
    out["gp_r"] = cf.pairwise_cost(torch.from_numpy(Xg), torch.from_numpy(c), torch.from_numpy(s),
                                   torch.from_numpy(d), torch.from_numpy(cal)).numpy()

    # camera-model tables
    for m in CameraModelId:
        if m.value < 0:
            continue
        info = get_camera_model_info(m)
        out[f"info_{m.value}_pp"] = np.array(info["pp"], dtype=np.int64)
        out[f"info_{m.value}_optimize"] = np.array(info["optimize"], dtype=np.int64)
        out[f"info_{m.value}_num_params"] = np.array(info["num_params"], dtype=np.int64)
        out[f"info_{m.value}_name"] = np.array(info["name"])

    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_cost_functions.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: v.shape for k, v in out.items() if k.startswith("proj")})


if __name__ == "__main__":
    main()
