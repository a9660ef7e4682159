"""Record an LM trajectory FROM THE REFERENCE'S OWN STACK: tests/golden/reference_lm_trajectory.npz.

    python tests/golden/make_lm_golden.py [--device cuda:0|cpu] [--steps 12]

Needs what this container does not have and cannot install offline: ``bae`` (HEAD of
github.com/zitongzhan/bae, README.md:67-70 of the reference) and ``pypose@bae``
(pyproject.toml:38), plus /root/reference.  Without them the script says so and exits 0 WITHOUT
writing anything -- the LM semantics of oracle/lm.py then stay "parity unpinned" (DESIGN.md section 7).

With them: the reference's ``instantsfm.processors.bundle_adjustment.TorchBA`` (unmodified) solves the
seeded scene of tests/test_ba_gpu.py::test_trajectory_matches_oracle; a recording visualizer
(``add_step`` is called after every ``optimizer.step``, bundle_adjustment.py:144-148) snapshots poses,
intrinsics and points per LM iteration.  tests/test_lm_golden_host.py replays the same scene through
oracle/ (fp64, same solver settings) and compares costs and parameters per iteration -- that pins
TrustRegion, Huber / FastTriggs, the SE3 tangent convention, the clamp / damping rule and the reject
rule in one go.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "reference_lm_trajectory.npz")
SCENE = dict(n_cam=24, n_pt=1500, n_obs=8000, seed=13)
OPTIONS = {"optimize_poses": True, "optimize_points": True, "min_num_view_per_track": 2, "thres_loss_function": 1.0,
           "function_tolerance": 0.0}


def reference_stack():
    try:
        import bae  # noqa: F401
        import pypose  # noqa: F401
    except Exception as e:   # noqa: BLE001
        return None, f"bae / pypose not importable ({e.__class__.__name__}: {e})"
    if not os.path.isdir("/root/reference/instantsfm"):
        return None, "/root/reference is absent"
    sys.path.insert(0, "/root/reference")
    try:
        from instantsfm.processors.bundle_adjustment import TorchBA
        from instantsfm.scene import defs
    except Exception as e:   # noqa: BLE001
        return None, f"the reference package does not import ({e.__class__.__name__}: {e})"
    return (TorchBA, defs), None


class Recorder:
    def __init__(self):
        self.poses, self.params, self.points = [], [], []

    def add_step(self, cameras, images, tracks, name=None):
        self.poses.append(np.stack([np.asarray(im.world2cam, dtype=np.float64) for im in images], 0))
        self.params.append(np.stack([np.asarray(c.params, dtype=np.float64) for c in cameras], 0))
        self.points.append(np.stack([np.asarray(t.xyz, dtype=np.float64) for t in tracks.values()], 0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--steps", type=int, default=12)
    args = ap.parse_args()
    stack, why = reference_stack()
    if stack is None:
        print(f"make_lm_golden: nothing written -- {why}.  The LM semantics stay parity-unpinned.")
        return 0
    TorchBA, defs = stack
    from instantsfm_b200.synthetic import ba_arrays_to_scene, make_ba_problem
    a = make_ba_problem(SCENE["n_cam"], SCENE["n_pt"], SCENE["n_obs"], seed=SCENE["seed"])
    cams, imgs, trks = ba_arrays_to_scene(a)
    cameras = [defs.Camera(id=c.id, model_id=defs.CameraModelId(c.model_id.value), params=list(c.params),
                           has_prior_focal_length=True) for c in cams]
    images = []
    for im in imgs:
        r = defs.Image(id=im.id, cam_id=im.cam_id, is_registered=True, world2cam=np.array(im.world2cam))
        r.features = np.array(im.features)
        images.append(r)
    tracks = {}
    for k, t in trks.items():
        r = defs.Track()
        r.id, r.xyz, r.observations = k, np.array(t.xyz), np.array(t.observations)
        tracks[k] = r
    rec = Recorder()
    opts = dict(OPTIONS, max_num_iterations=args.steps)
    TorchBA(visualizer=rec, device=args.device).Solve(cameras, images, tracks, opts)
    rec.add_step(cameras, images, tracks)   # the state after the final update() (bundle_adjustment.py:152-154)
    import bae
    import pypose
    np.savez_compressed(OUT, poses=np.stack(rec.poses, 0), params=np.stack(rec.params, 0), points=np.stack(rec.points, 0),
                        scene=np.array([SCENE["n_cam"], SCENE["n_pt"], SCENE["n_obs"], SCENE["seed"]]),
                        steps=np.array(args.steps), versions=np.array([getattr(bae, "__version__", "?"), getattr(pypose, "__version__", "?")]))
    print("wrote", OUT, "iterations recorded:", len(rec.poses))
    return 0


if __name__ == "__main__":
    sys.exit(main())
