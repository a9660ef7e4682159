"""Generate tests/golden/reference_camera_ops.npz FROM THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference and cv2):

    python tests/golden/make_camera_ops_golden.py

Loads /root/reference/instantsfm/scene/defs.py (Camera.cam2img / img2cam; imports cv2, scipy),
processors/image_undistortion.py and processors/track_filter.py UNMODIFIED by file path and runs

* ``Camera.cam2img`` and ``Camera.img2cam`` of all eleven camera models on seeded inputs,
* ``UndistortImages`` and ``FilterTracksByReprojection`` on seeded scenes
  (tests/helpers.make_pixel_scene; the scene's cameras are re-created as reference Camera objects),

storing inputs' checksums and the reference's outputs.  cv2.undistortPoints underneath is the
installed opencv-python (version stored in the file).
"""
import copy
import importlib.util
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.helpers import PIXEL_MODEL_PARAMS, make_pixel_scene  # noqa: E402

FILTER_CASES = [  # (name, threshold, scene kwargs)
    ("radial_4px", 4.0, dict(models=(3,), seed=21)),
    ("opencv_2px", 2.0, dict(models=(4,), seed=22, n_img=9, n_trk=300)),
    ("mixed_all_models_3px", 3.0, dict(models=tuple(range(11)), seed=23, n_img=22, n_trk=500, mean_len=6.0)),
    ("fisheye_1px", 1.0, dict(models=(5, 9), seed=24, n_img=10, n_trk=250)),
]
UNDISTORT_CASES = [("undist_all_models", dict(models=tuple(range(11)), seed=31, n_img=22, n_trk=120)),
                   ("undist_radial", dict(models=(3,), seed=32, n_img=6, n_trk=200))]


def point_inputs(model, n=64, seed=0):
    """Seeded camera-frame points (some behind / on the axis) and pixels for one model."""
    rng = np.random.default_rng(1000 + 17 * model + seed)
    uvw = np.column_stack([rng.normal(0, 0.6, n), rng.normal(0, 0.5, n), rng.uniform(0.8, 3.0, n)])
    uvw[0] = [0.0, 0.0, 2.0]          # on the optical axis
    uvw[1] = [0.3, -0.2, -1.5]        # behind the camera
    uvw[2] = [1e-9, -1e-9, 1.0]       # r below the fisheye clip
    xy = np.column_stack([rng.uniform(40, 1240, n), rng.uniform(30, 930, n)])
    xy[0] = [640.0, 480.0]            # the principal point (theta = 0 for the fisheye models)
    return uvw, xy


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join("/root/reference/instantsfm", rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def to_reference(defs, cameras, images):
    rc = [defs.Camera(id=c.id, model_id=defs.CameraModelId(c.model_id.value), width=c.width, height=c.height, params=list(c.params))
          for c in cameras]
    ri = []
    for im in images:
        r = defs.Image(id=im.id, cam_id=im.cam_id, is_registered=True, world2cam=np.array(im.world2cam))
        r.features = np.array(im.features)
        ri.append(r)
    return rc, ri


def snapshot(tracks):
    keys = np.array(list(tracks.keys()), dtype=np.int64)
    lens = np.array([len(tracks[k].observations) for k in keys], dtype=np.int64)
    obs = (np.concatenate([np.asarray(tracks[k].observations).reshape(-1, 2) for k in keys], 0)
           if len(keys) else np.zeros((0, 2), np.int64))
    return keys, lens, obs.astype(np.int64)


def main():
    import cv2
    defs = load("ref_defs", "scene/defs.py")
    undist = load("ref_undist", "processors/image_undistortion.py")
    tfilter = load("ref_track_filter", "processors/track_filter.py")
    out = {"cv2_version": np.array(cv2.__version__)}
    for m in range(11):
        cam = defs.Camera(id=0, model_id=defs.CameraModelId(m), width=1280, height=960, params=list(PIXEL_MODEL_PARAMS[m]))
        uvw, xy = point_inputs(m)
        with np.errstate(all="ignore"):
            out[f"cam2img/{m}"] = cam.cam2img(uvw.copy())
            out[f"img2cam/{m}"] = cam.img2cam(xy.copy())
        print("model", m, "cam2img", np.nanmax(np.abs(out[f"cam2img/{m}"])), "img2cam", np.nanmax(np.abs(out[f"img2cam/{m}"])))
    for name, kw in UNDISTORT_CASES:
        cameras, images, _ = make_pixel_scene(**kw)
        rc, ri = to_reference(defs, cameras, images)
        with np.errstate(all="ignore"):
            undist.UndistortImages(rc, ri)
        out[name + "/bearings"] = np.concatenate([im.features_undist for im in ri], 0)
        out[name + "/n_feat"] = np.array([len(im.features) for im in ri])
        print(name, out[name + "/bearings"].shape)
    for name, thr, kw in FILTER_CASES:
        cameras, images, tracks = make_pixel_scene(**kw)
        rc, ri = to_reference(defs, cameras, images)
        tracks = copy.deepcopy(tracks)
        n_before = sum(len(t.observations) for t in tracks.values())
        with redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            ret = tfilter.FilterTracksByReprojection(rc, ri, tracks, thr)
        keys, lens, obs = snapshot(tracks)
        out[name + "/keys"], out[name + "/lens"], out[name + "/obs"], out[name + "/ret"] = keys, lens, obs, np.array(int(ret))
        print(name, "obs", n_before, "->", int(lens.sum()), "ret", int(ret))
    np.savez_compressed(os.path.join(HERE, "reference_camera_ops.npz"), **out)


if __name__ == "__main__":
    main()
