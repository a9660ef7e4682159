"""Generate tests/golden/reference_track_filter.npz FROM THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_track_filter_golden.py

instantsfm/processors/track_filter.py imports numpy only, so it is loaded UNMODIFIED (by file
path -- the package __init__ chain pulls in cv2 / pyceres).  The three filters the global mapper
calls are run on seeded scenes from instantsfm_b200.synthetic.make_filter_scene (duck-typed
Image / Track objects); the file stores, per case, the surviving observations of every track,
the surviving track ids and the value each function returned.
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
from instantsfm_b200.synthetic import make_filter_scene  # noqa: E402

CASES = [  # (name, function, threshold, scene kwargs)
    ("angle_1deg", "FilterTracksByAngle", 1.0, dict(seed=11)),
    ("angle_5deg", "FilterTracksByAngle", 5.0, dict(seed=12, n_img=9, n_trk=200)),
    ("reproj_1e-2", "FilterTracksByReprojectionNormalized", 1e-2, dict(seed=13)),
    ("reproj_3e-2", "FilterTracksByReprojectionNormalized", 3e-2, dict(seed=14, n_img=20, n_trk=500, mean_len=6.0)),
    ("tri_1deg", "FilterTracksTriangulationAngle", 1.0, dict(seed=15)),
    ("tri_0.2deg", "FilterTracksTriangulationAngle", 0.2, dict(seed=16, n_img=30, n_trk=400, mean_len=8.0)),
]


def snapshot(tracks):
    keys = np.array(list(tracks.keys()), dtype=np.int64)
    lens = np.array([len(tracks[k].observations) for k in keys], dtype=np.int64)
    obs = (np.concatenate([np.asarray(tracks[k].observations).reshape(-1, 2) for k in keys], 0)
           if len(keys) else np.zeros((0, 2), np.int64))
    return keys, lens, obs.astype(np.int64)


def main():
    spec = importlib.util.spec_from_file_location("ref_track_filter", "/root/reference/instantsfm/processors/track_filter.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    for name, fn, thr, kw in CASES:
        cameras, images, tracks = make_filter_scene(**kw)
        tracks = copy.deepcopy(tracks)
        with redirect_stdout(io.StringIO()):
            ret = getattr(ref, fn)(cameras, images, tracks, thr)
        keys, lens, obs = snapshot(tracks)
        out[name + "/keys"], out[name + "/lens"], out[name + "/obs"] = keys, lens, obs
        out[name + "/ret"] = np.array(-1 if isinstance(ret, dict) else int(ret))
        print(name, "tracks", len(keys), "obs", int(lens.sum()), "ret", out[name + "/ret"])
    np.savez_compressed(os.path.join(HERE, "reference_track_filter.npz"), **out)


if __name__ == "__main__":
    main()
