"""Shared host-side helpers of the drop-in processors."""
import numpy as np

WINDOW = 4  # bundle_adjustment.py:128 / global_positioning.py:172


def device_index(device):
    """'cuda:0' / 'cuda' / torch.device / int -> CUDA ordinal.  There is no CPU path."""
    if isinstance(device, int):
        return device
    s = str(device)
    if not s.startswith("cuda"):
        raise RuntimeError(f"instantsfm_b200 runs on CUDA devices only (got device={device!r}); "
                           "the CPU restatement lives in oracle/ and is test infrastructure")
    return int(s.split(":")[1]) if ":" in s else 0


def should_stop(history, function_tolerance, identical_test):
    """Windowed relative-improvement stop rule, bundle_adjustment.py:134-141."""
    if len(history) < 2 * WINDOW:
        return False
    recent = np.mean(history[-WINDOW:])
    previous = np.mean(history[-2 * WINDOW:-WINDOW])
    improvement = (previous - recent) / previous
    if abs(improvement) < function_tolerance:
        return True
    return identical_test and history[-1] == history[-2]


def concat_features(images, attr):
    """One [sum n_i, k] array of every image's `attr` rows + per-image offsets, so that
    (image_id, feature_id) pairs can be looked up with a single fancy index."""
    counts = np.array([len(getattr(img, attr)) for img in images], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(counts)])
    rows = [np.asarray(getattr(img, attr), dtype=np.float64) for img in images if len(getattr(img, attr))]
    table = np.concatenate(rows, axis=0) if rows else np.zeros((0, 2))
    return table, offsets


def flatten_observations(tracks, track_ids):
    """Concatenated (image_id, feature_id, position-in-track_ids) of the given tracks, in the
    order of the reference's nested loops (bundle_adjustment.py:88-96)."""
    obs_list = [tracks[t].observations for t in track_ids]
    lengths = np.fromiter(map(len, obs_list), dtype=np.int64, count=len(obs_list))
    if lengths.sum() == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z, z
    try:       # the common case: every track holds an [k, 2] integer array -- one C-level concatenate
        obs = np.concatenate(obs_list, axis=0)
        if obs.ndim != 2 or obs.shape[1] != 2:
            raise ValueError
    except ValueError:   # lists of tuples, empty lists, 1-d arrays: normalise per track
        obs = np.concatenate([np.asarray(o).reshape(-1, 2) for o in obs_list], axis=0)
    obs = obs.astype(np.int64, copy=False)
    which = np.repeat(np.arange(len(track_ids), dtype=np.int64), lengths)
    return obs[:, 0], obs[:, 1], which


CAMERA_ROW = 16   # ISFM_CAMERA_ROW, include/isfm_b200.h
# model value -> (focal idx, pp idx, k idx (up to 6), p idx, omega idx, sx idx) into Camera.params:
# the attributes Camera.set_params derives (scene/defs.py:177-237)
_PARAM_LAYOUT = {
    0: ([0, 0], [1, 2], [], [], None, []),
    1: ([0, 1], [2, 3], [], [], None, []),
    2: ([0, 0], [1, 2], [3], [], None, []),
    3: ([0, 0], [1, 2], [3, 4], [], None, []),
    4: ([0, 1], [2, 3], [4, 5], [6, 7], None, []),
    5: ([0, 1], [2, 3], [4, 5, 6, 7], [], None, []),
    6: ([0, 1], [2, 3], [4, 5, 8, 9, 10, 11], [6, 7], None, []),
    7: ([0, 1], [2, 3], [], [], 4, []),
    8: ([0, 0], [1, 2], [3], [], None, []),
    9: ([0, 0], [1, 2], [3, 4], [], None, []),
    10: ([0, 1], [2, 3], [4, 5, 8, 9], [6, 7], None, [10, 11]),
}


def camera_table(cameras):
    """[n_cam, CAMERA_ROW] fp64 table of the scene-layer camera maps (isfm_filter_reprojection,
    isfm_undistort_features): model id, focal lengths, principal point, k[0..5], p, omega, sx."""
    table = np.zeros((len(cameras), CAMERA_ROW), dtype=np.float64)
    for i, cam in enumerate(cameras):
        model = cam.model_id.value if hasattr(cam.model_id, "value") else int(cam.model_id)
        if model not in _PARAM_LAYOUT:
            raise NotImplementedError
        f, pp, k, p, omega, sx = _PARAM_LAYOUT[model]
        prm = np.asarray(cam.params, dtype=np.float64)
        table[i, 0] = model
        table[i, 1:3] = prm[f]
        table[i, 3:5] = prm[pp]
        table[i, 5:5 + len(k)] = prm[k]
        if p:
            table[i, 11:13] = prm[p]
        if omega is not None:
            table[i, 13] = prm[omega]
        if sx:
            table[i, 14:16] = prm[sx]
    return table
