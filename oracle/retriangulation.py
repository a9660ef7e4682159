"""TEST INFRASTRUCTURE ONLY -- CPU restatement of ``complete_tracks``
(instantsfm/processors/track_retriangulation.py:18-108) in torch fp64, written the reference's
way: per-observation Python loops for the candidates, the oracle's ``reproject`` (pinned against
the reference's own reproject_* by tests/golden/reference_cost_functions.npz) and
``rotate_quat`` for the cheirality test.  The function itself cannot be imported from the
reference here (its module pulls in cv2 and bae at import time): parity of this row is pinned at
the level of the projection functions, not of the whole function.
"""
import numpy as np
import torch
from scipy.spatial.transform import Rotation

from .camera_models import PP_INDICES, reproject
from .lie import rotate_quat

EPSILON = 1e-7


def complete_tracks(cameras, images, tracks, tracks_orig, options):
    thr = options['complete_max_reproj_error']
    model = cameras[0].model_id
    model = model.value if hasattr(model, "value") else int(model)
    track_id2idx = {tid: i for i, tid in enumerate(tracks.keys())}
    idx2id = {i: tid for tid, i in track_id2idx.items()}
    feats, tidx, info = [], [], []
    for tid, obs in tracks_orig.items():                         # :50-57
        if tid not in track_id2idx:
            continue
        for img_id, feat_id in obs:
            feats.append(images[img_id].features[feat_id]); tidx.append(track_id2idx[tid]); info.append((img_id, feat_id))
    if not info:
        return 0
    observed = torch.tensor(np.array(feats), dtype=torch.float64)
    pidx = torch.tensor(tidx, dtype=torch.int64)
    info = np.array(info, dtype=np.int32)
    rows = [np.concatenate([img.world2cam[:3, 3], Rotation.from_matrix(img.world2cam[:3, :3]).as_quat(),
                            np.asarray(cameras[img.cam_id].params, dtype=np.float64)]) for img in images]   # :63-66
    cam = torch.tensor(np.stack(rows, 0), dtype=torch.float64)
    pts = torch.tensor(np.stack([t.xyz for t in tracks.values()], 0), dtype=torch.float64)[pidx]
    cam = cam[torch.tensor(info[:, 0].astype(np.int64))]
    pp_idx = [i + 7 for i in PP_INDICES[model]]
    rest = [i for i in range(cam.shape[1]) if i not in pp_idx]
    pps, cam = cam[:, pp_idx], cam[:, rest]
    valid = rotate_quat(pts, cam[:, :7])[:, 2] > EPSILON          # :81-82
    err = torch.norm(reproject(model, pts, cam, pps) - observed, dim=-1)   # :84-86
    passing = ((err <= thr) & valid).numpy()                      # :89-90
    info_p, pidx_p = info[passing], pidx.numpy()[passing]
    if pidx_p.size == 0:
        return 0
    split = [0] + (np.flatnonzero(np.diff(pidx_p)) + 1).tolist() + [pidx_p.size]
    n = 0
    for a, b in zip(split[:-1], split[1:]):                      # :101-106
        track = tracks[idx2id[int(pidx_p[a])]]
        n += abs((b - a) - track.observations.shape[0])
        track.observations = info_p[a:b]
    return n
