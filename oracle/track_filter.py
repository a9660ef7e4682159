"""TEST INFRASTRUCTURE ONLY -- CPU restatement of instantsfm/processors/track_filter.py
(the three filters the global mapper calls, global_mapper.py:105,118,123-124,144-145).

Plain per-observation Python loops in numpy fp64, written from the reference's formulas, for
small cases.  PINNED: tests/golden/reference_track_filter.npz was produced by importing the
reference's own track_filter.py unmodified (tests/golden/make_track_filter_golden.py) and this
module reproduces it exactly (tests/test_track_filter_host.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package.
"""
import numpy as np

EPSILON = 1e-10  # track_filter.py:3


def observation_mask_angle(images, tracks, max_angle_error):
    """track_filter.py:5-24 -> list of per-track boolean masks (True = observation kept)."""
    thres = np.cos(np.deg2rad(max_angle_error))
    out = {}
    for tid, track in tracks.items():
        m = np.zeros(len(track.observations), dtype=bool)
        for idx, (image_id, feature_id) in enumerate(track.observations):
            image = images[image_id]
            pt = image.world2cam[:3, :3] @ track.xyz + image.world2cam[:3, 3]          # :12
            if pt[2] < EPSILON:                                                        # :13
                continue
            pt = pt / np.linalg.norm(pt)                                               # :15
            m[idx] = np.dot(pt, image.features_undist[feature_id]) > thres             # :16
        out[tid] = m
    return out


def observation_mask_reprojection_normalized(images, tracks, max_reprojection_error):
    """track_filter.py:26-55 -> per-track boolean masks."""
    out = {}
    for tid, track in tracks.items():
        m = np.zeros(len(track.observations), dtype=bool)
        xyz1 = np.append(track.xyz, 1.0)
        for idx, (image_id, feature_id) in enumerate(track.observations):
            image = images[image_id]
            f = image.features_undist[feature_id]
            fu = f[:2] / (f[2:] + EPSILON)                                             # :48
            pt = (image.world2cam @ xyz1)[:3]                                          # :50-51
            reproj = pt[:2] / (pt[2:] + EPSILON)                                       # :53
            m[idx] = (pt[2] > EPSILON) and (np.linalg.norm(reproj - fu) < max_reprojection_error)   # :52,54-55
        out[tid] = m
    return out


def reprojection_counter(masks):
    """The counter the reference returns (:57-64): it tests the window of the flattened mask
    that FOLLOWS each track (``count`` is advanced before the test)."""
    flat = np.concatenate([masks[t] for t in masks]) if masks else np.zeros(0, bool)
    count = counter = 0
    for t in masks:
        n = len(masks[t])
        count += n
        if not np.all(flat[count:count + n]):
            counter += 1
    return counter


def triangulation_remove(images, tracks, min_angle):
    """track_filter.py:116-137 -> {track_id: True if the reference deletes it}."""
    thres = np.cos(np.deg2rad(min_angle))
    centers = np.array([img.world2cam[:3, :3].T @ -img.world2cam[:3, 3] for img in images])   # :119, defs.py:35
    out = {}
    for tid, track in tracks.items():
        ids = np.unique(np.asarray(track.observations).reshape(-1, 2)[:, 0]).astype(np.int64)  # :125
        v = track.xyz - centers[ids]                                                   # :126
        d = v / (np.linalg.norm(v, axis=1, keepdims=True) + EPSILON)                    # :127-128
        out[tid] = bool(np.all(d @ d.T > thres))                                       # :129-132
    return out


def apply_filters_like_reference(images, tracks, which, threshold):
    """Mutates ``tracks`` exactly as the reference function does; returns what it returns."""
    if which == "angle":
        masks = observation_mask_angle(images, tracks, threshold)
        for tid, m in masks.items():
            if not m.all():
                tracks[tid].observations = np.asarray(tracks[tid].observations)[np.flatnonzero(m)]
        return tracks
    if which == "reprojection_normalized":
        masks = observation_mask_reprojection_normalized(images, tracks, threshold)
        for tid, m in masks.items():
            tracks[tid].observations = np.asarray(tracks[tid].observations)[m]
        return reprojection_counter(masks)
    if which == "triangulation_angle":
        rem = triangulation_remove(images, tracks, threshold)
        n = 0
        for tid, r in rem.items():
            if r:
                del tracks[tid]
                n += 1
        return n
    raise ValueError(which)
