"""Drop-in for instantsfm/processors/track_filter.py: same function names,
arguments, in-place mutation of ``tracks`` and return values; the per-observation tests run as
CUDA kernels (include/isfm_b200.h: isfm_filter_observations, isfm_filter_triangulation_angle)
in fp64 like the reference's numpy, and the O(N_obs) Python loops that gather
``features_undist`` (:35-43) become one fancy index.

Reference quirks kept on purpose, because callers see them:
* ``FilterTracksByAngle`` returns ``tracks`` (:24), the others return a count;
* ``FilterTracksByReprojectionNormalized`` advances ``count`` BEFORE it tests
  ``np.all(valid_mask[count:count + obs_count])`` (:59-62), i.e. a track is counted as filtered
  when the window of mask entries that FOLLOWS it (sized like itself) holds a rejection --
  the returned / printed counter reproduces that;
* ``FilterTracksTriangulationAngle`` removes single-view and empty tracks (np.all of the 1x1 /
  0x0 dot-product matrix).
There is no CPU path: without the CUDA library / a GPU these functions raise.
"""
import numpy as np

from .. import _lib
from ._common import camera_table, concat_features, flatten_observations

EPSILON = 1e-10
_ANGLE, _REPROJ_NORM = 0, 1


def _ptr(a):
    return a.ctypes.data


def _flatten(images, tracks):
    keys = list(tracks.keys())
    image_id, feature_id, which = flatten_observations(tracks, keys)
    lengths = np.array([len(tracks[k].observations) for k in keys], dtype=np.int64)
    return keys, image_id, feature_id, which, lengths


def _observation_mask(mode, images, tracks, threshold):
    keys, image_id, feature_id, which, lengths = _flatten(images, tracks)
    n_obs = int(image_id.size)
    valid = np.zeros(n_obs, dtype=np.uint8)
    if n_obs:
        table, offsets = concat_features(images, "features_undist")
        feats = np.ascontiguousarray(table[offsets[image_id] + feature_id].reshape(-1, 3), dtype=np.float64)
        world2cam = np.ascontiguousarray(np.stack([np.asarray(img.world2cam, dtype=np.float64) for img in images], 0))
        xyz = np.ascontiguousarray(np.stack([np.asarray(tracks[k].xyz, dtype=np.float64) for k in keys], 0))
        ids = np.ascontiguousarray(image_id, dtype=np.int32)
        tix = np.ascontiguousarray(which, dtype=np.int32)
        _lib.check(_lib.load().isfm_filter_observations(mode, n_obs, len(images), len(keys), _ptr(world2cam), _ptr(xyz),
                                                        _ptr(feats), _ptr(ids), _ptr(tix), float(threshold), _ptr(valid), None))
    return keys, lengths, valid.astype(bool)


def _apply_mask_with_lookahead_counter(tracks, keys, lengths, valid):
    """Drop the rejected observations in place; the counter reproduces the reference's loop
    (track_filter.py:57-62 and :107-112), which advances ``count`` BEFORE testing
    ``np.all(valid_mask[count:count + obs_count])`` -- i.e. it looks at the NEXT window."""
    starts = np.concatenate([[0], np.cumsum(lengths)])
    invalid_prefix = np.concatenate([[0], np.cumsum(~valid)])
    n = valid.size
    counter = 0
    for i, k in enumerate(keys):
        lo, hi = int(starts[i]), int(starts[i + 1])
        if invalid_prefix[hi] != invalid_prefix[lo]:
            track = tracks[k]
            track.observations = np.asarray(track.observations)[valid[lo:hi]]
        w_lo, w_hi = min(hi, n), min(hi + (hi - lo), n)
        if invalid_prefix[w_hi] != invalid_prefix[w_lo]:
            counter += 1
    return counter


def FilterTracksByAngle(cameras, images, tracks, max_angle_error):
    """track_filter.py:5-24."""
    thres = np.cos(np.deg2rad(max_angle_error))
    keys, lengths, valid = _observation_mask(_ANGLE, images, tracks, thres)
    counter = 0
    starts = np.concatenate([[0], np.cumsum(lengths)])
    prefix = np.concatenate([[0], np.cumsum(valid)])
    kept = prefix[starts[1:]] - prefix[starts[:-1]]
    for i in np.flatnonzero((kept != lengths) & (lengths > 0)).tolist():
        track = tracks[keys[i]]
        track.observations = np.asarray(track.observations)[np.flatnonzero(valid[starts[i]:starts[i + 1]])]
        counter += 1
    print(f'Filtered {counter} / {len(tracks)} tracks by angle error')
    return tracks


def FilterTracksByReprojectionNormalized(cameras, images, tracks, max_reprojection_error):
    """track_filter.py:26-66."""
    keys, lengths, valid = _observation_mask(_REPROJ_NORM, images, tracks, max_reprojection_error)
    counter = _apply_mask_with_lookahead_counter(tracks, keys, lengths, valid)
    print(f'Filtered {counter} / {len(tracks)} tracks by reprojection error')
    return counter


def FilterTracksByReprojection(cameras, images, tracks, max_reprojection_error):
    """track_filter.py:68-114 (pixel space, through Camera.cam2img; called by filter_points,
    track_retriangulation.py:200-204).  All eleven camera models, one model per camera.  The
    returned counter has the same look-ahead quirk as FilterTracksByReprojectionNormalized
    (:109-112)."""
    keys, image_id, feature_id, which, lengths = _flatten(images, tracks)
    n_obs = int(image_id.size)
    valid = np.zeros(n_obs, dtype=np.uint8)
    if n_obs:
        table, offsets = concat_features(images, "features")
        feats = np.ascontiguousarray(table[offsets[image_id] + feature_id].reshape(-1, 2), dtype=np.float64)
        world2cam = np.ascontiguousarray(np.stack([np.asarray(img.world2cam, dtype=np.float64) for img in images], 0))
        image_cam = np.ascontiguousarray([img.cam_id for img in images], dtype=np.int32)
        cams = camera_table(cameras)
        xyz = np.ascontiguousarray(np.stack([np.asarray(tracks[k].xyz, dtype=np.float64) for k in keys], 0))
        ids = np.ascontiguousarray(image_id, dtype=np.int32)
        tix = np.ascontiguousarray(which, dtype=np.int32)
        _lib.check(_lib.load().isfm_filter_reprojection(n_obs, len(images), len(keys), len(cameras), _ptr(world2cam), _ptr(image_cam),
                                                        _ptr(cams), _ptr(xyz), _ptr(feats), _ptr(ids), _ptr(tix),
                                                        float(max_reprojection_error), _ptr(valid), None, None))
    counter = _apply_mask_with_lookahead_counter(tracks, keys, lengths, valid.astype(bool))
    print(f'Filtered {counter} / {len(tracks)} tracks by reprojection error')
    return counter


def FilterTracksTriangulationAngle(cameras, images, tracks, min_angle):
    """track_filter.py:116-137."""
    thres = np.cos(np.deg2rad(min_angle))
    keys, image_id, _, _, lengths = _flatten(images, tracks)
    counter = 0
    if keys:
        centers = np.ascontiguousarray(np.stack([
            np.asarray(img.world2cam, dtype=np.float64)[:3, :3].T @ -np.asarray(img.world2cam, dtype=np.float64)[:3, 3]
            for img in images], 0))
        xyz = np.ascontiguousarray(np.stack([np.asarray(tracks[k].xyz, dtype=np.float64) for k in keys], 0))
        track_off = np.ascontiguousarray(np.concatenate([[0], np.cumsum(lengths)]), dtype=np.int64)
        ids = np.ascontiguousarray(image_id if image_id.size else np.zeros(1), dtype=np.int32)
        remove = np.zeros(len(keys), dtype=np.uint8)
        _lib.check(_lib.load().isfm_filter_triangulation_angle(len(keys), int(image_id.size), len(images), _ptr(track_off),
                                                               _ptr(ids), _ptr(centers), _ptr(xyz), float(thres), _ptr(remove), None))
        for i in np.flatnonzero(remove).tolist():
            del tracks[keys[i]]
            counter += 1
    print(f'Filtered {counter} / {counter + len(tracks)} tracks by too small triangulation angle')
    return counter
