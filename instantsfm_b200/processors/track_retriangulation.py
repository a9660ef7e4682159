"""Drop-in for ``complete_tracks`` / ``complete_and_merge_tracks`` of
instantsfm/processors/track_retriangulation.py:18-108,205-213: same
arguments, same in-place mutation of ``tracks`` and the same returned count.  The batched
reprojection test (:81-91) runs as a CUDA kernel in fp64 (include/isfm_b200.h:
isfm_reprojection_test) with the camera models of the BA path; the per-observation Python loops
that gather the candidates (:50-57) are one fancy index.

``filter_points`` (:200-204) and ``RetriangulateTracks`` (:215-259) sequence this function,
points-only ``TorchBA.Solve`` (``optimize_poses=False``) and the track filters
(``FilterTracksByReprojection`` in pixel space + ``FilterTracksTriangulationAngle``), all of which
run through the C ABI.  ``merge_tracks`` is unused by the reference (:209-211: "not used in the
pipeline") and not provided.  There is no CPU path.
"""
import numpy as np

from .. import _lib
from ..geometry import matrices_to_pose7
from ._common import concat_features
from .bundle_adjustment import _PP, TorchBA, _model_value
from .track_filter import FilterTracksByReprojection, FilterTracksTriangulationAngle

EPSILON = 1e-7   # track_retriangulation.py:16


def _ptr(a):
    return a.ctypes.data


def complete_tracks(cameras, images, tracks, tracks_orig, TRIANGULATOR_OPTIONS):
    """track_retriangulation.py:18-108.  ``tracks_orig``: {track_id: int array [k, 2] of
    (image_id, feature_id)} -- the candidate observations of every track."""
    reproj_threshold = TRIANGULATOR_OPTIONS['complete_max_reproj_error']
    model_value = _model_value(cameras[0].model_id)          # one model for all cameras (:36)
    if model_value not in _PP:
        raise NotImplementedError("Unsupported camera model")
    track_ids = list(tracks.keys())
    track_id2idx = {tid: i for i, tid in enumerate(track_ids)}
    cand = [(track_id2idx[tid], np.asarray(obs).reshape(-1, 2)) for tid, obs in tracks_orig.items() if tid in track_id2idx]
    cand = [(i, o) for i, o in cand if o.shape[0] > 0]
    if not cand:
        return 0
    obs_info = np.concatenate([o for _, o in cand], axis=0).astype(np.int64)                     # [n, 2]
    point_idx = np.repeat(np.array([i for i, _ in cand], dtype=np.int64), [o.shape[0] for _, o in cand])
    table, offsets = concat_features(images, "features")
    observed = np.ascontiguousarray(table[offsets[obs_info[:, 0]] + obs_info[:, 1]].reshape(-1, 2), dtype=np.float64)

    poses = matrices_to_pose7(np.stack([np.asarray(img.world2cam, dtype=np.float64) for img in images], 0))   # :63-66
    intr = np.stack([np.asarray(cameras[img.cam_id].params, dtype=np.float64) for img in images], 0)
    full = np.concatenate([poses, intr], axis=1)
    pp_indices = np.array(_PP[model_value]) + 7
    remaining = np.array([i for i in range(full.shape[1]) if i not in pp_indices])
    cam = np.ascontiguousarray(full[:, remaining])
    pps = np.ascontiguousarray(full[:, pp_indices])
    pts = np.ascontiguousarray(np.stack([np.asarray(tracks[t].xyz, dtype=np.float64) for t in track_ids], 0))
    ci = np.ascontiguousarray(obs_info[:, 0], dtype=np.int32)
    pi = np.ascontiguousarray(point_idx, dtype=np.int32)
    passing = np.zeros(ci.size, dtype=np.uint8)
    code = _lib.load().isfm_reprojection_test(model_value, ci.size, cam.shape[0], pts.shape[0], _ptr(cam), _ptr(pps), _ptr(pts),
                                              _ptr(observed), _ptr(ci), _ptr(pi), float(reproj_threshold), EPSILON,
                                              _ptr(passing), None, None)
    if code == -2:
        raise NotImplementedError("Unsupported camera model")
    _lib.check(code)
    passing = passing.astype(bool)
    obs_pass, idx_pass = obs_info[passing].astype(np.int32), point_idx[passing]
    if idx_pass.size == 0:
        return 0
    # runs of equal track index among the passing candidates (:95-99)
    split = np.concatenate([[0], np.flatnonzero(np.diff(idx_pass)) + 1, [idx_pass.size]])
    num_completed = 0
    for a, b in zip(split[:-1].tolist(), split[1:].tolist()):
        track = tracks[track_ids[int(idx_pass[a])]]
        num_completed += abs((b - a) - np.asarray(track.observations).shape[0])
        track.observations = obs_pass[a:b]
    return num_completed


def complete_and_merge_tracks(cameras, images, tracks, tracks_orig, TRIANGULATOR_OPTIONS):
    """track_retriangulation.py:205-213 (merge_tracks is disabled in the reference)."""
    num_completed_observations = complete_tracks(cameras, images, tracks, tracks_orig, TRIANGULATOR_OPTIONS)
    print('Number of completed observations:', num_completed_observations)
    return num_completed_observations


def filter_points(cameras, images, tracks, TRIANGULATOR_OPTIONS):
    """track_retriangulation.py:200-204."""
    num_filtered = FilterTracksByReprojection(cameras, images, tracks, TRIANGULATOR_OPTIONS['filter_max_reproj_error'])
    num_filtered += FilterTracksTriangulationAngle(cameras, images, tracks, TRIANGULATOR_OPTIONS['filter_min_tri_angle'])
    return num_filtered


def RetriangulateTracks(cameras, images, tracks, tracks_orig, TRIANGULATOR_OPTIONS, BUNDLE_ADJUSTER_OPTIONS, ba_factory=TorchBA):
    """track_retriangulation.py:215-259: complete the tracks, then up to ``ba_global_max_refinements``
    rounds of points-only BA -> complete -> filter, until fewer than
    ``ba_global_max_refinement_change`` of the tracks changed.  Image registration flags are saved
    and restored around the loop like the reference does (:217,257-258).  ``ba_factory`` (default:
    this package's TorchBA) builds the solver of each round."""
    image_registered = [image.is_registered for image in images]
    complete_and_merge_tracks(cameras, images, tracks, tracks_orig, TRIANGULATOR_OPTIONS)
    rounds = TRIANGULATOR_OPTIONS['ba_global_max_refinements']
    for i in range(rounds):
        print(f'Running bundle adjustment iteration {i+1} / {rounds}')
        options = dict(BUNDLE_ADJUSTER_OPTIONS)
        options['optimize_poses'] = False
        ba_factory().Solve(cameras, images, tracks, options)
        changed = abs(complete_and_merge_tracks(cameras, images, tracks, tracks_orig, TRIANGULATOR_OPTIONS))
        changed += filter_points(cameras, images, tracks, TRIANGULATOR_OPTIONS)
        if changed / len(tracks) < TRIANGULATOR_OPTIONS['ba_global_max_refinement_change']:
            break
    for image, flag in zip(images, image_registered):
        image.is_registered = flag
