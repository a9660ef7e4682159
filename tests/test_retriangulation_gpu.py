"""complete_tracks (batched reprojection test through the C ABI, fp64) vs the oracle's
restatement of track_retriangulation.py:18-108, for several camera models."""
import copy
import io
from contextlib import redirect_stdout

import numpy as np
import pytest

from instantsfm_b200.processors.track_retriangulation import complete_and_merge_tracks, complete_tracks
from instantsfm_b200.synthetic import ba_arrays_to_scene, make_ba_problem
from oracle import retriangulation as orc

pytestmark = pytest.mark.gpu


def make_case(model_id, seed, n_cam=14, n_pt=700, n_obs=3600):
    a = make_ba_problem(n_cam, n_pt, n_obs, seed=seed, model_id=model_id)
    cameras, images, full = ba_arrays_to_scene(a)
    rng = np.random.default_rng(seed)
    tracks_orig = {tid: t.observations.copy() for tid, t in full.items()}
    tracks = {}
    for k, (tid, t) in enumerate(full.items()):
        if k % 11 == 0:
            continue                                   # track absent from the reconstruction: its candidates are skipped
        t = copy.deepcopy(t)
        keep = rng.random(len(t.observations)) < 0.7   # the reconstruction holds a subset of the candidates
        keep[0] = True
        t.observations = t.observations[keep]
        if k % 17 == 0:
            t.xyz = t.xyz + rng.normal(0, 5.0, 3)      # badly triangulated: nothing passes, track left untouched
        tracks[tid] = t
    tracks_orig[999999] = np.zeros((0, 2), dtype=np.int64)   # candidate list of a track that does not exist
    return cameras, images, tracks, tracks_orig


@pytest.mark.parametrize("model_id", [0, 1, 2, 3, 4, 5, 6, 8, 9])
def test_complete_tracks_matches_oracle(model_id):
    cameras, images, tracks, tracks_orig = make_case(model_id, seed=40 + model_id)
    opts = {'complete_max_reproj_error': 12.0}
    a, b = copy.deepcopy(tracks), copy.deepcopy(tracks)
    na = complete_tracks(cameras, images, a, tracks_orig, opts)
    nb = orc.complete_tracks(cameras, images, b, tracks_orig, opts)
    assert na == nb and na > 0
    changed = 0
    for tid in a:
        assert np.array_equal(np.asarray(a[tid].observations), np.asarray(b[tid].observations)), tid
        changed += len(a[tid].observations) != len(tracks[tid].observations)
    assert changed > 0


def test_threshold_monotone_and_wrapper():
    cameras, images, tracks, tracks_orig = make_case(3, seed=77, n_cam=30, n_pt=20000, n_obs=110000)
    loose, tight = copy.deepcopy(tracks), copy.deepcopy(tracks)
    with redirect_stdout(io.StringIO()):
        complete_and_merge_tracks(cameras, images, loose, tracks_orig, {'complete_max_reproj_error': 40.0})
        complete_and_merge_tracks(cameras, images, tight, tracks_orig, {'complete_max_reproj_error': 4.0})
    n_loose = sum(len(t.observations) for t in loose.values())
    n_tight = sum(len(t.observations) for t in tight.values())
    n_cand = sum(len(tracks_orig[t]) for t in tracks)
    assert n_tight < n_loose <= n_cand
    # a second pass with the same threshold changes nothing
    again = copy.deepcopy(loose)
    assert complete_tracks(cameras, images, again, tracks_orig, {'complete_max_reproj_error': 40.0}) == 0


def test_unsupported_model_raises():
    from instantsfm_b200.scene.defs import CameraModelId
    cameras, images, tracks, tracks_orig = make_case(3, seed=5)
    cameras[0].model_id = CameraModelId.FOV
    with pytest.raises(NotImplementedError):
        complete_tracks(cameras, images, tracks, tracks_orig, {'complete_max_reproj_error': 4.0})
