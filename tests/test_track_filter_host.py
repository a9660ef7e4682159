"""Oracle of the track filters vs the golden vectors generated from the reference's own
track_filter.py (tests/golden/make_track_filter_golden.py).  CPU only."""
import copy
import os

import numpy as np
import pytest

from instantsfm_b200.synthetic import make_filter_scene
from oracle import track_filter as orc
from tests.golden.make_track_filter_golden import CASES, snapshot

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_track_filter.npz"))
WHICH = {"FilterTracksByAngle": "angle", "FilterTracksByReprojectionNormalized": "reprojection_normalized",
         "FilterTracksTriangulationAngle": "triangulation_angle"}


def check_against_golden(name, tracks, ret):
    keys, lens, obs = snapshot(tracks)
    assert np.array_equal(keys, GOLDEN[name + "/keys"])
    assert np.array_equal(lens, GOLDEN[name + "/lens"])
    assert np.array_equal(obs, GOLDEN[name + "/obs"])
    want = int(GOLDEN[name + "/ret"])
    if want >= 0:
        assert int(ret) == want


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_reproduces_reference(case):
    name, fn, thr, kw = case
    _, images, tracks = make_filter_scene(**kw)
    tracks = copy.deepcopy(tracks)
    n_before = sum(len(t.observations) for t in tracks.values())
    ret = orc.apply_filters_like_reference(images, tracks, WHICH[fn], thr)
    check_against_golden(name, tracks, ret)
    # the fixture exercises the filter: something was removed, something survived
    n_after = sum(len(t.observations) for t in tracks.values())
    assert 0 < n_after < n_before or len(tracks) < kw.get("n_trk", 300)
