"""Host logic of the track-filter and complete_tracks drop-ins WITHOUT a GPU: the C-ABI calls are
replaced by numpy stand-ins that evaluate the oracle's per-observation tests on the flattened
arrays the drop-ins pass (so what is checked is the flattening, the mask -> tracks bookkeeping,
the returned counters and the reference's quirks), against the reference-generated golden
vectors and the oracle.  The CUDA kernels themselves are covered by the -m gpu tests."""
import copy
import ctypes
import io
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

from instantsfm_b200.processors import track_filter as tf
from instantsfm_b200.processors import track_retriangulation as tr
from instantsfm_b200.synthetic import ba_arrays_to_scene, make_ba_problem, make_filter_scene
from oracle import retriangulation as orc_rt
from oracle.camera_models import reproject
from oracle.lie import rotate_quat
from tests.golden.make_track_filter_golden import CASES
from tests.test_track_filter_host import check_against_golden

EPS = 1e-10


def _arr(ptr, shape, dtype):
    n = int(np.prod(shape))
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class FakeLib:
    """numpy stand-ins with the signatures of include/isfm_b200.h (host pointers only)."""

    def isfm_filter_observations(self, mode, n_obs, n_img, n_trk, w2c, xyz, feat, ids, tix, thr, out, stream):
        M = _arr(w2c, (n_img, 4, 4), np.float64)[_arr(ids, (n_obs,), np.int32)]
        X = _arr(xyz, (n_trk, 3), np.float64)[_arr(tix, (n_obs,), np.int32)]
        f = _arr(feat, (n_obs, 3), np.float64)
        p = np.einsum("nij,nj->ni", M[:, :3, :3], X) + M[:, :3, 3]
        if mode == 0:
            with np.errstate(invalid="ignore", divide="ignore"):
                d = np.einsum("ni,ni->n", p / np.linalg.norm(p, axis=1, keepdims=True), f)
            keep = (p[:, 2] >= EPS) & (d > thr)
        else:
            e = np.linalg.norm(p[:, :2] / (p[:, 2:] + EPS) - f[:, :2] / (f[:, 2:] + EPS), axis=1)
            keep = (p[:, 2] > EPS) & (e < thr)
        _arr(out, (n_obs,), np.uint8)[:] = keep
        return 0

    def isfm_filter_triangulation_angle(self, n_trk, n_obs, n_img, off, ids, centers, xyz, thr, out, stream):
        off = _arr(off, (n_trk + 1,), np.int64)
        ids = _arr(ids, (max(n_obs, 1),), np.int32)
        C = _arr(centers, (n_img, 3), np.float64)
        X = _arr(xyz, (n_trk, 3), np.float64)
        rem = _arr(out, (n_trk,), np.uint8)
        for t in range(n_trk):
            v = X[t] - C[ids[off[t]:off[t + 1]]]
            d = v / (np.linalg.norm(v, axis=1, keepdims=True) + EPS)
            rem[t] = np.all(d @ d.T > thr)
        return 0

    def isfm_reprojection_test(self, model, n_obs, n_cam, n_pt, cam, pp, pts, obs, ci, pi, max_err, min_depth, out, err, stream):
        ni = {0: 1, 1: 2, 2: 2, 3: 3, 4: 6, 5: 6, 6: 10, 8: 2, 9: 3}[model]
        ci_ = torch.from_numpy(_arr(ci, (n_obs,), np.int32).astype(np.int64))
        pi_ = torch.from_numpy(_arr(pi, (n_obs,), np.int32).astype(np.int64))
        cam_ = torch.from_numpy(_arr(cam, (n_cam, 7 + ni), np.float64).copy())[ci_]
        pp_ = torch.from_numpy(_arr(pp, (n_cam, 2), np.float64).copy())[ci_]
        X = torch.from_numpy(_arr(pts, (n_pt, 3), np.float64).copy())[pi_]
        o = torch.from_numpy(_arr(obs, (n_obs, 2), np.float64).copy())
        e = torch.norm(reproject(model, X, cam_, pp_) - o, dim=-1)
        z = rotate_quat(X, cam_[:, :7])[:, 2]
        _arr(out, (n_obs,), np.uint8)[:] = ((e <= max_err) & (z > min_depth)).numpy()
        return 0

    def isfm_last_error(self):
        return b""


@pytest.fixture
def fake_lib(monkeypatch):
    lib = FakeLib()
    monkeypatch.setattr(tf._lib, "load", lambda: lib)
    return lib


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_track_filter_host_logic_matches_reference_golden(case, fake_lib):
    name, fn, thr, kw = case
    _, images, tracks = make_filter_scene(**kw)
    tracks = copy.deepcopy(tracks)
    with redirect_stdout(io.StringIO()):
        ret = getattr(tf, fn)([], images, tracks, thr)
    if fn == "FilterTracksByAngle":
        assert ret is tracks
        ret = -1
    check_against_golden(name, tracks, ret)


@pytest.mark.parametrize("model_id", [0, 3, 6])
def test_complete_tracks_host_logic_matches_oracle(model_id, fake_lib):
    a = make_ba_problem(10, 300, 1500, seed=60 + model_id, model_id=model_id)
    cameras, images, full = ba_arrays_to_scene(a)
    tracks_orig = {tid: t.observations.copy() for tid, t in full.items()}
    tracks = {tid: copy.deepcopy(t) for k, (tid, t) in enumerate(full.items()) if k % 7}
    for t in tracks.values():
        t.observations = t.observations[:-1] if len(t.observations) > 2 else t.observations
    x, y = copy.deepcopy(tracks), copy.deepcopy(tracks)
    nx = tr.complete_tracks(cameras, images, x, tracks_orig, {'complete_max_reproj_error': 12.0})
    ny = orc_rt.complete_tracks(cameras, images, y, tracks_orig, {'complete_max_reproj_error': 12.0})
    assert nx == ny and nx > 0
    for tid in x:
        assert np.array_equal(np.asarray(x[tid].observations), np.asarray(y[tid].observations))
