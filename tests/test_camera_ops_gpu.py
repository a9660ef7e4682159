"""CUDA scene-layer camera maps (csrc/camera_ops.cu through the C ABI and the drop-in functions) vs
the golden vectors of the reference's own Camera / UndistortImages / FilterTracksByReprojection
(cv2 underneath), vs the oracle on other seeds, and size-independent properties at scale."""
import copy
import io
from contextlib import redirect_stdout

import numpy as np
import pytest

from instantsfm_b200 import _lib
from instantsfm_b200.processors import image_undistortion as iu
from instantsfm_b200.processors import track_filter as tf
from instantsfm_b200.processors import track_retriangulation as tr
from instantsfm_b200.processors._common import camera_table
from oracle import camera_ops as orc
from tests.golden.make_camera_ops_golden import FILTER_CASES, UNDISTORT_CASES, point_inputs, snapshot
from tests.helpers import PIXEL_MODEL_PARAMS, make_pixel_scene
from tests.test_camera_ops_host import GOLDEN, check_filter_golden, close

pytestmark = pytest.mark.gpu
FTOL = 1e-12


def _undistort(model, xy):
    row = orc.Intrinsics(model, PIXEL_MODEL_PARAMS[model]).row()[None].copy()
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    ci = np.zeros(len(xy), dtype=np.int32)
    out = np.zeros((len(xy), 3))
    _lib.check(_lib.load().isfm_undistort_features(len(xy), 1, row.ctypes.data, xy.ctypes.data, ci.ctypes.data, out.ctypes.data, None))
    return out


@pytest.mark.parametrize("model", range(11))
def test_img2cam_matches_reference_golden(model):
    """bearing = [img2cam(xy), 1]/norm  =>  img2cam = bearing.xy / bearing.z; golden from Camera.img2cam on cv2."""
    _, xy = point_inputs(model)
    b = _undistort(model, xy)
    with np.errstate(all="ignore"):
        close(b[:, :2] / b[:, 2:], GOLDEN[f"img2cam/{model}"], 1e-11)
        assert np.all(np.abs(np.linalg.norm(b[np.isfinite(b[:, 0])], axis=1) - 1) < 1e-14)


@pytest.mark.parametrize("model", range(11))
def test_cam2img_matches_reference_golden(model):
    """One observation per camera-frame point (identity pose, point = uvw): the error the kernel
    reports against a zero feature IS ||cam2img(uvw)||; against the golden pixels it is ~0."""
    uvw, _ = point_inputs(model)
    want = GOLDEN[f"cam2img/{model}"]
    n = len(uvw)
    row = orc.Intrinsics(model, PIXEL_MODEL_PARAMS[model]).row()[None].copy()
    w2c = np.eye(4)[None].copy()
    ic = np.zeros(1, np.int32)
    ids = np.zeros(n, np.int32)
    tix = np.arange(n, dtype=np.int32)
    xyz = np.ascontiguousarray(uvw)
    feat = np.ascontiguousarray(np.nan_to_num(want, nan=0.0, posinf=0.0, neginf=0.0))
    valid, err = np.zeros(n, np.uint8), np.zeros(n)
    _lib.check(_lib.load().isfm_filter_reprojection(n, 1, n, 1, w2c.ctypes.data, ic.ctypes.data, row.ctypes.data, xyz.ctypes.data,
                                                    feat.ctypes.data, ids.ctypes.data, tix.ctypes.data, 1e-6, valid.ctypes.data,
                                                    err.ctypes.data, None))
    fin = np.all(np.isfinite(want), axis=1)
    assert np.all(err[fin] <= 1e-9 * np.maximum(1.0, np.abs(want[fin]).max(axis=1))), err[fin].max()
    assert np.array_equal(valid.astype(bool), fin & (uvw[:, 2] > 1e-10) & (err < 1e-6))
    assert valid.sum() >= n - 4 and not valid[1]           # the point behind the camera is rejected


@pytest.mark.parametrize("case", FILTER_CASES, ids=[c[0] for c in FILTER_CASES])
def test_filter_reprojection_matches_reference_golden(case):
    name, thr, kw = case
    cameras, images, tracks = make_pixel_scene(**kw)
    tracks = copy.deepcopy(tracks)
    with redirect_stdout(io.StringIO()):
        ret = tf.FilterTracksByReprojection(cameras, images, tracks, thr)
    check_filter_golden(name, tracks, ret)           # bit-exact masks and counter


@pytest.mark.parametrize("seed", [201, 202])
def test_filter_reprojection_matches_oracle(seed):
    cameras, images, tracks = make_pixel_scene(models=tuple(range(11)), n_img=33, n_trk=600, mean_len=5.0, seed=seed)
    a, b = copy.deepcopy(tracks), copy.deepcopy(tracks)
    with redirect_stdout(io.StringIO()):
        ra = tf.FilterTracksByReprojection(cameras, images, a, 2.5)
    rb = orc.apply_filter_reprojection(cameras, images, b, 2.5)
    ka, la, oa = snapshot(a)
    kb, lb, ob = snapshot(b)
    assert np.array_equal(ka, kb) and np.array_equal(la, lb) and np.array_equal(oa, ob) and ra == rb


@pytest.mark.parametrize("case", UNDISTORT_CASES, ids=[c[0] for c in UNDISTORT_CASES])
def test_undistort_images_matches_reference_golden(case):
    name, kw = case
    cameras, images, _ = make_pixel_scene(**kw)
    iu.UndistortImages(cameras, images)
    assert [im.features_undist.shape for im in images] == [(n, 3) for n in GOLDEN[name + "/n_feat"]]
    close(np.concatenate([im.features_undist for im in images], 0), GOLDEN[name + "/bearings"], 1e-11)


def test_undistort_then_project_round_trip_at_scale():
    """Size-independent property on 2 M features: cam2img(img2cam(xy)) == xy for the models whose
    undistortion converges in 5 iterations at this distortion level (device pointers in and out)."""
    import torch
    n = 2_000_000
    rng = np.random.default_rng(5)
    models = [0, 1, 2, 3, 4, 6]
    cams = np.stack([orc.Intrinsics(m, PIXEL_MODEL_PARAMS[m]).row() for m in models], 0)
    xy = np.column_stack([rng.uniform(200, 1080, n), rng.uniform(150, 810, n)])
    ci = rng.integers(0, len(models), n).astype(np.int32)
    d_xy, d_ci, d_cams = torch.from_numpy(xy).cuda(), torch.from_numpy(ci).cuda(), torch.from_numpy(cams).cuda()
    d_out = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    lib = _lib.load()
    _lib.check(lib.isfm_undistort_features(n, len(models), d_cams.data_ptr(), d_xy.data_ptr(), d_ci.data_ptr(), d_out.data_ptr(), None))
    b = d_out
    assert torch.all(torch.abs(torch.linalg.norm(b, dim=1) - 1) < 1e-14)
    # project the bearings back: identity pose per camera model, one "image" per model, point = bearing
    w2c = torch.eye(4, dtype=torch.float64, device="cuda").repeat(len(models), 1, 1).contiguous()
    image_cam = torch.arange(len(models), dtype=torch.int32, device="cuda")
    tix = torch.arange(n, dtype=torch.int32, device="cuda")
    valid = torch.empty(n, dtype=torch.uint8, device="cuda")
    err = torch.empty(n, dtype=torch.float64, device="cuda")
    _lib.check(lib.isfm_filter_reprojection(n, len(models), n, len(models), w2c.data_ptr(), image_cam.data_ptr(), d_cams.data_ptr(),
                                            b.data_ptr(), d_xy.data_ptr(), d_ci.data_ptr(), tix.data_ptr(), 0.05, valid.data_ptr(),
                                            err.data_ptr(), None))
    assert float(err.max()) < 0.05 and int(valid.sum()) == n      # 5 fixed-point iterations: ~1e-3 px worst case here
    # pinhole models invert exactly up to cam2img's `z + 1e-10` (defs.py:375): ~1e-10 * 500 px / z
    assert float(err[torch.from_numpy(ci < 2).cuda()].max()) < 1e-6


def test_retriangulate_tracks_end_to_end():
    """RetriangulateTracks with the real points-only TorchBA on the GPU: runs the reference's loop,
    lowers the reprojection error of the points, keeps poses / intrinsics, restores the flags."""
    from instantsfm_b200.synthetic import ba_arrays_to_scene, make_ba_problem
    a = make_ba_problem(12, 500, 2800, seed=77, model_id=3)
    cameras, images, full = ba_arrays_to_scene(a)
    for im in images:
        im.is_registered = True
    tracks_orig = {tid: t.observations.copy() for tid, t in full.items()}
    tracks = {tid: copy.deepcopy(t) for k, (tid, t) in enumerate(full.items()) if k % 9}
    for t in tracks.values():
        t.observations = t.observations[:-1] if len(t.observations) > 3 else t.observations
    poses0 = [im.world2cam.copy() for im in images]
    params0 = [np.array(c.params, dtype=np.float64) for c in cameras]
    _, e0 = orc.filter_reprojection_mask(cameras, images, tracks, 1e9)
    topts = {'complete_max_reproj_error': 60.0, 'filter_max_reproj_error': 40.0, 'filter_min_tri_angle': 0.05,
             'ba_global_max_refinements': 3, 'ba_global_max_refinement_change': 0.0005}
    bopts = {'min_num_view_per_track': 2, 'optimize_poses': True, 'thres_loss_function': 1.0, 'max_num_iterations': 15,
             'function_tolerance': 1e-6}
    with redirect_stdout(io.StringIO()):
        tr.RetriangulateTracks(cameras, images, tracks, tracks_orig, topts, bopts)
    _, e1 = orc.filter_reprojection_mask(cameras, images, tracks, 1e9)
    m0 = np.median(np.concatenate(list(e0.values())))
    m1 = np.median(np.concatenate(list(e1.values())))
    assert m1 < 0.5 * m0, (m0, m1)
    for im, P in zip(images, poses0):
        assert np.allclose(im.world2cam, P, rtol=0, atol=1e-6)     # points-only BA
    for c, p in zip(cameras, params0):
        assert np.allclose(np.array(c.params, dtype=np.float64), p, rtol=1e-6)
    assert all(im.is_registered for im in images) and bopts['optimize_poses'] is True
