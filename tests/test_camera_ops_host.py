"""Oracle of the scene-layer camera maps (oracle/camera_ops.py) vs the golden vectors produced by the
reference's own Camera class / UndistortImages / FilterTracksByReprojection on cv2
(tests/golden/make_camera_ops_golden.py).  CPU only."""
import copy
import os

import numpy as np
import pytest

from oracle import camera_ops as orc
from tests.golden.make_camera_ops_golden import FILTER_CASES, UNDISTORT_CASES, point_inputs, snapshot
from tests.helpers import PIXEL_MODEL_PARAMS, make_pixel_scene

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_camera_ops.npz"))
FTOL = 1e-12   # fp64 outputs: libm (atan, tan, pow) and OpenCV's compiled arithmetic may differ in the last bits


def close(a, b, tol=FTOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    nan = np.isnan(b)
    assert np.array_equal(np.isnan(a), nan)
    fin = np.isfinite(b)
    assert np.array_equal(a[~fin & ~nan], b[~fin & ~nan])
    assert np.all(np.abs(a[fin] - b[fin]) <= tol * np.maximum(1.0, np.abs(b[fin]))), np.abs(a[fin] - b[fin]).max()


@pytest.mark.parametrize("model", range(11))
def test_cam2img_and_img2cam_match_reference(model):
    c = orc.Intrinsics(model, PIXEL_MODEL_PARAMS[model])
    uvw, xy = point_inputs(model)
    with np.errstate(all="ignore"):
        close(orc.cam2img(c, uvw), GOLDEN[f"cam2img/{model}"])
        close(orc.img2cam(c, xy), GOLDEN[f"img2cam/{model}"])


@pytest.mark.parametrize("case", UNDISTORT_CASES, ids=[c[0] for c in UNDISTORT_CASES])
def test_undistort_images_matches_reference(case):
    name, kw = case
    cameras, images, _ = make_pixel_scene(**kw)
    with np.errstate(all="ignore"):
        got = np.concatenate(orc.undistort_images(cameras, images), 0)
    assert np.array_equal(GOLDEN[name + "/n_feat"], [len(im.features) for im in images])
    close(got, GOLDEN[name + "/bearings"])


def check_filter_golden(name, tracks, ret):
    keys, lens, obs = snapshot(tracks)
    assert np.array_equal(keys, GOLDEN[name + "/keys"])
    assert np.array_equal(lens, GOLDEN[name + "/lens"])
    assert np.array_equal(obs, GOLDEN[name + "/obs"])
    assert int(ret) == int(GOLDEN[name + "/ret"])


@pytest.mark.parametrize("case", FILTER_CASES, ids=[c[0] for c in FILTER_CASES])
def test_filter_reprojection_matches_reference(case):
    name, thr, kw = case
    cameras, images, tracks = make_pixel_scene(**kw)
    tracks = copy.deepcopy(tracks)
    n_before = sum(len(t.observations) for t in tracks.values())
    ret = orc.apply_filter_reprojection(cameras, images, tracks, thr)
    check_filter_golden(name, tracks, ret)
    assert 0 < sum(len(t.observations) for t in tracks.values()) < n_before


def test_camera_table_row_layout():
    """Intrinsics.row() is the ISFM_CAMERA_ROW layout of include/isfm_b200.h."""
    r = orc.Intrinsics(10, PIXEL_MODEL_PARAMS[10]).row()
    p = PIXEL_MODEL_PARAMS[10]
    assert r.shape == (16,) and r[0] == 10
    assert list(r[1:5]) == p[0:4] and list(r[5:9]) == [p[4], p[5], p[8], p[9]] and list(r[11:13]) == p[6:8] and list(r[14:16]) == p[10:12]
    r = orc.Intrinsics(6, PIXEL_MODEL_PARAMS[6]).row()
    p = PIXEL_MODEL_PARAMS[6]
    assert list(r[5:11]) == [p[4], p[5], p[8], p[9], p[10], p[11]] and list(r[11:13]) == p[6:8]
    r = orc.Intrinsics(3, PIXEL_MODEL_PARAMS[3]).row()
    assert r[1] == r[2] == PIXEL_MODEL_PARAMS[3][0] and list(r[5:7]) == PIXEL_MODEL_PARAMS[3][3:5]


# ---------------------------------------------------------------------------------------------
# the DEVICE arithmetic of csrc/camera_ops.cu (csrc/camera_maps.cuh) compiled for the host
# ---------------------------------------------------------------------------------------------
def _hc(fn, row, pts, width):
    import ctypes
    from tests.hostcheck.build import load
    lib = load()
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    row = np.ascontiguousarray(row, dtype=np.float64)
    out = np.zeros((pts.shape[0], 2))
    p = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))   # noqa: E731
    getattr(lib, fn)(ctypes.c_long(pts.shape[0]), p(row), p(pts), p(out))
    assert pts.shape[1] == width
    return out


@pytest.mark.parametrize("model", range(11))
def test_device_camera_maps_on_host_match_reference_golden(model):
    """cam2img / img2cam exactly as the CUDA kernels inline them (same header, g++ without FMA
    contraction) against the vectors the reference's own Camera class produced on cv2 4.13."""
    row = orc.Intrinsics(model, PIXEL_MODEL_PARAMS[model]).row()
    uvw, xy = point_inputs(model)
    with np.errstate(all="ignore"):
        close(_hc("hc_cam2img", row, uvw, 3), GOLDEN[f"cam2img/{model}"], 1e-11)
        close(_hc("hc_img2cam", row, xy, 2), GOLDEN[f"img2cam/{model}"], 1e-11)


def test_device_camera_maps_on_host_match_oracle_on_random_cameras():
    rng = np.random.default_rng(9)
    for model in range(11):
        prm = np.array(PIXEL_MODEL_PARAMS[model]) * (1.0 + 0.02 * rng.normal(size=len(PIXEL_MODEL_PARAMS[model])))
        c = orc.Intrinsics(model, prm)
        uvw = np.column_stack([rng.normal(0, 0.5, 200), rng.normal(0, 0.4, 200), rng.uniform(0.7, 4.0, 200)])
        xy = np.column_stack([rng.uniform(100, 1180, 200), rng.uniform(80, 880, 200)])
        with np.errstate(all="ignore"):
            close(_hc("hc_cam2img", c.row(), uvw, 3), orc.cam2img(c, uvw), 1e-11)
            close(_hc("hc_img2cam", c.row(), xy, 2), orc.img2cam(c, xy), 1e-11)
