"""Host-side logic of the drop-in processors (no GPU): the vectorised tensor set-up of
TorchBA must produce bit-identical tensors to the reference's nested Python loops as
restated in oracle/ba.py (flatten_tracks, split_principal_point, cheirality_and_compact),
including unregistered images, short tracks and behind-camera observations."""
import copy

import numpy as np
import pytest

from instantsfm_b200.processors._common import should_stop
from instantsfm_b200.synthetic import ba_arrays_to_scene, make_ba_problem
from oracle import ba as oba
from oracle.lm import run_loop

OPTS = {"optimize_poses": True, "optimize_points": True, "min_num_view_per_track": 2, "thres_loss_function": 1.0,
        "max_num_iterations": 200, "function_tolerance": 5e-4}


def _scene(model_id=3, **kw):
    a = make_ba_problem(12, 300, 1500, seed=41, model_id=model_id)
    a.camera_pps = a.camera_pps + np.array([320.0, 240.0])
    # put a few points behind their first camera so that the cheirality filter drops observations
    a.points_3d[:5] *= -8.0
    return a, ba_arrays_to_scene(a, **kw)


@pytest.mark.parametrize("model_id", [0, 1, 2, 3, 4, 5, 6, 8, 9])
def test_build_matches_reference_loops(model_id):
    from instantsfm_b200.processors.bundle_adjustment import TorchBA
    a, (cameras, images, tracks) = _scene(model_id, unregistered=(3, 7), short_track_every=9)
    t = TorchBA()._build(cameras, images, tracks, OPTS, model_id)
    flat = oba.flatten_tracks(cameras, images, tracks, OPTS)
    cam, pps, rest, pp_idx = oba.split_principal_point(model_id, flat["camera_params"])
    comp = oba.cheirality_and_compact(cam, pps, flat["points_3d"], flat["points_2d"], flat["camera_indices"],
                                      flat["point_indices"])
    assert (~comp["keep"]).sum() > 0                       # the filter really dropped something
    assert 3 not in comp["unique_cameras"] and 7 not in comp["unique_cameras"]
    assert np.array_equal(t["unique_cameras"], comp["unique_cameras"])
    assert np.array_equal(t["unique_points"], comp["unique_points"])
    assert np.array_equal(t["camera_indices"], comp["camera_indices"])
    assert np.array_equal(t["point_indices"], comp["point_indices"])
    assert np.array_equal(t["points_2d"], comp["points_2d"])
    assert np.array_equal(t["points_3d"], comp["points_3d"])
    assert np.array_equal(t["camera_pps"], comp["camera_pps"])
    np.testing.assert_allclose(t["camera_params"], comp["camera_params"], rtol=0, atol=1e-15)
    assert np.array_equal(t["remaining"], rest) and np.array_equal(t["pp_indices"], pp_idx)
    assert t["camera_indices"].dtype == np.int32 and t["point_indices"].dtype == np.int32   # :99-100


def test_unsupported_models_raise():
    from instantsfm_b200.processors.bundle_adjustment import TorchBA
    from instantsfm_b200.scene.defs import Camera, CameraModelId
    for m, n in [(CameraModelId.FOV, 5), (CameraModelId.THIN_PRISM_FISHEYE, 12)]:
        with pytest.raises(NotImplementedError):
            TorchBA().Solve([Camera(model_id=m, params=[1.0] * n)], [], {}, OPTS)


def test_write_back_matches_reference_update():
    from instantsfm_b200.processors.bundle_adjustment import TorchBA, update
    a, (cameras, images, tracks) = _scene(3)
    t = TorchBA()._build(cameras, images, tracks, OPTS, 3)
    rng = np.random.default_rng(0)
    cam = t["camera_params"] + rng.normal(scale=1e-3, size=t["camera_params"].shape)
    cam[:, 3:7] /= np.linalg.norm(cam[:, 3:7], axis=1, keepdims=True)
    pts = t["points_3d"] + 0.1
    c2, i2, t2 = copy.deepcopy((cameras, images, tracks))
    update(cameras, images, tracks, t["track_keys"], t["unique_cameras"], t["unique_points"], t["remaining"],
           t["pp_indices"], cam, t["camera_pps"], pts)
    flat = oba.flatten_tracks(c2, i2, t2, OPTS)
    camf, pps, rest, pp_idx = oba.split_principal_point(3, flat["camera_params"])
    comp = oba.cheirality_and_compact(camf, pps, flat["points_3d"], flat["points_2d"], flat["camera_indices"], flat["point_indices"])
    pb = oba.BAProblem(3, comp["camera_params"], comp["camera_pps"], comp["points_3d"], comp["points_2d"],
                       comp["camera_indices"], comp["point_indices"])
    pb.cam, pb.pts = cam.copy(), pts.copy()
    oba.write_back(c2, i2, t2, flat, comp, rest, pp_idx, pb)
    for x, y in zip(images, i2):
        np.testing.assert_allclose(x.world2cam, y.world2cam, atol=1e-14)
    for x, y in zip(cameras, c2):
        np.testing.assert_allclose(np.asarray(x.params, float), np.asarray(y.params, float), atol=0)
    for k in tracks:
        np.testing.assert_array_equal(tracks[k].xyz, t2[k].xyz)


def test_stop_rule_matches_oracle_loop():
    class Fake:
        def __init__(self, seq):
            self.seq, self.i = seq, 0

        def step(self):
            self.i += 1
            return self.seq[self.i - 1]
    rng = np.random.default_rng(5)
    for trial in range(30):
        seq = list(np.cumsum(rng.uniform(0, 1, 40))[::-1] * rng.uniform(1e-4, 1) + 100)
        if trial % 3 == 0:
            seq[10] = seq[9]
        for ident in (True, False):
            ref = run_loop(Fake(seq), 40, 5e-4, ident)
            hist = []
            for v in seq:
                hist.append(v)
                if should_stop(hist, 5e-4, ident):
                    break
            assert hist == ref


def test_cpu_device_is_refused():
    from instantsfm_b200.processors._common import device_index
    assert device_index("cuda:3") == 3 and device_index("cuda") == 0
    with pytest.raises(RuntimeError):
        device_index("cpu")
