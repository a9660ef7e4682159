"""GPU tests of the drop-in processors: TorchBA.Solve / TorchGP.Optimize (reference
signatures, in-place mutation of cameras / images / tracks) against the oracle's restated
end-to-end Solve / Optimize."""
import copy

import numpy as np
import pytest

from instantsfm_b200.synthetic import ba_arrays_to_scene, gp_arrays_to_scene, make_ba_problem, make_gp_problem

pytestmark = pytest.mark.gpu

BA_OPTS = {"optimize_poses": True, "optimize_points": True, "min_num_view_per_track": 2, "thres_loss_function": 1.0,
           "max_num_iterations": 200, "function_tolerance": 5e-4}
GP_OPTS = {"min_num_view_per_track": 3, "thres_loss_function": 1e-1, "max_num_iterations": 100, "function_tolerance": 5e-4}


class Recorder:
    def __init__(self):
        self.calls = []

    def add_step(self, cameras, images, tracks, name=None):
        self.calls.append(name)


def _ba_scene():
    a = make_ba_problem(14, 500, 2600, seed=51)
    a.camera_pps = a.camera_pps + np.array([512.0, 384.0])
    a.points_3d[:4] *= -8.0
    return ba_arrays_to_scene(a, unregistered=(5,), short_track_every=11)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-7), (np.float32, 1e-4)])
def test_torchba_solve_matches_oracle(dtype, tol):
    from instantsfm_b200.processors import TorchBA
    from oracle import ba as oba
    cameras, images, tracks = _ba_scene()
    c2, i2, t2 = copy.deepcopy((cameras, images, tracks))
    opts = dict(BA_OPTS, max_num_iterations=25)
    ba = TorchBA(dtype=dtype, pcg_tol=1e-12 if dtype == np.float64 else 1e-6)
    assert ba.Solve(cameras, images, tracks, opts) is None
    hist, _, pb = oba.solve(c2, i2, t2, opts, solver="direct")
    assert len(ba.loss_history) == len(hist)
    if dtype == np.float64:
        np.testing.assert_allclose(ba.loss_history, hist, rtol=tol)
    else:
        # This scene is pathological on purpose: points on the wrong side of their cameras (near-
        # singular Hpp, residuals of 1e3 px).  The fp32 build tracks the fp64 reference to 1e-4 over
        # the well-conditioned first steps; where the cost falls 30 % per iteration a relative error
        # of 3e-3 in the step (condition 1e4 x fp32 storage of S = Hd - E) shows as ~1e-3 in the
        # cost, the trajectories meet again (1e-5 around iteration 10) and end on the same plateau
        # (DESIGN.md "fp32 conditioning"; BAL-shaped scenes hold 1e-4 throughout:
        # tests/test_baseline_configs_gpu.py).
        np.testing.assert_allclose(ba.loss_history[:4], hist[:4], rtol=tol)
        np.testing.assert_allclose(ba.loss_history[:14], hist[:14], rtol=2e-3)
        assert ba.loss_history[-1] <= 1.05 * hist[-1]
        return
    ptol = 1e-6
    for k in tracks:
        np.testing.assert_allclose(tracks[k].xyz, t2[k].xyz, atol=ptol * 10.0)
    for x, y in zip(images, i2):
        np.testing.assert_allclose(x.world2cam, y.world2cam, atol=ptol * 30.0)
    for x, y in zip(cameras, c2):
        np.testing.assert_allclose(np.asarray(x.params, float), np.asarray(y.params, float), rtol=ptol, atol=ptol)
    # untouched: unregistered image, principal points
    assert not images[5].is_registered and np.array_equal(images[5].world2cam, _ba_scene()[1][5].world2cam)


def test_torchba_points_only_and_visualizer():
    from instantsfm_b200.processors import TorchBA
    from oracle import ba as oba
    cameras, images, tracks = _ba_scene()
    c2, i2, t2 = copy.deepcopy((cameras, images, tracks))
    opts = dict(BA_OPTS, optimize_poses=False, max_num_iterations=6)
    vis = Recorder()
    ba = TorchBA(visualizer=vis, dtype=np.float64)
    ba.Solve(cameras, images, tracks, opts)
    hist, _, _ = oba.solve(c2, i2, t2, opts, solver="direct")
    np.testing.assert_allclose(ba.loss_history, hist, rtol=1e-8)
    assert vis.calls == ["bundle_adjustment"] * len(hist)
    for x, y in zip(images, c2 and i2):
        np.testing.assert_allclose(x.world2cam, y.world2cam, atol=1e-12)   # poses frozen


def test_torchgp_optimize_matches_oracle():
    from instantsfm_b200.processors import TorchGP
    from oracle import gp as ogp
    g = make_gp_problem(10, 160, 640, seed=53)
    cameras, images, tracks = gp_arrays_to_scene(g)
    # one short track that must be deleted, one image that loses all its tracks
    k0 = next(iter(tracks))
    tracks[k0].observations = tracks[k0].observations[:2]
    c2, i2, t2 = copy.deepcopy((cameras, images, tracks))
    opts = dict(GP_OPTS, max_num_iterations=10)
    gp = TorchGP(dtype=np.float64, pcg_tol=1e-12)
    gp.Optimize(cameras, images, tracks, None, opts)
    hist, _, pb = ogp.optimize(c2, i2, t2, None, opts, solver="direct")
    assert k0 not in tracks and k0 not in t2
    assert len(gp.loss_history) == len(hist)
    np.testing.assert_allclose(gp.loss_history, hist, rtol=1e-6)
    assert [im.is_registered for im in images] == [im.is_registered for im in i2]


def test_torchgp_convert_results_and_random_init():
    from instantsfm_b200.processors import TorchGP
    g = make_gp_problem(6, 40, 160, seed=3)
    cameras, images, tracks = gp_arrays_to_scene(g)
    gp = TorchGP()
    np.random.seed(0)
    gp.InitializeRandomPositions(cameras, images, tracks)
    assert all(np.abs(im.world2cam[:3, 3]).max() <= 100 for im in images)
    assert all(t.is_initialized for t in tracks.values())
    before = [im.world2cam.copy() for im in images]
    gp.ConvertResults(images)
    for b, im in zip(before, images):
        np.testing.assert_allclose(im.world2cam[:3, 3], -b[:3, :3] @ b[:3, 3])
