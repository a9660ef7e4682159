"""GPU edge cases of the BA path against the oracle (fp64 build, exact sparse solve):
degenerate index structures the reference tolerates (single-observation points, cameras
without observations, a single camera, very long tracks) and the zero Jacobian column of
OPENCV_FISHEYE's ignored k4 (diagonal clamp 1e-6)."""
import numpy as np
import pytest

from instantsfm_b200.synthetic import BAArrays, make_ba_problem

pytestmark = pytest.mark.gpu


def _run(a, steps=3, tol=1e-8, **kw):
    from instantsfm_b200.engine import BAEngine
    from oracle.ba import BAProblem, make_optimizer
    eng = BAEngine(a.model_id, dtype=np.float64, pcg_tol=1e-13, pcg_max_iter=5000, **kw)
    eng.set_problem(a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
    pb = BAProblem(a.model_id, a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices,
                   a.point_indices, kw.get("optimize_poses", True))
    opt = make_optimizer(pb, 1.0, solver="direct")
    for it in range(steps):
        ref = opt.step()
        loss, st = eng.step()
        assert abs(loss - ref) <= tol * ref + 1e-10, (it, loss, ref, st)
        assert st["trials"] == len(opt.trace[it]["trials"])
    cam, pts = eng.get_params()
    assert np.abs(cam - pb.cam).max() <= 1e-6 * max(1.0, np.abs(pb.cam).max())
    assert np.abs(pts - pb.pts).max() <= 1e-6 * max(1.0, np.abs(pb.pts).max())
    return eng


def _subset(a, keep):
    return BAArrays(a.model_id, a.camera_params, a.camera_pps, a.points_3d, a.points_2d[keep], a.camera_indices[keep],
                    a.point_indices[keep])


def test_single_camera_no_offdiagonal_blocks():
    a = make_ba_problem(6, 60, 240, seed=61)
    keep = a.camera_indices == 2
    b = _subset(a, keep)
    b.camera_params, b.camera_pps = a.camera_params[2:3], a.camera_pps[2:3]
    b.camera_indices = np.zeros(keep.sum(), np.int32)
    upt, inv = np.unique(b.point_indices, return_inverse=True)
    b.points_3d, b.point_indices = a.points_3d[upt], inv.astype(np.int32)
    eng = _run(b, steps=2)
    assert eng.schur_pattern()["nnzb"] == 1


def test_points_with_a_single_observation_and_unused_camera():
    """min_num_view_per_track is tested on the unfiltered track (bundle_adjustment.py:67-68), so
    single-observation points survive (rank-2 Hpp, held by clamp + damping); camera 3 ends up
    with no observation at all (zero Hcc block -> clamp)."""
    a = make_ba_problem(8, 120, 520, seed=63)
    keep = a.camera_indices != 3
    first = np.searchsorted(a.point_indices, np.arange(a.n_pt))
    drop_rest = np.zeros(a.n_obs, bool)
    for p in range(0, 30):                      # cut the first 30 tracks to one observation
        drop_rest[first[p] + 1:first[p + 1] if p + 1 < a.n_pt else a.n_obs] = True
    b = _subset(a, keep & ~drop_rest)
    counts = np.bincount(b.point_indices, minlength=a.n_pt)
    assert (counts == 1).sum() >= 10 and (np.bincount(b.camera_indices, minlength=8)[3] == 0)
    # a few points lose every observation (Hpp = 0: held by the diagonal clamp, like the reference)
    _run(b, steps=3, tol=1e-7)


def test_track_longer_than_the_fused_kernel_capacity():
    """One track seen by 300 cameras: the fused K1 (<= 256 observations per CTA) must hand over
    to the general kernels."""
    a = make_ba_problem(300, 400, 2000, seed=65)
    rng = np.random.default_rng(0)
    from instantsfm_b200.synthetic import project_numpy
    X = a.gt_points_3d[:1]
    cams = np.arange(300)
    obs, depth = project_numpy(a.model_id, np.repeat(X, 300, 0), a.gt_camera_params[cams], a.camera_pps[cams])
    assert depth.min() > 0.1
    obs += rng.normal(scale=0.5, size=obs.shape)
    keep = a.point_indices != 0
    b = BAArrays(a.model_id, a.camera_params, a.camera_pps, a.points_3d,
                 np.concatenate([obs, a.points_2d[keep]]), np.concatenate([cams.astype(np.int32), a.camera_indices[keep]]),
                 np.concatenate([np.zeros(300, np.int32), a.point_indices[keep]]))
    _run(b, steps=2, tol=1e-7)


def test_opencv_fisheye_ignored_k4_is_held_by_the_diagonal_clamp():
    a = make_ba_problem(8, 150, 700, seed=67, model_id=5)
    eng = _run(a, steps=2, tol=1e-7)
    cam, _ = eng.get_params()
    np.testing.assert_allclose(cam[:, -1], a.camera_params[:, -1], atol=1e-9)   # k4 has a zero Jacobian column


@pytest.mark.parametrize("model_id", [0, 1, 2, 4, 6, 8, 9])
def test_two_steps_for_every_other_model(model_id):
    a = make_ba_problem(8, 150, 700, seed=70 + model_id, model_id=model_id)
    _run(a, steps=2, tol=1e-7)


def test_fused_linearize_tma_and_plain_store_agree(monkeypatch):
    """The fused K1 exists twice: TMA kernel (record slab stored by bulk tensor copies, per-point
    sums by a warp-shuffle segmented scan) and the plain kernel (per-thread stores, per-point sums
    by a loop).  The two box sizes of the TMA kernel must agree bit for bit; the plain kernel sums
    the same numbers in another order, so it agrees to fp32 rounding."""
    import numpy as np
    from instantsfm_b200.engine import BAEngine
    from instantsfm_b200.synthetic import make_ba_problem
    a = make_ba_problem(24, 2500, 13000, seed=123)
    out = []
    for env in ({}, {"ISFM_TMA_BOX": "0"}, {"ISFM_NO_TMA": "1"}):
        for k in ("ISFM_NO_TMA", "ISFM_TMA_BOX"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng = BAEngine(a.model_id, dtype=np.float32)
        eng.set_problem(a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
        losses = [eng.step()[0] for _ in range(4)]
        cam, pts = eng.get_params()
        out.append((losses, cam, pts))
        eng.close()
    assert out[1][0] == out[0][0]
    assert np.array_equal(out[1][1], out[0][1]) and np.array_equal(out[1][2], out[0][2])
    assert np.allclose(out[2][0], out[0][0], rtol=1e-5)
    assert np.abs(out[2][1] - out[0][1]).max() <= 1e-3 * np.abs(out[0][1]).max()
    assert np.abs(out[2][2] - out[0][2]).max() <= 1e-3 * np.abs(out[0][2]).max()
