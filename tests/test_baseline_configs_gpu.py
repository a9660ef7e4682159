"""Parity on BASELINE.json configs at the north-star tolerance: per-iteration cost, final
reprojection RMSE and the optimised poses / points of the fp32 product build agree with the fp64
oracle within 1e-4 relative (BASELINE.json north_star).  Poses / points are compared in the
reference solution's GAUGE (tests/helpers.gauge_align: BA fixes the scene only up to a similarity
transform; the raw arrays are additionally held to RAW_TOL).  C1 runs the oracle live (exact solves
through the point Schur complement); C2 compares against the committed oracle trajectory
(tests/golden/ba_trajectory_C2.npz, written by tests/golden/make_ba_trajectory_golden.py)."""
import os

import numpy as np
import pytest

from instantsfm_b200.synthetic import make_config
from tests.helpers import gauge_align

pytestmark = pytest.mark.gpu
TOL = 1e-4   # BASELINE.json north_star, fp32
RAW_TOL = 1e-3   # un-aligned arrays: drift along the 7 gauge directions (cost-invisible) included
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _engine(a, dtype=np.float32, **kw):
    from instantsfm_b200.engine import BAEngine
    eng = BAEngine(a.model_id, dtype=dtype, **kw)
    eng.set_problem(a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
    return eng


def _nrel(x, ref):
    return float(np.linalg.norm(np.asarray(x, np.float64) - ref) / np.linalg.norm(ref))


def test_c1_full_trajectory_poses_points_fp32():
    from oracle.ba import BAProblem, make_optimizer
    a = make_config("C1")
    pb = BAProblem(a.model_id, a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
    opt = make_optimizer(pb, 1.0, solver="schur")
    eng = _engine(a)
    for it in range(12):
        ref = opt.step()
        loss, st = eng.step()
        assert abs(loss - ref) <= TOL * ref, (it, loss, ref)
        assert st["trials"] == len(opt.trace[-1]["trials"]), (it, st, opt.trace[-1])
    rob, sq = eng.cost()
    assert abs(np.sqrt(sq / a.n_obs) - pb.rmse()) <= TOL * pb.rmse()
    cam, pts = eng.get_params()
    assert _nrel(cam, pb.cam) <= RAW_TOL and _nrel(pts, pb.pts) <= RAW_TOL, (_nrel(cam, pb.cam), _nrel(pts, pb.pts))
    cam, pts, _ = gauge_align(cam, pts, pb.pts)
    assert _nrel(cam, pb.cam) <= TOL, _nrel(cam, pb.cam)
    assert _nrel(pts, pb.pts) <= TOL, _nrel(pts, pb.pts)
    # per-block view of the same bar: translations, quaternions, intrinsics
    assert _nrel(cam[:, :3], pb.cam[:, :3]) <= TOL and _nrel(cam[:, 3:7], pb.cam[:, 3:7]) <= TOL and _nrel(cam[:, 7:], pb.cam[:, 7:]) <= TOL


def test_c2_full_matches_committed_oracle_trajectory_fp32():
    path = os.path.join(GOLDEN, "ba_trajectory_C2.npz")
    g = np.load(path)
    a = make_config("C2")
    # the golden belongs to exactly this instance
    assert np.isclose(a.points_2d.sum(), g["obs_checksum"][0], rtol=0, atol=1e-6 * abs(g["obs_checksum"][0]))
    assert int(a.camera_indices.astype(np.int64).sum()) == int(g["obs_checksum"][1])
    eng = _engine(a)
    for it, ref in enumerate(g["costs"]):
        loss, st = eng.step()
        assert abs(loss - ref) <= TOL * ref, (it, loss, ref)
        assert st["trials"] == int(g["trials"][it]), (it, st["trials"], int(g["trials"][it]))
    rob, sq = eng.cost()
    rmse = np.sqrt(sq / a.n_obs)
    assert abs(rmse - g["rmse"][-1]) <= TOL * g["rmse"][-1], (rmse, g["rmse"][-1])
    cam, pts = eng.get_params()
    assert _nrel(cam, g["cam"]) <= RAW_TOL and _nrel(pts[g["point_sample"]], g["points"]) <= RAW_TOL
    cam, pts, _ = gauge_align(cam, pts, g["points"], sample=g["point_sample"])
    assert _nrel(cam, g["cam"]) <= TOL, _nrel(cam, g["cam"])
    assert _nrel(cam[:, :3], g["cam"][:, :3]) <= TOL and _nrel(cam[:, 3:7], g["cam"][:, 3:7]) <= TOL
    assert _nrel(pts[g["point_sample"]], g["points"]) <= TOL, _nrel(pts[g["point_sample"]], g["points"])


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-10), (np.float32, 1e-5)])
def test_persistent_kernel_matches_per_iteration_kernels(dtype, tol, monkeypatch):
    """The one-launch persistent PCG solve (grid barriers) and the four-kernels-per-iteration WHILE
    graph are the same algorithm: same LM trajectory, same iteration counts (fp64)."""
    from instantsfm_b200.synthetic import make_ba_problem
    a = make_ba_problem(48, 4000, 22000, seed=31)
    kw = dict(pcg_tol=1e-10 if dtype == np.float64 else 1e-6)
    monkeypatch.setenv("ISFM_TWO_LEVEL", "0")   # the per-iteration kernels know block-Jacobi only
    monkeypatch.delenv("ISFM_NO_PERSISTENT", raising=False)
    e1 = _engine(a, dtype, **kw)
    assert e1.pcg_phases()[1] == 0
    monkeypatch.setenv("ISFM_NO_PERSISTENT", "1")
    e2 = _engine(a, dtype, **kw)
    monkeypatch.delenv("ISFM_NO_PERSISTENT", raising=False)
    for it in range(8):
        l1, s1 = e1.step()
        l2, s2 = e2.step()
        assert abs(l1 - l2) <= tol * l2, (it, l1, l2)
        assert s1["trials"] == s2["trials"]
        if dtype == np.float64:
            assert abs(s1["pcg_iters"] - s2["pcg_iters"]) <= 1, (s1, s2)
        assert s1["pcg_status"] == 1 and s2["pcg_status"] == 1
    assert e1.pcg_phases()[1] > 0 and e2.pcg_phases()[1] == 0   # the persistent kernel ran on e1 only
    c1, p1 = e1.get_params()
    c2, p2 = e2.get_params()
    assert _nrel(c1, c2.astype(np.float64)) <= 100 * tol and _nrel(p1, p2.astype(np.float64)) <= 100 * tol


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_two_level_preconditioner_same_solution_fewer_iterations(dtype, monkeypatch):
    """Street scene (banded reduced camera system): block-Jacobi + the coarse level of cluster
    similarity modes solves the SAME systems (same LM trajectory) in several times fewer PCG
    iterations than block-Jacobi alone."""
    from instantsfm_b200.synthetic import make_ba_problem
    a = make_ba_problem(1200, 36000, 216000, window=32, seed=5)
    tol = 1e-10 if dtype == np.float64 else 1e-6
    monkeypatch.setenv("ISFM_TWO_LEVEL", "0")
    e0 = _engine(a, dtype, pcg_tol=tol, pcg_max_iter=20000)
    monkeypatch.setenv("ISFM_TWO_LEVEL", "1")
    e1 = _engine(a, dtype, pcg_tol=tol, pcg_max_iter=20000)
    monkeypatch.delenv("ISFM_TWO_LEVEL", raising=False)
    assert not e0.pcg_phases()[2] and e1.pcg_phases()[2]
    it0 = it1 = 0
    for it in range(6):
        l0, s0 = e0.step()
        l1, s1 = e1.step()
        assert abs(l0 - l1) <= (1e-8 if dtype == np.float64 else 1e-4) * l0, (it, l0, l1)
        assert s0["trials"] == s1["trials"] and s0["pcg_status"] == 1 and s1["pcg_status"] == 1, (s0, s1)
        it0 += s0["pcg_iters"]; it1 += s1["pcg_iters"]
    assert it1 * 2.5 <= it0, (it0, it1)
    if dtype == np.float64:
        c0, _ = e0.get_params()
        c1, _ = e1.get_params()
        assert _nrel(c1, c0) <= 1e-6


def test_two_level_default_policy(monkeypatch):
    """Default: the coarse level is on for every reduced camera system of >= 32 cameras (dense BAL-like
    ones included), off below."""
    from instantsfm_b200.synthetic import make_ba_problem
    monkeypatch.delenv("ISFM_TWO_LEVEL", raising=False)
    small = _engine(make_ba_problem(24, 1500, 8000, seed=13))
    dense = _engine(make_ba_problem(600, 6000, 36000, seed=3))           # every camera pair co-observes: dense S
    chain = _engine(make_ba_problem(2000, 40000, 240000, window=32, seed=4))
    assert not small.pcg_phases()[2]
    assert dense.pcg_phases()[2] and chain.pcg_phases()[2]
    for e in (dense, chain):
        for _ in range(3):
            _, st = e.step()
            assert st["pcg_status"] == 1


def test_two_level_stays_positive_definite_at_late_lm_damping(monkeypatch):
    """25 accepted LM steps double the trust-region radius 25 times: the damping falls to 3e-12 and the
    reduced system's seven gauge modes are left with eigenvalues far below the rounding noise of the
    stored fp32 blocks.  The coarse matrix (ridge, fp64 inverse) must stay positive definite -- same
    trajectory as block-Jacobi alone, no rejected trials, no breakdown, and far fewer iterations."""
    from instantsfm_b200.synthetic import make_ba_problem
    a = make_ba_problem(200, 12000, 80000, seed=17)
    monkeypatch.setenv("ISFM_TWO_LEVEL", "0")
    e0 = _engine(a)
    monkeypatch.setenv("ISFM_TWO_LEVEL", "1")
    e1 = _engine(a)
    monkeypatch.delenv("ISFM_TWO_LEVEL", raising=False)
    it0 = it1 = 0
    for it in range(25):
        l0, s0 = e0.step()
        l1, s1 = e1.step()
        assert abs(l0 - l1) <= 1e-4 * l0, (it, l0, l1)
        assert s1["pcg_status"] == 1 and s1["rejects"] == s0["rejects"], (it, s0, s1)
        it0 += s0["pcg_iters"]; it1 += s1["pcg_iters"]
    assert it1 * 1.5 <= it0, (it0, it1)


@pytest.mark.parametrize("model_id", [0, 2, 4, 6])   # D = 7, 8, 12, 16 (RADIAL, D = 9, is what the configs above run)
def test_two_level_every_block_size(model_id):
    """The persistent kernel with the coarse level for every camera-block size: 40 cameras (two-level on
    by default), fp64 vs the oracle's exact solves, fp32 at the north-star tolerance."""
    from instantsfm_b200.synthetic import make_ba_problem
    from oracle.ba import BAProblem, make_optimizer
    a = make_ba_problem(40, 1500, 9000, seed=80 + model_id, model_id=model_id)
    pb = BAProblem(a.model_id, a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
    opt = make_optimizer(pb, 1.0, solver="direct")
    e64 = _engine(a, np.float64, pcg_tol=1e-12)
    e32 = _engine(a, np.float32)
    assert e64.pcg_phases()[2] and e32.pcg_phases()[2]
    for it in range(3):   # (verified for 5; the exact oracle solves dominate the run time)
        ref = opt.step()
        l64, s64 = e64.step()
        l32, s32 = e32.step()
        assert abs(l64 - ref) <= 1e-8 * ref, (it, l64, ref)
        assert abs(l32 - ref) <= TOL * ref, (it, l32, ref)
        assert s64["pcg_status"] == 1 and s32["pcg_status"] == 1
    assert e64.pcg_phases()[1] > 0   # solved by the persistent kernel
