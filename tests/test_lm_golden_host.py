"""LM semantics of the oracle vs a trajectory recorded from the reference's own stack (real bae +
pypose@bae, tests/golden/make_lm_golden.py).  The file cannot be produced in this container (the two
packages are not installable offline): the test SKIPS, which is exactly the "parity unpinned" state
DESIGN.md section 7 declares; the generator itself must run and decline cleanly."""
import copy
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "reference_lm_trajectory.npz")


def test_generator_declines_cleanly_without_the_reference_stack():
    try:
        import bae  # noqa: F401
        import pypose  # noqa: F401
        pytest.skip("bae / pypose are importable here: run tests/golden/make_lm_golden.py and commit its output")
    except ImportError:
        pass
    before = os.path.exists(GOLDEN)
    out = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_lm_golden.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "nothing written" in out.stdout
    assert os.path.exists(GOLDEN) == before


def test_oracle_lm_matches_recorded_reference_trajectory():
    if not os.path.exists(GOLDEN):
        pytest.skip("parity unpinned: no trajectory recorded from real bae + pypose (not installable offline)")
    from instantsfm_b200.geometry import matrices_to_pose7
    from instantsfm_b200.synthetic import ba_arrays_to_scene, make_ba_problem
    from oracle import ba as oba
    g = np.load(GOLDEN)
    n_cam, n_pt, n_obs, seed = (int(x) for x in g["scene"])
    a = make_ba_problem(n_cam, n_pt, n_obs, seed=seed)
    cameras, images, tracks = ba_arrays_to_scene(a)
    steps = int(g["steps"])
    opts = {"optimize_poses": True, "optimize_points": True, "min_num_view_per_track": 2, "thres_loss_function": 1.0,
            "function_tolerance": 0.0, "max_num_iterations": steps}
    # replay step by step: after k LM iterations the oracle's scene must equal the recorded snapshot k
    for k in range(g["poses"].shape[0] - 1):
        c, i, t = copy.deepcopy((cameras, images, tracks))
        oba.solve(c, i, t, dict(opts, max_num_iterations=k + 1), solver="pcg", pcg_tol=1e-5)
        poses = np.stack([im.world2cam for im in i], 0)
        pts = np.stack([tr.xyz for tr in t.values()], 0)
        prm = np.stack([np.asarray(cm.params, float) for cm in c], 0)
        assert np.abs(matrices_to_pose7(poses) - matrices_to_pose7(g["poses"][k])).max() <= 1e-5, k
        assert np.abs(pts - g["points"][k]).max() <= 1e-4 * max(1.0, np.abs(g["points"][k]).max()), k
        assert np.abs(prm - g["params"][k]).max() <= 1e-5 * np.abs(g["params"][k]).max(), k
