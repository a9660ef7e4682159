"""oracle/torch_ba.py -- the torch restatement of the reference's loop that bench.py times as
`cpu_baseline` (device cpu) and `reference_gpu` (device cuda) -- against the scipy oracle
(oracle/lm.py) run the reference's way: full camera + point system, Jacobi PCG to 1e-5.  Both apply
the same algorithm; they must walk the same LM trajectory, in both ways the port applies J^T J."""
import numpy as np
import pytest

from instantsfm_b200.synthetic import make_ba_problem
from oracle.ba import BAProblem, make_optimizer
from oracle.torch_ba import TorchRefBA


@pytest.mark.parametrize("mode", ["sparse", "blocks"])
def test_torch_port_walks_the_oracle_trajectory(mode):
    a = make_ba_problem(12, 400, 2200, seed=33)
    args = (a.model_id, a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
    pb = BAProblem(*args)
    opt = make_optimizer(pb, 1.0, solver="pcg", pcg_tol=1e-5)
    port = TorchRefBA(*args, device="cpu", mode=mode)
    for it in range(6):
        ref = opt.step()
        got = port.step()
        # two PCG implementations stopped at the same 1e-5 residual: same cost to ~1e-6
        assert abs(got - ref) <= 2e-5 * ref, (it, got, ref)
    r = port.residuals().numpy()
    assert abs(np.sqrt((r * r).sum(-1).mean()) - pb.rmse()) <= 1e-4 * pb.rmse()
    assert np.abs(port.cam.numpy() - pb.cam).max() <= 1e-3 * np.abs(pb.cam).max()
    assert port.pcg_iters > 0


def test_torch_port_exact_solves_match_direct_oracle():
    """With a tight PCG tolerance the port agrees with the oracle's exact sparse solve."""
    a = make_ba_problem(8, 200, 1000, seed=34)
    args = (a.model_id, a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
    pb = BAProblem(*args)
    opt = make_optimizer(pb, 1.0, solver="direct")
    port = TorchRefBA(*args, device="cpu", pcg_tol=1e-13)
    for it in range(5):
        ref = opt.step()
        got = port.step()
        assert abs(got - ref) <= 1e-8 * ref, (it, got, ref)
