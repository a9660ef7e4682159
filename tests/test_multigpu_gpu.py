"""Launches tests/multigpu_check.py under torchrun when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("transport,split,push,extra", [
    ("peer", False, None, {}), ("nccl", False, None, {}), ("peer", True, None, {}), ("nccl", True, None, {}),
    ("peer", False, "grid", {}), ("peer", True, "grid", {}),
    ("peer", False, None, {"ISFM_TWO_LEVEL": "1"}),        # coarse level all-reduced across ranks (street problem)
    ("peer", False, None, {"ISFM_NO_PERSISTENT": "1"}),    # the per-iteration WHILE-graph path
    ("peer", False, "grid", {"ISFM_FULL_EXCHANGE": "1"})])  # whole-vector exchange instead of the touched row ring
def test_two_rank_solve_matches_single_rank(transport, split, push, extra):
    """Both transports of the per-iteration sum (peer-memory exchange fused into the PCG kernels,
    ncclAllReduce), with and without the split mat-vec (summed E reduce-scattered by unit ranges
    when the ranks share one block pattern), must reproduce the single-rank solve."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(here, "multigpu_check.py")]
    env = dict(os.environ)
    env.pop("ISFM_NO_PEER", None)
    if transport == "nccl":
        env["ISFM_NO_PEER"] = "1"
    env["ISFM_SPLIT_MATVEC"] = "1" if split else "0"
    env.pop("ISFM_PEER_PUSH", None)
    for k in ("ISFM_TWO_LEVEL", "ISFM_NO_PERSISTENT", "ISFM_FULL_EXCHANGE"):
        env.pop(k, None)
    env.update(extra)
    if push:
        env["ISFM_PEER_PUSH"] = push   # "grid": the grid-wide push kernel large camera systems use
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and "MULTIGPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    assert ("transport " + transport + (" split" if split else " nosplit")) in out.stdout, out.stdout[-2000:]
