"""N > 1 path on the CPU: world_size-2 (and 3) `gloo` process groups check that the
point-partitioned normal equations sum to the unpartitioned ones -- the exact quantities
the CUDA path all-reduces over NCCL (Hcc | g_c, diag(E) | e, and E p inside PCG) -- and that
every rank derives the same partition from the C ABI's integer entry point."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from instantsfm_b200.partition import shard_ba
from instantsfm_b200.synthetic import make_ba_problem
from oracle.ba import BAProblem
from oracle.index_prep import partition_points as ref_partition
from tests.helpers import assemble_reduced_system, normal_blocks, partial_blocks, schur_contribution


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a = make_ba_problem(9, 120, 560, seed=77)
        mu = 1.0 + 1e-4
        local, (p0, p1) = shard_ba(a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices,
                                   a.point_indices, rank, world)
        off = np.concatenate([[0], np.cumsum(np.bincount(a.point_indices, minlength=a.n_pt))])
        assert (p0, p1) == tuple(ref_partition(off, world)[rank:rank + 2])
        pb = BAProblem(a.model_id, *local)
        Hpp, gp, Hcc, gc, Hcp = partial_blocks(pb)
        E, e = schur_contribution(pb, Hpp, gp, Hcp, mu)
        p_vec = np.random.default_rng(3).normal(size=E.shape[0])      # same on every rank
        Ep = E @ p_vec
        n_obs = torch.tensor([pb.n_obs])
        bufs = [torch.from_numpy(x.copy()) for x in (Hcc, gc, E, e, Ep)]
        for b in bufs + [n_obs]:
            dist.all_reduce(b)
        assert int(n_obs) == a.n_obs                                   # every observation lives on exactly one rank
        Hcc_s, gc_s, E_s, e_s, Ep_s = (b.numpy() for b in bufs)
        S, rhs = assemble_reduced_system(Hcc_s, gc_s, E_s, e_s, mu)
        full = BAProblem(a.model_id, a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
        _, _, Hcc_f, gc_f, S_f, rhs_f = normal_blocks(full, mu)
        scale = np.abs(Hcc_f).max()
        assert np.abs(Hcc_s - Hcc_f).max() <= 1e-12 * scale
        assert np.abs(S - S_f).max() <= 1e-11 * scale
        assert np.abs(rhs - rhs_f).max() <= 1e-11 * np.abs(rhs_f).max()
        # PCG mat-vec: S p = damp(Hcc) p - allreduce(E_g p)
        Sp = (S + E_s) @ p_vec - Ep_s
        assert np.abs(Sp - S_f @ p_vec).max() <= 1e-10 * np.abs(S_f @ p_vec).max()
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_normal_equations_sum_to_the_whole(world):
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert sorted(ret.keys()) == list(range(world))


@pytest.mark.parametrize("config,scale", [("C5", 0.01), ("C3", 0.02)])
@pytest.mark.parametrize("world", [2, 8])
def test_sharded_generator_slices_one_instance(config, scale, world):
    """bench.py's strong-scaling legs rest on this: make_config(shard=(rank, world)) for all ranks
    concatenates to exactly the arrays of the single-rank instance (same cameras on every rank, the
    points / observations of contiguous point ranges), so N ranks solve the SAME problem as one."""
    from instantsfm_b200.synthetic import make_config
    whole = make_config(config, scale=scale, shard=(0, 1))
    parts = [make_config(config, scale=scale, shard=(r, world)) for r in range(world)]
    assert sum(p.n_obs for p in parts) == whole.n_obs == whole.n_obs_total
    assert all(p.n_obs_total == whole.n_obs_total and p.n_pt_total == whole.n_pt for p in parts)
    assert all(np.array_equal(p.camera_params, whole.camera_params) and np.array_equal(p.camera_pps, whole.camera_pps) for p in parts)
    assert np.array_equal(np.concatenate([p.points_3d for p in parts]), whole.points_3d)
    assert np.array_equal(np.concatenate([p.points_2d for p in parts]), whole.points_2d)
    assert np.array_equal(np.concatenate([p.camera_indices for p in parts]), whole.camera_indices)
    off = np.cumsum([0] + [p.n_pt for p in parts])
    assert np.array_equal(np.concatenate([p.point_indices + off[r] for r, p in enumerate(parts)]), whole.point_indices)
    # balanced by observation count
    assert max(p.n_obs for p in parts) - min(p.n_obs for p in parts) <= 64
