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
