"""Run under torchrun with >= 2 GPUs: the point-partitioned multi-rank BA / GP solve must
reproduce the single-rank solve (same per-iteration costs and parameters up to fp rounding
of the changed summation order).  Prints MULTIGPU_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instantsfm_b200.engine import BAEngine, Communicator, GPEngine  # noqa: E402
from instantsfm_b200.partition import point_offsets, shard_ba  # noqa: E402
from instantsfm_b200.engine import partition_points  # noqa: E402
from instantsfm_b200.synthetic import make_ba_problem, make_gp_problem  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Communicator()
    for dtype, tol in [(np.float64, 1e-9), (np.float32, 1e-4)]:
        a = make_ba_problem(20, 900, 4800, seed=91)
        local_t, (p0, p1) = shard_ba(a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices,
                                     a.point_indices, rank, world)
        multi = BAEngine(a.model_id, dtype=dtype, comm=comm, pcg_tol=1e-10 if dtype == np.float64 else 1e-6)
        multi.set_problem(*local_t)
        owned, total = multi.matvec_units()
        if os.environ.get("ISFM_SPLIT_MATVEC") == "1":
            # dense co-visibility: both ranks have the same block pattern, the mat-vec is split
            assert 0 < owned < total, (owned, total)
        else:
            assert owned == total
        single = BAEngine(a.model_id, dtype=dtype, pcg_tol=1e-10 if dtype == np.float64 else 1e-6)
        single.set_problem(a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
        for it in range(8):
            lm, sm = multi.step()
            ls, ss = single.step()
            assert abs(lm - ls) <= tol * ls, (dtype, it, lm, ls)
            assert sm["trials"] == ss["trials"]
        cm, pm = multi.get_params()
        cs, ps = single.get_params()
        ptol = 1e-7 if dtype == np.float64 else 2e-3
        assert np.abs(cm - cs).max() <= ptol * np.abs(cs).max()
        assert np.abs(pm - ps[p0:p1]).max() <= ptol * np.abs(ps).max()
        # every rank must hold bit-identical cameras (replicated PCG, identical all-reduce results)
        t = torch.from_numpy(cm.astype(np.float64)).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(t, ref), "camera parameters diverged across ranks"
    # banded (street) problem: most cameras have NO observation on a given rank
    a = make_ba_problem(400, 6000, 30000, window=24, seed=95)
    local_t, (p0, p1) = shard_ba(a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices,
                                 a.point_indices, rank, world)
    multi = BAEngine(a.model_id, dtype=np.float64, comm=comm, pcg_tol=1e-10, pcg_max_iter=20000)
    multi.set_problem(*local_t)
    single = BAEngine(a.model_id, dtype=np.float64, pcg_tol=1e-10, pcg_max_iter=20000)
    single.set_problem(a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
    for it in range(6):
        lm, sm = multi.step()
        ls, ss = single.step()
        assert abs(lm - ls) <= 1e-7 * ls, ("street", it, lm, ls, sm, ss)
    cm, _ = multi.get_params()
    cs, _ = single.get_params()
    assert np.abs(cm - cs).max() <= 1e-6 * np.abs(cs).max()
    rm, qm = multi.cost()
    rs, qs = single.cost()
    assert abs(qm - qs) <= 1e-7 * qs, (qm, qs)
    # global positioning
    g = make_gp_problem(16, 500, 2400, seed=93)
    begin = partition_points(point_offsets(g.point_indices, g.points_3d.shape[0]), world)
    p0, p1 = int(begin[rank]), int(begin[rank + 1])
    sel = np.flatnonzero((g.point_indices >= p0) & (g.point_indices < p1))
    multi = GPEngine(dtype=np.float64, comm=comm, pcg_tol=1e-11)
    multi.set_problem(g.camera_translations, g.points_3d[p0:p1], g.scales[sel], g.translations[sel], g.camera_indices[sel],
                      g.point_indices[sel] - p0, g.is_calibrated)
    single = GPEngine(dtype=np.float64, pcg_tol=1e-11)
    single.set_problem(g.camera_translations, g.points_3d, g.scales, g.translations, g.camera_indices, g.point_indices,
                       g.is_calibrated)
    for it in range(6):
        lm, _ = multi.step()
        ls, _ = single.step()
        assert abs(lm - ls) <= 1e-8 * ls, ("gp", it, lm, ls)
    dist.barrier()
    want_peer = not os.environ.get("ISFM_NO_PEER")
    if os.environ.get("ISFM_REQUIRE_PEER"):
        assert comm.peer_enabled == want_peer, "peer-memory exchange was expected to be %s" % ("on" if want_peer else "off")
    if rank == 0:
        print("MULTIGPU_OK world", world, "transport", "peer" if comm.peer_enabled else "nccl",
              "split" if os.environ.get("ISFM_SPLIT_MATVEC") == "1" else "nosplit")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
