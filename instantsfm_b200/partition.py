"""Multi-GPU sharding of a BA / GP problem: points (in compact sorted order) are split into
``world`` contiguous ranges balanced by observation count; every observation lives with
its point; cameras are replicated.  Pure integer host work on the C ABI's
``isfm_partition_points``; no collective is involved."""
import numpy as np

from .engine import partition_points


def point_offsets(point_indices, n_pt):
    counts = np.bincount(np.asarray(point_indices), minlength=n_pt).astype(np.int64)
    return np.concatenate([[0], np.cumsum(counts)])


def shard_ba(camera_params, camera_pps, points_3d, points_2d, camera_indices, point_indices, rank, world):
    """-> (local tensors for BAEngine.set_problem, (p0, p1) point range of this rank)."""
    point_indices = np.asarray(point_indices)
    begin = partition_points(point_offsets(point_indices, points_3d.shape[0]), world)
    p0, p1 = int(begin[rank]), int(begin[rank + 1])
    sel = np.flatnonzero((point_indices >= p0) & (point_indices < p1))
    local = (camera_params, camera_pps, points_3d[p0:p1], points_2d[sel], np.asarray(camera_indices)[sel],
             (point_indices[sel] - p0).astype(np.int32))
    return local, (p0, p1)
