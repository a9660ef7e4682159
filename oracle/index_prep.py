"""Integer oracle (numpy) for the index preparation of the CUDA path.  TEST INFRASTRUCTURE.

The reference does this with ``torch.unique(sorted=True, return_inverse=True)``
(bundle_adjustment.py:108-109) and bae's index tracing; the CUDA path needs more structure
(sort by point, sort by camera, pair lists, partition), all of it integer and bit-exact.
"""
import numpy as np


def structure(camera_indices, point_indices, n_cam, n_pt):
    perm = np.argsort(point_indices, kind="stable")
    cam_sorted = np.asarray(camera_indices)[perm]
    cperm = np.argsort(cam_sorted, kind="stable")
    return {"obs_perm": perm.astype(np.int32),
            "point_offsets": np.searchsorted(np.asarray(point_indices)[perm], np.arange(n_pt + 1)).astype(np.int64),
            "cam_perm": cperm.astype(np.int32),
            "cam_offsets": np.searchsorted(cam_sorted[cperm], np.arange(n_cam + 1)).astype(np.int64)}


def partition_points(point_offsets, world):
    """Contiguous point ranges balanced by observation count: boundary g = first point whose
    starting observation offset >= g * n_obs / world (exact integer arithmetic)."""
    off = np.asarray(point_offsets, dtype=np.int64)
    n_pt, n_obs = off.shape[0] - 1, int(off[-1])
    out = np.empty(world + 1, dtype=np.int64)
    for g in range(world + 1):
        # off[p] * world >= g * n_obs   (python ints: no overflow)
        lo, hi = 0, n_pt
        while lo < hi:
            mid = (lo + hi) // 2
            if int(off[mid]) * world < g * n_obs:
                lo = mid + 1
            else:
                hi = mid
        out[g] = n_pt if g == world else lo
    return out


def shard(problem_arrays, part_begin, rank):
    """Observations / points of one rank; cameras stay replicated.  Returns local arrays with
    point indices renumbered from 0."""
    ci, pi = np.asarray(problem_arrays["camera_indices"]), np.asarray(problem_arrays["point_indices"])
    p0, p1 = int(part_begin[rank]), int(part_begin[rank + 1])
    sel = np.flatnonzero((pi >= p0) & (pi < p1))
    out = dict(problem_arrays)
    out["points_3d"] = problem_arrays["points_3d"][p0:p1]
    out["points_2d"] = problem_arrays["points_2d"][sel]
    out["camera_indices"] = ci[sel]
    out["point_indices"] = (pi[sel] - p0).astype(pi.dtype)
    out["obs_sel"] = sel
    return out
