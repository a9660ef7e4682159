"""Array-level host API over the C ABI (include/isfm_b200.h).

``BAEngine`` / ``GPEngine`` own one library handle each and speak the tensors the
reference builds in TorchBA.Solve / TorchGP.Optimize before it enters ``optimizer.step``
(bundle_adjustment.py:111-126, global_positioning.py:108-168).  Inputs may be numpy arrays
(host) or CUDA ``torch.Tensor``s (device pointers are passed straight through).
All arithmetic happens in the CUDA library; nothing here computes on the CPU.
"""
import ctypes
from ctypes import byref, c_double, c_int32, c_int64, c_void_p

import numpy as np

from . import _lib
from ._lib import BADesc, GPDesc, StepStats, check

_DTYPES = {np.dtype(np.float32): 0, np.dtype(np.float64): 1}


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _buf(x, dtype, keep):
    """-> raw pointer of a contiguous array of `dtype`; host numpy or device torch."""
    if x is None:
        return None
    if _is_torch(x):
        import torch
        tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
               np.dtype(np.int32): torch.int32, np.dtype(np.uint8): torch.uint8}[np.dtype(dtype)]
        if x.dtype == torch.bool and np.dtype(dtype) == np.uint8:
            x = x.to(torch.uint8)
        t = x.to(tdt).contiguous()
        keep.append(t)
        return c_void_p(t.data_ptr())
    a = np.ascontiguousarray(x, dtype=dtype)
    keep.append(a)
    return a.ctypes.data_as(c_void_p)


def _current_stream():
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_stream().cuda_stream
    except Exception:
        pass
    return 0


class Communicator:
    """NCCL communicator of the library, one per process (rank = GPU).  Built from an
    initialised torch.distributed process group: rank 0 creates the NCCL unique id and it is
    broadcast with the group's own backend (gloo or nccl)."""

    def __init__(self, rank=None, world=None, unique_id=None):
        lib = _lib.load()
        import torch.distributed as dist
        if rank is None:
            rank, world = dist.get_rank(), dist.get_world_size()
        if unique_id is None:
            ident = (ctypes.c_uint8 * 128)()
            if rank == 0:
                check(lib.isfm_comm_unique_id(ident))
            obj = [bytes(ident)]
            dist.broadcast_object_list(obj, src=0)
            unique_id = obj[0]
        ident = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        self.handle = c_void_p()
        check(lib.isfm_comm_create(ident, rank, world, byref(self.handle)))
        self.rank, self.world = rank, world

    @property
    def peer_enabled(self):
        """True once the per-iteration sums travel over the peer-memory exchange (NVLink, CUDA
        IPC) instead of ncclAllReduce; decided at the first set_problem on this communicator."""
        return bool(self.handle) and bool(_lib.load().isfm_comm_peer_enabled(self.handle))

    def close(self):
        if self.handle:
            _lib.load().isfm_comm_destroy(self.handle)
            self.handle = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _EngineBase:
    _prefix = ""

    def _fn(self, name):
        return getattr(self.lib, f"isfm_{self._prefix}_{name}")

    def step(self):
        """One ``optimizer.step(input)``.  Returns (loss, stats dict)."""
        loss, st = c_double(), StepStats()
        check(self._fn("step")(self.handle, byref(loss), byref(st)))
        return loss.value, st.as_dict()

    def solve(self, max_iterations, function_tolerance):
        """The reference's outer loop with its windowed stop rule.  Returns the loss history."""
        hist = (c_double * max(max_iterations, 1))()
        n = c_int32()
        check(self._fn("solve")(self.handle, max_iterations, function_tolerance, hist, byref(n)))
        return [hist[i] for i in range(n.value)]

    def cost(self):
        """(sum rho(||r||^2), sum ||r||^2) at the current parameters."""
        a, b = c_double(), c_double()
        check(self._fn("cost")(self.handle, byref(a), byref(b)))
        return a.value, b.value

    def reset_timers(self, enable=True):
        check(self._fn("reset_timers")(self.handle, int(enable)))

    def timers(self):
        ms = (c_double * _lib.N_TIMERS)()
        n = (c_int64 * _lib.N_TIMERS)()
        check(self._fn("get_timers")(self.handle, ms, n))
        return {name: {"ms": ms[i], "launches": n[i]} for i, name in enumerate(_lib.timer_names()) if n[i]}

    def close(self):
        if getattr(self, "handle", None):
            self._fn("destroy")(self.handle)
            self.handle = c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BAEngine(_EngineBase):
    """Bundle adjustment LM (replaces ``bae.optim.LM`` over ``ReprojNonBatched``)."""
    _prefix = "ba"

    def __init__(self, model_id, optimize_poses=True, huber_delta=1.0, dtype=np.float32, pcg_tol=1e-6,
                 pcg_max_iter=0, tr_radius=1e4, tr_max=1e10, tr_up=2.0, tr_down=0.5 ** 4, reject=30,
                 comm=None, stream=None):
        self.lib = _lib.load()
        self.dtype = np.dtype(dtype)
        self.model_id = int(model_id)
        d = BADesc()
        self.lib.isfm_ba_default_desc(byref(d))
        d.dtype = _DTYPES[self.dtype]
        d.model_id = self.model_id
        d.optimize_poses = int(bool(optimize_poses))
        d.reject = reject
        d.huber_delta, d.tr_radius, d.tr_max, d.tr_up, d.tr_down = huber_delta, tr_radius, tr_max, tr_up, tr_down
        d.pcg_tol, d.pcg_max_iter = pcg_tol, pcg_max_iter
        d.stream = _current_stream() if stream is None else stream
        d.comm = comm.handle if comm is not None else None
        self._comm = comm
        self.handle = c_void_p()
        code = self.lib.isfm_ba_create(byref(d), byref(self.handle))
        if code == -2:
            raise NotImplementedError("Unsupported camera model")  # bundle_adjustment.py:47-50
        check(code)
        self.optimize_poses = bool(optimize_poses)
        self.n_cam = self.n_pt = self.n_obs = 0

    def set_problem(self, camera_params, camera_pps, points_3d, points_2d, camera_indices, point_indices):
        keep = []
        self.n_cam, self.n_pt, self.n_obs = int(camera_params.shape[0]), int(points_3d.shape[0]), int(points_2d.shape[0])
        self.cam_width = int(camera_params.shape[1])
        self.d = self.cam_width - 1
        check(self.lib.isfm_ba_set_problem(
            self.handle, self.n_cam, self.n_pt, self.n_obs,
            _buf(camera_params, self.dtype, keep), _buf(camera_pps, self.dtype, keep), _buf(points_3d, self.dtype, keep),
            _buf(points_2d, self.dtype, keep), _buf(camera_indices, np.int32, keep), _buf(point_indices, np.int32, keep)))

    def get_params(self):
        cam = np.empty((self.n_cam, self.cam_width), self.dtype)
        pts = np.empty((self.n_pt, 3), self.dtype)
        check(self.lib.isfm_ba_get_params(self.handle, cam.ctypes.data_as(c_void_p), pts.ctypes.data_as(c_void_p)))
        return cam, pts

    def set_params(self, camera_params=None, points_3d=None):
        keep = []
        check(self.lib.isfm_ba_set_params(self.handle, _buf(camera_params, self.dtype, keep), _buf(points_3d, self.dtype, keep)))

    def structure(self):
        perm = np.empty(self.n_obs, np.int32); cperm = np.empty(self.n_obs, np.int32)
        poff = np.empty(self.n_pt + 1, np.int64); coff = np.empty(self.n_cam + 1, np.int64)
        check(self.lib.isfm_ba_get_structure(self.handle, perm.ctypes.data_as(c_void_p), poff.ctypes.data_as(c_void_p),
                                             cperm.ctypes.data_as(c_void_p), coff.ctypes.data_as(c_void_p)))
        return {"obs_perm": perm, "point_offsets": poff, "cam_perm": cperm, "cam_offsets": coff}

    def matvec_units(self):
        """(owned, total) mat-vec work units of this rank per PCG iteration (owned < total when the
        ranks share one block pattern and the summed matrix is split across them)."""
        owned, total = c_int64(), c_int64()
        check(self.lib.isfm_ba_get_matvec_units(self.handle, byref(owned), byref(total)))
        return owned.value, total.value

    def pcg_phases(self):
        """(ms per phase of the persistent PCG kernel accumulated since creation [matvec, combine,
        exchange, update, coarse, direction, -, -], PCG solves run by it, two-level preconditioner on?)."""
        ms = (c_double * 8)()
        solves, two = c_int64(), c_int32()
        check(self.lib.isfm_ba_get_pcg_phases(self.handle, ms, byref(solves), byref(two)))
        return [ms[i] for i in range(8)], solves.value, bool(two.value)

    def schur_pattern(self):
        nnzb, npairs = c_int64(), c_int64()
        check(self.lib.isfm_ba_get_schur_pattern(self.handle, byref(nnzb), byref(npairs), None, None))
        rp = np.empty(self.n_cam + 1, np.int64); ci = np.empty(nnzb.value, np.int32)
        check(self.lib.isfm_ba_get_schur_pattern(self.handle, byref(nnzb), byref(npairs), rp.ctypes.data_as(c_void_p),
                                                 ci.ctypes.data_as(c_void_p)))
        return {"nnzb": nnzb.value, "n_pairs": npairs.value, "row_ptr": rp, "col_idx": ci}

    _DEBUG = {"residuals": (0, lambda s: (s.n_obs, 2)), "jac_cam": (1, lambda s: (s.n_obs, 2, s.d)),
              "jac_point": (2, lambda s: (s.n_obs, 2, 3)), "weighted_res": (3, lambda s: (s.n_obs, 2)),
              "hpp": (4, lambda s: (s.n_pt, 6)), "gp": (5, lambda s: (s.n_pt, 3)),
              "hcc": (6, lambda s: (s.n_cam, s.d, s.d)), "gc": (7, lambda s: (s.n_cam, s.d)),
              "schur_dense": (8, lambda s: (s.n_cam * s.d, s.n_cam * s.d)), "schur_rhs": (9, lambda s: (s.n_cam * s.d,)),
              "step_cam": (10, lambda s: (s.n_cam, s.d)), "step_point": (11, lambda s: (s.n_pt, 3))}

    def debug(self, what):
        code, shape = self._DEBUG[what]
        out = np.empty(shape(self), self.dtype)
        check(self.lib.isfm_ba_debug_get(self.handle, code, out.ctypes.data_as(c_void_p)))
        return out


class GPEngine(_EngineBase):
    """Global positioning LM (replaces ``bae.optim.LM`` over ``PairwiseNonBatched``)."""
    _prefix = "gp"

    def __init__(self, huber_delta=0.1, dtype=np.float32, pcg_tol=1e-6, pcg_max_iter=0, tr_radius=1e3, tr_max=1e8,
                 tr_up=2.0, tr_down=0.5 ** 4, reject=30, optimize_scales=True, comm=None, stream=None):
        self.lib = _lib.load()
        self.dtype = np.dtype(dtype)
        d = GPDesc()
        self.lib.isfm_gp_default_desc(byref(d))
        d.dtype = _DTYPES[self.dtype]
        d.reject = reject
        d.huber_delta, d.tr_radius, d.tr_max, d.tr_up, d.tr_down = huber_delta, tr_radius, tr_max, tr_up, tr_down
        d.pcg_tol, d.pcg_max_iter = pcg_tol, pcg_max_iter
        d.optimize_scales = int(bool(optimize_scales))
        d.stream = _current_stream() if stream is None else stream
        d.comm = comm.handle if comm is not None else None
        self._comm = comm
        self.handle = c_void_p()
        check(self.lib.isfm_gp_create(byref(d), byref(self.handle)))
        self.n_cam = self.n_pt = self.n_obs = 0

    def set_problem(self, camera_translations, points_3d, scales, translations, camera_indices, point_indices,
                    is_calibrated, scale_fixed=None):
        keep = []
        self.n_cam, self.n_pt, self.n_obs = int(camera_translations.shape[0]), int(points_3d.shape[0]), int(translations.shape[0])
        check(self.lib.isfm_gp_set_problem(
            self.handle, self.n_cam, self.n_pt, self.n_obs,
            _buf(camera_translations, self.dtype, keep), _buf(points_3d, self.dtype, keep),
            _buf(np.asarray(scales).reshape(-1) if not _is_torch(scales) else scales.reshape(-1), self.dtype, keep),
            _buf(translations, self.dtype, keep), _buf(camera_indices, np.int32, keep), _buf(point_indices, np.int32, keep),
            _buf(is_calibrated, np.uint8, keep), _buf(scale_fixed, np.uint8, keep)))

    def get_params(self):
        c = np.empty((self.n_cam, 3), self.dtype); p = np.empty((self.n_pt, 3), self.dtype)
        s = np.empty((self.n_obs, 1), self.dtype)
        check(self.lib.isfm_gp_get_params(self.handle, c.ctypes.data_as(c_void_p), p.ctypes.data_as(c_void_p),
                                          s.ctypes.data_as(c_void_p)))
        return c, p, s


def trim_cache():
    """Return the library's cached device blocks and its private memory pools to the driver
    (call at pipeline stage boundaries when other GPU stages need the memory)."""
    _lib.load().isfm_trim_cache()


def set_cache_limit(n_bytes):
    """Bound the process-wide cache of freed device blocks (default 4 GB; 0 disables it)."""
    _lib.load().isfm_set_cache_limit(int(n_bytes))


def partition_points(point_offsets, world):
    """Contiguous point ranges balanced by observation count (C ABI isfm_partition_points)."""
    lib = _lib.load()
    off = np.ascontiguousarray(point_offsets, dtype=np.int64)
    out = np.empty(world + 1, np.int64)
    check(lib.isfm_partition_points(off.ctypes.data_as(ctypes.POINTER(c_int64)), off.shape[0] - 1, world,
                                    out.ctypes.data_as(ctypes.POINTER(c_int64))))
    return out
