"""ctypes binding of include/isfm_b200.h.  There is no CPU fallback: importing this module
without the built CUDA library raises, and creating a handle without a GPU fails loudly."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# ISFM_LIB_PATH: an alternative build of the same library (kernel-variant A/B runs); never a fallback
LIB_PATH = os.environ.get("ISFM_LIB_PATH") or os.path.join(_HERE, "lib", "libisfm_b200.so")
N_TIMERS = 16


class IsfmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"isfm_b200 error {code}: {msg}")
        self.code = code


class StepStats(Structure):
    _fields_ = [("loss_before", c_double), ("loss", c_double), ("damping", c_double), ("quality", c_double),
                ("model_term", c_double), ("step_norm_cam", c_double), ("trials", c_int32), ("rejects", c_int32),
                ("pcg_iters", c_int32), ("accepted", c_int32), ("pcg_status", c_int32), ("reserved", c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class BADesc(Structure):
    _fields_ = [("dtype", c_int32), ("model_id", c_int32), ("optimize_poses", c_int32), ("reject", c_int32),
                ("huber_delta", c_double), ("tr_radius", c_double), ("tr_max", c_double), ("tr_up", c_double),
                ("tr_down", c_double), ("pcg_tol", c_double), ("pcg_max_iter", c_int32), ("reserved", c_int32),
                ("stream", c_void_p), ("comm", c_void_p)]


class GPDesc(Structure):
    _fields_ = [("dtype", c_int32), ("reject", c_int32), ("huber_delta", c_double), ("tr_radius", c_double),
                ("tr_max", c_double), ("tr_up", c_double), ("tr_down", c_double), ("pcg_tol", c_double),
                ("pcg_max_iter", c_int32), ("optimize_scales", c_int32), ("stream", c_void_p), ("comm", c_void_p)]


# name -> (restype, argtypes); must list every symbol include/isfm_b200.h declares
SIGNATURES = {
    "isfm_version": (c_char_p, []),
    "isfm_last_error": (c_char_p, []),
    "isfm_launch_count": (c_int64, []),
    "isfm_trim_cache": (None, []),
    "isfm_set_cache_limit": (None, [ctypes.c_uint64]),
    "isfm_ba_get_pcg_phases": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int64), POINTER(c_int32)]),
    "isfm_timer_name": (c_char_p, [c_int32]),
    "isfm_comm_unique_id": (c_int, [POINTER(c_uint8)]),
    "isfm_comm_create": (c_int, [POINTER(c_uint8), c_int, c_int, POINTER(c_void_p)]),
    "isfm_comm_destroy": (None, [c_void_p]),
    "isfm_comm_peer_enabled": (c_int, [c_void_p]),
    "isfm_partition_points": (c_int, [POINTER(c_int64), c_int64, c_int, POINTER(c_int64)]),
    "isfm_ba_default_desc": (None, [POINTER(BADesc)]),
    "isfm_ba_create": (c_int, [POINTER(BADesc), POINTER(c_void_p)]),
    "isfm_ba_destroy": (None, [c_void_p]),
    "isfm_ba_set_problem": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "isfm_ba_step": (c_int, [c_void_p, POINTER(c_double), POINTER(StepStats)]),
    "isfm_ba_solve": (c_int, [c_void_p, c_int32, c_double, POINTER(c_double), POINTER(c_int32)]),
    "isfm_ba_get_params": (c_int, [c_void_p, c_void_p, c_void_p]),
    "isfm_ba_set_params": (c_int, [c_void_p, c_void_p, c_void_p]),
    "isfm_ba_cost": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double)]),
    "isfm_ba_get_structure": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "isfm_ba_get_schur_pattern": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64), c_void_p, c_void_p]),
    "isfm_ba_get_matvec_units": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64)]),
    "isfm_ba_debug_get": (c_int, [c_void_p, c_int32, c_void_p]),
    "isfm_ba_get_timers": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int64)]),
    "isfm_ba_reset_timers": (c_int, [c_void_p, c_int32]),
    "isfm_gp_default_desc": (None, [POINTER(GPDesc)]),
    "isfm_gp_create": (c_int, [POINTER(GPDesc), POINTER(c_void_p)]),
    "isfm_gp_destroy": (None, [c_void_p]),
    "isfm_gp_set_problem": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "isfm_gp_step": (c_int, [c_void_p, POINTER(c_double), POINTER(StepStats)]),
    "isfm_gp_solve": (c_int, [c_void_p, c_int32, c_double, POINTER(c_double), POINTER(c_int32)]),
    "isfm_gp_get_params": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "isfm_gp_cost": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double)]),
    "isfm_gp_get_timers": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int64)]),
    "isfm_gp_reset_timers": (c_int, [c_void_p, c_int32]),
    "isfm_filter_observations": (c_int, [c_int32, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_double, c_void_p, c_void_p]),
    "isfm_reprojection_test": (c_int, [c_int32, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_double, c_double, c_void_p, c_void_p, c_void_p]),
    "isfm_filter_triangulation_angle": (c_int, [c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_double,
                                                c_void_p, c_void_p]),
    "isfm_filter_reprojection": (c_int, [c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_void_p]),
    "isfm_undistort_features": (c_int, [c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def load():
    """Load libisfm_b200.so (built by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  instantsfm_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise IsfmError(code, load().isfm_last_error().decode())


def timer_names():
    lib = load()
    return [lib.isfm_timer_name(i).decode() for i in range(N_TIMERS)]
