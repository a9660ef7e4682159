"""Projection residuals of the oracle (torch fp64).  TEST INFRASTRUCTURE.

Restates /root/reference/instantsfm/utils/cost_function.py:32-177 (BA reprojection, nine
implemented COLMAP models) and :22-29 (GP ``pairwise_cost``).  Every model shares
``y = rotate_quat(X, cam[:7]); u = y.xy / y.z`` followed by a model-specific distortion
and ``* focal + pp``.  Intrinsics are the camera row with the principal point removed
(bundle_adjustment.py:75-80), so their order is ``get_camera_model_info(...)['optimize']``
order (scene/defs.py:115-140) and the reference addresses them from the END of the row.

Pinned against the reference's own functions by tests/golden/reference_cost_functions.npz.
"""
import torch

from .lie import rotate_quat

# CameraModelId.value -> (name, number of intrinsics kept after removing pp)
# scene/defs.py:101-113 (ids) and :115-140 ('optimize' lists).
MODEL_INFO = {
    0: ("SIMPLE_PINHOLE", 1),
    1: ("PINHOLE", 2),
    2: ("SIMPLE_RADIAL", 2),
    3: ("RADIAL", 3),
    4: ("OPENCV", 6),
    5: ("OPENCV_FISHEYE", 6),
    6: ("FULL_OPENCV", 10),
    8: ("SIMPLE_RADIAL_FISHEYE", 2),
    9: ("RADIAL_FISHEYE", 3),
}
UNSUPPORTED = {7: "FOV", 10: "THIN_PRISM_FISHEYE"}  # cost_function.py:128,182 raise

# 'pp' and 'optimize' index lists per model, scene/defs.py:118-138
PP_INDICES = {0: [1, 2], 1: [2, 3], 2: [1, 2], 3: [1, 2], 4: [2, 3], 5: [2, 3], 6: [2, 3],
              8: [1, 2], 9: [1, 2]}
NUM_PARAMS = {0: 3, 1: 4, 2: 4, 3: 5, 4: 8, 5: 8, 6: 12, 8: 4, 9: 5}


def n_intrinsics(model_id):
    return MODEL_INFO[model_id][1]


def _fisheye(u, r2):
    # cost_function.py:96-98 / :159-161 / :173-175 : u * atan(r) / r   (0/0 at r == 0)
    r = torch.sqrt(r2)
    return u * torch.atan(r) / r


def _tangential(u, r2, p):
    # cost_function.py:78-81 / :117-120 : 2 p (u v) + flip(p) (r2 + 2 u^2)
    uv = (u[..., 0] * u[..., 1]).unsqueeze(-1)
    return 2 * p * uv + p.flip(-1) * (r2 + 2 * u * u)


def distort_and_scale(model_id, u, k):
    """u[...,2] normalised image point, k[...,n_intr] intrinsics (pp removed)."""
    r2 = (u * u).sum(-1, keepdim=True)
    if model_id == 0:      # :33-38
        return u * k[..., -1:]
    if model_id == 1:      # :41-46
        return u * k[..., -2:]
    if model_id == 2:      # :49-56
        return u * (1 + k[..., -1:] * r2) * k[..., -2:-1]
    if model_id == 3:      # :59-67
        return u * (1 + k[..., -2:-1] * r2 + k[..., -1:] * r2 ** 2) * k[..., -3:-2]
    if model_id == 4:      # :70-84
        ff, k1, k2, p = k[..., -6:-4], k[..., -4:-3], k[..., -3:-2], k[..., -2:]
        d = u * (k1 * r2 + k2 * r2 ** 2) + _tangential(u, r2, p)
        return (u + d) * ff
    if model_id == 5:      # :87-102  (k4 = k[..., -1] is ignored by the reference)
        ff, k1, k2, k3 = k[..., -6:-4], k[..., -4:-3], k[..., -3:-2], k[..., -2:-1]
        return _fisheye(u, r2) * (1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3) * ff
    if model_id == 6:      # :105-123
        ff, k1, k2, p = k[..., -10:-8], k[..., -8:-7], k[..., -7:-6], k[..., -6:-4]
        k3, k4, k5, k6 = k[..., -4:-3], k[..., -3:-2], k[..., -2:-1], k[..., -1:]
        radial = (1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3) / (1 + k4 * r2 + k5 * r2 ** 2 + k6 * r2 ** 3) - 1
        d = u * radial + _tangential(u, r2, p)
        return (u + d) * ff
    if model_id == 8:      # :153-163
        return _fisheye(u, r2) * (1 + k[..., -1:] * r2) * k[..., -2:-1]
    if model_id == 9:      # :166-177
        return _fisheye(u, r2) * (1 + k[..., -2:-1] * r2 + k[..., -1:] * r2 ** 2) * k[..., -3:-2]
    raise NotImplementedError("Unsupported camera model")  # bundle_adjustment.py:47-50


def reproject(model_id, points, camera_params, pp):
    """reproject_funcs[model_id](points, camera_params, pp) restated (cost_function.py:206-208)."""
    if model_id not in MODEL_INFO:
        raise NotImplementedError("Unsupported camera model")
    y = rotate_quat(points, camera_params[..., :7])
    u = y[..., :2] / y[..., 2:3]
    return distort_and_scale(model_id, u, camera_params[..., 7:]) + pp


def pairwise_cost(points, camera_translations, scales, translations, is_calibrated):
    """cost_function.py:22-29: w * (d - s (X - c)), w = 1 if calibrated else 0.5."""
    r = translations - scales * (points - camera_translations)
    w = torch.where(is_calibrated, 1.0, 0.5).unsqueeze(-1).to(r.dtype)
    return r * w
