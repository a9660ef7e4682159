"""SE(3) helpers of the oracle (torch, fp64, differentiable).  TEST INFRASTRUCTURE.

Storage convention pose7 = [tx, ty, tz, qx, qy, qz, qw], confirmed by the reference's
call sites: ``pp.mat2SE3(world2cam).tensor()`` (bundle_adjustment.py:71) and
``[world2cam[:3, 3], scipy as_quat() (xyzw)]`` (track_retriangulation.py:65-67).

Tangent convention (UNVERIFIED, public pypose): se3 = [tau(3), phi(3)],
``Exp([tau, phi]) = (J_l(phi) tau, exp_quat(phi))`` and the LM update is the left
retraction ``X <- Exp(delta) * X``.
"""
import torch


def quat_rotate(q, p):
    """Rotate p[...,3] by unit quaternion q[...,4] = (x, y, z, w)."""
    qv = q[..., :3]
    qw = q[..., 3:4]
    t = 2.0 * torch.cross(qv, p, dim=-1)
    return p + qw * t + torch.cross(qv, t, dim=-1)


def rotate_quat(points, pose7):
    """bae.utils.ba.rotate_quat restated: R(q) p + t  (cost_function.py:34 et al.)."""
    return quat_rotate(pose7[..., 3:7], points) + pose7[..., :3]


def quat_mul(a, b):
    """Hamilton product a (x) b, both (x, y, z, w)."""
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz,
    ], dim=-1)


def quat_to_mat(q):
    x, y, z, w = q.unbind(-1)
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1),
    ], dim=-2)


def _sinc_terms(theta2):
    """A = sin t / t, B = (1 - cos t) / t^2, C = (t - sin t) / t^3, series near 0."""
    small = theta2 < 1e-8
    t2 = torch.where(small, torch.ones_like(theta2), theta2)
    t = torch.sqrt(t2)
    A = torch.where(small, 1 - theta2 / 6, torch.sin(t) / t)
    B = torch.where(small, 0.5 - theta2 / 24, (1 - torch.cos(t)) / t2)
    C = torch.where(small, 1.0 / 6 - theta2 / 120, (t - torch.sin(t)) / (t2 * t))
    return A, B, C


def so3_exp_quat(phi):
    """exp: so(3) -> unit quaternion (x, y, z, w)."""
    theta2 = (phi * phi).sum(-1, keepdim=True)
    small = theta2 < 1e-8
    t2 = torch.where(small, torch.ones_like(theta2), theta2)
    half = 0.5 * torch.sqrt(t2)
    k = torch.where(small, 0.5 - theta2 / 48, torch.sin(half) / torch.sqrt(t2))
    w = torch.where(small, 1 - theta2 / 8, torch.cos(half))
    return torch.cat([k * phi, w], dim=-1)


def se3_exp(delta):
    """Exp([tau, phi]) -> pose7 = [J_l(phi) tau, exp_quat(phi)]."""
    tau, phi = delta[..., :3], delta[..., 3:6]
    theta2 = (phi * phi).sum(-1, keepdim=True)
    _, B, C = _sinc_terms(theta2)
    pxt = torch.cross(phi, tau, dim=-1)
    t = tau + B * pxt + C * torch.cross(phi, pxt, dim=-1)
    return torch.cat([t, so3_exp_quat(phi)], dim=-1)


def se3_mul(a, b):
    """Compose poses: (a * b)(p) = a(b(p))."""
    t = quat_rotate(a[..., 3:7], b[..., :3]) + a[..., :3]
    q = quat_mul(a[..., 3:7], b[..., 3:7])
    return torch.cat([t, q], dim=-1)


def se3_retract(pose7, delta6):
    """Left retraction X <- Exp(delta) * X used by the LM parameter update."""
    out = se3_mul(se3_exp(delta6), pose7)
    q = out[..., 3:7]
    return torch.cat([out[..., :3], q / q.norm(dim=-1, keepdim=True)], dim=-1)


def pose7_to_matrix(pose7):
    """pp.SE3(x).matrix() restated (bundle_adjustment.py:27): 4x4 world2cam."""
    R = quat_to_mat(pose7[..., 3:7])
    n = pose7.shape[:-1]
    M = torch.zeros(*n, 4, 4, dtype=pose7.dtype)
    M[..., :3, :3] = R
    M[..., :3, 3] = pose7[..., :3]
    M[..., 3, 3] = 1.0
    return M


def matrix_to_pose7(M):
    """pp.mat2SE3(world2cam).tensor() restated (bundle_adjustment.py:71), numpy in/out."""
    import numpy as np
    from scipy.spatial.transform import Rotation
    M = np.asarray(M, dtype=np.float64)
    q = Rotation.from_matrix(M[..., :3, :3]).as_quat()  # xyzw
    return np.concatenate([M[..., :3, 3], q], axis=-1)
