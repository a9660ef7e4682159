"""Host-side SE(3) helpers (numpy fp64) used by the drop-in processors and the synthetic
generator.  Storage convention pose7 = [t(3), q = (x, y, z, w)] as established by the
reference's callers (bundle_adjustment.py:71 ``pp.mat2SE3(world2cam).tensor()``)."""
import numpy as np
from scipy.spatial.transform import Rotation


def quat_to_mat(q):
    q = np.asarray(q, dtype=np.float64)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - z * w); R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w); R[..., 2, 1] = 2 * (y * z + x * w); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def quat_mul(a, b):
    ax, ay, az, aw = np.moveaxis(a, -1, 0)
    bx, by, bz, bw = np.moveaxis(b, -1, 0)
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def rotvec_to_quat(phi):
    return Rotation.from_rotvec(np.asarray(phi, dtype=np.float64)).as_quat()


def matrices_to_pose7(M):
    """4x4 world2cam (stack) -> [t, q_xyzw]; what pp.mat2SE3(M).tensor() returns."""
    M = np.asarray(M, dtype=np.float64)
    q = Rotation.from_matrix(M[..., :3, :3]).as_quat()
    return np.concatenate([M[..., :3, 3], q], axis=-1)


def pose7_to_matrices(pose7):
    """[t, q_xyzw] -> 4x4 world2cam; what pp.SE3(x).matrix() returns (bundle_adjustment.py:27)."""
    pose7 = np.asarray(pose7, dtype=np.float64)
    M = np.zeros(pose7.shape[:-1] + (4, 4))
    M[..., :3, :3] = quat_to_mat(pose7[..., 3:7])
    M[..., :3, 3] = pose7[..., :3]
    M[..., 3, 3] = 1.0
    return M


def transform_points(pose7, X):
    """R(q) X + t, row-wise."""
    R = quat_to_mat(pose7[..., 3:7])
    return np.einsum("...ij,...j->...i", R, X) + pose7[..., :3]
