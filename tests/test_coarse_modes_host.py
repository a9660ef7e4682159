"""The prolongation of the two-level PCG preconditioner (csrc/coarse.cuh: the seven similarity modes of
a camera cluster in the cameras' left-perturbation tangent) against the oracle's Jacobian: moving the
WHOLE scene by a similarity -- cameras by the mode formula, points along with them -- leaves every
residual unchanged, so J [P m ; dX(m)] = 0 for every mode m.  This pins the formula's tangent convention
(dphi = -R w, dtau = -R v - tt x (R w) + s tt, tt = t + R c0) to the oracle's retraction; the CUDA
kernel that evaluates it is covered on the GPU by the iteration counts it buys."""
import numpy as np
from scipy.spatial.transform import Rotation

from instantsfm_b200.synthetic import make_ba_problem
from oracle.ba import BAProblem


def similarity_modes(cam, c0):
    """[n_cam, 6, 7]: columns v (3), w (3), s (1); rows dtau (3), dphi (3) -- coarse.cuh header."""
    n = cam.shape[0]
    R = Rotation.from_quat(cam[:, 3:7]).as_matrix()
    tt = cam[:, :3] + np.einsum("nij,j->ni", R, c0)
    P = np.zeros((n, 6, 7))
    for k in range(3):
        e = np.zeros(3); e[k] = 1.0
        Re = np.einsum("nij,j->ni", R, e)
        P[:, :3, k] = -Re                                  # translation v = e_k: dtau = -R v
        P[:, :3, 3 + k] = -np.cross(tt, Re)                # rotation w = e_k:    dtau = -tt x (R w)
        P[:, 3:, 3 + k] = -Re                              #                      dphi = -R w
    P[:, :3, 6] = tt                                       # scale s:             dtau = s tt
    return P


def device_modes(cam, c0, dtype=np.float64):
    """math.cuh::similarity_modes -- what coarse_modes_kernel runs per camera -- compiled for the host."""
    import ctypes
    from tests.hostcheck.build import load
    lib = load()
    pose = np.ascontiguousarray(cam[:, :7], dtype=dtype)
    c0 = np.ascontiguousarray(c0, dtype=np.float64)
    P = np.zeros((pose.shape[0], 6, 7), dtype=dtype)
    fn = lib.hc_similarity_modes_f64 if dtype == np.float64 else lib.hc_similarity_modes_f32
    fn(ctypes.c_long(pose.shape[0]), pose.ctypes.data_as(ctypes.c_void_p), c0.ctypes.data_as(ctypes.c_void_p),
       P.ctypes.data_as(ctypes.c_void_p))
    return P


def test_device_mode_blocks_match_the_formula():
    a = make_ba_problem(10, 300, 1600, seed=44)
    c0 = np.array([0.3, -0.2, 0.1])
    want = similarity_modes(a.camera_params, c0)
    assert np.abs(device_modes(a.camera_params, c0) - want).max() <= 1e-13 * max(1.0, np.abs(want).max())
    assert np.abs(device_modes(a.camera_params, c0, np.float32) - want).max() <= 1e-6 * max(1.0, np.abs(want).max())


def test_global_similarity_modes_are_in_the_null_space_of_the_jacobian():
    a = make_ba_problem(10, 300, 1600, seed=44)
    pb = BAProblem(a.model_id, a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
    J = pb.jacobian()
    d, n_cam = pb.d, pb.n_cam
    c0 = np.array([0.3, -0.2, 0.1])          # any centre: the formula carries it
    P = device_modes(pb.cam, c0)             # the device arithmetic itself (host build)
    X = pb.pts - c0
    scale = np.abs(J).sum(axis=1).max()
    for m in range(7):
        delta = np.zeros(n_cam * d + 3 * pb.n_pt)
        Dc = np.zeros((n_cam, d)); Dc[:, :6] = P[:, :, m]
        delta[:n_cam * d] = Dc.reshape(-1)
        if m < 3:
            dX = np.tile(np.eye(3)[m], (pb.n_pt, 1))                      # X' = X + v
        elif m < 6:
            dX = np.cross(np.eye(3)[m - 3], X)                            # X' = c0 + Exp(w)(X - c0)
        else:
            dX = X                                                        # X' = c0 + s (X - c0)
        delta[n_cam * d:] = dX.reshape(-1)
        # scale: the camera frame is scaled too (x_cam -> s x_cam), the projection does not change
        r = J @ delta
        assert np.abs(r).max() <= 1e-9 * scale * max(1.0, np.abs(delta).max()), (m, np.abs(r).max())


def test_modes_match_a_finite_similarity():
    """Finite check of the same statement through the oracle's retraction: residuals after a small
    similarity of the whole scene equal the residuals before it (to second order)."""
    a = make_ba_problem(8, 200, 1000, seed=45)
    pb = BAProblem(a.model_id, a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
    r0 = pb.residuals().copy()
    c0 = pb.pts.mean(0)
    P = similarity_modes(pb.cam, c0)
    eps = 1e-6
    m = np.array([0.3, -0.1, 0.2, 0.5, -0.4, 0.3, 0.7]) * eps
    Dc = np.zeros((pb.n_cam, pb.d)); Dc[:, :6] = np.einsum("nij,j->ni", P, m)
    X = pb.pts - c0
    dX = m[:3] + np.cross(m[3:6], X) + m[6] * X
    pb.retract(np.concatenate([Dc.reshape(-1), dX.reshape(-1)]))
    assert np.abs(pb.residuals() - r0).max() <= 1e-8 * max(1.0, np.abs(r0).max())
