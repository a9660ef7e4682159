"""CPU check of the arithmetic the CUDA kernels inline (csrc/math.cuh compiled for the
host, tests/hostcheck) against the oracle: residuals + analytic/dual-number Jacobians for
all nine camera models vs torch autograd, SE(3) retraction, small SPD inverses, Huber."""
import ctypes

import numpy as np
import pytest
import torch

from instantsfm_b200.synthetic import make_ba_problem, N_INTR
from oracle import lie
from oracle.ba import BAProblem
from oracle.lm import huber_rho, huber_drho
from tests.hostcheck.build import load

lib = load()
P = ctypes.c_void_p


def _ptr(a):
    return a.ctypes.data_as(P)


@pytest.mark.parametrize("model_id", sorted(N_INTR))
def test_linearize_matches_autograd(model_id):
    a = make_ba_problem(8, 100, 400, seed=10 + model_id, model_id=model_id)
    pb = BAProblem(model_id, a.camera_params, a.camera_pps + 3.0, a.points_3d, a.points_2d,
                   a.camera_indices, a.point_indices)
    r_ref, Jc_ref, Jp_ref = pb.blocks()
    n, d = pb.n_obs, pb.d
    cam = np.ascontiguousarray(pb.cam[pb.ci]); pp = np.ascontiguousarray(pb.pps[pb.ci])
    X = np.ascontiguousarray(pb.pts[pb.pi]); obs = pb.obs
    r = np.empty((n, 2)); Jc = np.empty((n, 2, d)); Jp = np.empty((n, 2, 3))
    rc = lib.hc_linearize_f64(model_id, ctypes.c_long(n), _ptr(cam), _ptr(pp), _ptr(X), _ptr(obs), _ptr(r), _ptr(Jc), _ptr(Jp))
    assert rc == 0
    np.testing.assert_allclose(r, r_ref, rtol=1e-11, atol=1e-9)
    scale = np.abs(Jc_ref).max()
    np.testing.assert_allclose(Jc, Jc_ref, rtol=1e-9, atol=1e-9 * scale)
    np.testing.assert_allclose(Jp, Jp_ref, rtol=1e-9, atol=1e-9 * np.abs(Jp_ref).max())
    r2 = np.empty((n, 2))
    lib.hc_residual_f64(model_id, ctypes.c_long(n), _ptr(cam), _ptr(pp), _ptr(X), _ptr(obs), _ptr(r2))
    np.testing.assert_allclose(r2, r_ref, rtol=1e-11, atol=1e-9)
    # fp32 build of the same code: 1e-4 relative (the north-star tolerance) with margin
    c32, p32, X32, o32 = (v.astype(np.float32) for v in (cam, pp, X, obs))
    r32 = np.empty((n, 2), np.float32); Jc32 = np.empty((n, 2, d), np.float32); Jp32 = np.empty((n, 2, 3), np.float32)
    lib.hc_linearize_f32(model_id, ctypes.c_long(n), _ptr(c32), _ptr(p32), _ptr(X32), _ptr(o32), _ptr(r32), _ptr(Jc32), _ptr(Jp32))
    # the fp32 inputs themselves are rounded: compare against the oracle at the rounded inputs
    pb32 = BAProblem(model_id, a.camera_params.astype(np.float32), (a.camera_pps + 3.0).astype(np.float32),
                     a.points_3d.astype(np.float32), a.points_2d.astype(np.float32), a.camera_indices, a.point_indices)
    _, Jc_r32, Jp_r32 = pb32.blocks()
    assert np.abs(Jc32 - Jc_r32).max() <= 2e-4 * np.abs(Jc_r32).max()
    assert np.abs(Jp32 - Jp_r32).max() <= 2e-4 * np.abs(Jp_r32).max()


def test_unsupported_models_rejected():
    z = np.zeros(16)
    for model_id in (7, 10, -1, 11):
        assert lib.hc_residual_f64(model_id, ctypes.c_long(0), _ptr(z), _ptr(z), _ptr(z), _ptr(z), _ptr(z)) == -2


def test_se3_retract_matches_oracle():
    rng = np.random.default_rng(3)
    n = 200
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    pose = np.concatenate([rng.normal(size=(n, 3)), q], 1)
    delta = rng.normal(size=(n, 6)) * np.repeat([1.0, 1e-3, 1e-6, 0.0], n // 4)[:, None]
    out = np.empty((n, 7))
    lib.hc_se3_retract_f64(ctypes.c_long(n), _ptr(pose), _ptr(delta), _ptr(out))
    ref = lie.se3_retract(torch.from_numpy(pose), torch.from_numpy(delta)).numpy()
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-12)
    out32 = np.empty((n, 7), np.float32)
    p32, d32 = pose.astype(np.float32), delta.astype(np.float32)
    lib.hc_se3_retract_f32(ctypes.c_long(n), _ptr(p32), _ptr(d32), _ptr(out32))
    np.testing.assert_allclose(out32, ref, rtol=0, atol=5e-6)


def test_retraction_is_first_order_consistent_with_jacobian():
    """r(x (+) eps delta) - r(x) ~ eps J delta: pins tangent convention and column order."""
    a = make_ba_problem(5, 30, 100, seed=2)
    pb = BAProblem(3, a.camera_params, a.camera_pps, a.points_3d, a.points_2d, a.camera_indices, a.point_indices)
    J = pb.jacobian()
    rng = np.random.default_rng(0)
    D = rng.normal(size=J.shape[1])
    r0 = pb.residuals().reshape(-1)
    eps = 1e-6
    pb.retract(eps * D)
    r1 = pb.residuals().reshape(-1)
    np.testing.assert_allclose((r1 - r0) / eps, J @ D, rtol=1e-4, atol=1e-4 * np.abs(J @ D).max())


@pytest.mark.parametrize("D", [3, 7, 8, 9, 12, 16])
def test_spd_inverse(D):
    rng = np.random.default_rng(D)
    for _ in range(20):
        B = rng.normal(size=(D, 2 * D))
        A = B @ B.T + 1e-3 * np.eye(D)
        M = A.copy()
        assert lib.hc_spd_inverse(D, _ptr(M)) == 1
        np.testing.assert_allclose(M, np.linalg.inv(A), rtol=1e-8, atol=1e-10 * np.abs(np.linalg.inv(A)).max())


def test_sym3_inverse_and_huber():
    rng = np.random.default_rng(1)
    B = rng.normal(size=(50, 3, 4))
    A = B @ B.transpose(0, 2, 1)
    h = np.ascontiguousarray(np.stack([A[:, 0, 0], A[:, 0, 1], A[:, 0, 2], A[:, 1, 1], A[:, 1, 2], A[:, 2, 2]], 1))
    inv = np.empty_like(h)
    lib.hc_sym3_inverse_f64(ctypes.c_long(50), _ptr(h), _ptr(inv))
    Ai = np.linalg.inv(A)
    ref = np.stack([Ai[:, 0, 0], Ai[:, 0, 1], Ai[:, 0, 2], Ai[:, 1, 1], Ai[:, 1, 2], Ai[:, 2, 2]], 1)
    np.testing.assert_allclose(inv, ref, rtol=1e-9)
    s = np.concatenate([rng.uniform(0, 4, 100), [0.0, 1.0]])
    rho = np.empty_like(s); w = np.empty_like(s)
    lib.hc_huber_f64(ctypes.c_long(s.size), _ptr(s), ctypes.c_double(1.0), _ptr(rho), _ptr(w))
    np.testing.assert_allclose(rho, huber_rho(s, 1.0), rtol=1e-14)
    np.testing.assert_allclose(w, np.sqrt(huber_drho(s, 1.0)), rtol=1e-14)
