"""The reference's BA loop restated with torch tensors on ANY device (fp64).  TEST / BASELINE
INFRASTRUCTURE -- never imported by the product package.

This is the denominator of BASELINE.json's ">= 20x the reference's GPU BA" target
(last bullet): `bundle_adjustment.py:115-142` -- LM(strategy=TrustRegion,
solver=PCG(tol=1e-5), kernel=Huber, reject=30).step(input) in a loop -- with the same
semantics as oracle/lm.py (full camera + point system, FastTriggs, multiplicative cumulative
damping of the clamped diagonal, scalar-Jacobi PCG, x0 = 0, ||r|| < tol ||b||), but every
operation is a torch op, so `device="cuda"` runs it on the GPU the way the reference's eager
PyTorch path does.  bae itself cannot be installed here (no network).  Two ways to apply the
normal matrix inside PCG, same arithmetic:

* ``mode="sparse"`` (default, the GPU denominator): J and J^T as torch sparse CSR tensors, two
  cuSPARSE SpMVs per PCG iteration (A p = J^T (J p)) -- the library route bae's block-sparse
  J^T J product takes, with FEWER non-zeros streamed per iteration (2 x 12 per residual row
  instead of the ~54 per observation of the explicit J^T J), i.e. a denominator that flatters the
  reference, never the other way round;
* ``mode="blocks"``: gather + batched 2x9 / 2x3 products + index_add_ on the per-observation dense
  blocks (eager-torch style; on CUDA its index_add_ serialises on fp64 atomics -- much slower).

Jacobians: torch.autograd through the left retraction of oracle/lie.py (one backward pass per
residual component), as in oracle/ba.py.
"""
import time

import torch

from . import lie
from .camera_models import n_intrinsics, reproject


class TorchRefBA:
    def __init__(self, model_id, camera_params, camera_pps, points_3d, points_2d, camera_indices, point_indices,
                 huber_delta=1.0, pcg_tol=1e-5, reject=30, device="cpu", tr_radius=1e4, tr_max=1e10, tr_up=2.0,
                 tr_down=0.5 ** 4, mode="sparse"):
        self.mode = mode
        dev, f64 = torch.device(device), torch.float64
        self.dev = dev
        self.model_id = int(model_id)
        self.ni = n_intrinsics(self.model_id)
        self.d = 6 + self.ni
        self.cam = torch.as_tensor(camera_params, dtype=f64).to(dev).clone()
        self.pps = torch.as_tensor(camera_pps, dtype=f64).to(dev)
        self.pts = torch.as_tensor(points_3d, dtype=f64).to(dev).clone()
        self.obs = torch.as_tensor(points_2d, dtype=f64).to(dev)
        self.ci = torch.as_tensor(camera_indices).to(dev).long()
        self.pi = torch.as_tensor(point_indices).to(dev).long()
        self.n_cam, self.n_pt, self.n_obs = self.cam.shape[0], self.pts.shape[0], self.obs.shape[0]
        self.delta, self.pcg_tol, self.reject = float(huber_delta), float(pcg_tol), int(reject)
        # pp.optim.strategy.TrustRegion (bundle_adjustment.py:116); defaults as in oracle/lm.py
        self.damping, self.tr_max, self.tr_up, self.tr_down0, self.tr_down = 1.0 / tr_radius, tr_max, tr_up, tr_down, tr_down
        self.loss = None
        self.pcg_iters = 0
        self.pcg_seconds = 0.0

    # -- model ---------------------------------------------------------------------------
    def residuals(self, cam=None, pts=None):
        cam = self.cam if cam is None else cam
        pts = self.pts if pts is None else pts
        with torch.no_grad():
            return reproject(self.model_id, pts[self.pi], cam[self.ci], self.pps[self.ci]) - self.obs

    def blocks(self):
        ni, d = self.ni, self.d
        cam, X, pp = self.cam[self.ci], self.pts[self.pi], self.pps[self.ci]
        z = torch.cat([torch.zeros(self.n_obs, 6, dtype=torch.float64, device=self.dev), cam[:, 7:], X], dim=1)
        z.requires_grad_(True)
        pose = lie.se3_retract(cam[:, :7], z[:, :6])
        proj = reproject(self.model_id, z[:, 6 + ni:], torch.cat([pose, z[:, 6:6 + ni]], dim=1), pp)
        J = torch.stack([torch.autograd.grad(proj[:, k].sum(), z, retain_graph=(k == 0))[0] for k in range(2)], dim=1)
        r = (proj.detach() - self.obs)
        return r, J[:, :, :d].contiguous(), J[:, :, d:].contiguous()

    def _rho(self, r):
        s = (r * r).sum(-1)
        rs = torch.sqrt(s)
        return torch.where(rs < self.delta, s, 2.0 * self.delta * rs - self.delta ** 2).sum()

    # -- linear algebra on the per-observation blocks ---------------------------------------
    def _J(self, Jc, Jp, dc, dp):
        return torch.einsum("nkd,nd->nk", Jc, dc[self.ci]) + torch.einsum("nkd,nd->nk", Jp, dp[self.pi])

    def _JT(self, Jc, Jp, v):
        gc = torch.zeros(self.n_cam, self.d, dtype=v.dtype, device=self.dev).index_add_(0, self.ci, torch.einsum("nkd,nk->nd", Jc, v))
        gp = torch.zeros(self.n_pt, 3, dtype=v.dtype, device=self.dev).index_add_(0, self.pi, torch.einsum("nkd,nk->nd", Jp, v))
        return gc, gp

    def _sparse_J(self, Jc, Jp):
        """J [2N, d n_cam + 3 n_pt] and J^T as sparse CSR (rows: residual components; 12 entries per row at d = 9)."""
        n, d, w = self.n_obs, self.d, self.d + 3
        ccol = (self.ci * d)[:, None] + torch.arange(d, device=self.dev)[None, :]
        pcol = (self.n_cam * d + self.pi * 3)[:, None] + torch.arange(3, device=self.dev)[None, :]
        col = torch.cat([ccol, pcol], dim=1)                                   # [N, w]
        col = col[:, None, :].expand(n, 2, w).reshape(-1)
        val = torch.cat([Jc, Jp], dim=2).reshape(-1)                           # [N, 2, w]
        crow = torch.arange(0, 2 * n + 1, device=self.dev, dtype=torch.int64) * w
        J = torch.sparse_csr_tensor(crow, col, val, size=(2 * n, self.n_cam * d + self.n_pt * 3))
        JT = J.to_sparse_coo().t().coalesce().to_sparse_csr()
        return J, JT

    def _pcg_sparse(self, J, JT, dg, dd, b):
        """Jacobi-PCG on A = J^T J with diagonal dg replaced by dd; flat vectors; two SpMVs per iteration."""
        t0 = time.perf_counter()
        x = torch.zeros_like(b)
        r = b.clone()
        atol2 = (self.pcg_tol ** 2) * float(b @ b)
        mi = 1.0 / dd
        shift = dd - dg
        rho_prev, p = None, None
        it, maxiter = 0, 10 * b.numel()
        while it < maxiter:
            if float(r @ r) < atol2:       # one host sync per iteration, like `.item()` in eager torch
                break
            z = mi * r
            rho = r @ z
            p = z.clone() if it == 0 else z + (rho / rho_prev) * p
            q = torch.mv(JT, torch.mv(J, p)) + shift * p
            alpha = rho / (p @ q)
            x += alpha * p
            r -= alpha * q
            rho_prev = rho
            it += 1
        if self.dev.type == "cuda":
            torch.cuda.synchronize(self.dev)
        self.pcg_iters += it
        self.pcg_seconds += time.perf_counter() - t0
        return x

    def _pcg(self, Jc, Jp, dgc, dgp, ddc, ddp, bc, bp):
        """Jacobi-PCG on A = J^T J with its diagonal (dgc | dgp) replaced by the damped one (ddc | ddp)."""
        t0 = time.perf_counter()
        xc, xp = torch.zeros_like(bc), torch.zeros_like(bp)
        rc, rp = bc.clone(), bp.clone()
        atol2 = (self.pcg_tol ** 2) * float((bc * bc).sum() + (bp * bp).sum())
        maxiter = 10 * (bc.numel() + bp.numel())
        mic, mip = 1.0 / ddc, 1.0 / ddp
        rho_prev, pc, pp_ = None, None, None
        it = 0
        while it < maxiter:
            if float((rc * rc).sum() + (rp * rp).sum()) < atol2:   # one host sync per iteration, like `.item()` in eager torch
                break
            zc, zp = mic * rc, mip * rp
            rho = (rc * zc).sum() + (rp * zp).sum()
            if it == 0:
                pc, pp_ = zc.clone(), zp.clone()
            else:
                beta = rho / rho_prev
                pc, pp_ = zc + beta * pc, zp + beta * pp_
            qc, qp = self._JT(Jc, Jp, self._J(Jc, Jp, pc, pp_))
            qc = qc + (ddc - dgc) * pc
            qp = qp + (ddp - dgp) * pp_
            alpha = rho / ((pc * qc).sum() + (pp_ * qp).sum())
            xc += alpha * pc; xp += alpha * pp_
            rc -= alpha * qc; rp -= alpha * qp
            rho_prev = rho
            it += 1
        if self.dev.type == "cuda":
            torch.cuda.synchronize(self.dev)
        self.pcg_iters += it
        self.pcg_seconds += time.perf_counter() - t0
        return xc, xp

    def _retract(self, dc, dp):
        pose = lie.se3_retract(self.cam[:, :7], dc[:, :6].contiguous())
        cam = torch.cat([pose, self.cam[:, 7:] + dc[:, 6:]], dim=1)
        return cam, self.pts + dp

    # -- one optimizer.step(input) (oracle/lm.py::LM.step restated on tensors) --------------
    def step(self):
        r, Jc, Jp = self.blocks()
        s = (r * r).sum(-1)
        rs = torch.sqrt(s)
        w = torch.sqrt(torch.where(rs < self.delta, torch.ones_like(rs), self.delta / torch.where(rs > 0, rs, torch.ones_like(rs))))
        R = r * w[:, None]
        Jc = Jc * w[:, None, None]
        Jp = Jp * w[:, None, None]
        if self.loss is None:
            self.loss = float(self._rho(r))
        last = self.loss
        gc, gp = self._JT(Jc, Jp, R)
        dgc = torch.zeros(self.n_cam, self.d, dtype=torch.float64, device=self.dev).index_add_(0, self.ci, (Jc * Jc).sum(1))
        dgp = torch.zeros(self.n_pt, 3, dtype=torch.float64, device=self.dev).index_add_(0, self.pi, (Jp * Jp).sum(1))
        ddc, ddp = dgc.clamp(1e-6, 1e32), dgp.clamp(1e-6, 1e32)
        rejects = 0
        if self.mode == "sparse":
            Js, JTs = self._sparse_J(Jc, Jp)
        while last <= self.loss:
            lam = self.damping
            ddc, ddp = ddc * (1.0 + lam), ddp * (1.0 + lam)
            if self.mode == "sparse":
                nc = self.n_cam * self.d
                x = self._pcg_sparse(Js, JTs, torch.cat([dgc.reshape(-1), dgp.reshape(-1)]), torch.cat([ddc.reshape(-1), ddp.reshape(-1)]),
                                     -torch.cat([gc.reshape(-1), gp.reshape(-1)]))
                dc, dp = x[:nc].reshape(self.n_cam, self.d), x[nc:].reshape(self.n_pt, 3)
            else:
                dc, dp = self._pcg(Jc, Jp, dgc, dgp, ddc, ddp, -gc, -gp)
            cam_new, pts_new = self._retract(dc, dp)
            self.loss = float(self._rho(self.residuals(cam_new, pts_new)))
            JD = self._J(Jc, Jp, dc, dp)
            denom = -float((JD * (2.0 * R + JD)).sum())
            quality = (last - self.loss) / denom if denom != 0.0 else 0.0
            radius = 1.0 / self.damping
            if quality > 0.5:
                radius, self.tr_down = self.tr_up * radius, self.tr_down0
            elif quality > 1e-3:
                self.tr_down = self.tr_down0
            else:
                radius, self.tr_down = radius * self.tr_down, self.tr_down * 0.5
            self.tr_down = max(1e-6, min(self.tr_down, self.tr_max))
            radius = max(1e-6, min(radius, self.tr_max))
            self.damping = 1.0 / radius
            if last < self.loss and rejects < self.reject:
                self.loss = last
                rejects += 1
            else:
                self.cam, self.pts = cam_new, pts_new
                break
        return self.loss
