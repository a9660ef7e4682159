"""Levenberg-Marquardt of the oracle (numpy / scipy.sparse fp64).  TEST INFRASTRUCTURE.

Restates what ``bae.optim.LM(model, strategy=TrustRegion, solver=PCG(tol=1e-5),
kernel=Huber(delta), reject=30).step(input)`` does at the reference's call sites
(bundle_adjustment.py:116-119,132; global_positioning.py:158-161,176).  bae/pypose are
not available: the semantics below are the public pypose ``LevenbergMarquardt`` /
``TrustRegion`` / ``Huber`` / ``FastTriggs`` definitions as summarised in SURVEY.md
9.3-9.5 -- UNVERIFIED, see the assumption ledger in oracle/__init__.py.

A problem object supplies
    residuals() -> [N, rd] array at the current parameters
    jacobian()  -> scipy.sparse.csr_matrix [N*rd, n_unknowns] (tangent-space Jacobian)
    retract(D)  -> apply x <- x (+) D in place
    snapshot() / restore(s)
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def huber_rho(s, delta):
    """pypose Huber on s = ||r||^2: s if sqrt(s) < delta else 2 delta sqrt(s) - delta^2."""
    rs = np.sqrt(s)
    return np.where(rs < delta, s, 2.0 * delta * rs - delta * delta)


def huber_drho(s, delta):
    """rho'(s): 1 inside, delta / sqrt(s) outside."""
    rs = np.sqrt(s)
    return np.where(rs < delta, 1.0, delta / np.where(rs > 0, rs, 1.0))


def robust_cost(r, delta):
    """sum_i rho(||r_i||^2) -- the number the reference prints (bundle_adjustment.py:142)."""
    return float(huber_rho((r * r).sum(-1), delta).sum())


def triggs_scale(r, delta):
    """FastTriggs weights sqrt(rho'(||r_i||^2)), one per residual block."""
    return np.sqrt(huber_drho((r * r).sum(-1), delta))


class TrustRegion:
    """pp.optim.strategy.TrustRegion(radius, max, up, down) restated.

    BA: radius=1e4, max=1e10, up=2, down=0.5**4 (bundle_adjustment.py:116);
    GP: radius=1e3, max=1e8 (global_positioning.py:158).  Defaults UNVERIFIED.
    """

    def __init__(self, radius=1e6, high=0.5, low=1e-3, up=2.0, down=0.5, factor=0.5,
                 max=1e6, min=1e-6):
        self.radius, self.high, self.low, self.up = radius, high, low, up
        self.down0, self.down, self.factor, self.max, self.min = down, down, factor, max, min
        self.damping = 1.0 / radius

    def update(self, last, loss, JD, R):
        denom = -float(JD @ (2.0 * R + JD))
        quality = (last - loss) / denom if denom != 0.0 else 0.0
        radius = 1.0 / self.damping
        if quality > self.high:
            radius = self.up * radius
            self.down = self.down0
        elif quality > self.low:
            self.down = self.down0
        else:
            radius = radius * self.down
            self.down = self.down * self.factor
        self.down = max(self.min, min(self.down, self.max))
        radius = max(self.min, min(radius, self.max))
        self.radius = radius
        self.damping = 1.0 / radius
        return quality


def pcg_jacobi(A, b, tol=1e-5, maxiter=None):
    """bae.utils.pysolvers.PCG(tol) restated: scalar-Jacobi CG, x0 = 0, ||r|| < tol ||b||."""
    n = b.shape[0]
    maxiter = 10 * n if maxiter is None else maxiter
    x = np.zeros_like(b)
    r = b.copy()
    atol = tol * np.linalg.norm(b)
    Minv = 1.0 / A.diagonal()
    rho_prev, p = None, None
    iters = 0
    for it in range(maxiter):
        if np.linalg.norm(r) < atol:
            break
        z = Minv * r
        rho = float(r @ z)
        p = z.copy() if it == 0 else z + (rho / rho_prev) * p
        q = A @ p
        alpha = rho / float(p @ q)
        x += alpha * p
        r -= alpha * q
        rho_prev = rho
        iters = it + 1
    return x, iters


def direct_solve(A, b):
    """Exact sparse solve of the damped normal equations (ground truth for parity)."""
    if A.shape[0] <= 3000:
        return np.linalg.solve(A.toarray(), b), 0
    lu = spla.splu(A.tocsc(), permc_spec="COLAMD", diag_pivot_thresh=0.0,
                   options=dict(SymmetricMode=True))
    return lu.solve(b), 0


def schur_direct(A, b, n_lead, block=3):
    """Exact solve of the damped normal equations through the Schur complement of the trailing
    block-diagonal part (points, ``block`` x ``block``): same solution as ``direct_solve`` up to
    rounding, but feasible at BASELINE config 2 size (the full sparse LU is not).  The reduced
    system is formed densely and solved by LAPACK."""
    A = A.tocsr()
    n = A.shape[0]
    App = A[n_lead:, n_lead:].tocoo()
    nb = (n - n_lead) // block
    blocks = np.zeros((nb, block, block))
    np.add.at(blocks, (App.row // block, App.row % block, App.col % block), App.data)
    inv = np.linalg.inv(blocks)
    Hinv = sp.bsr_matrix((inv, np.arange(nb), np.arange(nb + 1)), shape=(n - n_lead, n - n_lead)).tocsr()
    Acp = A[:n_lead, n_lead:].tocsr()
    W = (Acp @ Hinv).tocsr()
    S = A[:n_lead, :n_lead].toarray() - (W @ Acp.T).toarray()
    rhs = b[:n_lead] - W @ b[n_lead:]
    xc = np.linalg.solve(S, rhs)
    xp = Hinv @ (b[n_lead:] - Acp.T @ xc)
    return np.concatenate([xc, xp]), 0


class LM:
    """One ``step`` = one linearisation + >=1 damped solves / trial evaluations."""

    def __init__(self, problem, strategy, huber_delta, solver="pcg", pcg_tol=1e-5,
                 reject=30, dmin=1e-6, dmax=1e32):
        self.problem, self.strategy, self.delta = problem, strategy, huber_delta
        self.solver, self.pcg_tol, self.reject = solver, pcg_tol, reject
        self.dmin, self.dmax = dmin, dmax
        self.loss = None
        self.trace = []  # one dict per step, for trajectory parity tests

    def _solve(self, A, b):
        if self.solver == "pcg":
            return pcg_jacobi(A, b, self.pcg_tol)
        if self.solver == "schur":   # exact, via the point Schur complement (problem.n_lead leading unknowns stay)
            return schur_direct(A, b, self.problem.n_lead)
        return direct_solve(A, b)

    def step(self):
        pb = self.problem
        r = pb.residuals()
        J = pb.jacobian()
        w = triggs_scale(r, self.delta)
        rd = r.shape[1]
        R = (r * w[:, None]).reshape(-1)
        J = sp.diags(np.repeat(w, rd)) @ J
        if self.loss is None:
            self.loss = robust_cost(r, self.delta)
        last = self.loss
        A = (J.T @ J).tocsr()
        g = J.T @ R
        diag = np.clip(A.diagonal(), self.dmin, self.dmax)
        A = A - sp.diags(A.diagonal()) + sp.diags(diag)
        rejects, trials, lin_iters = 0, 0, 0
        info = {"loss_before": last, "trials": []}
        while last <= self.loss:
            lam = self.strategy.damping
            diag = diag * (1.0 + lam)
            A = (A - sp.diags(A.diagonal()) + sp.diags(diag)).tocsr()
            D, its = self._solve(A, -g)
            lin_iters += its
            snap = pb.snapshot()
            pb.retract(D)
            self.loss = robust_cost(pb.residuals(), self.delta)
            JD = J @ D
            q = self.strategy.update(last, self.loss, JD, R)
            trials += 1
            info["trials"].append({"lambda": lam, "loss": self.loss, "quality": q,
                                   "step_norm": float(np.linalg.norm(D))})
            if last < self.loss and rejects < self.reject:
                pb.restore(snap)
                self.loss = last
                rejects += 1
            else:
                break
        info.update(loss=self.loss, rejects=rejects, lin_iters=lin_iters)
        self.trace.append(info)
        return self.loss


def run_loop(optimizer, max_num_iterations, function_tolerance, stop_on_identical):
    """Outer loop + stop rule: bundle_adjustment.py:128-141 / global_positioning.py:172-183."""
    window = 4
    history = []
    for _ in range(max_num_iterations):
        history.append(optimizer.step())
        if len(history) >= 2 * window:
            recent = np.mean(history[-window:])
            previous = np.mean(history[-2 * window:-window])
            improvement = (previous - recent) / previous
            if abs(improvement) < function_tolerance:
                break
            if stop_on_identical and history[-1] == history[-2]:
                break
    return history
