"""Oracle-side helpers shared by the parity tests (numpy fp64, brute force, small sizes)."""
import numpy as np

from oracle.lm import triggs_scale


def weighted_blocks(pb, delta=1.0):
    """r (unweighted), R, Jc, Jp (Triggs-weighted) of an oracle BAProblem."""
    r, Jc, Jp = pb.blocks()
    w = triggs_scale(r, delta)
    return r, r * w[:, None], Jc * w[:, None, None], Jp * w[:, None, None]


def _damp(H, mu):
    H = H.copy()
    idx = np.arange(H.shape[1])
    H[:, idx, idx] = np.clip(H[:, idx, idx], 1e-6, 1e32) * mu
    return H


def partial_blocks(pb, delta=1.0):
    """Undamped Hpp, g_p, Hcc, g_c and Hcp of (a shard of) a problem."""
    _, R, Jc, Jp = weighted_blocks(pb, delta)
    n_cam, n_pt, d = pb.n_cam, pb.n_pt, pb.d
    Hpp = np.zeros((n_pt, 3, 3)); gp = np.zeros((n_pt, 3)); Hcc = np.zeros((n_cam, d, d)); gc = np.zeros((n_cam, d))
    np.add.at(Hpp, pb.pi, np.einsum("nki,nkj->nij", Jp, Jp))
    np.add.at(gp, pb.pi, np.einsum("nki,nk->ni", Jp, R))
    np.add.at(Hcc, pb.ci, np.einsum("nki,nkj->nij", Jc, Jc))
    np.add.at(gc, pb.ci, np.einsum("nki,nk->ni", Jc, R))
    return Hpp, gp, Hcc, gc, np.einsum("nki,nkj->nij", Jc, Jp)


def schur_contribution(pb, Hpp, gp, Hcp, mu):
    """E = sum_p Hcp Hpp^-1 Hcp^T (dense) and e = sum Hcp Hpp^-1 g_p for this shard's points."""
    n_cam, n_pt, d = pb.n_cam, pb.n_pt, pb.d
    inv = np.linalg.inv(_damp(Hpp, mu))
    W = np.einsum("nij,njk->nik", Hcp, inv[pb.pi])
    E = np.zeros((n_cam * d, n_cam * d))
    e = np.zeros((n_cam, d))
    np.add.at(e, pb.ci, np.einsum("nij,nj->ni", W, gp[pb.pi]))
    order = np.argsort(pb.pi, kind="stable")
    off = np.searchsorted(pb.pi[order], np.arange(n_pt + 1))
    for p in range(n_pt):
        obs = order[off[p]:off[p + 1]]
        for x in obs:
            for y in obs:
                i, j = pb.ci[x], pb.ci[y]
                E[i * d:(i + 1) * d, j * d:(j + 1) * d] += W[x] @ Hcp[y].T
    return E, e


def assemble_reduced_system(Hcc, gc, E, e, mu):
    """S = damp(Hcc) - E,  b = -(g_c - e)."""
    n_cam, d = Hcc.shape[0], Hcc.shape[1]
    S = -E.copy()
    Hd = _damp(Hcc, mu)
    for i in range(n_cam):
        S[i * d:(i + 1) * d, i * d:(i + 1) * d] += Hd[i]
    return S, -(gc - e).reshape(-1)


def normal_blocks(pb, mu, delta=1.0):
    Hpp, gp, Hcc, gc, Hcp = partial_blocks(pb, delta)
    E, e = schur_contribution(pb, Hpp, gp, Hcp, mu)
    S, rhs = assemble_reduced_system(Hcc, gc, E, e, mu)
    return Hpp, gp, Hcc, gc, S, rhs
