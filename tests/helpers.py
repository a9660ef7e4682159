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


# ---------------------------------------------------------------------------------------------
# pixel-space scenes for the scene-layer camera maps (FilterTracksByReprojection, UndistortImages)
# ---------------------------------------------------------------------------------------------
PIXEL_MODEL_PARAMS = {   # plausible intrinsics per CameraModelId.value (scene/defs.py:177-237 layouts)
    0: [1000.0, 640.0, 480.0],
    1: [1010.0, 990.0, 640.0, 480.0],
    2: [1000.0, 640.0, 480.0, -0.08],
    3: [1000.0, 640.0, 480.0, -0.08, 0.015],
    4: [1010.0, 990.0, 640.0, 480.0, -0.07, 0.02, 1e-3, -8e-4],
    5: [1010.0, 990.0, 640.0, 480.0, -0.03, 0.01, -2e-3, 5e-4],
    6: [1010.0, 990.0, 640.0, 480.0, -0.07, 0.02, 1e-3, -8e-4, 3e-3, 0.01, -4e-3, 1e-3],
    7: [1010.0, 990.0, 640.0, 480.0, 0.35],
    8: [1000.0, 640.0, 480.0, -0.03],
    9: [1000.0, 640.0, 480.0, -0.03, 0.008],
    10: [1010.0, 990.0, 640.0, 480.0, -0.03, 0.01, 1e-3, -8e-4, -2e-3, 5e-4, 2e-3, -1e-3],
}


def make_pixel_scene(models=(3,), n_img=12, n_trk=200, mean_len=4.0, seed=0, outlier_frac=0.15, behind_frac=0.05, noise_px=0.4):
    """Seeded scene with PIXEL features: one camera per entry of ``models`` (images cycle through
    them), poses on a ring looking inwards, tracks whose features are the oracle's cam2img of the
    point plus noise; gross outliers, points behind a camera, a single-view and an empty track.
    Every image also carries unreferenced random features so that UndistortImages covers the frame.
    Returns (cameras, images, tracks) of instantsfm_b200.scene.defs types."""
    from instantsfm_b200.scene.defs import Camera, CameraModelId, Image, Track
    from oracle.camera_ops import Intrinsics, cam2img
    rng = np.random.default_rng(seed)
    cameras = []
    for ci, m in enumerate(models):
        p = np.array(PIXEL_MODEL_PARAMS[m]) * (1.0 + 0.01 * rng.normal(size=len(PIXEL_MODEL_PARAMS[m])))
        cameras.append(Camera(id=ci, model_id=CameraModelId(m), width=1280, height=960, params=[float(x) for x in p]))
    intr = [Intrinsics(c.model_id.value, c.params) for c in cameras]
    images = []
    for i in range(n_img):
        ang = 2 * np.pi * i / n_img
        c = np.array([12 * np.cos(ang), 12 * np.sin(ang), rng.normal(0, 0.5)])
        z = -c / np.linalg.norm(c) + rng.normal(0, 0.05, 3)
        z /= np.linalg.norm(z)
        x = np.cross([0, 0, 1.0], z); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R = np.stack([x, y, z], 0)
        w2c = np.eye(4); w2c[:3, :3] = R; w2c[:3, 3] = -R @ c
        images.append(Image(id=i, cam_id=i % len(cameras), is_registered=True, world2cam=w2c))
    feats = [[rng.uniform([0, 0], [1280, 960]) for _ in range(5)] for _ in range(n_img)]
    tracks = {}
    for t in range(n_trk):
        X = rng.normal(0, 2.0, 3)
        k = 0 if t == 7 else (1 if t % 23 == 0 else min(n_img, 2 + rng.geometric(1.0 / max(mean_len - 1.0, 1.0))))
        ids = rng.choice(n_img, size=k, replace=False)
        if k and rng.random() < behind_frac:
            c = images[ids[0]].center()
            X = c + (c - X) * 0.5
        obs = []
        for i in ids:
            p = images[i].world2cam[:3, :3] @ X + images[i].world2cam[:3, 3]
            with np.errstate(all="ignore"):
                px = cam2img(intr[images[i].cam_id], p[None])[0]
            if not np.all(np.isfinite(px)):
                px = np.array([640.0, 480.0])
            px = px + rng.normal(0, noise_px, 2)
            if rng.random() < outlier_frac:
                px = px + rng.normal(0, 25.0, 2)
            obs.append((int(i), len(feats[i])))
            feats[i].append(px)
        tracks[500 + 7 * t] = Track(id=500 + 7 * t, xyz=X, observations=np.array(obs, dtype=np.int64).reshape(-1, 2))
    for i, img in enumerate(images):
        img.features = np.array(feats[i]).reshape(-1, 2)
    return cameras, images, tracks


# ---------------------------------------------------------------------------------------------
# gauge alignment
# ---------------------------------------------------------------------------------------------
def gauge_align(cam, pts, pts_ref, sample=None):
    """Bundle adjustment determines the scene up to a similarity transform of the world (7 gauge
    freedoms): the cost, the residuals and the RMSE do not see it, and the damped LM step pins it only
    through eigenvalues of the size of the damping (1e-4 here), so two solvers whose linear solves
    agree to 1e-6 in the residual still drift apart ALONG the gauge by orders of magnitude more than
    across it.  Poses / points are therefore compared after the least-squares similarity
    X -> s R X + t (Umeyama) that maps this solution's points onto the reference's; the cameras
    [t, q_xyzw, intrinsics] follow: R_c -> R_c R^T, t_c -> s t_c - R_c R^T t (projection unchanged).
    ``sample``: indices of ``pts`` that correspond to ``pts_ref`` (default: all).  Returns
    (cam_aligned, pts_aligned, (s, R, t))."""
    from scipy.spatial.transform import Rotation
    cam = np.asarray(cam, np.float64); pts = np.asarray(pts, np.float64); ref = np.asarray(pts_ref, np.float64)
    src = pts if sample is None else pts[sample]
    mu_s, mu_r = src.mean(0), ref.mean(0)
    A, B = src - mu_s, ref - mu_r
    U, S, Vt = np.linalg.svd(B.T @ A / len(src))
    d = np.sign(np.linalg.det(U @ Vt))
    Dm = np.diag([1.0, 1.0, d])
    R = U @ Dm @ Vt
    s = float((S * np.diag(Dm)).sum() / (A * A).sum() * len(src))
    t = mu_r - s * R @ mu_s
    pts_al = s * pts @ R.T + t
    cam_al = cam.copy()
    Rc = Rotation.from_quat(cam[:, 3:7]).as_matrix()
    Rn = Rc @ R.T
    cam_al[:, :3] = s * cam[:, :3] - np.einsum("nij,j->ni", Rn, t)
    q = Rotation.from_matrix(Rn).as_quat()
    flip = np.sign(np.einsum("ni,ni->n", q, cam[:, 3:7]))
    cam_al[:, 3:7] = q * np.where(flip == 0, 1.0, flip)[:, None]
    return cam_al, pts_al, (s, R, t)
