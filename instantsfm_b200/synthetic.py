"""Seeded synthetic BA / GP problems shaped like the BASELINE.json configs.

Everything is vectorised numpy so that the 5 M- and 60 M-observation configs are
generated in seconds to a minute on the host.  Geometry: cameras orbit a point cloud on a
closed loop looking inward (depth ~15..45), point ``p`` is observed by ``k_p`` distinct
cameras drawn from a window of ``window`` consecutive loop positions around a random
anchor, so camera pairs further apart than ``window`` share no point and the reduced
camera system S is banded; ``window >= n_cam`` gives the dense S of the BAL sets.
Track length ``k_p = 2 + Geometric`` (cap 200) trimmed so that the totals hit the
requested n_obs exactly.  Observations come out sorted by point, as the reference's
flattening loop produces them (bundle_adjustment.py:88-96).
"""
from dataclasses import dataclass, field

import numpy as np

from .geometry import quat_mul, quat_to_mat, rotvec_to_quat, matrices_to_pose7

# number of intrinsics once the principal point is removed, per CameraModelId.value
N_INTR = {0: 1, 1: 2, 2: 2, 3: 3, 4: 6, 5: 6, 6: 10, 8: 2, 9: 3}
N_FOCAL = {0: 1, 1: 2, 2: 1, 3: 1, 4: 2, 5: 2, 6: 2, 8: 1, 9: 1}

CONFIGS = {
    # name: (n_cam, n_pt, n_obs, window, seed)
    "C1": (64, 10_000, 60_000, 64, 1001),
    "C2": (1_723, 156_000, 678_000, 1_723, 1002),
    "C3": (1_778, 993_000, 5_000_000, 1_778, 1003),
    "C5": (20_000, 10_000_000, 60_000_000, 256, 1005),
}
GP_CONFIGS = {"C4": (2_500, 500_000, 3_000_000, 2_500, 1004)}


@dataclass
class BAArrays:
    """Flat BA problem in the compacted index space of bundle_adjustment.py:108-113."""
    model_id: int
    camera_params: np.ndarray      # [Nc, 7 + n_intr] = [t, q_xyzw, intrinsics without pp]
    camera_pps: np.ndarray         # [Nc, 2]
    points_3d: np.ndarray          # [Np, 3]
    points_2d: np.ndarray          # [N, 2]
    camera_indices: np.ndarray     # int32 [N]
    point_indices: np.ndarray      # int32 [N], non-decreasing
    gt_camera_params: np.ndarray = field(default=None, repr=False)
    gt_points_3d: np.ndarray = field(default=None, repr=False)
    point_range: tuple = None          # (p0, p1) of the global instance held by this object
    n_pt_total: int = 0                # sizes of the global instance
    n_obs_total: int = 0

    @property
    def n_cam(self):
        return self.camera_params.shape[0]

    @property
    def n_pt(self):
        return self.points_3d.shape[0]

    @property
    def n_obs(self):
        return self.points_2d.shape[0]


def distort_numpy(model_id, u, k):
    """Distortion * focal for the nine implemented models (cost_function.py:32-177)."""
    r2 = (u * u).sum(-1, keepdims=True)

    def fisheye(u):
        r = np.sqrt(r2)
        return u * np.arctan(r) / r

    def tangential(p):
        uv = u[:, :1] * u[:, 1:2]
        return 2 * p * uv + p[:, ::-1] * (r2 + 2 * u * u)

    if model_id == 0:
        return u * k[:, 0:1]
    if model_id == 1:
        return u * k[:, 0:2]
    if model_id == 2:
        return u * (1 + k[:, 1:2] * r2) * k[:, 0:1]
    if model_id == 3:
        return u * (1 + k[:, 1:2] * r2 + k[:, 2:3] * r2 ** 2) * k[:, 0:1]
    if model_id == 4:
        return (u + u * (k[:, 2:3] * r2 + k[:, 3:4] * r2 ** 2) + tangential(k[:, 4:6])) * k[:, 0:2]
    if model_id == 5:
        return fisheye(u) * (1 + k[:, 2:3] * r2 + k[:, 3:4] * r2 ** 2 + k[:, 4:5] * r2 ** 3) * k[:, 0:2]
    if model_id == 6:
        rad = (1 + k[:, 2:3] * r2 + k[:, 3:4] * r2 ** 2 + k[:, 6:7] * r2 ** 3) / \
              (1 + k[:, 7:8] * r2 + k[:, 8:9] * r2 ** 2 + k[:, 9:10] * r2 ** 3) - 1
        return (u + u * rad + tangential(k[:, 4:6])) * k[:, 0:2]
    if model_id == 8:
        return fisheye(u) * (1 + k[:, 1:2] * r2) * k[:, 0:1]
    if model_id == 9:
        return fisheye(u) * (1 + k[:, 1:2] * r2 + k[:, 2:3] * r2 ** 2) * k[:, 0:1]
    raise NotImplementedError("Unsupported camera model")


def project_numpy(model_id, X, cam, pp, chunk=4_000_000):
    """Row-wise projection of X[N,3] with camera rows cam[N,7+ni], pp[N,2]; chunked."""
    out = np.empty((X.shape[0], 2))
    depth = np.empty(X.shape[0])
    for s in range(0, X.shape[0], chunk):
        e = min(s + chunk, X.shape[0])
        R = quat_to_mat(cam[s:e, 3:7])
        y = np.einsum("nij,nj->ni", R, X[s:e]) + cam[s:e, :3]
        u = y[:, :2] / y[:, 2:3]
        out[s:e] = distort_numpy(model_id, u, cam[s:e, 7:]) + pp[s:e]
        depth[s:e] = y[:, 2]
    return out, depth


def _prev_prime(n):
    def is_p(m):
        if m < 2:
            return False
        i = 2
        while i * i <= m:
            if m % i == 0:
                return False
            i += 1
        return True
    while not is_p(n):
        n -= 1
    return n


def _track_lengths(rng, n_pt, n_obs, kmax, kmin=2):
    mean_extra = n_obs / n_pt - kmin
    assert mean_extra >= 0, "too few observations per point"
    assert n_obs <= n_pt * kmax, "too many observations per point for this window"
    if mean_extra == 0:
        return np.full(n_pt, kmin, dtype=np.int64)
    p = 1.0 / (1.0 + mean_extra)
    k = kmin + (rng.geometric(p, size=n_pt) - 1)
    k = np.minimum(k, kmax).astype(np.int64)
    diff = int(n_obs - k.sum())
    while diff != 0:
        if diff > 0:
            cand = np.flatnonzero(k < kmax)
            pick = rng.choice(cand, size=min(diff, cand.size), replace=False)
            k[pick] += 1
        else:
            cand = np.flatnonzero(k > kmin)
            pick = rng.choice(cand, size=min(-diff, cand.size), replace=False)
            k[pick] -= 1
        diff = int(n_obs - k.sum())
    return k


def _orbit_cameras(rng, n_cam, radius=30.0):
    ang = 2 * np.pi * np.arange(n_cam) / n_cam
    C = np.stack([radius * np.cos(ang), radius * np.sin(ang), rng.uniform(-3, 3, n_cam)], 1)
    C[:, :2] *= rng.uniform(0.9, 1.1, (n_cam, 1))
    target = rng.normal(scale=1.5, size=(n_cam, 3))
    z = target - C
    z /= np.linalg.norm(z, axis=1, keepdims=True)
    up = np.array([0.0, 0.0, 1.0])
    x = np.cross(z, up)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    y = np.cross(z, x)
    R = np.stack([x, y, z], axis=1)                    # rows = camera axes in world coords
    M = np.zeros((n_cam, 4, 4))
    M[:, :3, :3] = R
    M[:, :3, 3] = -np.einsum("nij,nj->ni", R, C)
    M[:, 3, 3] = 1
    return matrices_to_pose7(M), C


POINT_CHUNK = 1 << 16   # points per generation chunk: the unit of the per-chunk random streams


def _visibility(rng, n_cam, n_pt, n_obs, window, kmin=2, anchor_range=None):
    """Sorted-by-point observation lists: camera_indices, point_indices (int32).  (GP problems.)

    Slots inside the window are visited with a per-point stride modulo a prime W, so the
    k_p cameras of a point are distinct."""
    W = _prev_prime(min(window, n_cam))
    k = _track_lengths(rng, n_pt, n_obs, min(200, W), kmin)
    # strides >= W / 8 keep even two-view tracks at a useful baseline
    stride = rng.integers(max(1, W // 8), W, size=n_pt) if W > 2 else np.ones(n_pt, dtype=np.int64)
    if anchor_range is None:
        anchor = rng.integers(0, n_cam, size=n_pt)
    else:
        anchor = np.sort(rng.integers(anchor_range[0], anchor_range[1], size=n_pt))
    start = rng.integers(0, W, size=n_pt)
    pt = np.repeat(np.arange(n_pt, dtype=np.int64), k)
    offs = np.cumsum(k) - k
    j = np.arange(pt.shape[0], dtype=np.int64) - offs[pt]
    slot = (start[pt] + j * stride[pt]) % W
    cam = (anchor[pt] + slot - W // 2) % n_cam
    return cam.astype(np.int32), pt.astype(np.int32)


def _street_cameras(rng, n_cam, window):
    """Cameras along a closed street (ring of circumference n_cam * spacing) looking at the
    facade ~30 units inside the ring.  `window` consecutive cameras span ~25 units, so the
    tracks of a banded problem keep wide baselines (depth 15..40)."""
    spacing = 25.0 / window
    radius = n_cam * spacing / (2 * np.pi)
    ang = 2 * np.pi * np.arange(n_cam) / n_cam
    C = np.stack([radius * np.cos(ang), radius * np.sin(ang), rng.uniform(-1, 1, n_cam)], 1)
    inward = -np.stack([np.cos(ang), np.sin(ang), np.zeros(n_cam)], 1)
    z = inward * 30.0 + rng.normal(scale=1.5, size=(n_cam, 3))
    z /= np.linalg.norm(z, axis=1, keepdims=True)
    up = np.array([0.0, 0.0, 1.0])
    x = np.cross(z, up)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    y = np.cross(z, x)
    R = np.stack([x, y, z], axis=1)
    M = np.zeros((n_cam, 4, 4))
    M[:, :3, :3] = R
    M[:, :3, 3] = -np.einsum("nij,nj->ni", R, C)
    M[:, 3, 3] = 1
    return matrices_to_pose7(M), radius, spacing


def partition_by_observations(point_offsets, world):
    """numpy twin of isfm_partition_points (include/isfm_b200.h): boundary g = first point whose
    starting observation offset >= g * n_obs / world (integer arithmetic)."""
    off = np.asarray(point_offsets, dtype=np.int64)
    n_obs = int(off[-1])
    out = np.empty(world + 1, dtype=np.int64)
    for g in range(world + 1):
        # first p with off[p] * world >= g * n_obs
        out[g] = np.searchsorted(off[:-1] * world, g * n_obs, side="left")
    out[world] = off.shape[0] - 1
    return out


def make_ba_problem(n_cam, n_pt, n_obs, window=None, seed=0, model_id=3, noise_px=0.5,
                    outlier_frac=0.01, perturb=1.0, point_range=None, shard=None):
    """Build a BAL-shaped problem; ``perturb`` scales the initial-state perturbation.

    ``window >= n_cam``: object-centric orbit, every camera may see every point (dense reduced
    camera system, BAL-like).  ``window < n_cam``: street geometry, a point is seen only from
    ``window`` consecutive cameras (banded system, city-scale).

    ONE instance per (sizes, seed): the cameras, the track lengths and (street) the anchors come
    from global random streams; everything else about a point -- visibility pattern, position,
    observation noise, outliers, initial perturbation -- comes from the stream of its chunk of
    ``POINT_CHUNK`` consecutive points.  ``point_range = (p0, p1)`` or ``shard = (rank, world)``
    (contiguous point ranges balanced by observation count, the library's partition) therefore
    return exactly the slice of the global problem that a full generation would contain:
    N ranks solve the same problem as one.  Shard results carry ``point_range`` and local point
    indices; cameras are always the full set."""
    rng = np.random.default_rng([seed, 0])
    window = n_cam if window is None else window
    street = window < n_cam
    if street:
        pose_gt, radius, spacing = _street_cameras(rng, n_cam, window)
    else:
        pose_gt, _ = _orbit_cameras(rng, n_cam)
    ni, nf = N_INTR[model_id], N_FOCAL[model_id]
    intr_gt = np.empty((n_cam, ni))
    intr_gt[:, :nf] = rng.uniform(800, 1200, (n_cam, 1))
    if ni > nf:
        scales = np.array([1e-2, 1e-3] + [1e-4] * 8)[: ni - nf]
        intr_gt[:, nf:] = rng.normal(size=(n_cam, ni - nf)) * scales
    cam_gt = np.concatenate([pose_gt, intr_gt], 1)
    pps = np.zeros((n_cam, 2))                          # BAL convention
    # perturbed initial cameras: rotation 0.5 deg rms, translation 1 % of scene scale (30), focal 1 %
    cam0 = cam_gt.copy()
    dq = rotvec_to_quat(rng.normal(scale=perturb * np.deg2rad(0.5) / np.sqrt(3), size=(n_cam, 3)))
    cam0[:, 3:7] = quat_mul(dq, cam_gt[:, 3:7])
    cam0[:, :3] += rng.normal(scale=perturb * 0.3 / np.sqrt(3), size=(n_cam, 3))
    cam0[:, 7:7 + nf] *= 1 + rng.normal(scale=perturb * 0.01, size=(n_cam, nf))

    # global per-point streams: track lengths (they fix the partition) and street anchors
    W = _prev_prime(min(window, n_cam))
    k_all = _track_lengths(np.random.default_rng([seed, 1]), n_pt, n_obs, min(200, W), 2)
    off_all = np.concatenate([[0], np.cumsum(k_all)])
    anchor_all = np.sort(np.random.default_rng([seed, 2]).integers(0, n_cam, size=n_pt)) if street else None
    if shard is not None:
        rank, world = shard
        b = partition_by_observations(off_all, world)
        point_range = (int(b[rank]), int(b[rank + 1]))
    p0, p1 = (0, n_pt) if point_range is None else point_range

    ci_l, pi_l, X_l, X0_l, obs_l = [], [], [], [], []
    for c in range(p0 // POINT_CHUNK, (max(p1, p0 + 1) - 1) // POINT_CHUNK + 1):
        a, b = c * POINT_CHUNK, min((c + 1) * POINT_CHUNK, n_pt)
        if a >= b:
            continue
        crng = np.random.default_rng([seed, 3, c])
        m = b - a
        k = k_all[a:b]
        stride = crng.integers(max(1, W // 8), W, size=m) if W > 2 else np.ones(m, dtype=np.int64)
        anchor = anchor_all[a:b] if street else crng.integers(0, n_cam, size=m)
        start = crng.integers(0, W, size=m)
        pt = np.repeat(np.arange(m, dtype=np.int64), k)
        offs = np.cumsum(k) - k
        j = np.arange(pt.shape[0], dtype=np.int64) - offs[pt]
        slot = (start[pt] + j * stride[pt]) % W
        ci = (anchor[pt] + slot - W // 2) % n_cam
        if street:
            # a point sits on the facade opposite the middle of its camera window, 15..40 deep
            mid = ci[offs]                                   # any camera of the track fixes the arc position
            rel = ((ci - mid[pt] + n_cam // 2) % n_cam) - n_cam // 2
            centre = mid + np.round(np.bincount(pt, weights=rel, minlength=m) / k)
            theta = 2 * np.pi * (centre + crng.uniform(-0.5, 0.5, m)) / n_cam
            rad = radius - crng.uniform(15, 40, m)
            X = np.stack([rad * np.cos(theta), rad * np.sin(theta), crng.uniform(-6, 6, m)], 1)
        else:
            # points in a ball of radius 10 around the origin
            X = crng.normal(size=(m, 3))
            X *= (10.0 * crng.uniform(0, 1, (m, 1)) ** (1 / 3)) / np.linalg.norm(X, axis=1, keepdims=True)
        obs, depth = project_numpy(model_id, X[pt], cam_gt[ci], pps[ci])
        assert depth.min() > 0.1
        obs += crng.normal(scale=noise_px, size=obs.shape)
        out = crng.uniform(size=obs.shape[0]) < outlier_frac
        obs[out] += crng.uniform(-20, 20, size=(int(out.sum()), 2))
        X0 = X + crng.normal(scale=perturb * 0.6 / np.sqrt(3), size=X.shape)   # 2 % of depth
        # cut the chunk to the requested point range
        lo, hi = max(p0, a) - a, min(p1, b) - a
        sel = slice(int(offs[lo]) if lo < m else pt.shape[0], int(offs[hi]) if hi < m else pt.shape[0])
        ci_l.append(ci[sel].astype(np.int32)); pi_l.append((pt[sel] + a - p0).astype(np.int32))
        X_l.append(X[lo:hi]); X0_l.append(X0[lo:hi]); obs_l.append(obs[sel])
    cat = lambda xs, shape: np.concatenate(xs, axis=0) if xs else np.zeros(shape)
    out = BAArrays(model_id, cam0, pps, cat(X0_l, (0, 3)), cat(obs_l, (0, 2)), cat(ci_l, (0,)).astype(np.int32),
                   cat(pi_l, (0,)).astype(np.int32), cam_gt, cat(X_l, (0, 3)))
    out.point_range = (p0, p1)
    out.n_pt_total, out.n_obs_total = int(n_pt), int(n_obs)
    return out


def make_config(name, scale=1.0, model_id=3, shard=None):
    """BASELINE.json config by name ('C1','C2','C3','C5'); ``scale`` shrinks / grows points and
    observations (cameras fixed).  ``shard = (rank, world)``: this rank's contiguous range of the
    points of the ONE global instance (balanced by observation count)."""
    n_cam, n_pt, n_obs, window, seed = CONFIGS[name]
    if scale != 1.0:
        n_pt = max(int(n_pt * scale), 16)
        n_obs = max(int(n_obs * scale), 2 * n_pt)
    return make_ba_problem(n_cam, n_pt, n_obs, window, seed, model_id, shard=shard)


@dataclass
class GPArrays:
    """Flat global-positioning problem (global_positioning.py:101-152)."""
    camera_translations: np.ndarray   # [Nc, 3] camera centres (initial)
    points_3d: np.ndarray             # [Np, 3] (initial)
    translations: np.ndarray          # [N, 3] world-frame unit rays
    camera_indices: np.ndarray        # int32 [N]
    point_indices: np.ndarray         # int32 [N]
    is_calibrated: np.ndarray         # bool [Nc]
    scales: np.ndarray                # [N, 1]
    gt_centres: np.ndarray = field(default=None, repr=False)
    gt_points_3d: np.ndarray = field(default=None, repr=False)


def make_gp_problem(n_cam, n_pt, n_obs, window=None, seed=0, ray_noise_deg=0.2, outlier_frac=0.02):
    rng = np.random.default_rng(seed)
    window = n_cam if window is None else window
    _, C = _orbit_cameras(rng, n_cam)
    X = rng.normal(size=(n_pt, 3))
    X *= (10.0 * rng.uniform(0, 1, (n_pt, 1)) ** (1 / 3)) / np.linalg.norm(X, axis=1, keepdims=True)
    ci, pi = _visibility(rng, n_cam, n_pt, n_obs, window, kmin=3)
    d = X[pi] - C[ci]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    dq = rotvec_to_quat(rng.normal(scale=np.deg2rad(ray_noise_deg) / np.sqrt(3), size=d.shape))
    d = np.einsum("nij,nj->ni", quat_to_mat(dq), d)
    n_out = int(outlier_frac * d.shape[0])
    if n_out:
        which = rng.choice(d.shape[0], size=n_out, replace=False)
        o = rng.normal(size=(n_out, 3))
        d[which] = o / np.linalg.norm(o, axis=1, keepdims=True)
    cal = rng.uniform(size=n_cam) < 0.9
    c0 = 100.0 * rng.uniform(-1, 1, (n_cam, 3))       # global_positioning.py:31-35, seeded here
    X0 = 100.0 * rng.uniform(-1, 1, (n_pt, 3))
    return GPArrays(c0, X0, d, ci, pi, cal, np.ones((d.shape[0], 1)), C, X)


def make_gp_config(name="C4", scale=1.0):
    n_cam, n_pt, n_obs, window, seed = GP_CONFIGS[name]
    if scale != 1.0:
        n_pt = max(int(n_pt * scale), 16)
        n_obs = max(int(n_obs * scale), 3 * n_pt)
    return make_gp_problem(n_cam, n_pt, n_obs, window, seed)


# ---------------------------------------------------------------------------------------
# scene objects (cameras / images / tracks) for the drop-in processors
# ---------------------------------------------------------------------------------------
_PP_IDX = {0: [1, 2], 1: [2, 3], 2: [1, 2], 3: [1, 2], 4: [2, 3], 5: [2, 3], 6: [2, 3], 8: [1, 2], 9: [1, 2]}


def ba_arrays_to_scene(a, unregistered=(), short_track_every=0, rng=None):
    """cameras (one per image), images, tracks for TorchBA.Solve from a flat problem.

    ``unregistered``: image ids flagged is_registered=False (their observations must be
    skipped); ``short_track_every``: every n-th track is cut to one observation (must be
    dropped by min_num_view_per_track).  Track keys are deliberately not 0..n-1."""
    from .geometry import pose7_to_matrices
    from .scene.defs import Camera, CameraModelId, Image, Track
    model = CameraModelId(a.model_id)
    n_full = a.camera_params.shape[1] - 7 + 2
    pp_idx = _PP_IDX[a.model_id]
    rest = [i for i in range(n_full) if i not in pp_idx]
    mats = pose7_to_matrices(a.camera_params[:, :7])
    cameras, images = [], []
    feat_id = np.zeros(a.n_obs, dtype=np.int64)
    order = np.argsort(a.camera_indices, kind="stable")
    cam_sorted = a.camera_indices[order]
    starts = np.searchsorted(cam_sorted, np.arange(a.n_cam + 1))
    feat_id[order] = np.arange(a.n_obs) - starts[cam_sorted]
    for i in range(a.n_cam):
        full = np.zeros(n_full)
        full[rest] = a.camera_params[i, 7:]
        full[pp_idx] = a.camera_pps[i]
        cameras.append(Camera(id=i, model_id=model, params=full, has_prior_focal_length=True))
        feats = a.points_2d[order[starts[i]:starts[i + 1]]]
        images.append(Image(id=i, cam_id=i, is_registered=i not in set(unregistered), world2cam=mats[i].copy(),
                            features=feats.copy()))
    tracks = {}
    pt_starts = np.searchsorted(a.point_indices, np.arange(a.n_pt + 1))
    for p in range(a.n_pt):
        sl = slice(pt_starts[p], pt_starts[p + 1])
        obs = np.stack([a.camera_indices[sl].astype(np.int64), feat_id[sl]], axis=1)
        if short_track_every and p % short_track_every == 0:
            obs = obs[:1]
        tracks[10 + 3 * p] = Track(id=10 + 3 * p, xyz=a.points_3d[p].copy(), observations=obs)
    return cameras, images, tracks


def gp_arrays_to_scene(g, seed=0):
    """cameras, images (random rotations, features_undist = R d), tracks for TorchGP.Optimize."""
    from .geometry import quat_to_mat
    from .scene.defs import Camera, CameraModelId, Image, Track
    rng = np.random.default_rng(seed)
    n_cam, n_obs = g.camera_translations.shape[0], g.translations.shape[0]
    q = rng.normal(size=(n_cam, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    R = quat_to_mat(q)
    order = np.argsort(g.camera_indices, kind="stable")
    cam_sorted = g.camera_indices[order]
    starts = np.searchsorted(cam_sorted, np.arange(n_cam + 1))
    feat_id = np.zeros(n_obs, dtype=np.int64)
    feat_id[order] = np.arange(n_obs) - starts[cam_sorted]
    cameras, images = [], []
    for i in range(n_cam):
        cameras.append(Camera(id=i, model_id=CameraModelId.SIMPLE_PINHOLE, params=[1000.0, 0.0, 0.0],
                              has_prior_focal_length=bool(g.is_calibrated[i])))
        M = np.eye(4)
        M[:3, :3] = R[i]
        M[:3, 3] = g.camera_translations[i]          # holds the camera CENTRE at this stage (:31-35)
        d = g.translations[order[starts[i]:starts[i + 1]]]
        images.append(Image(id=i, cam_id=i, is_registered=True, world2cam=M, features_undist=d @ R[i].T))
    tracks = {}
    n_pt = g.points_3d.shape[0]
    pt_starts = np.searchsorted(g.point_indices, np.arange(n_pt + 1))
    for p in range(n_pt):
        sl = slice(pt_starts[p], pt_starts[p + 1])
        tracks[5 + 2 * p] = Track(id=5 + 2 * p, xyz=g.points_3d[p].copy(),
                                  observations=np.stack([g.camera_indices[sl].astype(np.int64), feat_id[sl]], axis=1))
    return cameras, images, tracks


def make_filter_scene(n_img=12, n_trk=300, mean_len=4.0, seed=0, outlier_frac=0.15, behind_frac=0.05):
    """Seeded scene for the track filters (processors/track_filter.py): images with poses and unit
    bearings (``features_undist``), tracks with points and (image_id, feature_id) observations.
    Bearings are the exact direction of the point in the camera plus noise; a fraction are gross
    outliers, a fraction of the points sit behind one of their cameras, some tracks have a single
    view, one is empty, one sees the same image twice.  Returns (cameras, images, tracks)."""
    from .scene.defs import Image, Track
    rng = np.random.default_rng(seed)
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
        images.append(Image(id=i, cam_id=0, is_registered=True, world2cam=w2c, features_undist=[]))
    feats = [[] for _ in range(n_img)]
    tracks = {}
    for t in range(n_trk):
        X = rng.normal(0, 2.0, 3)
        far = rng.random() < 0.2
        if far:   # far points: small triangulation angles
            X = X / np.linalg.norm(X) * rng.uniform(300, 3000)
        k = 0 if t == 7 else (1 if t % 23 == 0 else min(n_img, 2 + rng.geometric(1.0 / max(mean_len - 1.0, 1.0))))
        ids = rng.choice(n_img, size=k, replace=False) if not far else (rng.integers(0, n_img) + np.arange(k)) % n_img
        if t == 11 and k >= 2:
            ids[1] = ids[0]   # the same image twice in one track
        if k and rng.random() < behind_frac:
            c = images[ids[0]].center()
            X = c + (c - X) * 0.5   # behind (or very near) the first camera
        obs = []
        for i in ids:
            p = images[i].world2cam[:3, :3] @ X + images[i].world2cam[:3, 3]
            b = p / np.linalg.norm(p) + rng.normal(0, 2e-3, 3)
            if rng.random() < outlier_frac:
                b = b + rng.normal(0, 0.2, 3)
            b /= np.linalg.norm(b)
            obs.append((int(i), len(feats[i])))
            feats[i].append(b)
        tracks[1000 + 3 * t] = Track(id=1000 + 3 * t, xyz=X, observations=np.array(obs, dtype=np.int64).reshape(-1, 2))
    for i, img in enumerate(images):
        img.features_undist = np.array(feats[i]).reshape(-1, 3)
    return [], images, tracks
