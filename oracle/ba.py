"""Bundle adjustment of the oracle.  TEST INFRASTRUCTURE.

Restates /root/reference/instantsfm/processors/bundle_adjustment.py:
    :66-100  track flattening (A1)           -> flatten_tracks
    :75-80   principal-point split (A2)      -> split_principal_point
    :102-113 cheirality filter + compaction  -> cheirality_and_compact
    :51-64   ReprojNonBatched.forward        -> BAProblem.residuals
    :116-142 optimiser set-up, loop, stop    -> solve_arrays
    :18-36   write-back                      -> write_back
Jacobians are taken by torch.func (vmap(jacrev)) through the left retraction of
oracle/lie.py, i.e. they are by construction consistent with ``retract``.
"""
import numpy as np
import scipy.sparse as sp
import torch

from . import lie
from .camera_models import PP_INDICES, NUM_PARAMS, n_intrinsics, reproject
from .lm import LM, TrustRegion, run_loop


class BAProblem:
    """Flat-array BA problem in the compacted index space of bundle_adjustment.py:108-113."""

    def __init__(self, model_id, camera_params, camera_pps, points_3d, points_2d,
                 camera_indices, point_indices, optimize_poses=True):
        self.model_id = int(model_id)
        self.ni = n_intrinsics(self.model_id)
        self.d = 6 + self.ni
        self.cam = np.array(camera_params, dtype=np.float64)
        self.pps = np.array(camera_pps, dtype=np.float64)
        self.pts = np.array(points_3d, dtype=np.float64)
        self.obs = np.array(points_2d, dtype=np.float64)
        self.ci = np.asarray(camera_indices, dtype=np.int64)
        self.pi = np.asarray(point_indices, dtype=np.int64)
        self.optimize_poses = bool(optimize_poses)
        assert self.cam.shape[1] == 7 + self.ni
        self.n_cam, self.n_pt, self.n_obs = self.cam.shape[0], self.pts.shape[0], self.obs.shape[0]
        self.n_lead = self.n_cam * self.d if self.optimize_poses else 0   # unknowns ahead of the point blocks

    # -- residuals: bundle_adjustment.py:59-64 -------------------------------------------
    def residuals(self):
        with torch.no_grad():
            proj = reproject(self.model_id, torch.from_numpy(self.pts[self.pi]),
                             torch.from_numpy(self.cam[self.ci]), torch.from_numpy(self.pps[self.ci]))
        return proj.numpy() - self.obs

    # -- per-observation Jacobian blocks (unweighted) ------------------------------------
    def blocks(self):
        """Returns r[N,2], Jc[N,2,d] (pose tangent 6 then intrinsics), Jp[N,2,3]."""
        ni, mid = self.ni, self.model_id
        cam = torch.from_numpy(self.cam[self.ci])
        X = torch.from_numpy(self.pts[self.pi])
        pp = torch.from_numpy(self.pps[self.ci])
        z = torch.cat([torch.zeros(self.n_obs, 6, dtype=torch.float64), cam[:, 7:], X], dim=1)
        z.requires_grad_(True)
        pose = lie.se3_retract(cam[:, :7], z[:, :6])
        proj = reproject(mid, z[:, 6 + ni:], torch.cat([pose, z[:, 6:6 + ni]], dim=1), pp)
        # rows are independent, so d(sum_i proj[i,k]) / dz[i] is row i of the Jacobian
        J = torch.stack([torch.autograd.grad(proj[:, k].sum(), z, retain_graph=(k == 0))[0]
                         for k in range(2)], dim=1).numpy()
        return self.residuals(), J[:, :, :self.d], J[:, :, self.d:]

    def jacobian(self):
        _, Jc, Jp = self.blocks()
        n, d = self.n_obs, self.d
        rows = (2 * np.arange(n)[:, None] + np.arange(2)[None, :])  # [n, 2]
        if self.optimize_poses:
            off = self.n_cam * d
            ccols = self.ci[:, None] * d + np.arange(d)[None, :]       # [n, d]
            r_c = np.broadcast_to(rows[:, :, None], (n, 2, d)).reshape(-1)
            c_c = np.broadcast_to(ccols[:, None, :], (n, 2, d)).reshape(-1)
        else:
            off = 0
        pcols = off + self.pi[:, None] * 3 + np.arange(3)[None, :]
        r_p = np.broadcast_to(rows[:, :, None], (n, 2, 3)).reshape(-1)
        c_p = np.broadcast_to(pcols[:, None, :], (n, 2, 3)).reshape(-1)
        if self.optimize_poses:
            r_all = np.concatenate([r_c, r_p]); c_all = np.concatenate([c_c, c_p])
            v_all = np.concatenate([Jc.reshape(-1), Jp.reshape(-1)])
        else:
            r_all, c_all, v_all = r_p, c_p, Jp.reshape(-1)
        return sp.csr_matrix((v_all, (r_all, c_all)), shape=(2 * n, off + 3 * self.n_pt))

    # -- update: Exp(delta) * X for the pose, += for the rest ----------------------------
    def retract(self, D):
        d = self.d
        if self.optimize_poses:
            Dc = D[:self.n_cam * d].reshape(self.n_cam, d)
            pose = lie.se3_retract(torch.from_numpy(self.cam[:, :7]), torch.from_numpy(Dc[:, :6].copy()))
            self.cam[:, :7] = pose.numpy()
            self.cam[:, 7:] += Dc[:, 6:]
            Dp = D[self.n_cam * d:]
        else:
            Dp = D
        self.pts += Dp.reshape(self.n_pt, 3)

    def snapshot(self):
        return self.cam.copy(), self.pts.copy()

    def restore(self, s):
        self.cam, self.pts = s[0].copy(), s[1].copy()

    def rmse(self):
        r = self.residuals()
        return float(np.sqrt((r * r).sum(-1).mean()))


# ---------------------------------------------------------------------------------------
# tracks / images / cameras level (python structures, small problems only)
# ---------------------------------------------------------------------------------------

def flatten_tracks(cameras, images, tracks, options):
    """bundle_adjustment.py:66-100.  Returns dict of numpy arrays (pre-filter)."""
    track_keys = list(tracks.keys())
    track_lengths = np.array([len(tracks[k].observations) for k in track_keys])
    valid = track_lengths >= options["min_num_view_per_track"]
    registered = np.array([img.is_registered for img in images], dtype=bool)
    rows = []
    for img in images:
        pose = lie.matrix_to_pose7(img.world2cam) if img.is_registered else np.array([0, 0, 0, 0, 0, 0, 1.0])
        rows.append(np.concatenate([pose, np.asarray(cameras[img.cam_id].params, dtype=np.float64)]))
    camera_params = np.stack(rows, 0)
    points_3d = np.stack([t.xyz for t in tracks.values()], 0).astype(np.float64)
    p2d, ci, pi = [], [], []
    for tid in valid.nonzero()[0]:
        for image_id, feature_id in tracks[track_keys[tid]].observations:
            if not registered[image_id]:
                continue
            p2d.append(images[image_id].features[feature_id])
            ci.append(image_id)
            pi.append(tid)
    return {"track_keys": track_keys, "camera_params": camera_params, "points_3d": points_3d,
            "points_2d": np.array(p2d, dtype=np.float64).reshape(-1, 2),
            "camera_indices": np.array(ci, dtype=np.int32), "point_indices": np.array(pi, dtype=np.int32)}


def split_principal_point(model_id, camera_params):
    """bundle_adjustment.py:75-80."""
    pp_idx = np.array(PP_INDICES[model_id]) + 7
    rest = np.array([i for i in range(camera_params.shape[1]) if i not in pp_idx])
    return camera_params[:, rest], camera_params[:, pp_idx], rest, pp_idx


def cheirality_and_compact(camera_params, camera_pps, points_3d, points_2d, camera_indices, point_indices):
    """bundle_adjustment.py:102-113: keep z > 0.1 (evaluated once), then torch.unique compaction."""
    y = lie.rotate_quat(torch.from_numpy(points_3d[point_indices]),
                        torch.from_numpy(camera_params[camera_indices][:, :7])).numpy()
    keep = y[:, 2] > 0.1
    p2d, ci, pi = points_2d[keep], camera_indices[keep], point_indices[keep]
    ucam, ci_ = np.unique(ci, return_inverse=True)
    upt, pi_ = np.unique(pi, return_inverse=True)
    return {"keep": keep, "unique_cameras": ucam, "unique_points": upt,
            "camera_indices": ci_.astype(np.int64), "point_indices": pi_.astype(np.int64),
            "points_2d": p2d, "camera_params": camera_params[ucam], "camera_pps": camera_pps[ucam],
            "points_3d": points_3d[upt]}


def make_optimizer(problem, huber_delta, solver="pcg", pcg_tol=1e-5):
    """bundle_adjustment.py:116-119."""
    strategy = TrustRegion(radius=1e4, max=1e10, up=2.0, down=0.5 ** 4)
    return LM(problem, strategy, huber_delta, solver=solver, pcg_tol=pcg_tol, reject=30)


def solve_arrays(problem, options, solver="pcg", pcg_tol=1e-5):
    """bundle_adjustment.py:115-142 on an already compacted problem.  Returns (history, optimizer)."""
    opt = make_optimizer(problem, options["thres_loss_function"], solver, pcg_tol)
    hist = run_loop(opt, options["max_num_iterations"], options["function_tolerance"], stop_on_identical=True)
    return hist, opt


def write_back(cameras, images, tracks, flat, comp, rest, pp_idx, problem):
    """bundle_adjustment.py:18-36 (images sharing a camera overwrite each other; last wins)."""
    full = np.zeros((problem.n_cam, problem.cam.shape[1] + 2))
    full[:, rest] = problem.cam
    full[:, pp_idx] = problem.pps
    mats = lie.pose7_to_matrix(torch.from_numpy(full[:, :7])).numpy()
    for i, orig in enumerate(comp["unique_points"].tolist()):
        tracks[flat["track_keys"][orig]].xyz = problem.pts[i].copy()
    for i, image_id in enumerate(comp["unique_cameras"].tolist()):
        img = images[image_id]
        img.world2cam = mats[i]
        cameras[img.cam_id].set_params(full[i, 7:])


def solve(cameras, images, tracks, options, solver="pcg", pcg_tol=1e-5):
    """TorchBA.Solve restated end to end (bundle_adjustment.py:44-154); mutates its inputs."""
    model_id = cameras[0].model_id.value if hasattr(cameras[0].model_id, "value") else int(cameras[0].model_id)
    if model_id not in NUM_PARAMS:
        raise NotImplementedError("Unsupported camera model")
    flat = flatten_tracks(cameras, images, tracks, options)
    cam, pps, rest, pp_idx = split_principal_point(model_id, flat["camera_params"])
    comp = cheirality_and_compact(cam, pps, flat["points_3d"], flat["points_2d"],
                                  flat["camera_indices"], flat["point_indices"])
    pb = BAProblem(model_id, comp["camera_params"], comp["camera_pps"], comp["points_3d"], comp["points_2d"],
                   comp["camera_indices"], comp["point_indices"], options["optimize_poses"])
    hist, opt = solve_arrays(pb, options, solver, pcg_tol)
    write_back(cameras, images, tracks, flat, comp, rest, pp_idx, pb)
    return hist, opt, pb
