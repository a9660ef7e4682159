"""Global positioning of the oracle.  TEST INFRASTRUCTURE.

Restates /root/reference/instantsfm/processors/global_positioning.py:
    :85-99   short-track / unused-image pruning     -> prune
    :101-152 tensor construction                    -> flatten
    :51-71   PairwiseNonBatched.forward             -> GPProblem.residuals
    :158-184 optimiser set-up, loop, stop rule      -> solve_arrays
    :199-206, :41-43 write-back + ConvertResults    -> write_back
Unknown ordering [translations, points, scales] (nn.Parameter registration order,
global_positioning.py:54-56); rows of ``scales`` outside ``optimize_indices`` (:57-59) have
no Jacobian column.  The full system is solved, as bae does (no elimination).
"""
import numpy as np
import scipy.sparse as sp
import torch

from .camera_models import pairwise_cost
from .lm import LM, TrustRegion, run_loop


class GPProblem:
    def __init__(self, camera_translations, points_3d, scales, translations, camera_indices, point_indices,
                 is_calibrated, scale_fixed=None, depth_only=False):
        self.c = np.array(camera_translations, dtype=np.float64)
        self.X = np.array(points_3d, dtype=np.float64)
        self.s = np.array(scales, dtype=np.float64).reshape(-1, 1)
        self.d = np.array(translations, dtype=np.float64)
        self.ci = np.asarray(camera_indices, dtype=np.int64)
        self.pi = np.asarray(point_indices, dtype=np.int64)
        self.cal = np.asarray(is_calibrated, dtype=bool)
        n = self.d.shape[0]
        fixed = np.zeros(n, bool) if scale_fixed is None else np.asarray(scale_fixed, dtype=bool)
        if depth_only:                      # PairwiseNonBatchedDepthOnly :73-83
            fixed = np.ones(n, bool)
        self.free = np.flatnonzero(~fixed)  # scales.optimize_indices
        self.n_cam, self.n_pt, self.n_obs = self.c.shape[0], self.X.shape[0], n

    def residuals(self):
        r = pairwise_cost(torch.from_numpy(self.X[self.pi]), torch.from_numpy(self.c[self.ci]),
                          torch.from_numpy(self.s), torch.from_numpy(self.d), torch.from_numpy(self.cal[self.ci]))
        return r.numpy()

    def jacobian(self):
        n = self.n_obs
        w = np.where(self.cal[self.ci], 1.0, 0.5)
        ws = w * self.s[:, 0]
        e = self.X[self.pi] - self.c[self.ci]
        rows = 3 * np.arange(n)[:, None] + np.arange(3)[None, :]            # [n,3]
        ccols = 3 * self.ci[:, None] + np.arange(3)[None, :]
        pcols = 3 * self.n_cam + 3 * self.pi[:, None] + np.arange(3)[None, :]
        r_all = [rows.reshape(-1), rows.reshape(-1)]
        c_all = [ccols.reshape(-1), pcols.reshape(-1)]
        v_all = [np.repeat(ws, 3), -np.repeat(ws, 3)]                        # dr/dc = +w s I, dr/dX = -w s I
        col_of = -np.ones(n, dtype=np.int64)
        col_of[self.free] = 3 * (self.n_cam + self.n_pt) + np.arange(self.free.size)
        f = self.free
        r_all.append(rows[f].reshape(-1))
        c_all.append(np.repeat(col_of[f], 3))
        v_all.append((-w[f, None] * e[f]).reshape(-1))                       # dr/ds = -w (X - c)
        n_unk = 3 * (self.n_cam + self.n_pt) + self.free.size
        return sp.csr_matrix((np.concatenate(v_all), (np.concatenate(r_all), np.concatenate(c_all))), shape=(3 * n, n_unk))

    def retract(self, D):
        a, b = 3 * self.n_cam, 3 * (self.n_cam + self.n_pt)
        self.c += D[:a].reshape(-1, 3)
        self.X += D[a:b].reshape(-1, 3)
        self.s[self.free, 0] += D[b:]

    def snapshot(self):
        return self.c.copy(), self.X.copy(), self.s.copy()

    def restore(self, snap):
        self.c, self.X, self.s = snap[0].copy(), snap[1].copy(), snap[2].copy()


def make_optimizer(problem, huber_delta, solver="pcg", pcg_tol=1e-5):
    """global_positioning.py:158-161."""
    strategy = TrustRegion(radius=1e3, max=1e8, up=2.0, down=0.5 ** 4)
    return LM(problem, strategy, huber_delta, solver=solver, pcg_tol=pcg_tol, reject=30)


def solve_arrays(problem, options, solver="pcg", pcg_tol=1e-5):
    opt = make_optimizer(problem, options["thres_loss_function"], solver, pcg_tol)
    hist = run_loop(opt, options["max_num_iterations"], options["function_tolerance"], stop_on_identical=False)
    return hist, opt


# -- tracks / images level (python structures, small problems only) ----------------------

def prune(images, tracks, options):
    """global_positioning.py:85-99 (mutates tracks and images[*].is_registered)."""
    for tid in list(tracks.keys()):
        if tracks[tid].observations.shape[0] < options["min_num_view_per_track"]:
            del tracks[tid]
    used = np.zeros(len(images), dtype=bool)
    for t in tracks.values():
        used[np.unique(t.observations[:, 0])] = True
        if all(used):
            break
    for i, img in enumerate(images):
        if not used[i]:
            img.is_registered = False


def flatten(cameras, images, tracks, depths=None, depth_only=False):
    """global_positioning.py:101-152."""
    id2idx, idx2id = {}, {}
    for i, img in enumerate(images):
        if not img.is_registered:
            continue
        id2idx[i] = len(id2idx)
        idx2id[len(idx2id)] = i
    centres = np.stack([img.world2cam[:3, 3] for img in images if img.is_registered], 0).astype(np.float64)
    points = np.stack([t.xyz for t in tracks.values()], 0).astype(np.float64)
    rays, ci, pi, inv_depth, avail = [], [], [], [], []
    for tidx, t in enumerate(tracks.values()):
        for image_id, feature_id in t.observations:
            img = images[image_id]
            if not img.is_registered:
                continue
            if depths is not None:
                dep = img.depths[feature_id]
                if depth_only and not dep:
                    continue
                avail.append(bool(dep))
                inv_depth.append(1.0 / (dep if dep else 1.0))
            rays.append(img.world2cam[:3, :3].T @ img.features_undist[feature_id])
            ci.append(id2idx[image_id])
            pi.append(tidx)
    cal = np.array([cameras[img.cam_id].has_prior_focal_length for img in images if img.is_registered], dtype=bool)
    n = len(rays)
    if depths is None:
        scales, fixed = np.ones((n, 1)), None
    else:
        scales, fixed = np.array(inv_depth).reshape(-1, 1), np.array(avail, dtype=bool)
    return {"centres": centres, "points": points, "rays": np.array(rays).reshape(-1, 3), "camera_indices": np.array(ci, np.int32),
            "point_indices": np.array(pi, np.int32), "is_calibrated": cal, "scales": scales, "scale_fixed": fixed,
            "idx2id": idx2id}


def write_back(images, tracks, flat, problem):
    """global_positioning.py:199-206 + ConvertResults :41-43 (applied to EVERY image)."""
    for tidx, t in enumerate(tracks.values()):
        t.xyz = problem.X[tidx].copy()
    for idx in range(problem.n_cam):
        images[flat["idx2id"][idx]].world2cam[:3, 3] = problem.c[idx]
    for img in images:
        img.world2cam[:3, 3] = -(img.world2cam[:3, :3] @ img.world2cam[:3, 3])


def optimize(cameras, images, tracks, depths, options, depth_only=False, solver="pcg", pcg_tol=1e-5):
    """TorchGP.Optimize restated end to end; mutates its inputs."""
    if depth_only and depths is None:
        return None
    prune(images, tracks, options)
    flat = flatten(cameras, images, tracks, depths, depth_only)
    pb = GPProblem(flat["centres"], flat["points"], flat["scales"], flat["rays"], flat["camera_indices"],
                   flat["point_indices"], flat["is_calibrated"], flat["scale_fixed"], depth_only)
    hist, opt = solve_arrays(pb, options, solver, pcg_tol)
    write_back(images, tracks, flat, pb)
    return hist, opt, pb
