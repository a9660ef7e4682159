"""Drop-in for instantsfm/processors/bundle_adjustment.py: same class, same signatures, same
in-place mutation of ``cameras`` / ``images`` / ``tracks`` -- the LM solve runs in the CUDA
library (include/isfm_b200.h) instead of bae + pypose.

Differences a caller can observe: (i) arithmetic is fp32 by default (``dtype=np.float64``
selects the validation build), (ii) the O(N_obs) Python loops of the reference
(:85-100, :29-36) are vectorised numpy producing the identical tensors, (iii) the linear
system is solved by Schur complement + block-Jacobi PCG (tolerance 1e-6 on the reduced
system) instead of full-system Jacobi PCG (1e-5).
"""
import numpy as np

from ..engine import BAEngine
from ..geometry import matrices_to_pose7, pose7_to_matrices, quat_to_mat as quat_rows_to_matrices
from ._common import concat_features, device_index, should_stop

# CameraModelId.value -> principal-point indices inside Camera.params (scene/defs.py:115-140).
# FOV (7) and THIN_PRISM_FISHEYE (10) are listed by the reference but their cost functions
# raise NotImplementedError (utils/cost_function.py:128,182).
_PP = {0: [1, 2], 1: [2, 3], 2: [1, 2], 3: [1, 2], 4: [2, 3], 5: [2, 3], 6: [2, 3], 8: [1, 2], 9: [1, 2]}


def _model_value(model_id):
    return model_id.value if hasattr(model_id, "value") else int(model_id)


def update(cameras, images, tracks, track_keys, unique_cameras, unique_points, remaining_indices, pp_indices,
           camera_params, camera_pps, points_3d, write_cameras=True):
    """bundle_adjustment.py:18-36: write the optimised tensors back into the scene objects
    (images sharing a camera overwrite each other's intrinsics; the last one wins).
    ``write_cameras=False`` (points-only BA, ``optimize_poses=False``): poses and intrinsics were
    not variables, so the scene keeps its fp64 values untouched instead of a round trip through
    the solver's precision."""
    track_objs = list(tracks.values()) if len(unique_points) == len(track_keys) else None
    if track_objs is not None and np.array_equal(unique_points, np.arange(len(track_keys))):
        for track, xyz in zip(track_objs, points_3d):       # every track is a variable: no key look-ups
            track.xyz = xyz
    else:
        for key, xyz in zip([track_keys[i] for i in unique_points.tolist()], points_3d):
            tracks[key].xyz = xyz
    if not write_cameras:
        return
    full = np.zeros((camera_params.shape[0], camera_params.shape[1] + 2))
    full[:, remaining_indices] = camera_params
    full[:, pp_indices] = camera_pps
    pose_matrices = pose7_to_matrices(full[:, :7])
    for i, image_id in enumerate(unique_cameras.tolist()):
        image = images[image_id]
        image.world2cam = pose_matrices[i]
        cameras[image.cam_id].set_params(full[i, 7:])


class TorchBA:
    def __init__(self, visualizer=None, device="cuda:0", dtype=np.float32, pcg_tol=1e-6):
        self.device = device
        self.visualizer = visualizer
        self.dtype = dtype
        self.pcg_tol = pcg_tol
        self.loss_history = []
        self.last_stats = []
        self.last_timing = {}      # seconds per host phase of the last Solve (flatten, set-up, steps, write-back)

    # -- tensor set-up, bundle_adjustment.py:66-113 ---------------------------------------
    def _build(self, cameras, images, tracks, options, model_value):
        # ONE pass over the track objects (a million of them on a C3-sized scene: every Python-level
        # touch costs ~0.3 s there): keys, observation arrays and points are collected together, the
        # observations of ALL tracks are concatenated once and the invalid tracks masked out afterwards
        track_keys = list(tracks.keys())
        track_objs = list(tracks.values())
        obs_list = [t.observations for t in track_objs]
        track_lengths = np.fromiter(map(len, obs_list), dtype=np.int64, count=len(obs_list))
        is_track_valid = track_lengths >= options["min_num_view_per_track"]
        registered = np.array([img.is_registered for img in images], dtype=bool)

        poses = np.tile(np.array([0, 0, 0, 0, 0, 0, 1.0]), (len(images), 1))        # pp.identity_SE3()
        if registered.any():
            mats = np.stack([images[i].world2cam for i in np.flatnonzero(registered)], 0)
            poses[registered] = matrices_to_pose7(mats)
        intr = np.stack([np.asarray(cameras[img.cam_id].params, dtype=np.float64) for img in images], 0)
        camera_params = np.concatenate([poses, intr], axis=1)
        pp_indices = np.array(_PP[model_value]) + 7
        remaining = np.array([i for i in range(camera_params.shape[1]) if i not in pp_indices])
        camera_pps = camera_params[:, pp_indices]
        camera_params = camera_params[:, remaining]
        points_3d = np.array([t.xyz for t in track_objs], dtype=np.float64).reshape(len(track_objs), 3)

        valid_ids = np.flatnonzero(is_track_valid)
        if track_lengths.sum() == 0:
            image_id = feature_id = which = np.zeros(0, dtype=np.int64)
        else:
            try:       # the common case: every track holds an [k, 2] integer array
                obs = np.concatenate(obs_list, axis=0)
                if obs.ndim != 2 or obs.shape[1] != 2:
                    raise ValueError
            except ValueError:   # lists of tuples, empty lists, 1-d arrays: normalise per track
                obs = np.concatenate([np.asarray(o).reshape(-1, 2) for o in obs_list], axis=0)
            obs = obs.astype(np.int64, copy=False)
            owner = np.repeat(np.arange(len(obs_list), dtype=np.int64), track_lengths)
            sel = is_track_valid[owner]
            image_id, feature_id, which = obs[sel, 0], obs[sel, 1], owner[sel]
        keep = registered[image_id] if image_id.size else np.zeros(0, bool)
        image_id, feature_id, point_idx = image_id[keep], feature_id[keep], which[keep]
        table, offsets = concat_features(images, "features")
        points_2d = table[offsets[image_id] + feature_id].reshape(-1, 2)

        # cheirality, evaluated once before the solve (:102-107): only the depth y.z = R[2,:] X + t_z is
        # needed, with ONE rotation matrix per image instead of one per observation
        if image_id.size:
            R = quat_rows_to_matrices(camera_params[:, 3:7])
            z = np.einsum("nj,nj->n", R[image_id, 2, :], points_3d[point_idx]) + camera_params[image_id, 2]
            valid = z > 0.1
        else:
            valid = np.zeros(0, bool)
        points_2d, image_id, point_idx = points_2d[valid], image_id[valid], point_idx[valid]
        # torch.unique(sorted=True, return_inverse=True) (:108-109) without a sort: images through a
        # presence table, points through the run boundaries of the (non-decreasing) point index
        present = np.zeros(len(images), dtype=bool)
        present[image_id] = True
        unique_cameras = np.flatnonzero(present)
        cam_inv = (np.cumsum(present) - 1)[image_id]
        if point_idx.size and np.all(point_idx[1:] >= point_idx[:-1]):
            new_run = np.concatenate([[True], point_idx[1:] != point_idx[:-1]])
            unique_points = point_idx[new_run]
            pt_inv = np.cumsum(new_run) - 1
        else:
            unique_points, pt_inv = np.unique(point_idx, return_inverse=True)
        return {"track_keys": track_keys, "unique_cameras": unique_cameras, "unique_points": unique_points,
                "remaining": remaining, "pp_indices": pp_indices,
                "camera_params": camera_params[unique_cameras], "camera_pps": camera_pps[unique_cameras],
                "points_3d": points_3d[unique_points], "points_2d": points_2d,
                "camera_indices": cam_inv.astype(np.int32), "point_indices": pt_inv.astype(np.int32)}

    def Solve(self, cameras, images, tracks, BUNDLE_ADJUSTER_OPTIONS):
        self.camera_model = cameras[0].model_id  # assume all cameras are under the same model
        model_value = _model_value(self.camera_model)
        if model_value not in _PP:
            raise NotImplementedError("Unsupported camera model")
        opts = BUNDLE_ADJUSTER_OPTIONS
        import time
        t0 = time.perf_counter()
        t = self._build(cameras, images, tracks, opts, model_value)
        self.last_timing = {"flatten": time.perf_counter() - t0}
        if t["points_2d"].shape[0] == 0:
            return

        import torch
        with torch.cuda.device(device_index(self.device)):
            t0 = time.perf_counter()
            # the handle is destroyed on every exit path, under the device it was created on
            with BAEngine(model_value, optimize_poses=opts["optimize_poses"], huber_delta=opts["thres_loss_function"],
                          dtype=self.dtype, pcg_tol=self.pcg_tol) as engine:
                engine.set_problem(t["camera_params"], t["camera_pps"], t["points_3d"], t["points_2d"],
                                   t["camera_indices"], t["point_indices"])
                self.last_timing["set_problem"] = time.perf_counter() - t0
                # the solver works in self.dtype; the scene keeps its fp64 values and receives the
                # solver's CHANGE (output - input, both in solver precision): parameters the solve left
                # untouched come back bit-identical instead of rounded through fp32
                cam_in = t["camera_params"].astype(self.dtype)
                pts_in = t["points_3d"].astype(self.dtype)

                def write_back():
                    cam, pts = engine.get_params()
                    cam64 = t["camera_params"] + (cam.astype(np.float64) - cam_in.astype(np.float64))
                    pts64 = t["points_3d"] + (pts.astype(np.float64) - pts_in.astype(np.float64))
                    update(cameras, images, tracks, t["track_keys"], t["unique_cameras"], t["unique_points"],
                           t["remaining"], t["pp_indices"], cam64, t["camera_pps"], pts64,
                           write_cameras=bool(opts["optimize_poses"]))

                t0 = time.perf_counter()
                self.loss_history, self.last_stats = [], []
                for _ in range(opts["max_num_iterations"]):
                    loss, stats = engine.step()
                    self.loss_history.append(loss)
                    self.last_stats.append(stats)
                    if should_stop(self.loss_history, opts["function_tolerance"], identical_test=True):
                        break
                    if self.visualizer:
                        write_back()
                        self.visualizer.add_step(cameras, images, tracks, "bundle_adjustment")
                self.last_timing["steps"] = time.perf_counter() - t0
                t0 = time.perf_counter()
                write_back()
                self.last_timing["write_back"] = time.perf_counter() - t0
