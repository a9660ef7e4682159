"""Drop-in for instantsfm/processors/global_positioning.py (TorchGP): same class, same
signatures and in-place mutations; the LM solve runs in the CUDA library."""
import numpy as np

from ..engine import GPEngine
from ._common import concat_features, device_index, flatten_observations, should_stop


class TorchGP:
    def __init__(self, visualizer=None, device="cuda:0", dtype=np.float32, pcg_tol=1e-6):
        self.device = device
        self.visualizer = visualizer
        self.dtype = dtype
        self.pcg_tol = pcg_tol
        self.loss_history = []

    def InitializeRandomPositions(self, cameras, images, tracks, depths=None):
        """global_positioning.py:23-39 (unseeded np.random, exactly as the reference)."""
        scene_scale = 100
        if depths is not None:
            valid_depths = depths[depths > 0]
            if len(valid_depths):
                scene_scale = np.mean(valid_depths) * 4.0
        for image in images:
            image.world2cam[:3, 3] = scene_scale * np.random.uniform(-1, 1, 3)
        for track in tracks.values():
            track.xyz = scene_scale * np.random.uniform(-1, 1, 3)
            track.is_initialized = True
        if self.visualizer:
            self.visualizer.add_step(cameras, images, tracks)

    def ConvertResults(self, images):
        """global_positioning.py:41-43: camera centre -> translation, t = -R c."""
        for image in images:
            image.world2cam[:3, 3] = -(image.world2cam[:3, :3] @ image.world2cam[:3, 3])

    def Optimize(self, cameras, images, tracks, depths, GLOBAL_POSITIONER_OPTIONS, depth_only=False):
        opts = GLOBAL_POSITIONER_OPTIONS
        if depth_only and depths is None:
            print("Warning: No depth maps provided, skip depth-only optimization.")
            return
        # :85-99 prune short tracks, unregister images without tracks (mutates the inputs)
        for track_id in list(tracks.keys()):
            if tracks[track_id].observations.shape[0] < opts["min_num_view_per_track"]:
                del tracks[track_id]
        image_used = np.zeros(len(images), dtype=bool)
        track_keys = list(tracks.keys())
        image_id, feature_id, which = flatten_observations(tracks, track_keys)
        image_used[np.unique(image_id)] = True
        for i, image in enumerate(images):
            if not image_used[i]:
                image.is_registered = False

        # :101-152 tensors
        registered = np.array([img.is_registered for img in images], dtype=bool)
        reg_ids = np.flatnonzero(registered)
        image_id2idx = -np.ones(len(images), dtype=np.int64)
        image_id2idx[reg_ids] = np.arange(reg_ids.size)
        centres = np.stack([images[i].world2cam[:3, 3] for i in reg_ids], 0).astype(np.float64)
        points_3d = np.stack([np.asarray(t.xyz, dtype=np.float64) for t in tracks.values()], 0)
        keep = registered[image_id]
        scales = scale_fixed = None
        if depths is not None:
            dtable, doff = concat_features(images, "depths")
            depth = dtable.reshape(-1)[doff[image_id] + feature_id]
            available = depth != 0
            if depth_only:
                keep &= available
            scales = 1.0 / np.where(available, depth, 1.0)
            scale_fixed = available
        image_id, feature_id, which = image_id[keep], feature_id[keep], which[keep]
        rot = np.stack([img.world2cam[:3, :3] for img in images], 0)
        ftable, foff = concat_features(images, "features_undist")
        feats = ftable[foff[image_id] + feature_id].reshape(-1, 3)
        translations = np.einsum("nji,nj->ni", rot[image_id], feats)                 # R^T f  (:135)
        n = translations.shape[0]
        if depths is None:
            scales_t = np.ones((n, 1))
            fixed_t = None
        else:
            scales_t = scales[keep].reshape(-1, 1)
            fixed_t = scale_fixed[keep]
        is_calibrated = np.array([cameras[images[i].cam_id].has_prior_focal_length for i in reg_ids], dtype=bool)
        if n == 0:
            return

        import torch
        with torch.cuda.device(device_index(self.device)):
            # the handle is destroyed on every exit path, under the device it was created on
            with GPEngine(huber_delta=opts["thres_loss_function"], dtype=self.dtype, optimize_scales=not depth_only,
                          pcg_tol=self.pcg_tol) as engine:
                engine.set_problem(centres, points_3d, scales_t, translations, image_id2idx[image_id].astype(np.int32),
                                   which.astype(np.int32), is_calibrated, fixed_t)

                def write_back():
                    c, X, _ = engine.get_params()
                    c, X = c.astype(np.float64), X.astype(np.float64)
                    for track, xyz in zip(tracks.values(), X):
                        track.xyz = xyz
                    for idx, iid in enumerate(reg_ids.tolist()):
                        images[iid].world2cam[:3, 3] = c[idx]
                    self.ConvertResults(images)

                self.loss_history = []
                for _ in range(opts["max_num_iterations"]):
                    loss, _ = engine.step()
                    self.loss_history.append(loss)
                    if should_stop(self.loss_history, opts["function_tolerance"], identical_test=False):
                        break
                    if self.visualizer:
                        write_back()
                        self.visualizer.add_step(cameras, images, tracks, "global_positioning")
                write_back()
