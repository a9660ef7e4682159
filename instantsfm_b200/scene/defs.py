"""Minimal mirror of the scene types the BA / GP boundary marshals.

Only what the hot path touches: ``CameraModelId``,
``get_camera_model_info`` (/root/reference/instantsfm/scene/defs.py:101-140),
``Camera.params / set_params / model_id / has_prior_focal_length`` (:142-237),
``Image.world2cam / features / features_undist / depths / is_registered / cam_id`` (:8-33)
and ``Track.xyz / observations`` (:414-423).  The drop-in processors are duck-typed: the
reference's own classes work unchanged; these exist so tests and benches can build scenes
without the reference tree (which is absent on the GPU box).
"""
from enum import Enum

import numpy as np


class CameraModelId(Enum):
    INVALID = -1
    SIMPLE_PINHOLE = 0
    PINHOLE = 1
    SIMPLE_RADIAL = 2
    RADIAL = 3
    OPENCV = 4
    OPENCV_FISHEYE = 5
    FULL_OPENCV = 6
    FOV = 7
    SIMPLE_RADIAL_FISHEYE = 8
    RADIAL_FISHEYE = 9
    THIN_PRISM_FISHEYE = 10


# model -> (num_params, focal idx, pp idx, optimise idx); pp is never optimised.
_MODEL_TABLE = {
    CameraModelId.SIMPLE_PINHOLE: (3, [0], [1, 2], [0]),
    CameraModelId.PINHOLE: (4, [0, 1], [2, 3], [0, 1]),
    CameraModelId.SIMPLE_RADIAL: (4, [0], [1, 2], [0, 3]),
    CameraModelId.RADIAL: (5, [0], [1, 2], [0, 3, 4]),
    CameraModelId.OPENCV: (8, [0, 1], [2, 3], [0, 1, 4, 5, 6, 7]),
    CameraModelId.OPENCV_FISHEYE: (8, [0, 1], [2, 3], [0, 1, 4, 5, 6, 7]),
    CameraModelId.FULL_OPENCV: (12, [0, 1], [2, 3], [0, 1, 4, 5, 6, 7, 8, 9, 10, 11]),
    CameraModelId.FOV: (5, [0, 1], [2, 3], [0, 1, 4]),
    CameraModelId.SIMPLE_RADIAL_FISHEYE: (4, [0], [1, 2], [0, 3]),
    CameraModelId.RADIAL_FISHEYE: (5, [0], [1, 2], [0, 3, 4]),
    CameraModelId.THIN_PRISM_FISHEYE: (12, [0, 1], [2, 3], [0, 1, 4, 5, 6, 7, 8, 9, 10, 11]),
}


def get_camera_model_info(model_id):
    """Subset of the reference dict the BA path reads: name, num_params, focal, pp, optimize."""
    if model_id not in _MODEL_TABLE:
        raise NotImplementedError
    n, focal, pp, opt = _MODEL_TABLE[model_id]
    return {"name": model_id.name, "num_params": n, "focal": list(focal), "pp": list(pp),
            "optimize": list(opt)}


class Camera:
    def __init__(self, id=-1, model_id=CameraModelId.INVALID, width=0, height=0, params=None,
                 has_prior_focal_length=False):
        self.id, self.model_id, self.width, self.height = id, model_id, width, height
        self.has_prior_focal_length = has_prior_focal_length
        self.focal_length = np.zeros(2)
        self.principal_point = np.zeros(2)
        self.params = []
        if params is not None:
            self.set_params(params)

    def set_params(self, params):
        info = get_camera_model_info(self.model_id)
        assert len(params) == info["num_params"]
        self.params = params
        f = [params[i] for i in info["focal"]]
        self.focal_length = np.array([f[0], f[-1]])
        self.principal_point = np.array([params[i] for i in info["pp"]])


class Image:
    def __init__(self, id=-1, cam_id=-1, is_registered=False, world2cam=None, features=None,
                 depths=None, features_undist=None):
        self.id, self.cam_id, self.is_registered = id, cam_id, is_registered
        self.world2cam = np.eye(4) if world2cam is None else world2cam
        self.features = [] if features is None else features
        self.depths = [] if depths is None else depths
        self.features_undist = [] if features_undist is None else features_undist

    def center(self):
        """scene/defs.py:35-36."""
        return self.world2cam[:3, :3].T @ -self.world2cam[:3, 3]


class Track:
    def __init__(self, **kwargs):
        self.id = -1
        self.xyz = np.zeros(3)
        self.is_initialized = False
        self.observations = np.zeros((0, 2), dtype=np.int64)
        for k, v in kwargs.items():
            setattr(self, k, v)
