"""Drop-in for instantsfm/processors/image_undistortion.py:3-9: same function
names, arguments and in-place effect (``image.features_undist`` of every image is replaced by the
[n, 3] unit bearings of its features).  The reference calls ``Camera.img2cam`` per image (numpy /
``cv2.undistortPoints``); here the features of ALL images go through one CUDA kernel launch
(include/isfm_b200.h: isfm_undistort_features, csrc/camera_ops.cu), fp64, all eleven camera models.
There is no CPU path: without the CUDA library / a GPU these functions raise."""
import numpy as np

from .. import _lib
from ._common import camera_table


def _undistort_batch(cameras, images):
    counts = np.array([len(img.features) for img in images], dtype=np.int64)
    total = int(counts.sum())
    out = np.zeros((total, 3), dtype=np.float64)
    if total:
        feats = np.ascontiguousarray(np.concatenate(
            [np.asarray(img.features, dtype=np.float64).reshape(-1, 2) for img, n in zip(images, counts) if n], axis=0))
        cam_idx = np.ascontiguousarray(np.repeat(np.array([img.cam_id for img in images], dtype=np.int32), counts))
        cams = camera_table(cameras)
        _lib.check(_lib.load().isfm_undistort_features(total, len(cameras), cams.ctypes.data, feats.ctypes.data,
                                                       cam_idx.ctypes.data, out.ctypes.data, None))
    return np.split(out, np.cumsum(counts)[:-1]) if len(images) else []


def undistort_process(image, cam):
    """image_undistortion.py:3-6 (one image)."""
    cameras = [cam] * (image.cam_id + 1)   # the kernel indexes the table with image.cam_id
    image.features_undist = _undistort_batch(cameras, [image])[0]


def UndistortImages(cameras, images):
    """image_undistortion.py:7-9."""
    for image, bearings in zip(images, _undistort_batch(cameras, images)):
        image.features_undist = bearings
