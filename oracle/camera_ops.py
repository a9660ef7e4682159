"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the scene layer's pixel-space camera maps:

* ``cam2img``  -- Camera.cam2img, instantsfm/scene/defs.py:371-412 (with Distortion :255-313 and
  fisheye_from_normal :244-248),
* ``img2cam``  -- Camera.img2cam, scene/defs.py:315-369; the distorted models call
  ``cv2.undistortPoints`` there.  OpenCV (opencv-python, pyproject.toml; 4.13.0 in the build
  container) is a third-party dependency: its published algorithm (calib3d
  ``cvUndistortPointsInternal``: normalise with K, then FIVE fixed-point iterations -- the default
  criteria of the 6-argument overload is TermCriteria(MAX_ITER, 5, 0.01)) is restated in
  ``undistort_points_opencv`` below,
* ``filter_reprojection_mask`` -- FilterTracksByReprojection, processors/track_filter.py:68-114,
* ``undistort_images`` -- UndistortImages, processors/image_undistortion.py:3-9.

Scalar per-point loops in fp64 for small cases.  PINNED: tests/golden/reference_camera_ops.npz was
produced by the reference's own Camera class (cv2 4.13 underneath) and FilterTracksByReprojection
(tests/golden/make_camera_ops_golden.py); tests/test_camera_ops_host.py checks this module against it.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package.
"""
import math

import numpy as np

# model ids: scene/defs.py:101-113
SIMPLE_PINHOLE, PINHOLE, SIMPLE_RADIAL, RADIAL, OPENCV, OPENCV_FISHEYE, FULL_OPENCV, FOV, SIMPLE_RADIAL_FISHEYE, \
    RADIAL_FISHEYE, THIN_PRISM_FISHEYE = range(11)
FISHEYE = (OPENCV_FISHEYE, SIMPLE_RADIAL_FISHEYE, RADIAL_FISHEYE, THIN_PRISM_FISHEYE)
TWO_FOCALS = (PINHOLE, OPENCV, OPENCV_FISHEYE, FULL_OPENCV, THIN_PRISM_FISHEYE)   # cam2img multiplies by `ff` (:380-410)


class Intrinsics:
    """What Camera.set_params derives from ``params`` (scene/defs.py:177-237)."""

    def __init__(self, model, params):
        p = [float(x) for x in params]
        self.model = int(model)
        self.k, self.p, self.omega, self.sx = [0.0] * 6, [0.0, 0.0], 0.0, [0.0, 0.0]
        m = self.model
        if m in (SIMPLE_PINHOLE, SIMPLE_RADIAL, RADIAL, SIMPLE_RADIAL_FISHEYE, RADIAL_FISHEYE):
            self.f, self.c = [p[0], p[0]], [p[1], p[2]]
            for i, v in enumerate(p[3:]):
                self.k[i] = v
        else:
            self.f, self.c = [p[0], p[1]], [p[2], p[3]]
            if m == OPENCV:
                self.k[:2], self.p = p[4:6], p[6:8]
            elif m == OPENCV_FISHEYE:
                self.k[:4] = p[4:8]
            elif m == FULL_OPENCV:
                self.k, self.p = [p[4], p[5], p[8], p[9], p[10], p[11]], p[6:8]
            elif m == FOV:
                self.omega = p[4]
            elif m == THIN_PRISM_FISHEYE:
                self.k[:4], self.p, self.sx = [p[4], p[5], p[8], p[9]], p[6:8], p[10:12]
            elif m != PINHOLE:
                raise NotImplementedError

    def row(self):
        """The camera-table row of include/isfm_b200.h (ISFM_CAMERA_ROW doubles)."""
        return np.array([self.model, self.f[0], self.f[1], self.c[0], self.c[1], *self.k, *self.p, self.omega, *self.sx], dtype=np.float64)


def _distortion(c, u, v):
    """Camera.Distortion (:255-313) for one point."""
    r2 = u * u + v * v
    m = c.model
    if m in (SIMPLE_RADIAL, SIMPLE_RADIAL_FISHEYE):
        return u * c.k[0] * r2, v * c.k[0] * r2
    if m in (RADIAL, RADIAL_FISHEYE):
        return u * c.k[0] * r2 + u * c.k[1] * r2 ** 2, v * c.k[0] * r2 + v * c.k[1] * r2 ** 2
    if m in (OPENCV, FULL_OPENCV, THIN_PRISM_FISHEYE):
        uv = u * v
        if m == OPENCV:
            radial = c.k[0] * r2 + c.k[1] * r2 ** 2
        elif m == THIN_PRISM_FISHEYE:
            radial = c.k[0] * r2 + c.k[1] * r2 ** 2 + c.k[2] * r2 ** 3
        else:
            radial = (1 + c.k[0] * r2 + c.k[1] * r2 ** 2 + c.k[2] * r2 ** 3) / (1 + c.k[3] * r2 + c.k[4] * r2 ** 2 + c.k[5] * r2 ** 3) - 1
        du, dv = u * radial + 2 * c.p[0] * uv, v * radial + 2 * c.p[1] * uv
        du, dv = du + c.p[1] * (r2 + 2 * u * u), dv + c.p[0] * (r2 + 2 * v * v)
        if m == THIN_PRISM_FISHEYE:
            du, dv = du + c.sx[0] * r2, dv + c.sx[1] * r2
        return du, dv
    if m == OPENCV_FISHEYE:
        radial = c.k[0] * r2 + c.k[1] * r2 ** 2 + c.k[2] * r2 ** 3
        return u * radial, v * radial
    if m == FOV:
        omega, eps = c.omega, 1e-4
        omega2 = omega * omega
        if omega2 < eps:
            factor = (omega2 * r2) / 3 - omega2 / 12 + 1
        elif r2 < eps:
            th = math.tan(omega / 2)
            factor = (-2 * th * (4 * r2 * th ** 2 - 3)) / (3 * omega)
        else:
            radius = math.sqrt(r2)
            factor = math.atan(radius * 2 * math.tan(omega / 2)) / (radius * omega)
        return u * factor, v * factor
    raise NotImplementedError


def cam2img(c, uvw):
    """Camera.cam2img (:371-412).  uvw [n, 3] -> pixels [n, 2]."""
    uvw = np.asarray(uvw, dtype=np.float64).reshape(-1, 3)
    out = np.zeros((uvw.shape[0], 2))
    f = (c.f[0] + c.f[1]) / 2.0
    for i, (X, Y, Z) in enumerate(uvw):
        with np.errstate(all="ignore"):
            zz = np.float64(Z) + 1e-10
            u, v = float(np.float64(X) / zz), float(np.float64(Y) / zz)
        if not (math.isfinite(u) and math.isfinite(v)):
            out[i] = np.nan
            continue
        if c.model in FISHEYE:
            r = max(math.sqrt(u * u + v * v), 1e-8)
            theta = math.atan(r)
            u, v = u * theta / r, v * theta / r
        if c.model == FOV:
            u, v = _distortion(c, u, v)
        elif c.model >= SIMPLE_RADIAL:
            du, dv = _distortion(c, u, v)
            u, v = u + du, v + dv
        fx, fy = (c.f[0], c.f[1]) if c.model in TWO_FOCALS else (f, f)
        out[i] = (u * fx + c.c[0], v * fy + c.c[1])
    return out


def undistort_points_opencv(c, kk, px, py):
    """cv2.undistortPoints(xy, K, dist) for one pixel; kk = (k1, k2, p1, p2, k3, k4, k5, k6, s1, s2, s3, s4)."""
    ifx, ify = 1.0 / c.f[0], 1.0 / c.f[1]
    x0, y0 = (px - c.c[0]) * ifx, (py - c.c[1]) * ify
    x, y = x0, y0
    for _ in range(5):
        r2 = x * x + y * y
        icdist = (1 + ((kk[7] * r2 + kk[6]) * r2 + kk[5]) * r2) / (1 + ((kk[4] * r2 + kk[1]) * r2 + kk[0]) * r2)
        if icdist < 0:
            return x0, y0
        dx = 2 * kk[2] * x * y + kk[3] * (r2 + 2 * x * x) + kk[8] * r2 + kk[9] * r2 * r2
        dy = kk[2] * (r2 + 2 * y * y) + 2 * kk[3] * x * y + kk[10] * r2 + kk[11] * r2 * r2
        x, y = (x0 - dx) * icdist, (y0 - dy) * icdist
    return x, y


def img2cam(c, xy):
    """Camera.img2cam (:315-369).  xy [n, 2] pixels -> normalised coordinates [n, 2]."""
    xy = np.asarray(xy, dtype=np.float64).reshape(-1, 2)
    out = np.zeros_like(xy)
    m = c.model
    kk = [0.0] * 12
    if m in (SIMPLE_RADIAL, SIMPLE_RADIAL_FISHEYE):
        kk[0] = c.k[0]
    elif m in (RADIAL, RADIAL_FISHEYE):
        kk[0], kk[1] = c.k[0], c.k[1]
    elif m == OPENCV:
        kk[:4] = [c.k[0], c.k[1], c.p[0], c.p[1]]
    elif m == OPENCV_FISHEYE:
        kk[:5] = [c.k[0], c.k[1], 0.0, 0.0, c.k[2]]
    elif m == FULL_OPENCV:
        kk[:8] = [c.k[0], c.k[1], c.p[0], c.p[1], c.k[2], c.k[3], c.k[4], c.k[5]]
    elif m == THIN_PRISM_FISHEYE:
        kk[:5] = [c.k[0], c.k[1], c.p[0], c.p[1], c.k[2]]
        kk[8], kk[9] = c.sx[0], c.sx[1]
    for i, (px, py) in enumerate(xy):
        px, py = float(px), float(py)
        if m == SIMPLE_PINHOLE:
            f = (c.f[0] + c.f[1]) / 2.0
            u, v = (px - c.c[0]) / f, (py - c.c[1]) / f
        elif m == PINHOLE:
            u, v = (px - c.c[0]) / c.f[0], (py - c.c[1]) / c.f[1]
        elif m == FOV:
            omega, eps = c.omega, 1e-4
            omega2, r2 = omega * omega, px * px + py * py      # r2 from the raw pixels (:344)
            if omega2 < eps:
                factor = (omega2 * r2) / 3 - omega2 / 12 + 1
            elif r2 < eps:
                factor = (omega * (omega2 * r2 + 3)) / (6 * math.tan(omega / 2))
            else:
                radius = math.sqrt(r2)
                factor = math.tan(radius * omega) / (radius * 2 * math.tan(omega / 2))
            u, v = (px - c.c[0]) / c.f[0] * factor, (py - c.c[1]) / c.f[1] * factor
        else:
            u, v = undistort_points_opencv(c, kk, px, py)
            if m in FISHEYE:
                theta = math.sqrt(u * u + v * v)
                tc = theta * math.cos(theta)
                s = math.sin(theta)
                u, v = (u * s / tc, v * s / tc) if tc != 0.0 else (math.nan, math.nan)
        out[i] = (u, v)
    return out


def undistort_images(cameras, images):
    """UndistortImages (image_undistortion.py:3-9) -> list of [n_i, 3] unit bearings, one per image."""
    out = []
    for image in images:
        cam = cameras[image.cam_id]
        c = Intrinsics(cam.model_id.value, cam.params)
        uv = img2cam(c, np.asarray(image.features, dtype=np.float64).reshape(-1, 2))
        b = np.hstack([uv, np.ones((uv.shape[0], 1))])
        out.append(b / np.linalg.norm(b, axis=1, keepdims=True))
    return out


def filter_reprojection_mask(cameras, images, tracks, max_reprojection_error):
    """FilterTracksByReprojection (track_filter.py:68-107) -> {track_id: boolean mask} and the errors."""
    EPSILON = 1e-10
    intr = [Intrinsics(cam.model_id.value, cam.params) for cam in cameras]
    masks, errors = {}, {}
    for tid, track in tracks.items():
        obs = np.asarray(track.observations).reshape(-1, 2)
        m, e = np.zeros(obs.shape[0], dtype=bool), np.zeros(obs.shape[0])
        xyz1 = np.append(np.asarray(track.xyz, dtype=np.float64), 1.0)
        for idx, (image_id, feature_id) in enumerate(obs):
            image = images[image_id]
            M = np.asarray(image.world2cam, dtype=np.float64)
            p = np.array([((M[r, 0] * xyz1[0] + M[r, 1] * xyz1[1]) + M[r, 2] * xyz1[2]) + M[r, 3] for r in range(3)])
            with np.errstate(all="ignore"):
                px = cam2img(intr[image.cam_id], p[None])[0]
                e[idx] = np.linalg.norm(px - np.asarray(image.features[feature_id], dtype=np.float64))
            m[idx] = (p[2] > EPSILON) and (e[idx] < max_reprojection_error)
        masks[tid], errors[tid] = m, e
    return masks, errors


def apply_filter_reprojection(cameras, images, tracks, max_reprojection_error):
    """The whole function incl. in-place mutation and the returned counter (:109-114: the counter
    tests the window of the mask that FOLLOWS each track -- ``count`` is advanced first)."""
    masks, _ = filter_reprojection_mask(cameras, images, tracks, max_reprojection_error)
    flat = np.concatenate([masks[t] for t in masks]) if masks else np.zeros(0, bool)
    count = counter = 0
    for tid, track in tracks.items():
        n = len(masks[tid])
        track.observations = np.asarray(track.observations).reshape(-1, 2)[masks[tid]]
        count += n
        if not np.all(flat[count:count + n]):
            counter += 1
    return counter
