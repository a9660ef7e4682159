// camera_maps.cuh -- host+device scalar forms of the scene layer's pixel-space camera maps
// (Camera.cam2img / Camera.img2cam, instantsfm/scene/defs.py:244-412) used by camera_ops.cu.
// __host__ __device__ like math.cuh, so that tests/hostcheck can run the exact arithmetic of the
// kernels on the CPU against the reference-generated golden vectors (test-only; the product path
// never runs it on the host).  Compile WITHOUT FMA contraction (nvcc --fmad=false, g++
// -ffp-contract=off): the reference computes in numpy, one rounded ufunc per operation.
#pragma once
#include <cmath>

#include "math.cuh"

namespace isfm {

constexpr int CAMROW = 16;   // doubles per camera-table row (ISFM_CAMERA_ROW, include/isfm_b200.h)
// row layout: [0] model id, [1] fx, [2] fy, [3] cx, [4] cy, [5..10] k[0..5], [11] p0, [12] p1,
//             [13] omega, [14] sx0, [15] sx1  -- the attributes Camera.set_params derives (defs.py:177-237)

struct CamRow {
  int model;
  double fx, fy, cx, cy, k[6], p0, p1, omega, sx0, sx1;
};

ISFM_HD CamRow load_row(const double* __restrict__ t) {
  CamRow c;
  c.model = (int)t[0]; c.fx = t[1]; c.fy = t[2]; c.cx = t[3]; c.cy = t[4];
#pragma unroll
  for (int i = 0; i < 6; ++i) c.k[i] = t[5 + i];
  c.p0 = t[11]; c.p1 = t[12]; c.omega = t[13]; c.sx0 = t[14]; c.sx1 = t[15];
  return c;
}

ISFM_HD double sq(double x) { return x * x; }

// Camera.fisheye_from_normal (defs.py:244-248)
ISFM_HD void fisheye_from_normal(double& u, double& v) {
  double r = sqrt(u * u + v * v);
  r = fmax(r, 1e-8);
  const double theta = atan(r);
  u = u * theta / r; v = v * theta / r;
}

// Camera.Distortion (defs.py:255-313): returns the additive term d (FOV: the multiplied point itself)
ISFM_HD void distortion(const CamRow& c, double u, double v, double& du, double& dv) {
  const double r2 = u * u + v * v;
  switch (c.model) {
    case 2: case 8: du = u * c.k[0] * r2; dv = v * c.k[0] * r2; break;
    case 3: case 9: {
      const double r4 = r2 * r2;
      du = u * c.k[0] * r2 + u * c.k[1] * r4; dv = v * c.k[0] * r2 + v * c.k[1] * r4; break;
    }
    case 4: case 6: case 10: {
      const double uv = u * v;
      double radial;
      if (c.model == 4) radial = c.k[0] * r2 + c.k[1] * (r2 * r2);
      else if (c.model == 10) radial = c.k[0] * r2 + c.k[1] * (r2 * r2) + c.k[2] * pow(r2, 3.0);
      else radial = (1 + c.k[0] * r2 + c.k[1] * (r2 * r2) + c.k[2] * pow(r2, 3.0)) /
                    (1 + c.k[3] * r2 + c.k[4] * (r2 * r2) + c.k[5] * pow(r2, 3.0)) - 1;
      du = u * radial + 2 * c.p0 * uv; dv = v * radial + 2 * c.p1 * uv;
      du += c.p1 * (r2 + 2 * sq(u)); dv += c.p0 * (r2 + 2 * sq(v));     // self.p[::-1] * (r2 + 2 uv^2)
      if (c.model == 10) { du += c.sx0 * r2; dv += c.sx1 * r2; }
      break;
    }
    case 5: {
      const double radial = c.k[0] * r2 + c.k[1] * (r2 * r2) + c.k[2] * pow(r2, 3.0);   // k3 ignored (:277)
      du = u * radial; dv = v * radial; break;
    }
    case 7: {
      const double omega = c.omega, omega2 = omega * omega, eps = 1e-4;
      double factor;
      if (omega2 < eps) {
        factor = (omega2 * r2) / 3 - omega2 / 12 + 1;
      } else if (r2 < eps) {
        const double th = tan(omega / 2);
        factor = (-2 * th * (4 * r2 * (th * th) - 3)) / (3 * omega);
      } else {
        const double radius = sqrt(r2);
        factor = atan(radius * 2 * tan(omega / 2)) / (radius * omega);
      }
      du = u * factor; dv = v * factor; break;
    }
    default: du = 0; dv = 0; break;
  }
}

// Camera.cam2img (defs.py:371-412)
ISFM_HD void cam2img(const CamRow& c, double X, double Y, double Z, double& px, double& py) {
  const double zz = Z + 1e-10;
  double u = X / zz, v = Y / zz, du, dv;
  const double f = (c.fx + c.fy) / 2.0;   // np.mean(self.focal_length)
  const bool fisheye = c.model == 5 || c.model == 8 || c.model == 9 || c.model == 10;
  if (fisheye) fisheye_from_normal(u, v);
  if (c.model == 7) {
    distortion(c, u, v, du, dv); u = du; v = dv;
  } else if (c.model >= 2) {
    distortion(c, u, v, du, dv); u += du; v += dv;
  }
  const bool two_focals = c.model == 1 || c.model == 4 || c.model == 5 || c.model == 6 || c.model == 10;
  px = u * (two_focals ? c.fx : f) + c.cx;
  py = v * (two_focals ? c.fy : f) + c.cy;
}

// cv2.undistortPoints(xy, K, dist) with dist = (k1, k2, p1, p2, k3, k4, k5, k6, s1, s2, s3, s4)
ISFM_HD void undistort_opencv(const CamRow& c, const double* __restrict__ kk, double px, double py,
                                                 double& x, double& y) {
  const double ifx = 1.0 / c.fx, ify = 1.0 / c.fy;
  const double x0 = (px - c.cx) * ifx, y0 = (py - c.cy) * ify;
  x = x0; y = y0;
  for (int j = 0; j < 5; ++j) {
    const double r2 = x * x + y * y;
    const double icdist = (1 + ((kk[7] * r2 + kk[6]) * r2 + kk[5]) * r2) / (1 + ((kk[4] * r2 + kk[1]) * r2 + kk[0]) * r2);
    if (icdist < 0) { x = x0; y = y0; break; }
    const double dX = 2 * kk[2] * x * y + kk[3] * (r2 + 2 * x * x) + kk[8] * r2 + kk[9] * r2 * r2;
    const double dY = kk[2] * (r2 + 2 * y * y) + 2 * kk[3] * x * y + kk[10] * r2 + kk[11] * r2 * r2;
    x = (x0 - dX) * icdist;
    y = (y0 - dY) * icdist;
  }
}

// Camera.normal_from_fisheye (defs.py:250-253); theta == 0 gives 0/0 = NaN like the reference
ISFM_HD void normal_from_fisheye(double& u, double& v) {
  const double theta = sqrt(u * u + v * v);
  const double tc = theta * cos(theta), s = sin(theta);
  u = u * s / tc; v = v * s / tc;
}

// Camera.img2cam (defs.py:315-369)
ISFM_HD void img2cam(const CamRow& c, double px, double py, double& u, double& v) {
  double kk[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  switch (c.model) {
    case 0: { const double f = (c.fx + c.fy) / 2.0; u = (px - c.cx) / f; v = (py - c.cy) / f; return; }
    case 1: u = (px - c.cx) / c.fx; v = (py - c.cy) / c.fy; return;
    case 2: case 8: kk[0] = c.k[0]; break;
    case 3: case 9: kk[0] = c.k[0]; kk[1] = c.k[1]; break;
    case 4: kk[0] = c.k[0]; kk[1] = c.k[1]; kk[2] = c.p0; kk[3] = c.p1; break;
    case 5: kk[0] = c.k[0]; kk[1] = c.k[1]; kk[4] = c.k[2]; break;                       // (k0, k1, 0, 0, k2)
    case 6: kk[0] = c.k[0]; kk[1] = c.k[1]; kk[2] = c.p0; kk[3] = c.p1; kk[4] = c.k[2]; kk[5] = c.k[3]; kk[6] = c.k[4]; kk[7] = c.k[5]; break;
    case 10: kk[0] = c.k[0]; kk[1] = c.k[1]; kk[2] = c.p0; kk[3] = c.p1; kk[4] = c.k[2]; kk[8] = c.sx0; kk[9] = c.sx1; break;
    case 7: {
      // r2 is taken from the RAW pixel coordinates (defs.py:344), as the reference does
      const double omega = c.omega, omega2 = omega * omega, eps = 1e-4, r2 = px * px + py * py;
      double factor;
      if (omega2 < eps) {
        factor = (omega2 * r2) / 3 - omega2 / 12 + 1;
      } else if (r2 < eps) {
        factor = (omega * (omega2 * r2 + 3)) / (6 * tan(omega / 2));
      } else {
        const double radius = sqrt(r2);
        factor = tan(radius * omega) / (radius * 2 * tan(omega / 2));
      }
      u = (px - c.cx) / c.fx * factor; v = (py - c.cy) / c.fy * factor; return;
    }
    default: u = v = NAN; return;
  }
  undistort_opencv(c, kk, px, py, u, v);
  if (c.model == 5 || c.model == 8 || c.model == 9 || c.model == 10) normal_from_fisheye(u, v);
}

}  // namespace isfm
