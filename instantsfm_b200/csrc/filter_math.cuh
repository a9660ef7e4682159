// filter_math.cuh -- per-observation tests of the inter-BA track filters (filters.cu), host+device so
// that tests/hostcheck can run the kernels' exact arithmetic on the CPU against the golden vectors the
// reference's own track_filter.py produced (test-only; the product path never runs it on the host).
// fp64 with the reference's order of operations and WITHOUT fused multiply-adds: explicit
// round-to-nearest intrinsics on the device, plain operators on the host (build with
// -ffp-contract=off).
#pragma once
#include <cmath>

#include "math.cuh"

namespace isfm {

constexpr double FILTER_EPS = 1e-10;   // track_filter.py:3

#if defined(__CUDA_ARCH__)
#define ISFM_MUL_RN(a, b) __dmul_rn((a), (b))
#define ISFM_ADD_RN(a, b) __dadd_rn((a), (b))
#define ISFM_DIV_RN(a, b) __ddiv_rn((a), (b))
#else
#define ISFM_MUL_RN(a, b) ((a) * (b))
#define ISFM_ADD_RN(a, b) ((a) + (b))
#define ISFM_DIV_RN(a, b) ((a) / (b))
#endif

ISFM_HD double dot3_nofma(double a0, double a1, double a2, double b0, double b1, double b2) {
  return ISFM_ADD_RN(ISFM_ADD_RN(ISFM_MUL_RN(a0, b0), ISFM_MUL_RN(a1, b1)), ISFM_MUL_RN(a2, b2));
}

// MODE 0: FilterTracksByAngle (track_filter.py:5-24)
//   pt = R X + t; reject if pt.z < EPS; pt /= ||pt||; keep iff dot(pt, f) > cos(max_angle)
// MODE 1: FilterTracksByReprojectionNormalized (track_filter.py:26-66)
//   pt = [R|t] [X;1]; valid = pt.z > EPS; e = || pt.xy/(pt.z+EPS) - f.xy/(f.z+EPS) ||; keep iff valid && e < thr
// M: the image's 4x4 world2cam, row-major.
template <int MODE>
ISFM_HD bool filter_keep(const double* __restrict__ M, double x, double y, double z, double f0, double f1, double f2, double thr) {
  double p[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    // (R X) first, then + t: `R @ xyz + t` (:12) and einsum over [x, y, z, 1] (:50) agree on this order
    p[r] = ISFM_ADD_RN(dot3_nofma(M[4 * r], M[4 * r + 1], M[4 * r + 2], x, y, z), M[4 * r + 3]);
  }
  bool keep;
  if (MODE == 0) {
    if (p[2] < FILTER_EPS) {
      keep = false;
    } else {
      const double n = sqrt(dot3_nofma(p[0], p[1], p[2], p[0], p[1], p[2]));   // np.linalg.norm
      keep = dot3_nofma(ISFM_DIV_RN(p[0], n), ISFM_DIV_RN(p[1], n), ISFM_DIV_RN(p[2], n), f0, f1, f2) > thr;
    }
  } else {
    const double pz = ISFM_ADD_RN(p[2], FILTER_EPS), fz = ISFM_ADD_RN(f2, FILTER_EPS);
    const double d0 = ISFM_ADD_RN(ISFM_DIV_RN(p[0], pz), -ISFM_DIV_RN(f0, fz));
    const double d1 = ISFM_ADD_RN(ISFM_DIV_RN(p[1], pz), -ISFM_DIV_RN(f1, fz));
    const double e = sqrt(ISFM_ADD_RN(ISFM_MUL_RN(d0, d0), ISFM_MUL_RN(d1, d1)));
    keep = (p[2] > FILTER_EPS) && (e < thr);
  }
  return keep;
}

}  // namespace isfm
