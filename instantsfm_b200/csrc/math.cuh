// math.cuh -- host+device scalar math shared by every kernel of the BA / GP path.
// Everything here is __host__ __device__ so that tests/hostcheck can exercise the exact
// same arithmetic on the CPU (test-only; the product path never runs it on the host).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define ISFM_HD __host__ __device__ __forceinline__
#else
#define ISFM_HD inline
#endif

namespace isfm {

// --------------------------------------------------------------------------------------
// forward-mode dual numbers: exact derivatives of the distortion polynomials
// (utils/cost_function.py:32-177) without nine hand-derived Jacobians.
// --------------------------------------------------------------------------------------
template <typename T, int N>
struct Dual {
  T v;
  T d[N];
  ISFM_HD Dual() {}
  ISFM_HD Dual(T c) : v(c) {
#pragma unroll
    for (int i = 0; i < N; ++i) d[i] = T(0);
  }
  ISFM_HD static Dual seed(T c, int k) {
    Dual r(c);
    r.d[k] = T(1);
    return r;
  }
};
template <typename T, int N> ISFM_HD Dual<T, N> operator+(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r; r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <typename T, int N> ISFM_HD Dual<T, N> operator-(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r; r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <typename T, int N> ISFM_HD Dual<T, N> operator*(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r; r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <typename T, int N> ISFM_HD Dual<T, N> operator/(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r; T inv = T(1) / b.v; r.v = a.v * inv;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
template <typename T, int N> ISFM_HD Dual<T, N> operator+(const Dual<T, N>& a, T b) { Dual<T, N> r = a; r.v += b; return r; }
template <typename T, int N> ISFM_HD Dual<T, N> operator+(T b, const Dual<T, N>& a) { Dual<T, N> r = a; r.v += b; return r; }
template <typename T, int N> ISFM_HD Dual<T, N> operator-(const Dual<T, N>& a, T b) { Dual<T, N> r = a; r.v -= b; return r; }
template <typename T, int N> ISFM_HD Dual<T, N> operator*(const Dual<T, N>& a, T b) {
  Dual<T, N> r; r.v = a.v * b;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b;
  return r;
}
template <typename T, int N> ISFM_HD Dual<T, N> operator*(T b, const Dual<T, N>& a) { return a * b; }
template <typename T, int N> ISFM_HD Dual<T, N> dsqrt(const Dual<T, N>& a) {
  Dual<T, N> r; r.v = sqrt(a.v); T k = T(0.5) / r.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * k;
  return r;
}
template <typename T, int N> ISFM_HD Dual<T, N> datan(const Dual<T, N>& a) {
  Dual<T, N> r; r.v = atan(a.v); T k = T(1) / (T(1) + a.v * a.v);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * k;
  return r;
}
ISFM_HD float dsqrt(float a) { return sqrtf(a); }
ISFM_HD double dsqrt(double a) { return sqrt(a); }
ISFM_HD float datan(float a) { return atanf(a); }
ISFM_HD double datan(double a) { return atan(a); }
template <typename T, int N> ISFM_HD T value_of(const Dual<T, N>& a) { return a.v; }
ISFM_HD float value_of(float a) { return a; }
ISFM_HD double value_of(double a) { return a; }

// --------------------------------------------------------------------------------------
// camera models.  MODEL = CameraModelId.value (scene/defs.py:101-113).  Intrinsics k[] are
// the camera row with the principal point removed (bundle_adjustment.py:75-80), i.e. in
// get_camera_model_info(...)['optimize'] order.
// --------------------------------------------------------------------------------------
template <int MODEL> struct ModelTraits;
template <> struct ModelTraits<0> { static constexpr int NI = 1; };   // SIMPLE_PINHOLE  f
template <> struct ModelTraits<1> { static constexpr int NI = 2; };   // PINHOLE         fx fy
template <> struct ModelTraits<2> { static constexpr int NI = 2; };   // SIMPLE_RADIAL   f k
template <> struct ModelTraits<3> { static constexpr int NI = 3; };   // RADIAL          f k1 k2
template <> struct ModelTraits<4> { static constexpr int NI = 6; };   // OPENCV          fx fy k1 k2 p1 p2
template <> struct ModelTraits<5> { static constexpr int NI = 6; };   // OPENCV_FISHEYE  fx fy k1 k2 k3 k4(unused)
template <> struct ModelTraits<6> { static constexpr int NI = 10; };  // FULL_OPENCV     fx fy k1 k2 p1 p2 k3 k4 k5 k6
template <> struct ModelTraits<8> { static constexpr int NI = 2; };   // SIMPLE_RADIAL_FISHEYE f k
template <> struct ModelTraits<9> { static constexpr int NI = 3; };   // RADIAL_FISHEYE  f k1 k2

inline int model_n_intr(int model_id) {
  switch (model_id) {
    case 0: return 1; case 1: return 2; case 2: return 2; case 3: return 3; case 4: return 6;
    case 5: return 6; case 6: return 10; case 8: return 2; case 9: return 3; default: return -1;
  }
}

// u * atan(r) / r with r^2 computed from the UNDISTORTED u (cost_function.py:95-98).  The
// reference evaluates 0/0 at r == 0; we take the limit there (documented deviation).
template <typename S> ISFM_HD void fisheye_scale(const S& r2, S& fac) {
  if (value_of(r2) < 1e-16) {
    fac = S(1.0f) - r2 * decltype(value_of(r2))(1.0 / 3.0);
  } else {
    S r = dsqrt(r2);
    fac = datan(r) / r;
  }
}

// out = distortion(u; k) * focal   (everything of reproject_* after the perspective divide
// except "+ pp").
template <int MODEL, typename S>
ISFM_HD void distort(const S& u0, const S& u1, const S* k, S& o0, S& o1) {
  S r2 = u0 * u0 + u1 * u1;
  if (MODEL == 0) {            // cost_function.py:33-38
    o0 = u0 * k[0]; o1 = u1 * k[0];
  } else if (MODEL == 1) {     // :41-46
    o0 = u0 * k[0]; o1 = u1 * k[1];
  } else if (MODEL == 2) {     // :49-56
    S g = (k[1] * r2 + decltype(value_of(r2))(1)) * k[0];
    o0 = u0 * g; o1 = u1 * g;
  } else if (MODEL == 3) {     // :59-67
    S g = (k[1] * r2 + k[2] * r2 * r2 + decltype(value_of(r2))(1)) * k[0];
    o0 = u0 * g; o1 = u1 * g;
  } else if (MODEL == 4 || MODEL == 6) {   // :70-84, :105-123
    using T = decltype(value_of(r2));
    S radial;
    const S* p;
    if (MODEL == 4) {
      radial = k[2] * r2 + k[3] * r2 * r2;
      p = k + 4;
    } else {
      S r4 = r2 * r2, r6 = r4 * r2;
      radial = (k[2] * r2 + k[3] * r4 + k[6] * r6 + T(1)) / (k[7] * r2 + k[8] * r4 + k[9] * r6 + T(1)) - T(1);
      p = k + 4;
    }
    S uv = u0 * u1;
    S d0 = u0 * radial + p[0] * uv * T(2) + p[1] * (r2 + u0 * u0 * T(2));
    S d1 = u1 * radial + p[1] * uv * T(2) + p[0] * (r2 + u1 * u1 * T(2));
    o0 = (u0 + d0) * k[0]; o1 = (u1 + d1) * k[1];
  } else if (MODEL == 5) {     // :87-102 (k4 ignored)
    using T = decltype(value_of(r2));
    S fac; fisheye_scale(r2, fac);
    S radial = k[2] * r2 + k[3] * r2 * r2 + k[4] * r2 * r2 * r2 + T(1);
    o0 = u0 * fac * radial * k[0]; o1 = u1 * fac * radial * k[1];
  } else if (MODEL == 8) {     // :153-163
    using T = decltype(value_of(r2));
    S fac; fisheye_scale(r2, fac);
    S g = fac * (k[1] * r2 + T(1)) * k[0];
    o0 = u0 * g; o1 = u1 * g;
  } else if (MODEL == 9) {     // :166-177
    using T = decltype(value_of(r2));
    S fac; fisheye_scale(r2, fac);
    S g = fac * (k[1] * r2 + k[2] * r2 * r2 + T(1)) * k[0];
    o0 = u0 * g; o1 = u1 * g;
  }
}

// --------------------------------------------------------------------------------------
// rigid motion
// --------------------------------------------------------------------------------------
template <typename T> ISFM_HD void quat_to_rot(const T* q, T R[9]) {
  T x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

// y = R(q) X + t : bae.utils.ba.rotate_quat (cost_function.py:34)
template <typename T> ISFM_HD void transform_point(const T* cam7, const T* X, T R[9], T y[3]) {
  quat_to_rot(cam7 + 3, R);
  y[0] = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + cam7[0];
  y[1] = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + cam7[1];
  y[2] = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + cam7[2];
}

// residual only: proj - obs  (ReprojNonBatched.forward, bundle_adjustment.py:59-64), in the
// arithmetic of S.
template <int MODEL, typename S>
ISFM_HD void ba_residual_in(const S* cam, const S* pp, const S* X, const S* obs, S r[2]) {
  S R[9], y[3];
  transform_point(cam, X, R, y);
  S iz = S(1) / y[2];
  S u0 = y[0] * iz, u1 = y[1] * iz, o0, o1;
  distort<MODEL, S>(u0, u1, cam + 7, o0, o1);
  r[0] = o0 + pp[0] - obs[0];
  r[1] = o1 + pp[1] - obs[1];
}

// The residual is ALWAYS evaluated in double, also in the fp32 build: proj ~ 1e3 px against
// residuals of ~0.5 px means that fp32 arithmetic leaves ~1e-4 RELATIVE noise on every residual,
// i.e. on the gradient J^T r -- which moves weakly constrained points by several 1e-4 of their
// norm and is the difference between meeting and missing the 1e-4 parity bar on poses / points.
// ~80 double operations per observation next to ~800 fp32 ones for the Jacobian blocks; the
// parameters themselves stay fp32 (the minimiser is sought on the fp32 grid, evaluated exactly).
template <int MODEL, typename T>
ISFM_HD void ba_residual(const T* cam, const T* pp, const T* X, const T* obs, T r[2]) {
  constexpr int CW = 7 + ModelTraits<MODEL>::NI;
  double c[CW], p2[2], x[3], o[2], rd[2];
#pragma unroll
  for (int i = 0; i < CW; ++i) c[i] = (double)cam[i];
  p2[0] = (double)pp[0]; p2[1] = (double)pp[1];
  x[0] = (double)X[0]; x[1] = (double)X[1]; x[2] = (double)X[2];
  o[0] = (double)obs[0]; o[1] = (double)obs[1];
  ba_residual_in<MODEL, double>(c, p2, x, o, rd);
  r[0] = (T)rd[0]; r[1] = (T)rd[1];
}

// residual + Jacobian blocks (unweighted), in the arithmetic of T.  Jc[2][D] with D = 6 + NI:
// columns [d tau (3), d phi (3), intrinsics] for the left perturbation X <- Exp(delta) X; Jp[2][3].
template <int MODEL, typename T>
ISFM_HD void ba_linearize_in(const T* cam, const T* pp, const T* X, const T* obs, T r[2],
                             T* Jc /* 2*D */, T Jp[6]) {
  constexpr int NI = ModelTraits<MODEL>::NI;
  constexpr int D = 6 + NI;
  constexpr int ND = 2 + NI;
  typedef Dual<T, ND> S;
  T R[9], y[3];
  transform_point(cam, X, R, y);
  T iz = T(1) / y[2];
  T u0 = y[0] * iz, u1 = y[1] * iz;
  S k[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) k[i] = S::seed(cam[7 + i], 2 + i);
  S o0, o1;
  distort<MODEL, S>(S::seed(u0, 0), S::seed(u1, 1), k, o0, o1);
  r[0] = o0.v + pp[0] - obs[0];
  r[1] = o1.v + pp[1] - obs[1];
  // Jy = dproj/du * du/dy,  du/dy = (1/z) [1 0 -u0; 0 1 -u1]
  T Jy[6];
  Jy[0] = o0.d[0] * iz; Jy[1] = o0.d[1] * iz; Jy[2] = -(o0.d[0] * u0 + o0.d[1] * u1) * iz;
  Jy[3] = o1.d[0] * iz; Jy[4] = o1.d[1] * iz; Jy[5] = -(o1.d[0] * u0 + o1.d[1] * u1) * iz;
#pragma unroll
  for (int row = 0; row < 2; ++row) {
    const T* j = Jy + 3 * row;
    T* c = Jc + D * row;
    c[0] = j[0]; c[1] = j[1]; c[2] = j[2];                       // d/d tau
    c[3] = j[2] * y[1] - j[1] * y[2];                            // d/d phi = -Jy [y]x
    c[4] = j[0] * y[2] - j[2] * y[0];
    c[5] = j[1] * y[0] - j[0] * y[1];
    const S& o = row == 0 ? o0 : o1;
#pragma unroll
    for (int i = 0; i < NI; ++i) c[6 + i] = o.d[2 + i];
    Jp[3 * row + 0] = j[0] * R[0] + j[1] * R[3] + j[2] * R[6];   // Jy R
    Jp[3 * row + 1] = j[0] * R[1] + j[1] * R[4] + j[2] * R[7];
    Jp[3 * row + 2] = j[0] * R[2] + j[1] * R[5] + j[2] * R[8];
  }
}

// The linearisation is ALWAYS evaluated in double, also in the fp32 build (the blocks are then
// stored as fp32).  Two measured reasons (C1, 12 LM steps against the fp64 oracle): (i) proj ~ 1e3
// px against residuals of ~0.5 px leaves ~1e-4 relative noise on every fp32 residual; (ii) the
// rotation columns -Jy [y]x are differences of products ~1e3 and carry ~1e-6 relative error in
// fp32.  Either one perturbs the gradient J^T r by an amount that the softest modes of the
// reduced system (lambda ~ 1e-4 .. 1e-6 of the diagonal) amplify into ~1e-4 relative errors of
// rotations and weakly constrained points -- the difference between missing and meeting the 1e-4
// parity bar.  ~10^3 double operations per observation, once per LM step.
template <int MODEL, typename T>
ISFM_HD void ba_linearize(const T* cam, const T* pp, const T* X, const T* obs, T r[2],
                          T* Jc /* 2*D */, T Jp[6]) {
  if (sizeof(T) == 8) {
    ba_linearize_in<MODEL, T>(cam, pp, X, obs, r, Jc, Jp);
  } else {
    constexpr int NI = ModelTraits<MODEL>::NI;
    constexpr int D = 6 + NI, CW = 7 + NI;
    double c[CW], p2[2], x[3], o[2], rd[2], jc[2 * D], jp[6];
#pragma unroll
    for (int i = 0; i < CW; ++i) c[i] = (double)cam[i];
    p2[0] = (double)pp[0]; p2[1] = (double)pp[1];
    x[0] = (double)X[0]; x[1] = (double)X[1]; x[2] = (double)X[2];
    o[0] = (double)obs[0]; o[1] = (double)obs[1];
    ba_linearize_in<MODEL, double>(c, p2, x, o, rd, jc, jp);
    r[0] = (T)rd[0]; r[1] = (T)rd[1];
#pragma unroll
    for (int i = 0; i < 2 * D; ++i) Jc[i] = (T)jc[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) Jp[i] = (T)jp[i];
  }
}

// Huber on s = ||r||^2 (pypose.optim.kernel.Huber): rho and sqrt(rho').
template <typename T> ISFM_HD void huber(T s, T delta, T& rho, T& w) {
  T rs = sqrt(s);
  if (rs < delta) { rho = s; w = T(1); }
  else { rho = T(2) * delta * rs - delta * delta; w = sqrt(delta / rs); }
}

// Left retraction of a pose: Exp([tau, phi]) * (t, q)   
template <typename T> ISFM_HD void se3_retract(const T* pose7, const T* delta6, T* out7) {
  const T* tau = delta6; const T* phi = delta6 + 3;
  T th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  T B, C, kq, wq;
  if (th2 < T(1e-8)) {
    B = T(0.5) - th2 / 24; C = T(1.0 / 6) - th2 / 120; kq = T(0.5) - th2 / 48; wq = T(1) - th2 / 8;
  } else {
    T th = sqrt(th2);
    B = (T(1) - cos(th)) / th2; C = (th - sin(th)) / (th2 * th);
    kq = sin(T(0.5) * th) / th; wq = cos(T(0.5) * th);
  }
  // t_delta = J_l(phi) tau = tau + B phi x tau + C phi x (phi x tau)
  T c1[3] = {phi[1] * tau[2] - phi[2] * tau[1], phi[2] * tau[0] - phi[0] * tau[2], phi[0] * tau[1] - phi[1] * tau[0]};
  T c2[3] = {phi[1] * c1[2] - phi[2] * c1[1], phi[2] * c1[0] - phi[0] * c1[2], phi[0] * c1[1] - phi[1] * c1[0]};
  T td[3] = {tau[0] + B * c1[0] + C * c2[0], tau[1] + B * c1[1] + C * c2[1], tau[2] + B * c1[2] + C * c2[2]};
  T qd[4] = {kq * phi[0], kq * phi[1], kq * phi[2], wq};
  T Rd[9];
  quat_to_rot(qd, Rd);
  const T* t = pose7; const T* q = pose7 + 3;
  out7[0] = Rd[0] * t[0] + Rd[1] * t[1] + Rd[2] * t[2] + td[0];
  out7[1] = Rd[3] * t[0] + Rd[4] * t[1] + Rd[5] * t[2] + td[1];
  out7[2] = Rd[6] * t[0] + Rd[7] * t[1] + Rd[8] * t[2] + td[2];
  T ax = qd[0], ay = qd[1], az = qd[2], aw = qd[3], bx = q[0], by = q[1], bz = q[2], bw = q[3];
  T qx = aw * bx + ax * bw + ay * bz - az * by;
  T qy = aw * by - ax * bz + ay * bw + az * bx;
  T qz = aw * bz + ax * by - ay * bx + az * bw;
  T qw = aw * bw - ax * bx - ay * by - az * bz;
  T inv = T(1) / sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
  out7[3] = qx * inv; out7[4] = qy * inv; out7[5] = qz * inv; out7[6] = qw * inv;
}

// --------------------------------------------------------------------------------------
// small SPD blocks (K3)
// --------------------------------------------------------------------------------------
// Damped diagonal: clamp(d, 1e-6, 1e32) * mu  (pypose LM: A.diagonal().clamp_(min,max) then
// cumulative diag += diag * damping; mu = prod (1 + lambda_j)).
template <typename T> ISFM_HD T damp_diag(T d, T mu) {
  T c = d < T(1e-6) ? T(1e-6) : (d > T(1e32) ? T(1e32) : d);
  return c * mu;
}

// inverse of a symmetric 3x3 given as (xx xy xz yy yz zz); computed in double on the matrix
// scaled by its largest diagonal entry (no overflow of the determinant under heavy damping).
template <typename T> ISFM_HD void sym3_inverse(const T h[6], T inv[6]) {
  double is = 1.0;
  if (sizeof(T) == 8) {   // float entries (<= 1e32 * 1e24) cannot overflow a double determinant: no scaling needed
    double s = (double)h[0];
    if ((double)h[3] > s) s = (double)h[3];
    if ((double)h[5] > s) s = (double)h[5];
    is = s > 0.0 ? 1.0 / s : 1.0;
  }
  double a = h[0] * is, b = h[1] * is, c = h[2] * is, d = h[3] * is, e = h[4] * is, f = h[5] * is;
  double A = d * f - e * e, Bc = c * e - b * f, Cc = b * e - c * d;
  double det = a * A + b * Bc + c * Cc;
  double id = is / det;
  inv[0] = T(A * id); inv[1] = T(Bc * id); inv[2] = T(Cc * id);
  inv[3] = T((a * f - c * c) * id); inv[4] = T((b * c - a * e) * id);
  inv[5] = T((a * d - b * b) * id);
}

// In-place inverse of an SPD DxD matrix (row-major, full storage) by Cholesky, in double.
// Returns false if a pivot is not positive (pivot is then floored).
template <int D> ISFM_HD bool spd_inverse(double* A) {
  bool ok = true;
  // Cholesky A = L L^T, L stored in the lower triangle
  for (int j = 0; j < D; ++j) {
    double s = A[j * D + j];
    for (int k = 0; k < j; ++k) s -= A[j * D + k] * A[j * D + k];
    if (!(s > 0.0)) { ok = false; s = 1e-30; }
    double l = sqrt(s);
    A[j * D + j] = l;
    double il = 1.0 / l;
    for (int i = j + 1; i < D; ++i) {
      double t = A[i * D + j];
      for (int k = 0; k < j; ++k) t -= A[i * D + k] * A[j * D + k];
      A[i * D + j] = t * il;
    }
  }
  // invert L in place (lower triangular)
  for (int j = 0; j < D; ++j) {
    A[j * D + j] = 1.0 / A[j * D + j];
    for (int i = j + 1; i < D; ++i) {
      double t = 0.0;
      for (int k = j; k < i; ++k) t -= A[i * D + k] * A[k * D + j];
      A[i * D + j] = t / A[i * D + i];
    }
  }
  // A^-1 = L^-T L^-1 ; write the full symmetric result
  for (int i = 0; i < D; ++i)
    for (int j = 0; j <= i; ++j) {
      double t = 0.0;
      for (int k = i; k < D; ++k) t += A[k * D + i] * A[k * D + j];
      A[j * D + i] = t;  // upper triangle (j <= i) is free to overwrite: L^-1 lives below
    }
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < i; ++j) A[i * D + j] = A[j * D + i];
  return ok;
}

// --------------------------------------------------------------------------------------
// Prolongation block of the two-level PCG preconditioner (coarse.cuh): the seven similarity modes
// of a rigid piece of the scene about the centre c0 -- translation v, rotation w, scale s --
// expressed in one camera's left-perturbation tangent [dtau, dphi] (pose c = [t, q_xyzw, ...]):
//   dphi = -R w,   dtau = -R v - tt x (R w) + s tt,   tt = t + R c0.
// P is [6][7] row-major: rows dtau (3), dphi (3); columns v (3), w (3), s.
// --------------------------------------------------------------------------------------
template <typename T>
ISFM_HD void similarity_modes(const T* c, const double* o, T* P) {
  constexpr int NM = 7;
  double R[9];
  const double q[4] = {(double)c[3], (double)c[4], (double)c[5], (double)c[6]};
  quat_to_rot(q, R);
  double tt[3];
  for (int k = 0; k < 3; ++k) tt[k] = (double)c[k] + R[k * 3 + 0] * o[0] + R[k * 3 + 1] * o[1] + R[k * 3 + 2] * o[2];
  const double tx[9] = {0, -tt[2], tt[1], tt[2], 0, -tt[0], -tt[1], tt[0], 0};
  for (int r = 0; r < 3; ++r) {
    for (int k = 0; k < 3; ++k) {
      P[r * NM + k] = (T)(-R[r * 3 + k]);                                                                   // dtau / v
      P[r * NM + 3 + k] = (T)(-(tx[r * 3 + 0] * R[0 * 3 + k] + tx[r * 3 + 1] * R[1 * 3 + k] + tx[r * 3 + 2] * R[2 * 3 + k]));   // dtau / w
      P[(3 + r) * NM + k] = T(0);                                                                           // dphi / v
      P[(3 + r) * NM + 3 + k] = (T)(-R[r * 3 + k]);                                                         // dphi / w
    }
    P[r * NM + 6] = (T)tt[r];                                                                               // dtau / s
    P[(3 + r) * NM + 6] = T(0);
  }
}

}  // namespace isfm
