// TEST-ONLY host build of instantsfm_b200/csrc/math.cuh.
// Lets the CPU test-suite (no GPU in the build container) check the per-observation
// arithmetic the CUDA kernels inline -- camera models + Jacobians, SE(3) retraction,
// small SPD inverses -- against the oracle.  Never loaded by the product package.
#include "../../instantsfm_b200/csrc/math.cuh"
#include "../../instantsfm_b200/csrc/camera_maps.cuh"
#include "../../instantsfm_b200/csrc/filter_math.cuh"

using namespace isfm;

template <int MODEL, typename T>
static void lin_loop(long n, const T* cam, const T* pp, const T* X, const T* obs, T* r, T* Jc, T* Jp) {
  constexpr int NI = ModelTraits<MODEL>::NI;
  constexpr int D = 6 + NI;
  for (long i = 0; i < n; ++i)
    ba_linearize<MODEL, T>(cam + i * (7 + NI), pp + 2 * i, X + 3 * i, obs + 2 * i, r + 2 * i, Jc + 2 * D * i, Jp + 6 * i);
}
template <int MODEL, typename T>
static void res_loop(long n, const T* cam, const T* pp, const T* X, const T* obs, T* r) {
  constexpr int NI = ModelTraits<MODEL>::NI;
  for (long i = 0; i < n; ++i) ba_residual<MODEL, T>(cam + i * (7 + NI), pp + 2 * i, X + 3 * i, obs + 2 * i, r + 2 * i);
}

#define DISPATCH(FN, ...)                                  \
  switch (model) {                                         \
    case 0: FN<0, T>(__VA_ARGS__); break;                  \
    case 1: FN<1, T>(__VA_ARGS__); break;                  \
    case 2: FN<2, T>(__VA_ARGS__); break;                  \
    case 3: FN<3, T>(__VA_ARGS__); break;                  \
    case 4: FN<4, T>(__VA_ARGS__); break;                  \
    case 5: FN<5, T>(__VA_ARGS__); break;                  \
    case 6: FN<6, T>(__VA_ARGS__); break;                  \
    case 8: FN<8, T>(__VA_ARGS__); break;                  \
    case 9: FN<9, T>(__VA_ARGS__); break;                  \
    default: return -2;                                    \
  }

template <typename T>
static int lin(int model, long n, const T* cam, const T* pp, const T* X, const T* obs, T* r, T* Jc, T* Jp) {
  DISPATCH(lin_loop, n, cam, pp, X, obs, r, Jc, Jp);
  return 0;
}
template <typename T>
static int res(int model, long n, const T* cam, const T* pp, const T* X, const T* obs, T* r) {
  DISPATCH(res_loop, n, cam, pp, X, obs, r);
  return 0;
}

extern "C" {
int hc_linearize_f64(int model, long n, const double* cam, const double* pp, const double* X, const double* obs,
                     double* r, double* Jc, double* Jp) { return lin<double>(model, n, cam, pp, X, obs, r, Jc, Jp); }
int hc_linearize_f32(int model, long n, const float* cam, const float* pp, const float* X, const float* obs,
                     float* r, float* Jc, float* Jp) { return lin<float>(model, n, cam, pp, X, obs, r, Jc, Jp); }
int hc_residual_f64(int model, long n, const double* cam, const double* pp, const double* X, const double* obs,
                    double* r) { return res<double>(model, n, cam, pp, X, obs, r); }
void hc_se3_retract_f64(long n, const double* pose, const double* delta, double* out) {
  for (long i = 0; i < n; ++i) se3_retract<double>(pose + 7 * i, delta + 6 * i, out + 7 * i);
}
void hc_se3_retract_f32(long n, const float* pose, const float* delta, float* out) {
  for (long i = 0; i < n; ++i) se3_retract<float>(pose + 7 * i, delta + 6 * i, out + 7 * i);
}
void hc_sym3_inverse_f64(long n, const double* h, double* inv) {
  for (long i = 0; i < n; ++i) sym3_inverse<double>(h + 6 * i, inv + 6 * i);
}
int hc_spd_inverse(int D, double* A) {
  switch (D) {
    case 3: return spd_inverse<3>(A); case 7: return spd_inverse<7>(A); case 8: return spd_inverse<8>(A);
    case 9: return spd_inverse<9>(A); case 12: return spd_inverse<12>(A); case 16: return spd_inverse<16>(A);
    default: return -1;
  }
}
void hc_huber_f64(long n, const double* s, double delta, double* rho, double* w) {
  for (long i = 0; i < n; ++i) huber<double>(s[i], delta, rho[i], w[i]);
}
}

// scene-layer camera maps (csrc/camera_maps.cuh): one camera-table row, n points
extern "C" {
void hc_cam2img(long n, const double* row, const double* uvw, double* out) {
  const CamRow c = load_row(row);
  for (long i = 0; i < n; ++i) cam2img(c, uvw[3 * i], uvw[3 * i + 1], uvw[3 * i + 2], out[2 * i], out[2 * i + 1]);
}
void hc_img2cam(long n, const double* row, const double* xy, double* out) {
  const CamRow c = load_row(row);
  for (long i = 0; i < n; ++i) img2cam(c, xy[2 * i], xy[2 * i + 1], out[2 * i], out[2 * i + 1]);
}
}

// prolongation blocks of the two-level preconditioner (math.cuh::similarity_modes): n poses [n][7], one centre
extern "C" {
void hc_similarity_modes_f64(long n, const double* pose, const double* c0, double* P) {
  for (long i = 0; i < n; ++i) similarity_modes<double>(pose + 7 * i, c0, P + 42 * i);
}
void hc_similarity_modes_f32(long n, const float* pose, const double* c0, float* P) {
  for (long i = 0; i < n; ++i) similarity_modes<float>(pose + 7 * i, c0, P + 42 * i);
}
}

// per-observation track-filter tests (csrc/filter_math.cuh), the loop of filter_observations_kernel
extern "C" int hc_filter_observations(int mode, long n_obs, const double* world2cam, const double* xyz, const double* feat,
                                      const int* image_ids, const int* track_idx, double thr, unsigned char* valid_out) {
  for (long i = 0; i < n_obs; ++i) {
    const double* M = world2cam + (long)image_ids[i] * 16;
    const double* X = xyz + (long)track_idx[i] * 3;
    const bool k = mode == 0 ? filter_keep<0>(M, X[0], X[1], X[2], feat[3 * i], feat[3 * i + 1], feat[3 * i + 2], thr)
                             : filter_keep<1>(M, X[0], X[1], X[2], feat[3 * i], feat[3 * i + 1], feat[3 * i + 2], thr);
    valid_out[i] = k ? 1 : 0;
  }
  return 0;
}

