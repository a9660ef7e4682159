// Standalone timing harness for fused_linearize_kernel with phase ablation (tools, not product).
// nvcc -DISFM_KBENCH -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o build/kbench tools/kbench.cu
#include <cstdio>
#include <random>
#include <vector>
#include "../instantsfm_b200/csrc/ba_kernels.cuh"
namespace isfm { int64_t g_launch_count = 0; void set_last_error(const std::string&) {} const char* get_last_error() { return ""; } }
using namespace isfm;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main(int argc, char** argv) {
  const int n_cam = 1778, n_pt = 993000; const bool wide = argc > 2 ? atoi(argv[2]) != 0 : true;
  std::mt19937 rng(1);
  std::vector<int> off(n_pt + 1, 0);
  std::geometric_distribution<int> geo(1.0 / 4.03);
  for (int p = 0; p < n_pt; ++p) off[p + 1] = off[p] + std::min(2 + geo(rng), 200);
  const int64_t n_obs = off[n_pt];
  std::vector<int> cam_of(n_obs), pt_of(n_obs);
  std::vector<float> cam(n_cam * 12, 0.f), pp(n_cam * 2, 0.f), pts(n_pt * 3), obs(n_obs * 2);
  std::uniform_real_distribution<float> U(-1.f, 1.f);
  for (int c = 0; c < n_cam; ++c) { float* r = &cam[c * 12]; r[0] = U(rng); r[1] = U(rng); r[2] = 20.f + U(rng); r[3] = 0.01f * U(rng); r[4] = 0.01f * U(rng); r[5] = 0.01f * U(rng); r[6] = 1.f; r[7] = 1000.f; r[8] = 0.01f * U(rng); r[9] = 0.001f * U(rng); }
  for (int p = 0; p < n_pt; ++p) { pts[3 * p] = 5 * U(rng); pts[3 * p + 1] = 5 * U(rng); pts[3 * p + 2] = 5 * U(rng); }
  for (int p = 0; p < n_pt; ++p) { int base = (int)((int64_t)p * n_cam / n_pt); for (int a = off[p]; a < off[p + 1]; ++a) { pt_of[a] = p; cam_of[a] = wide ? (int)(rng() % n_cam) : (base + (a - off[p]) * 7 + rng() % 50) % n_cam; obs[2 * a] = 100 * U(rng); obs[2 * a + 1] = 100 * U(rng); } }
  std::vector<int> cta; cta.push_back(0); int start = 0;
  for (int p = 0; p < n_pt; ++p) if (off[p + 1] - off[start] > FUSED_TPB) { cta.push_back(p); start = p; }
  cta.push_back(n_pt);
  const int n_cta = (int)cta.size() - 1;
  printf("n_obs %lld n_cta %d\n", (long long)n_obs, n_cta);
  int *d_off, *d_cam_of, *d_pt_of; int4* d_cta; std::vector<int4> tiles; for (int i = 0; i < n_cta; ++i) tiles.push_back(make_int4(cta[i], cta[i + 1], off[cta[i]], off[cta[i + 1]] - off[cta[i]])); float *d_cam, *d_pp, *d_pts, *d_obs, *R, *OBS, *HPP, *GPT, *HPPINV, *TP; double *pa, *pb;
  CK(cudaMalloc(&d_off, off.size() * 4)); CK(cudaMalloc(&d_cam_of, n_obs * 4)); CK(cudaMalloc(&d_pt_of, n_obs * 4)); CK(cudaMalloc(&d_cta, tiles.size() * 16));
  CK(cudaMalloc(&d_cam, cam.size() * 4)); CK(cudaMalloc(&d_pp, pp.size() * 4)); CK(cudaMalloc(&d_pts, pts.size() * 4)); CK(cudaMalloc(&d_obs, obs.size() * 4));
  CK(cudaMalloc(&R, n_obs * 8)); CK(cudaMalloc(&OBS, n_obs * 128)); CK(cudaMalloc(&HPP, n_pt * 24)); CK(cudaMalloc(&GPT, n_pt * 12)); CK(cudaMalloc(&HPPINV, n_pt * 24)); CK(cudaMalloc(&TP, n_pt * 12));
  CK(cudaMalloc(&pa, n_cta * 8)); CK(cudaMalloc(&pb, n_cta * 8));
  CK(cudaMemcpy(d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_cam_of, cam_of.data(), n_obs * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_pt_of, pt_of.data(), n_obs * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_cta, tiles.data(), tiles.size() * 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_cam, cam.data(), cam.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_pp, pp.data(), pp.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_pts, pts.data(), pts.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_obs, obs.data(), obs.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(fused_linearize_kernel<float, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FusedCfg<float, 9>::SMEM));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  // plain write / copy references
  for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); cudaMemsetAsync(OBS, 0, n_obs * 128); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); printf("memset 640MB: %.1f us (%.0f GB/s)\n", ms * 1000, n_obs * 128 / ms / 1e6); }
  int flush = argc > 1 ? atoi(argv[1]) : 1; printf("L2 flush between repetitions: %d\n", flush);
  char* scratch; CK(cudaMalloc(&scratch, 512u << 20));
  int variants[] = {0, 64, 1, 2, 3, 4, 8, 32, 63};
  for (int v : variants) {
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      if (flush) cudaMemsetAsync(scratch, rep, 512u << 20);
      cudaEventRecord(e0);
      fused_linearize_kernel<float, 3><<<n_cta, FUSED_TPB, FusedCfg<float, 9>::SMEM>>>(d_cta, d_off, d_cam, d_pts, d_obs, d_cam_of, d_pt_of, 1.0f, 1.0001f, R, OBS, HPP, GPT, HPPINV, TP, pa, pb, (v & 64) ? 1 : 0, v);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
    }
    printf("dbg %2d: %.1f us\n", v, best * 1000);
  }
  CK(cudaGetLastError());
  return 0;
}
