// float64 (validation) instantiation of the BA solver: same kernels, double arithmetic.
#include "ba_solver.cuh"
namespace isfm {
BASolverBase* make_ba_solver_f64(const isfm_ba_desc& d) { return make_ba_solver_t<double>(d); }
}  // namespace isfm
