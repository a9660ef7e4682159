// float32 (product) and float64 (validation) instantiations of the global-positioning solver.
#include "gp_solver.cuh"
namespace isfm {
GPSolverBase* make_gp_solver_f32(const isfm_gp_desc& d) { return new GPSolver<float>(d); }
GPSolverBase* make_gp_solver_f64(const isfm_gp_desc& d) { return new GPSolver<double>(d); }
}  // namespace isfm
