// float32 (product) instantiation of the BA solver for all nine camera models.
#include "ba_solver.cuh"
namespace isfm {
BASolverBase* make_ba_solver_f32(const isfm_ba_desc& d) { return make_ba_solver_t<float>(d); }
}  // namespace isfm
