// step_many_kernel<.., MLP = true>: K env-steps per launch with the reference's Q-network evaluated
// in the kernel (aqua_rollout_policy); see aqua_many_launch.cuh
#include "aqua_many_launch.cuh"
namespace aqua {
cudaError_t many_tu_mlp_f32(int tab, const ManyCall& c) { return many_mlp<float>(tab, c); }
cudaError_t many_tu_mlp_f64(int tab, const ManyCall& c) { return many_mlp<double>(tab == 1 ? 0 : tab, c); }
}  // namespace aqua
