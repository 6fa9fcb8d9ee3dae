// step_many_kernel<float, 1, ...>: see aqua_many_launch.cuh
#include "aqua_many_launch.cuh"
namespace aqua {
cudaError_t many_tu_f32v1(int tab, int inloop, int act, int tape, const ManyCall& c) {
    return many_tab<float, 1>(tab, inloop, act, tape, c);
}
}  // namespace aqua
