// step_many_kernel<float, 2, ...>: see aqua_many_launch.cuh
#include "aqua_many_launch.cuh"
namespace aqua {
cudaError_t many_tu_f32v2(int tab, int inloop, int act, int tape, const ManyCall& c) {
    return many_tab<float, 2>(tab, inloop, act, tape, c);
}
}  // namespace aqua
