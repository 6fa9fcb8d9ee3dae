// step_many_kernel<double, 1 | 2, ...>: see aqua_many_launch.cuh
#include "aqua_many_launch.cuh"
namespace aqua {
cudaError_t many_tu_f64(int v, int tab, int inloop, int act, int tape, const ManyCall& c) {
    if (v == 1) return many_tab<double, 1>(tab, inloop, act, tape, c);
    if (v == 2) return many_tab<double, 2>(tab, inloop, act, tape, c);
    return cudaErrorInvalidValue;
}
}  // namespace aqua
