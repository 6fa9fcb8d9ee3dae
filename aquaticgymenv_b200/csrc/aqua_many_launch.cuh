// aqua_many_launch.cuh -- instantiation of K1m (step_many_kernel) split over several translation
// units (aqua_many_f32v1.cu, ..f32v2.cu, ..f32v4.cu, ..f64.cu compile side by side): the kernel
// exists once per (precision, envs per thread, action dtype, tape / random policy, obstacle test).
#pragma once

#include "aqua_step_many.cuh"

namespace aqua {

// what a translation unit is asked to do with one instantiation
struct ManyCall {
    int kind;  // 0: launch (a, g, s); 1: occupancy query (threads, use_grid -> *out)
    const StepArgs* a;
    LaunchGeom g;
    cudaStream_t s;
    int threads, use_grid;
    int* out;
};

// dynamic shared memory of a launch: culling grid, refill queues, stash records, the lanes' records in the making
template <int V> static size_t many_smem(int threads, int use_grid) {
    return cell_bytes(use_grid) + (size_t)(threads / 32) * sizeof(RefillQueue<V>) + (size_t)threads * V * 32 +
           (size_t)threads * sizeof(LaneWork);
}

template <typename R, int V, int ACT, bool TAPE, int TAB, bool INLOOP> static cudaError_t many_call(const ManyCall& c) {
    if (c.kind == 1)
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(c.out, step_many_kernel<R, V, ACT, TAPE, TAB, INLOOP>, c.threads,
                                                             many_smem<V>(c.threads, c.use_grid));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)c.g.blocks);
    cfg.blockDim = dim3((unsigned)c.g.threads);
    cfg.dynamicSmemBytes = many_smem<V>(c.g.threads, TAB == 2);
    cfg.stream = c.s;
    cudaLaunchAttribute attr[1];  // consecutive launches overlap their launch latency and prologue
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, step_many_kernel<R, V, ACT, TAPE, TAB, INLOOP>, *c.a);
}

template <typename R, int V, int TAB, bool INLOOP> static cudaError_t many_act(int act, int tape, const ManyCall& c) {
    if (!tape)  // random policy: the action dtype only tells discrete from continuous
        return act == AQUA_ACT_F32X2 ? many_call<R, V, AQUA_ACT_F32X2, false, TAB, INLOOP>(c)
                                     : many_call<R, V, AQUA_ACT_I32, false, TAB, INLOOP>(c);
    switch (act) {
        case AQUA_ACT_I32: return many_call<R, V, AQUA_ACT_I32, true, TAB, INLOOP>(c);
        case AQUA_ACT_I64: return many_call<R, V, AQUA_ACT_I64, true, TAB, INLOOP>(c);
        case AQUA_ACT_U8: return many_call<R, V, AQUA_ACT_U8, true, TAB, INLOOP>(c);
        case AQUA_ACT_F32X2: return many_call<R, V, AQUA_ACT_F32X2, true, TAB, INLOOP>(c);
        default: return cudaErrorInvalidValue;
    }
}

// tab: 0 brute force, 1 small table (float only), 2 culling grid; inloop: see step_many_kernel
template <typename R, int V> static cudaError_t many_tab(int tab, int inloop, int act, int tape, const ManyCall& c) {
    if (tab == 2) return many_act<R, V, 2, true>(act, tape, c);
    if constexpr (sizeof(R) == 4) {
        if (tab == 1) return inloop ? many_act<R, V, 1, true>(act, tape, c) : many_act<R, V, 1, false>(act, tape, c);
    }
    return inloop ? many_act<R, V, 0, true>(act, tape, c) : many_act<R, V, 0, false>(act, tape, c);
}

// the in-kernel policy: one env per thread, discrete actions, records drawn in the step loop
template <typename R> static cudaError_t many_mlp(int tab, const ManyCall& c) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)c.g.blocks);
    cfg.blockDim = dim3((unsigned)c.g.threads);
    cfg.dynamicSmemBytes = many_smem<1>(c.g.threads, tab == 2);
    cfg.stream = c.s;
    if (tab == 2) return cudaLaunchKernelEx(&cfg, step_many_kernel<R, 1, AQUA_ACT_I32, false, 2, true, true>, *c.a);
    if constexpr (sizeof(R) == 4) {
        if (tab == 1) return cudaLaunchKernelEx(&cfg, step_many_kernel<R, 1, AQUA_ACT_I32, false, 1, true, true>, *c.a);
    }
    return cudaLaunchKernelEx(&cfg, step_many_kernel<R, 1, AQUA_ACT_I32, false, 0, true, true>, *c.a);
}

// one per translation unit
cudaError_t many_tu_mlp_f32(int tab, const ManyCall& c);
cudaError_t many_tu_mlp_f64(int tab, const ManyCall& c);
cudaError_t many_tu_f32v1(int tab, int inloop, int act, int tape, const ManyCall& c);
cudaError_t many_tu_f32v2(int tab, int inloop, int act, int tape, const ManyCall& c);
cudaError_t many_tu_f32v4(int tab, int inloop, int act, int tape, const ManyCall& c);
cudaError_t many_tu_f64(int v, int tab, int inloop, int act, int tape, const ManyCall& c);

}  // namespace aqua
