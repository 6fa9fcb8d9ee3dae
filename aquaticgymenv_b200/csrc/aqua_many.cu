// aqua_many.cu -- launcher of K1m (step_many_kernel, aqua_step_many.cuh): K env-steps per launch
// from an action tape (aqua_step_many) or under the synthetic random policy (aqua_rollout_random).
// Its own translation unit so that the two kernel families compile side by side.
#include "aqua_step_many.cuh"

namespace aqua {

template <typename R, int V, typename Op> static cudaError_t many_act(int act, int tape, const Op& op) {
    if (!tape)  // random policy: the action dtype only tells discrete from continuous
        return act == AQUA_ACT_F32X2 ? op.template run<R, V, AQUA_ACT_F32X2, false>()
                                     : op.template run<R, V, AQUA_ACT_I32, false>();
    switch (act) {
        case AQUA_ACT_I32: return op.template run<R, V, AQUA_ACT_I32, true>();
        case AQUA_ACT_I64: return op.template run<R, V, AQUA_ACT_I64, true>();
        case AQUA_ACT_U8: return op.template run<R, V, AQUA_ACT_U8, true>();
        case AQUA_ACT_F32X2: return op.template run<R, V, AQUA_ACT_F32X2, true>();
        default: return cudaErrorInvalidValue;
    }
}
template <typename Op> static cudaError_t many_dispatch(int precision, int v, int act, int tape, const Op& op) {
    if (precision == AQUA_F64) {
        if (v == 1) return many_act<double, 1>(act, tape, op);
        if (v == 2) return many_act<double, 2>(act, tape, op);
        return cudaErrorInvalidValue;
    }
    if (v == 1) return many_act<float, 1>(act, tape, op);
    if (v == 2) return many_act<float, 2>(act, tape, op);
    if (v == 4) return many_act<float, 4>(act, tape, op);
    return cudaErrorInvalidValue;
}

// dynamic shared memory of a launch: culling grid, refill queues, stash records
template <int V> static size_t many_smem(int threads, int use_grid) {
    return cell_bytes(use_grid) + (size_t)(threads / 32) * sizeof(RefillQueue<V>) + (size_t)threads * V * 32;
}

struct ManyLaunchOp {
    const StepArgs& a;
    LaunchGeom g;
    cudaStream_t s;
    template <typename R, int V, int ACT, bool TAPE> cudaError_t run() const {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)g.blocks);
        cfg.blockDim = dim3((unsigned)g.threads);
        cfg.dynamicSmemBytes = many_smem<V>(g.threads, a.w.use_grid);
        cfg.stream = s;
        cudaLaunchAttribute attr[1];  // consecutive launches overlap their launch latency and prologue
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, step_many_kernel<R, V, ACT, TAPE>, a);
    }
};

// `a`: env_begin / env_end, actions / reward / flags (row 0), row_stride, steps, step0, stash set by
// the caller; the base pointers the kernel addresses with are filled here
cudaError_t launch_step_many(int precision, int action_dtype, int tape, const StepArgs& a, LaunchGeom g, cudaStream_t s) {
    const int64_t n = a.env_end - a.env_begin;
    if (n <= 0 || n > ((int64_t)1 << 31)) return cudaErrorInvalidValue;
    StepArgs b = a;
    const size_t es = precision == AQUA_F64 ? 8 : 4;
    size_t ab = 4;
    if (action_dtype == AQUA_ACT_I64 || action_dtype == AQUA_ACT_F32X2) ab = 8;
    else if (action_dtype == AQUA_ACT_U8) ab = 1;
    for (int pl = 0; pl < AQUA_N_PLANES; ++pl)
        b.pl0[pl] = static_cast<const char*>(a.planes) + ((size_t)pl * (size_t)a.ld) * es +
                    (size_t)a.env_begin * (pl == AQUA_PLANE_TIME ? 4 : es);
    b.act0 = a.actions != nullptr ? static_cast<const char*>(a.actions) + (size_t)a.env_begin * ab : nullptr;
    b.rew0 = a.reward != nullptr ? static_cast<char*>(a.reward) + (size_t)a.env_begin * es : nullptr;
    b.flg0 = a.flags != nullptr ? a.flags + a.env_begin : nullptr;
    b.n_stream = (uint32_t)n;
    b.env_id0 = (uint32_t)(a.w.env_id_base + (uint64_t)a.env_begin);
    const int64_t per_cta = (int64_t)g.threads * g.v;
    g.blocks = (int)((n + per_cta - 1) / per_cta);
    return many_dispatch(precision, g.v, action_dtype, tape, ManyLaunchOp{b, g, s});
}

struct ManyOccupancyOp {
    int threads, use_grid;
    int* out;
    template <typename R, int V, int ACT, bool TAPE> cudaError_t run() const {
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, step_many_kernel<R, V, ACT, TAPE>, threads,
                                                             many_smem<V>(threads, use_grid));
    }
};
cudaError_t many_occupancy(int precision, int v, int action_dtype, int tape, int threads, int use_grid, int* blocks_per_sm) {
    return many_dispatch(precision, v, action_dtype, tape, ManyOccupancyOp{threads, use_grid, blocks_per_sm});
}

cudaError_t launch_stash_init(uint4* stash, int64_t ld, cudaStream_t s) {
    const int threads = 256;
    stash_init_kernel<<<(int)((ld + threads - 1) / threads), threads, 0, s>>>(stash, ld);
    return cudaGetLastError();
}

}  // namespace aqua
