// aqua_many.cu -- launcher of K1m (step_many_kernel, aqua_step_many.cuh): K env-steps per launch
// from an action tape (aqua_step_many) or under the synthetic random policy (aqua_rollout_random).
// Its own translation unit so that the two kernel families compile side by side.
#include "aqua_many_launch.cuh"

namespace aqua {

static int many_tab_kind(const World& w) { return w.small_mode != 0 ? 1 : (w.use_grid ? 2 : 0); }

static cudaError_t many_dispatch(int precision, int v, int tab, int inloop, int act, int tape, const ManyCall& c) {
    if (precision == AQUA_F64) return many_tu_f64(v, tab == 1 ? 0 : tab, inloop, act, tape, c);
    if (v == 1) return many_tu_f32v1(tab, inloop, act, tape, c);
    if (v == 2) return many_tu_f32v2(tab, inloop, act, tape, c);
    if (v == 4) return many_tu_f32v4(tab, inloop, act, tape, c);
    return cudaErrorInvalidValue;
}

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
    ManyCall c = {};
    c.kind = 0;
    c.a = &b;
    c.g = g;
    c.s = s;
    const int tab = many_tab_kind(a.w);
    return many_dispatch(precision, g.v, tab, tab == 2 || a.steps > kRingStartOnlySteps, action_dtype, tape, c);
}

cudaError_t launch_step_policy(int precision, const StepArgs& a, int threads, cudaStream_t s) {
    const int64_t n = a.env_end - a.env_begin;
    if (n <= 0 || n > ((int64_t)1 << 31) || a.mlp_w == nullptr) return cudaErrorInvalidValue;
    StepArgs b = a;
    const size_t es = precision == AQUA_F64 ? 8 : 4;
    for (int pl = 0; pl < AQUA_N_PLANES; ++pl)
        b.pl0[pl] = static_cast<const char*>(a.planes) + ((size_t)pl * (size_t)a.ld) * es +
                    (size_t)a.env_begin * (pl == AQUA_PLANE_TIME ? 4 : es);
    b.act0 = nullptr;
    b.rew0 = a.reward != nullptr ? static_cast<char*>(a.reward) + (size_t)a.env_begin * es : nullptr;
    b.flg0 = a.flags != nullptr ? a.flags + a.env_begin : nullptr;
    if (b.act_out != nullptr) b.act_out += a.env_begin;
    b.n_stream = (uint32_t)n;
    b.env_id0 = (uint32_t)(a.w.env_id_base + (uint64_t)a.env_begin);
    ManyCall c = {};
    c.kind = 0;
    c.a = &b;
    c.g.threads = threads;
    c.g.v = 1;
    c.g.blocks = (int)((n + threads - 1) / threads);
    c.s = s;
    const int tab = many_tab_kind(a.w);
    return precision == AQUA_F64 ? many_tu_mlp_f64(tab, c) : many_tu_mlp_f32(tab, c);
}

// `tab`: 0 brute force, 1 small table, 2 culling grid (as the launcher derives it from the World)
cudaError_t many_occupancy(int precision, int v, int action_dtype, int tape, int threads, int tab, int* blocks_per_sm) {
    ManyCall c = {};
    c.kind = 1;
    c.threads = threads;
    c.use_grid = tab == 2;
    c.out = blocks_per_sm;
    return many_dispatch(precision, v, tab, tab == 2, action_dtype, tape, c);
}

// Fresh ring: no slot holds an episode.  The records' own tag words count as well (a record is taken
// on the strength of the tag it carries): the whole allocation is filled with 0xff = kStashEmpty, so
// that no record a previous owner of the memory left behind can pass for one of this env's.
cudaError_t launch_stash_init(void* ring, size_t bytes, cudaStream_t s) {
    static_assert(kStashEmpty == 0xffffffffu, "memset pattern");
    return cudaMemsetAsync(ring, 0xff, bytes, s);
}

int stash_ring_depth() { return kRingDepth; }

}  // namespace aqua
