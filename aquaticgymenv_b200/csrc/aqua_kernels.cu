// aqua_kernels.cu -- the sm_100a kernels of the batched AquaGym environment step.
//
// Replaces AquaEnv.step()/reset() of the reference (gym_aqua/envs/aqua.py:100-126, 135-213) for
// N independent env instances.  One thread owns V consecutive envs (V = 4 float / 2 double) so
// that every state plane is read and written with one 16-byte access; the obstacle table sits in
// shared memory; episode statistics are folded with warp reductions into per-CTA slots.
// The step is elementwise: there is no contraction, hence no tensor-core work; the bound is HBM
// bandwidth (65 B per discrete env-step).
//
// The kernels live in headers by family -- aqua_step_stream.cuh (K1s, large slabs),
// aqua_step_narrow.cuh (K1n, small slabs), aqua_step_general.cuh (K1 general step),
// aqua_step_many.cuh (K1m, K steps per launch; launched from aqua_many.cu), helpers in
// aqua_kernel_util.cuh, device arithmetic in aqua_device.cuh and aqua_packed.cuh; this file holds
// the small kernels (K2 reset, K3 statistics, K4 state access) and the single-step launchers.
// Host-side tables: aqua_tables.cu.
#include <algorithm>
#include <cstring>

#include "aqua_step_general.cuh"
#include "aqua_step_narrow.cuh"
#include "aqua_step_stream.cuh"

namespace aqua {

// ------------------------------------------------------------------------------------------
// K2: env.reset() -- aqua.py:100-126.  One thread per env (not a bandwidth path).
// ------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(kMaxThreads) reset_kernel(const __grid_constant__ StepArgs p,
                                                            const uint8_t* __restrict__ mask) {
    const int64_t i = p.env_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.env_end) return;
    if (mask != nullptr && mask[i] == 0) return;
    uint32_t* const ptw = plane_tw<R>(p.planes, p.ld);
    const uint32_t ep = ((ptw[i] >> AQUA_TIME_BITS) + 1u) & kEpisodeMask;
    const uint32_t env_id = (uint32_t)(p.w.env_id_base + (uint64_t)i);
    const ResetResult rr = reset_env(p.w.keys, env_id, ep, p.w.tab64, p.w.k_obst,
                                     WorldFlags{p.w.waves != 0, p.w.random_boat != 0, p.w.random_goal != 0}, p.w.forced);
    // the draws are float32 values: exact in either precision
    plane<R>(p.planes, p.ld, AQUA_PLANE_X)[i] = (R)rr.x;
    plane<R>(p.planes, p.ld, AQUA_PLANE_Y)[i] = (R)rr.y;
    plane<R>(p.planes, p.ld, AQUA_PLANE_THETA)[i] = (R)rr.th;
    plane<R>(p.planes, p.ld, AQUA_PLANE_WAVE_X)[i] = (R)rr.wx;
    plane<R>(p.planes, p.ld, AQUA_PLANE_WAVE_Y)[i] = (R)rr.wy;
    plane<R>(p.planes, p.ld, AQUA_PLANE_GOAL_X)[i] = (R)rr.gx;
    plane<R>(p.planes, p.ld, AQUA_PLANE_GOAL_Y)[i] = (R)rr.gy;
    ptw[i] = ep << AQUA_TIME_BITS;
}

// fresh handle: no episode opened yet (the first reset opens episode 0), pose planes zero
template <typename R>
__global__ void init_state_kernel(void* planes, int64_t ld) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ld) return;
#pragma unroll
    for (int pl = 0; pl < 7; ++pl) plane<R>(planes, ld, pl)[i] = R(0);
    plane_tw<R>(planes, ld)[i] = AQUA_EPISODE_MASK << AQUA_TIME_BITS;
}

// synthetic random policy: `action_space.sample()` (test_random.py:19-20) from the ACTION stream
template <int ACT>
__global__ void random_actions_kernel(const __grid_constant__ World w, int64_t n, uint64_t step, void* actions) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t env_id = (uint32_t)(w.env_id_base + (uint64_t)i);
    const uint2 r = draw_action_words(w.keys, env_id, step);
    if constexpr (ACT == AQUA_ACT_F32X2) {
        reinterpret_cast<float2*>(actions)[i] = make_float2(draw_thrust(r.x), draw_thrust(r.y));
    } else {
        const uint32_t a = __umulhi(r.x, 3u);
        if constexpr (ACT == AQUA_ACT_I32) reinterpret_cast<int32_t*>(actions)[i] = (int32_t)a;
        else if constexpr (ACT == AQUA_ACT_I64) reinterpret_cast<long long*>(actions)[i] = (long long)a;
        else reinterpret_cast<uint8_t*>(actions)[i] = (uint8_t)a;
    }
}

// canonical double[9][n] {x, y, theta, wave_x, wave_y, goal_x, goal_y, time, episode} <-> planes
__host__ __device__ constexpr int canonical_plane(int k) {
    return k < 3 ? k : (k < 5 ? AQUA_PLANE_WAVE_X + (k - 3) : AQUA_PLANE_GOAL_X + (k - 5));
}
template <typename R>
__global__ void get_state_kernel(const void* planes, int64_t ld, int64_t n, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    void* pl = const_cast<void*>(planes);
#pragma unroll
    for (int k = 0; k < 7; ++k) out[(int64_t)k * n + i] = (double)plane<R>(pl, ld, canonical_plane(k))[i];
    const uint32_t tw = plane_tw<R>(pl, ld)[i];
    out[(int64_t)7 * n + i] = (double)(tw & kTimeMask);
    out[(int64_t)8 * n + i] = (double)(tw >> AQUA_TIME_BITS);
}
template <typename R>
__global__ void set_state_kernel(void* planes, int64_t ld, int64_t n, const double* __restrict__ in) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < 7; ++k) plane<R>(planes, ld, canonical_plane(k))[i] = (R)in[(int64_t)k * n + i];
    const uint32_t t = min((uint32_t)in[(int64_t)7 * n + i], kTimeMask);
    const uint32_t ep = (uint32_t)in[(int64_t)8 * n + i] & kEpisodeMask;
    plane_tw<R>(planes, ld)[i] = (ep << AQUA_TIME_BITS) | t;
}

// Observation rows [n][5] float32 for a policy network, optionally scaled like
// AquaStateNormalizer (main/impl/utils.py:15-27): obs / (high - low) with the bounds of
// aqua.py:46-49, angle shifted by +0.5.  One thread per env; a warp reads 5 coalesced plane
// segments and writes 640 contiguous bytes.
template <typename R>
__global__ void observe_kernel(const void* planes, int64_t ld, int64_t n, float* __restrict__ out, int normalized) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    void* pl = const_cast<void*>(planes);
    float v[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) v[k] = (float)plane<R>(pl, ld, k)[i];  // planes 0..4 = x, y, theta, gx, gy
    if (normalized) {
        const float inv_pos = 1.0f / 100.0f, inv_ang = 1.0f / (2.0f * 3.14159265358979323846f);
        v[0] *= inv_pos;
        v[1] *= inv_pos;
        v[2] = fmaf(v[2], inv_ang, 0.5f);
        v[3] *= inv_pos;
        v[4] *= inv_pos;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) out[i * 5 + k] = v[k];
}

// Host-buffer step: the goal planes of the caller's host observation only change for envs whose
// episode changed.  When that buffer is pinned (hence mapped into the device's address space) the
// new goals of the finished envs of a chunk go straight to it: a few thousand 4-byte posted writes
// over PCIe per million envs instead of two whole planes.
template <typename R>
__global__ void goal_updates_kernel(const void* planes, int64_t ld, const uint8_t* __restrict__ flags, int64_t lo,
                                    int64_t len, R* __restrict__ h_gx, R* __restrict__ h_gy) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= len) return;
    const int64_t i = lo + k;
    if ((flags[i] & AQUA_FLAG_DONE) == 0) return;
    void* pl = const_cast<void*>(planes);
    h_gx[i] = plane<R>(pl, ld, AQUA_PLANE_GOAL_X)[i];
    h_gy[i] = plane<R>(pl, ld, AQUA_PLANE_GOAL_Y)[i];
}

// The outputs of a host step written by the SMs straight into the caller's MAPPED pinned buffers
// (posted writes over PCIe, 16 bytes per thread and plane): planes x, y, angle (and the goal planes,
// or -- sparse_goals -- only the goals of the envs that finished), reward, flags of the envs
// [lo, lo + len).  One launch instead of three copy-engine transfers and the goal kernel.  A thread
// moves 16 / sizeof(R) consecutive envs; `lo` is a multiple of AQUA_TILE, every base pointer 16-byte
// aligned, and the tail of a slab is moved element by element.
template <typename R>
__global__ void pack_host_kernel(const void* planes, int64_t ld, const R* __restrict__ reward,
                                 const uint8_t* __restrict__ flags, int64_t lo, int64_t len, int64_t n, R* h_obs, R* h_rew,
                                 uint8_t* h_flg, int sparse_goals) {
    constexpr int W = 16 / (int)sizeof(R);
    void* pl = const_cast<void*>(planes);
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * W;
    if (k >= len) return;
    const int64_t i = lo + k;
    const bool full = k + W <= len;
    uint8_t fl[W];
#pragma unroll
    for (int j = 0; j < W; ++j) fl[j] = (full || k + j < len) ? flags[i + j] : (uint8_t)0;
    if (h_obs != nullptr) {
        const int n_planes = sparse_goals ? 3 : 5;
        for (int p = 0; p < n_planes; ++p) {
            const R* src = plane<R>(pl, ld, p) + i;
            R* dst = h_obs + (int64_t)p * n + i;
            if (full && (((uintptr_t)dst) & 15u) == 0) {
                *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
            } else {
                for (int j = 0; j < W && k + j < len; ++j) dst[j] = src[j];
            }
        }
        if (sparse_goals) {
#pragma unroll
            for (int j = 0; j < W; ++j)
                if (fl[j] & AQUA_FLAG_DONE) {
                    h_obs[(int64_t)3 * n + i + j] = plane<R>(pl, ld, AQUA_PLANE_GOAL_X)[i + j];
                    h_obs[(int64_t)4 * n + i + j] = plane<R>(pl, ld, AQUA_PLANE_GOAL_Y)[i + j];
                }
        }
    }
    if (h_rew != nullptr) {
        if (full && (((uintptr_t)(h_rew + i)) & 15u) == 0) {
            *reinterpret_cast<uint4*>(h_rew + i) = *reinterpret_cast<const uint4*>(reward + i);
        } else {
            for (int j = 0; j < W && k + j < len; ++j) h_rew[i + j] = reward[i + j];
        }
    }
    if (h_flg != nullptr) {
        if (full && W == 4 && (((uintptr_t)(h_flg + i)) & 3u) == 0) {
            *reinterpret_cast<uint32_t*>(h_flg + i) = (uint32_t)fl[0] | ((uint32_t)fl[1] << 8) | ((uint32_t)fl[W > 2 ? 2 : 0] << 16) |
                                                       ((uint32_t)fl[W > 3 ? 3 : 0] << 24);
        } else {
            for (int j = 0; j < W && k + j < len; ++j) h_flg[i + j] = fl[j];
        }
    }
}

// K3: fold the per-CTA slots (kSlotStats counters each) and the forced-reset counter into
// int64[AQUA_N_STATS]
__global__ void stats_reduce_kernel(unsigned long long* slots, unsigned int* forced, int64_t* out, int clear) {
    __shared__ unsigned long long s[kSlotStats];
    if (threadIdx.x < kSlotStats) s[threadIdx.x] = 0ull;
    __syncthreads();
    unsigned long long acc[kSlotStats];
#pragma unroll
    for (int k = 0; k < kSlotStats; ++k) acc[k] = 0ull;
    for (int slot = threadIdx.x; slot < kStatSlots; slot += blockDim.x) {
#pragma unroll
        for (int k = 0; k < kSlotStats; ++k) {
            acc[k] += slots[slot * kSlotStats + k];
            if (clear) slots[slot * kSlotStats + k] = 0ull;
        }
    }
#pragma unroll
    for (int k = 0; k < kSlotStats; ++k) {
        unsigned long long v = acc[k];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s[k], v);
    }
    __syncthreads();
    if (threadIdx.x < kSlotStats) out[threadIdx.x] = (int64_t)s[threadIdx.x];
    if (threadIdx.x == 0) {
        out[AQUA_STAT_FORCED_RESETS] = (int64_t)*forced;
        if (clear) *forced = 0u;
    }
}

// ------------------------------------------------------------------------------------------
// launchers.  Every step / rollout kernel exists for V = Prec<R>::V envs per thread (wide: 16-byte
// plane accesses, the bandwidth layout) and for V = 1 (narrow: four times the threads and a
// quarter of the per-thread latency, for slabs too small to fill the GPU with wide threads).
// ------------------------------------------------------------------------------------------
template <typename R, int V, typename Op> static cudaError_t dispatch_act(int act, const Op& op) {
    switch (act) {
        case AQUA_ACT_I32: return op.template run<R, V, AQUA_ACT_I32>();
        case AQUA_ACT_I64: return op.template run<R, V, AQUA_ACT_I64>();
        case AQUA_ACT_U8: return op.template run<R, V, AQUA_ACT_U8>();
        case AQUA_ACT_F32X2: return op.template run<R, V, AQUA_ACT_F32X2>();
        default: return cudaErrorInvalidValue;
    }
}
template <typename Op> static cudaError_t dispatch_step(int precision, int v, int act, const Op& op) {
    if (precision == AQUA_F64)
        return v == 1 ? dispatch_act<double, 1>(act, op) : dispatch_act<double, Prec<double>::V>(act, op);
    return v == 1 ? dispatch_act<float, 1>(act, op) : dispatch_act<float, Prec<float>::V>(act, op);
}
struct StepLaunchOp {
    const StepArgs& a;
    LaunchGeom g;
    cudaStream_t s;
    template <typename R, int V, int ACT> cudaError_t run() const {
        const size_t smem = stage_bytes<R, V>() + cell_bytes(a.w.use_grid);
        if (a.noise_x != nullptr) step_kernel<R, V, ACT, true><<<g.blocks, g.threads, smem, s>>>(a);
        else step_kernel<R, V, ACT, false><<<g.blocks, g.threads, smem, s>>>(a);
        return cudaGetLastError();
    }
};
// streaming kernel: wide threads only; (precision, action dtype) pairs whose per-thread action
// footprint cp.async cannot move (double + uint8: 2 bytes) stay with step_kernel
struct StreamLaunchOp {
    const StepArgs& a;
    LaunchGeom g;
    cudaStream_t s;
    template <typename R, int V, int ACT> cudaError_t run() const {
        if constexpr (stream_supported<ACT, V>()) {
            const size_t smem = stage_bytes<R, V>() + cell_bytes(a.w.use_grid);
            // bulk (TMA) staging for launches large enough to stream from HBM, per-lane cp.async below
            const bool tma = AQUA_TMA == 2 || (AQUA_TMA == 1 && (int64_t)a.n_stream >= kTmaMinEnvs);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)g.blocks);
            cfg.blockDim = dim3((unsigned)g.threads);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = s;
            cudaLaunchAttribute attr[1];  // consecutive steps overlap launch latency and prologue
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = AQUA_PDL;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            if (a.w.waves) {
                if (tma) return cudaLaunchKernelEx(&cfg, step_stream_kernel<R, V, ACT, true, true>, a);
                return cudaLaunchKernelEx(&cfg, step_stream_kernel<R, V, ACT, true, false>, a);
            }
            if (tma) return cudaLaunchKernelEx(&cfg, step_stream_kernel<R, V, ACT, false, true>, a);
            return cudaLaunchKernelEx(&cfg, step_stream_kernel<R, V, ACT, false, false>, a);
        } else {
            return cudaErrorNotSupported;
        }
    }
};
struct StreamSupportedOp {
    bool* out;
    template <typename R, int V, int ACT> cudaError_t run() const {
        *out = stream_supported<ACT, V>();
        return cudaSuccess;
    }
};
static bool has_stream_kernel(int precision, int action_dtype) {
    bool yes = false;
    dispatch_step(precision, envs_per_thread(precision), action_dtype, StreamSupportedOp{&yes});
    return yes;
}

// narrow kernel: one env per thread, one CTA of kNarrowThreads per 128 envs (not persistent)
struct NarrowLaunchOp {
    const StepArgs& a;
    int blocks;
    cudaStream_t s;
    template <typename R, int V, int ACT> cudaError_t run() const {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)blocks);
        cfg.blockDim = dim3((unsigned)kNarrowThreads);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = AQUA_PDL;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (a.w.waves) return cudaLaunchKernelEx(&cfg, step_narrow_kernel<R, ACT, true>, a);
        return cudaLaunchKernelEx(&cfg, step_narrow_kernel<R, ACT, false>, a);
    }
};

// the fields the streaming and narrow kernels address with: base pointers at env_begin
static void fill_offsets(StepArgs& b, const StepArgs& a, int precision, int action_dtype) {
    const size_t es = precision == AQUA_F64 ? 8 : 4;
    size_t ab = 4;
    if (action_dtype == AQUA_ACT_I64 || action_dtype == AQUA_ACT_F32X2) ab = 8;
    else if (action_dtype == AQUA_ACT_U8) ab = 1;
    for (int pl = 0; pl < AQUA_N_PLANES; ++pl)
        b.pl0[pl] = static_cast<const char*>(a.planes) + ((size_t)pl * (size_t)a.ld) * es +
                    (size_t)a.env_begin * (pl == AQUA_PLANE_TIME ? 4 : es);
    b.act0 = static_cast<const char*>(a.actions) + (size_t)a.env_begin * ab;
    b.rew0 = static_cast<char*>(a.reward) + (size_t)a.env_begin * es;
    b.flg0 = a.flags + a.env_begin;
    b.tile_lane_mask = 0u;
    b.env_id0 = (uint32_t)(a.w.env_id_base + (uint64_t)a.env_begin);
}

// Narrow geometry: the narrow kernel.  Wide geometry: whole warps of wide threads go to the
// streaming kernel, what is left of the range (fewer than 32 V envs) to the general kernel.  Noise
// injection (parity tests) always runs the general kernel.
cudaError_t launch_step(int precision, int action_dtype, const StepArgs& a, LaunchGeom g, cudaStream_t s) {
    const int wide = envs_per_thread(precision);
    const int64_t n = a.env_end - a.env_begin;
    if (g.v == 1 && a.noise_x == nullptr && n > 0 && n <= ((int64_t)1 << 31)) {
        StepArgs b = a;
        fill_offsets(b, a, precision, action_dtype);
        b.n_stream = (uint32_t)n;
        const int blocks = (int)((n + kNarrowThreads - 1) / kNarrowThreads);
        return dispatch_step(precision, 1, action_dtype, NarrowLaunchOp{b, blocks, s});
    }
    const int64_t warp_envs = 32 * wide;
    int64_t n_stream = 0;
    // (the bulk copies of the streaming kernel want every plane and the actions 16-byte aligned at
    // env_begin: 16 envs cover the narrowest element, a uint8 action)
    if (g.v == wide && wide > 1 && a.noise_x == nullptr && n <= ((int64_t)1 << 31) &&
        a.env_begin % 16 == 0 && has_stream_kernel(precision, action_dtype))
        n_stream = n / warp_envs * warp_envs;
    if (n_stream == 0) return dispatch_step(precision, g.v, action_dtype, StepLaunchOp{a, g, s});

    StepArgs b = a;
    b.env_end = a.env_begin + n_stream;
    fill_offsets(b, a, precision, action_dtype);
    b.n_stream = (uint32_t)n_stream;
    LaunchGeom gs = g;
    const int64_t tile = (int64_t)g.threads * wide;
    gs.blocks = (int)std::max<int64_t>(1, std::min<int64_t>(g.blocks, (n_stream + tile - 1) / tile));
    cudaError_t e = dispatch_step(precision, wide, action_dtype, StreamLaunchOp{b, gs, s});
    if (e != cudaSuccess || n_stream == n) return e;
    StepArgs c = a;
    c.env_begin = a.env_begin + n_stream;
    return dispatch_step(precision, wide, action_dtype, StepLaunchOp{c, LaunchGeom{g.threads, 1, wide}, s});
}

struct StepConfigureOp {  // dynamic shared memory above the 48 KiB default limit needs an opt-in
    template <typename R, int V, int ACT> cudaError_t run() const {
        constexpr int kMaxDyn = stage_bytes<R, V>() + (int)kCellBytes;  // staging buffers + culling grid
        cudaError_t e = cudaFuncSetAttribute(step_kernel<R, V, ACT, false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(step_kernel<R, V, ACT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kMaxDyn);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(step_kernel<R, V, ACT, false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(step_kernel<R, V, ACT, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared);
        if constexpr (stream_supported<ACT, V>()) {
            auto opt_in = [&](auto* kernel) {
                if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
                if (e == cudaSuccess)
                    e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             cudaSharedmemCarveoutMaxShared);
            };
            opt_in(step_stream_kernel<R, V, ACT, false, false>);
            opt_in(step_stream_kernel<R, V, ACT, true, false>);
            opt_in(step_stream_kernel<R, V, ACT, false, true>);
            opt_in(step_stream_kernel<R, V, ACT, true, true>);
        }
        return e;
    }
};
cudaError_t configure_kernels(int precision) {
    const int wide = precision == AQUA_F64 ? Prec<double>::V : Prec<float>::V;
    for (int v : {wide, 1})
        for (int act = 0; act < 4; ++act) {
            const cudaError_t e = dispatch_step(precision, v, act, StepConfigureOp{});
            if (e != cudaSuccess) return e;
        }
    return cudaSuccess;
}

struct StepOccupancyOp {
    int threads, use_grid;
    int* out;
    template <typename R, int V, int ACT> cudaError_t run() const {
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, step_kernel<R, V, ACT, false>, threads,
                                                             stage_bytes<R, V>() + cell_bytes(use_grid));
    }
};
cudaError_t step_occupancy(int precision, int v, int action_dtype, int threads, int use_grid, int* blocks_per_sm) {
    return dispatch_step(precision, v, action_dtype, StepOccupancyOp{threads, use_grid, blocks_per_sm});
}

struct StreamOccupancyOp {
    int threads, waves, use_grid;
    int* out;
    template <typename R, int V, int ACT> cudaError_t run() const {
        if constexpr (stream_supported<ACT, V>()) {
            const size_t smem = stage_bytes<R, V>() + cell_bytes(use_grid);
            if (waves)
                return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, step_stream_kernel<R, V, ACT, true, true>,
                                                                     threads, smem);
            return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, step_stream_kernel<R, V, ACT, false, true>,
                                                                 threads, smem);
        } else {
            *out = 0;
            return cudaSuccess;
        }
    }
};
cudaError_t stream_occupancy(int precision, int action_dtype, int waves, int threads, int use_grid, int* blocks_per_sm) {
    return dispatch_step(precision, envs_per_thread(precision), action_dtype,
                         StreamOccupancyOp{threads, waves, use_grid, blocks_per_sm});
}

static inline int blocks_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

cudaError_t launch_reset(int precision, const StepArgs& a, const uint8_t* mask, cudaStream_t s) {
    const int64_t n = a.env_end - a.env_begin;
    if (n <= 0) return cudaSuccess;
    const int threads = 128;
    if (precision == AQUA_F64) reset_kernel<double><<<blocks_for(n, threads), threads, 0, s>>>(a, mask);
    else reset_kernel<float><<<blocks_for(n, threads), threads, 0, s>>>(a, mask);
    return cudaGetLastError();
}

cudaError_t launch_init_state(int precision, void* planes, int64_t ld, cudaStream_t s) {
    const int threads = 256;
    if (precision == AQUA_F64) init_state_kernel<double><<<blocks_for(ld, threads), threads, 0, s>>>(planes, ld);
    else init_state_kernel<float><<<blocks_for(ld, threads), threads, 0, s>>>(planes, ld);
    return cudaGetLastError();
}

cudaError_t launch_random_actions(const World& w, int64_t n, uint64_t step_index, void* actions,
                                  int action_dtype, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int threads = 256, blocks = blocks_for(n, threads);
    switch (action_dtype) {
        case AQUA_ACT_I32: random_actions_kernel<AQUA_ACT_I32><<<blocks, threads, 0, s>>>(w, n, step_index, actions); break;
        case AQUA_ACT_I64: random_actions_kernel<AQUA_ACT_I64><<<blocks, threads, 0, s>>>(w, n, step_index, actions); break;
        case AQUA_ACT_U8: random_actions_kernel<AQUA_ACT_U8><<<blocks, threads, 0, s>>>(w, n, step_index, actions); break;
        case AQUA_ACT_F32X2: random_actions_kernel<AQUA_ACT_F32X2><<<blocks, threads, 0, s>>>(w, n, step_index, actions); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_get_state(int precision, const void* planes, int64_t ld, int64_t n, double* out,
                             cudaStream_t s) {
    const int threads = 256;
    if (precision == AQUA_F64) get_state_kernel<double><<<blocks_for(n, threads), threads, 0, s>>>(planes, ld, n, out);
    else get_state_kernel<float><<<blocks_for(n, threads), threads, 0, s>>>(planes, ld, n, out);
    return cudaGetLastError();
}
cudaError_t launch_set_state(int precision, void* planes, int64_t ld, int64_t n, const double* in,
                             cudaStream_t s) {
    const int threads = 256;
    if (precision == AQUA_F64) set_state_kernel<double><<<blocks_for(n, threads), threads, 0, s>>>(planes, ld, n, in);
    else set_state_kernel<float><<<blocks_for(n, threads), threads, 0, s>>>(planes, ld, n, in);
    return cudaGetLastError();
}

cudaError_t launch_observe(int precision, const void* planes, int64_t ld, int64_t n, float* out, int normalized,
                           cudaStream_t s) {
    const int threads = 256;
    if (precision == AQUA_F64) observe_kernel<double><<<blocks_for(n, threads), threads, 0, s>>>(planes, ld, n, out, normalized);
    else observe_kernel<float><<<blocks_for(n, threads), threads, 0, s>>>(planes, ld, n, out, normalized);
    return cudaGetLastError();
}

// planes 0..4 (the observation, aqua.py:213) of envs [0, n) into `out` = [5][n] of the state's element
// type, by ONE thread block; then `*done = seq` (system scope).  `out` and `done` are mapped host
// memory (small-slab host step): the host thread polls `done` instead of synchronising the stream.
template <typename R>
__global__ void copy_obs_planes_kernel(const void* planes, int64_t ld, int64_t n, R* out, volatile uint32_t* done,
                                       uint32_t seq) {
    void* pl = const_cast<void*>(planes);
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
#pragma unroll
        for (int k = 0; k < 5; ++k) out[(int64_t)k * n + i] = plane<R>(pl, ld, k)[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        *done = seq;
        __threadfence_system();
    }
}

cudaError_t launch_copy_obs_planes(int precision, const void* planes, int64_t ld, int64_t n, void* out, uint32_t* done,
                                   uint32_t seq, cudaStream_t s) {
    const int threads = n >= 1024 ? 1024 : (n > 32 ? 256 : 32);
    if (precision == AQUA_F64)
        copy_obs_planes_kernel<double><<<1, threads, 0, s>>>(planes, ld, n, static_cast<double*>(out), done, seq);
    else
        copy_obs_planes_kernel<float><<<1, threads, 0, s>>>(planes, ld, n, static_cast<float*>(out), done, seq);
    return cudaGetLastError();
}

cudaError_t launch_goal_updates(int precision, const void* planes, int64_t ld, const uint8_t* flags, int64_t lo,
                                int64_t len, void* h_gx, void* h_gy, cudaStream_t s) {
    const int threads = 256;
    if (precision == AQUA_F64)
        goal_updates_kernel<double><<<blocks_for(len, threads), threads, 0, s>>>(planes, ld, flags, lo, len,
                                                                                static_cast<double*>(h_gx), static_cast<double*>(h_gy));
    else
        goal_updates_kernel<float><<<blocks_for(len, threads), threads, 0, s>>>(planes, ld, flags, lo, len,
                                                                               static_cast<float*>(h_gx), static_cast<float*>(h_gy));
    return cudaGetLastError();
}

cudaError_t launch_pack_host(int precision, const void* planes, int64_t ld, const void* reward, const uint8_t* flags,
                             int64_t lo, int64_t len, int64_t n, void* h_obs, void* h_rew, uint8_t* h_flg, int sparse_goals,
                             cudaStream_t s) {
    const int threads = 256;
    if (precision == AQUA_F64)
        pack_host_kernel<double><<<blocks_for((len + 1) / 2, threads), threads, 0, s>>>(
            planes, ld, static_cast<const double*>(reward), flags, lo, len, n, static_cast<double*>(h_obs),
            static_cast<double*>(h_rew), h_flg, sparse_goals);
    else
        pack_host_kernel<float><<<blocks_for((len + 3) / 4, threads), threads, 0, s>>>(
            planes, ld, static_cast<const float*>(reward), flags, lo, len, n, static_cast<float*>(h_obs),
            static_cast<float*>(h_rew), h_flg, sparse_goals);
    return cudaGetLastError();
}

cudaError_t launch_stats_reduce(unsigned long long* slots, unsigned int* forced, int64_t* out, int clear, cudaStream_t s) {
    stats_reduce_kernel<<<1, 256, 0, s>>>(slots, forced, out, clear);
    return cudaGetLastError();
}

}  // namespace aqua
