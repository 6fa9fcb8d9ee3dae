// aqua_kernels.cu -- the sm_100a kernels of the batched AquaGym environment step.
//
// Replaces AquaEnv.step()/reset() of the reference (gym_aqua/envs/aqua.py:100-126, 135-213) for
// N independent env instances.  One thread owns V consecutive envs (V = 4 float / 2 double) so
// that every state plane is read and written with one 16-byte access; the obstacle table sits in
// shared memory; episode statistics are folded with warp reductions into per-CTA slots.
// The step is elementwise: there is no contraction, hence no tensor-core work; the bound is HBM
// bandwidth (65 B per discrete env-step).
#include "aqua_device.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace aqua {

// ------------------------------------------------------------------------------------------
// vector access helpers
// ------------------------------------------------------------------------------------------
template <typename T, int V> struct alignas(sizeof(T) * V <= 16 ? sizeof(T) * V : 16) Pack {
    T v[V];
};
template <typename T, int V> __device__ __forceinline__ Pack<T, V> ld_pack(const T* p) {
    return *reinterpret_cast<const Pack<T, V>*>(p);
}
template <typename T, int V> __device__ __forceinline__ void st_pack(T* p, const Pack<T, V>& v) {
    *reinterpret_cast<Pack<T, V>*>(p) = v;
}
#ifndef AQUA_MEMBENCH
#define AQUA_MEMBENCH 0
#endif
// AQUA_SCHED_WARP=1: the streaming kernel runs one extra warp per CTA that does all tile claims and
// feeds the worker warps through shared-memory rings (see step_stream_kernel), so that no worker
// waits for the reply of a global atomic.  Correct (all GPU tests pass with it) and NOT faster:
// 8.335e10 vs 8.304e10 env-steps/s on one box, 3 % slower on the 64-obstacle map
// (profiles/r01L_scheduler_warp_ab.txt) -- the time a worker spends waiting for its claim is time
// the other warps of its scheduler use.  Off by default.
#ifndef AQUA_SCHED_WARP
#define AQUA_SCHED_WARP 0
#endif
constexpr int kSchedThreads = AQUA_SCHED_WARP ? 32 : 0;  // added to the CTA of the streaming kernel
constexpr int kRing = 4;                                 // tiles queued per worker warp
constexpr uint32_t kRingEmpty = 0xffffffffu, kRingDone = 0xfffffffeu;
// Stores of the streaming kernel.  AQUA_STORE_CS=1 makes them evict-first (st.global.cs).  Measured
// on one B200 box (tools/tune.py, 2^24 envs): plain 8.20e10 env-steps/s, evict-first 7.97e10 -- the
// deferred reset writes 4-byte values into lines the owner stored a few tiles earlier, and those
// should still be in L2 by then.  Hence plain stores.
#ifndef AQUA_STORE_CS
#define AQUA_STORE_CS 0
#endif
template <typename T, int V> __device__ __forceinline__ void st_pack_cs(T* p, const Pack<T, V>& v) {
    if constexpr (!AQUA_STORE_CS) {
        st_pack<T, V>(p, v);
    } else if constexpr (sizeof(T) * V == 16) {
        uint4 r;
        memcpy(&r, &v, 16);
        __stcs(reinterpret_cast<uint4*>(p), r);
    } else if constexpr (sizeof(T) * V == 8) {
        uint2 r;
        memcpy(&r, &v, 8);
        __stcs(reinterpret_cast<uint2*>(p), r);
    } else if constexpr (sizeof(T) * V == 4) {
        unsigned r;
        memcpy(&r, &v, 4);
        __stcs(reinterpret_cast<unsigned*>(p), r);
    } else {
        st_pack<T, V>(p, v);
    }
}
// read-only (non-coherent) path for the caller's action / noise buffers
template <typename T, int V> __device__ __forceinline__ Pack<T, V> ldg_pack(const T* p) {
    Pack<T, V> out;
    if constexpr (sizeof(T) * V == 16) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
        memcpy(&out, &r, 16);
    } else if constexpr (sizeof(T) * V == 32) {
        const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(p));
        const uint4 r1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
        memcpy(&out, &r0, 16);
        memcpy(reinterpret_cast<char*>(&out) + 16, &r1, 16);
    } else if constexpr (sizeof(T) * V == 8) {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
        memcpy(&out, &r, 8);
    } else if constexpr (sizeof(T) * V == 4) {
        const uint32_t r = __ldg(reinterpret_cast<const uint32_t*>(p));
        memcpy(&out, &r, 4);
    } else if constexpr (sizeof(T) * V == 2) {
        const unsigned short r = __ldg(reinterpret_cast<const unsigned short*>(p));
        memcpy(&out, &r, 2);
    } else {
        static_assert(sizeof(T) * V == 1, "unsupported pack");
        const unsigned char r = __ldg(reinterpret_cast<const unsigned char*>(p));
        memcpy(&out, &r, 1);
    }
    return out;
}

template <typename R> __device__ __forceinline__ R* plane(void* planes, int64_t ld, int p) {
    return reinterpret_cast<R*>(planes) + (int64_t)p * ld;
}
template <typename R> __device__ __forceinline__ uint32_t* plane_tw(void* planes, int64_t ld) {
    return reinterpret_cast<uint32_t*>(reinterpret_cast<R*>(planes) + (int64_t)7 * ld);
}

// ------------------------------------------------------------------------------------------
// action decode for V envs starting at absolute env index `base`
// ------------------------------------------------------------------------------------------
// The actions of a thread's V envs as 32-bit words: discrete indices (an int64 index outside 32
// bits maps to 0xffffffff, i.e. invalid) or the 2 V thrust floats.
template <int ACT, int V> struct ActVals {
    static constexpr int N = (ACT == AQUA_ACT_F32X2) ? 2 * V : V;
    uint32_t w[N];
};

template <bool SMEM, typename T, int V> __device__ __forceinline__ Pack<T, V> act_pack(const T* src) {
    if constexpr (SMEM) return *reinterpret_cast<const Pack<T, V>*>(src);
    else return ldg_pack<T, V>(src);
}

// `src` points at the thread's V actions (global memory, or its staging slot when SMEM)
template <bool SMEM, typename T, int V>
__device__ __forceinline__ void fetch_discrete(const void* src_, int nvalid, uint32_t (&a)[V]) {
    const T* src = reinterpret_cast<const T*>(src_);
    T raw[V];
    if (nvalid == V) {
        const Pack<T, V> pk = act_pack<SMEM, T, V>(src);
#pragma unroll
        for (int j = 0; j < V; ++j) raw[j] = pk.v[j];
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) raw[j] = (j < nvalid) ? __ldg(src + j) : (T)2;
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
        if constexpr (sizeof(T) == 8) {
            const unsigned long long u = (unsigned long long)raw[j];
            a[j] = (u >> 32) ? 0xffffffffu : (uint32_t)u;
        } else {
            a[j] = (uint32_t)raw[j];
        }
    }
}

template <int ACT> struct ActElem { using T = int32_t; static constexpr int kPerEnv = 1; };
template <> struct ActElem<AQUA_ACT_I64> { using T = long long; static constexpr int kPerEnv = 1; };
template <> struct ActElem<AQUA_ACT_U8> { using T = uint8_t; static constexpr int kPerEnv = 1; };
template <> struct ActElem<AQUA_ACT_F32X2> { using T = float; static constexpr int kPerEnv = 2; };
template <int ACT, int V> constexpr int act_bytes() { return (int)sizeof(typename ActElem<ACT>::T) * ActElem<ACT>::kPerEnv * V; }

// address of the actions of the V envs starting at absolute env index `base`
template <int ACT> __device__ __forceinline__ const void* act_ptr(const void* actions, int64_t base) {
    return reinterpret_cast<const typename ActElem<ACT>::T*>(actions) + base * ActElem<ACT>::kPerEnv;
}

// loads the actions of V envs from `src` (padding lanes: valid values)
template <int ACT, int V, bool SMEM = false>
__device__ __forceinline__ ActVals<ACT, V> fetch_actions(const void* src_, int nvalid) {
    ActVals<ACT, V> out;
    if constexpr (ACT == AQUA_ACT_F32X2) {
        const float* src = reinterpret_cast<const float*>(src_);
        if (nvalid == V) {
            const Pack<float, 2 * V> pk = act_pack<SMEM, float, 2 * V>(src);
#pragma unroll
            for (int j = 0; j < 2 * V; ++j) out.w[j] = __float_as_uint(pk.v[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 2 * V; ++j) out.w[j] = __float_as_uint((j < 2 * nvalid) ? __ldg(src + j) : 0.5f);
        }
    } else {
        fetch_discrete<SMEM, typename ActElem<ACT>::T, V>(src_, nvalid, out.w);
    }
    return out;
}

// the same from the raw bytes of the thread's actions (as staged in shared memory)
template <int ACT, int V, int NW>
__device__ __forceinline__ ActVals<ACT, V> actions_from_raw(const uint32_t (&raw)[NW]) {
    ActVals<ACT, V> out;
    if constexpr (ACT == AQUA_ACT_F32X2 || ACT == AQUA_ACT_I32) {
#pragma unroll
        for (int j = 0; j < ActVals<ACT, V>::N; ++j) out.w[j] = raw[j];
    } else if constexpr (ACT == AQUA_ACT_I64) {
#pragma unroll
        for (int j = 0; j < V; ++j) out.w[j] = raw[2 * j + 1] ? 0xffffffffu : raw[2 * j];
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) out.w[j] = (raw[0] >> (8 * j)) & 0xffu;
    }
    return out;
}

template <typename R, int ACT, int V>
__device__ __forceinline__ void decode_drives(const ActVals<ACT, V>& a, Drive<R> (&dr)[V], uint32_t& bad) {
    if constexpr (ACT == AQUA_ACT_F32X2) {
#pragma unroll
        for (int j = 0; j < V; ++j)  // padding lanes hold 0.5: never counted
            dr[j] = drive_continuous<R>(__uint_as_float(a.w[2 * j]), __uint_as_float(a.w[2 * j + 1]), bad);
    } else {
        uint32_t worst = 0u;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            dr[j] = drive_from_index<R>(a.w[j]);
            worst = max(worst, a.w[j]);
        }
        if (worst > 2u) {  // rare: an index outside {0, 1, 2} (padding lanes hold 2)
#pragma unroll
            for (int j = 0; j < V; ++j) bad += (a.w[j] > 2u) ? 1u : 0u;
        }
    }
}

// ------------------------------------------------------------------------------------------
// asynchronous global -> shared copies (LDGSTS): each thread stages the 8 plane slices of ITS
// next tile in its own shared-memory slots while it computes the current tile, so no barrier is
// involved and the memory pipe stays busy through the compute phase
// ------------------------------------------------------------------------------------------
template <int BYTES> __device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
    else if constexpr (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
    else {
        static_assert(BYTES == 4, "cp.async moves 4, 8 or 16 bytes");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
#ifndef AQUA_STAGES
#define AQUA_STAGES 2
#endif
constexpr int kStages = AQUA_STAGES;                        // 2: double buffer; 1: refill after the reads
constexpr int kStageSlots = AQUA_N_PLANES + 2;              // 8 state planes + up to two slots of actions
// one slot = the bytes one thread moves per plane (16 at V = 4 float / V = 2 double); slots are
// kMaxThreads apart whatever blockDim is, so that every slot offset is an immediate
template <typename R, int V> constexpr int stage_slot() { return (int)sizeof(R) * V; }
template <typename R, int V> constexpr int stage_pitch() { return kMaxThreads * stage_slot<R, V>(); }
template <typename R, int V> constexpr int stage_bytes() { return V > 1 ? kStages * kStageSlots * stage_pitch<R, V>() : 0; }

// ------------------------------------------------------------------------------------------
// per-thread episode statistics and their reduction
// ------------------------------------------------------------------------------------------
// The flag bytes of the V envs of a thread arrive packed in one word (byte j = env j).  A finished
// env carries exactly one of COLLIDED (2), TIME (4), SUCCESS (8) next to DONE (aqua.py:194-211), so
// three masked byte sums (one LOP3 + one IDP.4A each) count everything: the accumulators hold
// 2 x collided, 4 x timeouts, 8 x successes, and done is their sum.
struct ThreadStats {
    uint32_t coll2 = 0, time4 = 0, succ8 = 0;
    uint32_t length = 0, steps = 0, bad = 0;
    long long reward_fx = 0;  // sum of step rewards, fixed point 2^-20

    // w: flags of the thread's envs, byte j = env j (bytes of padding envs must be 0)
    __device__ __forceinline__ void add_flags(uint32_t w, int nvalid) {
        coll2 = __dp4a(w & 0x02020202u, 0x01010101u, coll2);
        time4 = __dp4a(w & 0x04040404u, 0x01010101u, time4);
        succ8 = __dp4a(w & 0x08080808u, 0x01010101u, succ8);
        steps += (uint32_t)nvalid;
    }
    // without auto-reset the episode length is taken here; with it, phase B adds it (the reset
    // pass reads the terminal time word anyway)
    __device__ __forceinline__ void add_length(uint32_t flags, uint32_t time_after) {
        length += (flags & 1u) ? time_after : 0u;
    }
    template <typename R> __device__ __forceinline__ void add_env(R reward) {
        int fx;
        if constexpr (sizeof(R) == 8) fx = __double2int_rn(reward * R(AQUA_REWARD_FIXED_POINT));
        else fx = __float2int_rn(reward * R(AQUA_REWARD_FIXED_POINT));
        reward_fx += fx;
    }
};

// warp reduce (REDUX) -> shared slots -> one global atomic per counter per CTA
__device__ __forceinline__ void flush_stats(ThreadStats& t, unsigned long long* s_stats,
                                            unsigned long long* g_slots) {
    const unsigned full = 0xffffffffu;
    const uint32_t c = __reduce_add_sync(full, t.coll2 >> 1);
    const uint32_t o = __reduce_add_sync(full, t.time4 >> 2);
    const uint32_t s = __reduce_add_sync(full, t.succ8 >> 3);
    const uint32_t d = c + o + s;
    const uint32_t l = __reduce_add_sync(full, t.length);
    const uint32_t n = __reduce_add_sync(full, t.steps);
    const uint32_t b = __reduce_add_sync(full, t.bad);
    long long r = t.reward_fx;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) r += __shfl_xor_sync(full, r, off);
    if ((threadIdx.x & 31) == 0) {
        if (d) atomicAdd(&s_stats[AQUA_STAT_EPISODES], (unsigned long long)d);
        if (c) atomicAdd(&s_stats[AQUA_STAT_COLLIDED], (unsigned long long)c);
        if (o) atomicAdd(&s_stats[AQUA_STAT_TIMEOUT], (unsigned long long)o);
        if (s) atomicAdd(&s_stats[AQUA_STAT_SUCCESS], (unsigned long long)s);
        if (l) atomicAdd(&s_stats[AQUA_STAT_LENGTH_SUM], (unsigned long long)l);
        if (n) atomicAdd(&s_stats[AQUA_STAT_STEPS], (unsigned long long)n);
        if (b) atomicAdd(&s_stats[AQUA_STAT_BAD_ACTIONS], (unsigned long long)b);
        if (r) atomicAdd(&s_stats[AQUA_STAT_REWARD_SUM], (unsigned long long)r);
    }
    __syncthreads();
    if (threadIdx.x < AQUA_N_STATS) {
        const unsigned long long v = s_stats[threadIdx.x];
        if (v) atomicAdd(&g_slots[(blockIdx.x % kStatSlots) * AQUA_N_STATS + threadIdx.x], v);
    }
}

// dynamic shared memory a launch needs for the culling grid (dense maps only), placed behind the
// staging buffers
constexpr size_t kCellBytes = sizeof(unsigned long long) * kGridCells;
__host__ __device__ inline size_t cell_bytes(int use_grid) { return use_grid ? kCellBytes : 0; }

// Copies the obstacle table of the launch (global ObstacleTable<R>) into shared memory: the boxes
// into `s_box`, the cells of a dense map into `s_cell`; returns the view the kernels read.  The
// caller synchronises the CTA.
template <typename R>
__device__ __forceinline__ ObstacleView<R> stage_table(ObstacleBox<R>* s_box, unsigned char* s_cell, const void* g_tab,
                                                       int k_obst, bool use_grid) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(g_tab);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_box);
    constexpr int kWordsPerBox = sizeof(ObstacleBox<R>) / 4;
    for (int i = threadIdx.x; i < k_obst * kWordsPerBox; i += blockDim.x) dst[i] = __ldg(src + i);
    if (use_grid) {  // dense map: the culling grid as well
        constexpr int kCellOff = sizeof(ObstacleBox<R>) * AQUA_MAX_OBSTACLES / 4;
        uint32_t* cdst = reinterpret_cast<uint32_t*>(s_cell);
        for (int i = threadIdx.x; i < kGridCells * 2; i += blockDim.x) cdst[i] = __ldg(src + kCellOff + i);
    }
    return ObstacleView<R>{s_box, use_grid ? reinterpret_cast<const unsigned long long*>(s_cell) : nullptr};
}

// ------------------------------------------------------------------------------------------
// Phase B of a step, per warp: the warp resets the envs its lanes finished, kCoopLanes lanes
// per env, 32 / kCoopLanes envs per pass (no CTA barrier).  Finished envs are first queued per
// warp (env index relative to env_begin); the step kernel drains the queue only in full passes
// and flushes the remainder before it exits, so that every pass runs with all groups busy.
// The owner has already stored the terminal state and time word; the fresh episode goes
// straight to the state planes (aqua.py:100-126).  Both functions are warp-collective.
// ------------------------------------------------------------------------------------------
constexpr int kResetGroups = 32 / kCoopLanes;
template <int V> struct WarpResetQueue {
    uint32_t rel[32 * V + kResetGroups];  // env index relative to env_begin
    uint32_t tw[32 * V + kResetGroups];   // its terminal time word (episode, length)
};

// fw: flags of the lane's V envs, one byte each (0 for padding / inactive lanes); rel_base: index
// of the lane's first env relative to env_begin; tw: the envs' time words after the step.
// Returns the new queue length.
template <int V>
__device__ __forceinline__ int reset_queue_push(WarpResetQueue<V>& wq, int n, uint32_t fw, uint32_t rel_base,
                                                const uint32_t (&tw)[V]) {
    const unsigned full = 0xffffffffu;
    const unsigned lt = (1u << (threadIdx.x & 31u)) - 1u;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const bool d = ((fw >> (8 * j)) & 1u) != 0u;
        const unsigned mj = __ballot_sync(full, d);
        if (d) {
            const int at = n + __popc(mj & lt);
            wq.rel[at] = rel_base + (uint32_t)j;
            wq.tw[at] = tw[j];
        }
        n += __popc(mj);
    }
    return n;
}

// Resets queued envs in passes of kResetGroups, from the tail, while at least `min_batch` are
// queued (min_batch = kResetGroups: full passes only; 1: flush).  Returns the new queue length.
template <typename R, int V>
__device__ __forceinline__ int reset_queue_drain(const StepArgs& p, const ObstacleView<R>& tab,
                                                 const WarpResetQueue<V>& wq, int n, int min_batch, uint32_t& length_sum) {
    const unsigned lane = threadIdx.x & 31u;
    const WorldFlags wf{p.w.waves != 0, p.w.random_boat != 0, p.w.random_goal != 0};
    uint32_t* const ptw = plane_tw<R>(p.planes, p.ld);
    const int grp = (int)(lane / kCoopLanes);
    const bool leader = (lane % kCoopLanes) == 0u;
    __syncwarp();  // queue entries are visible to the whole warp
    while (n >= min_batch && n > 0) {
        const int take = min(n, kResetGroups);
        const int first = n - take;
        const bool valid = grp < take;
        const int64_t i = p.env_begin + (int64_t)(valid ? wq.rel[first + grp] : 0u);
        const uint32_t tw = valid ? wq.tw[first + grp] : 0u;
        const uint32_t ep = ((tw >> AQUA_TIME_BITS) + 1u) & kEpisodeMask;
        const uint32_t env_id = (uint32_t)(p.w.env_id_base + (uint64_t)i);
        const CoopResetResult rr = coop_reset<R>(p.w.keys, tab, p.w.tab64, p.w.k_circ, p.w.k_obst,
                                                 p.w.reset_band, wf, valid, env_id, ep);
        if (valid && leader) {
            length_sum += tw & kTimeMask;  // AQUA_STAT_LENGTH_SUM: the finished episode's length
            plane<R>(p.planes, p.ld, AQUA_PLANE_X)[i] = (R)rr.x;
            plane<R>(p.planes, p.ld, AQUA_PLANE_Y)[i] = (R)rr.y;
            plane<R>(p.planes, p.ld, AQUA_PLANE_THETA)[i] = (R)rr.th;
            plane<R>(p.planes, p.ld, AQUA_PLANE_GOAL_X)[i] = (R)rr.gx;
            plane<R>(p.planes, p.ld, AQUA_PLANE_GOAL_Y)[i] = (R)rr.gy;
            plane<R>(p.planes, p.ld, AQUA_PLANE_WAVE_X)[i] = (R)rr.wx;
            plane<R>(p.planes, p.ld, AQUA_PLANE_WAVE_Y)[i] = (R)rr.wy;
            ptw[i] = ep << AQUA_TIME_BITS;
        }
        n = first;
    }
    __syncwarp();
    return n;
}

// the V flag bytes as one word (byte j = env j)
template <int V> __device__ __forceinline__ uint32_t flags_word(const Pack<uint8_t, V>& f) {
    uint32_t w = 0;
    memcpy(&w, &f, V);
    return w;
}

// ------------------------------------------------------------------------------------------
// K1: env.step(action) for a range of envs -- aqua.py:135-213 (+ reset aqua.py:100-126 when an
// env finishes and autoreset is on)
// ------------------------------------------------------------------------------------------
extern __shared__ __align__(16) unsigned char s_stage[];

template <typename R, int V, int ACT, bool INJECT>
__global__ void __launch_bounds__(kMaxThreads, V == 1 ? 4 : kMinBlocks) step_kernel(const __grid_constant__ StepArgs p) {
    __shared__ ObstacleBox<R> s_box[AQUA_MAX_OBSTACLES];
    __shared__ unsigned long long s_stats[AQUA_N_STATS];
    __shared__ WarpResetQueue<V> s_wq[kMaxThreads / 32];
    const ObstacleView<R> tab = stage_table<R>(s_box, s_stage + stage_bytes<R, V>(), p.w.tab, p.w.k_obst, p.w.use_grid != 0);
    if (threadIdx.x < AQUA_N_STATS) s_stats[threadIdx.x] = 0ull;
    __syncthreads();

    const int64_t ld = p.ld;
    R* const pl0 = reinterpret_cast<R*>(p.planes);
    uint32_t* const ptw_ = plane_tw<R>(p.planes, ld);
    const bool waves = p.w.waves != 0;
    const bool autoreset = p.w.autoreset != 0;
    const int k_obst = p.w.k_obst, k_circ = p.w.k_circ;
    const int64_t tile_envs = (int64_t)blockDim.x * V;
    const int64_t stride = (int64_t)gridDim.x * tile_envs;
    const int64_t warp_off = (int64_t)(threadIdx.x & ~31u) * V;  // first env of this warp within the tile

    // this thread's staging slots: [stage][plane][thread] x 16 bytes (conflict-free 16-byte LDS)
    constexpr int kSlot = stage_slot<R, V>(), kPitch = stage_pitch<R, V>();
    unsigned char* const my_stage = s_stage + (size_t)threadIdx.x * kSlot;
    auto slot = [&](int stage, int pl) { return my_stage + (size_t)stage * (kStageSlots * kPitch) + pl * kPitch; };
    // actions ride along when their per-thread footprint is a legal cp.async size
    constexpr int kActBytes = act_bytes<ACT, V>();
    constexpr bool kActStaged = kActBytes >= 4;                       // 4, 8, 16 or 32
    constexpr int kActPiece = kActBytes < kSlot ? kActBytes : kSlot;  // bytes per action slot
    constexpr int kActPieces = kActBytes >= 4 ? kActBytes / kActPiece : 0;
    static_assert(kActPieces <= 2, "the actions fit two slots");
    // narrow threads (V = 1) serve small, latency-bound slabs with one tile per thread: nothing to
    // pipeline, so they load directly and the kernel is launched without the staging buffer
    constexpr bool kStaging = V > 1;
    auto prefetch = [&](int stage, int64_t b) {  // the 8 plane slices + actions of the V envs starting at b
        const R* g = pl0 + b;  // running pointer: one 64-bit add per plane
#pragma unroll
        for (int pl = 0; pl < 7; ++pl, g += ld) cp_async<kSlot>(slot(stage, pl), g);
        cp_async<4 * V>(slot(stage, AQUA_PLANE_TIME), ptw_ + b);  // uint32 plane: its own index scale
        if constexpr (kActStaged) {
            const char* a = reinterpret_cast<const char*>(act_ptr<ACT>(p.actions, b));
#pragma unroll
            for (int q = 0; q < kActPieces; ++q) cp_async<kActPiece>(slot(stage, AQUA_N_PLANES + q), a + q * kActPiece);
        }
    };
    auto staged_actions = [&](int stage) {
        constexpr int kWords = kActBytes >= 4 ? kActBytes / 4 : 1;
        constexpr int kPieceWords = kActPiece >= 4 ? kActPiece / 4 : 1;
        uint32_t raw[kWords] = {};
#pragma unroll
        for (int q = 0; q < kActPieces; ++q) {
            const Pack<uint32_t, kPieceWords> pk =
                *reinterpret_cast<const Pack<uint32_t, kPieceWords>*>(slot(stage, AQUA_N_PLANES + q));
#pragma unroll
            for (int i = 0; i < kPieceWords; ++i) raw[q * kPieceWords + i] = pk.v[i];
        }
        return actions_from_raw<ACT, V, kWords>(raw);
    };

    ThreadStats st;
    WarpResetQueue<V>& wq = s_wq[threadIdx.x >> 5];
    int n_queued = 0;  // warp-uniform

    int64_t tile_base = p.env_begin + (int64_t)blockIdx.x * tile_envs;
    // prologue: stage the first tile (full threads only; a partial tail thread loads directly)
    bool staged = false;
    ActVals<ACT, V> act_next;  // only used when the actions cannot be staged (2-byte footprint)
    {
        const int64_t b0 = tile_base + (int64_t)threadIdx.x * V;
        if (kStaging && b0 + V <= p.env_end) {
            prefetch(0, b0);
            if constexpr (!kActStaged) act_next = fetch_actions<ACT, V>(act_ptr<ACT>(p.actions, b0), V);
            staged = true;
        }
        cp_async_commit();
    }
    // the loop is uniform over each warp (phase B is warp-collective)
    for (int it = 0; tile_base + warp_off < p.env_end; tile_base += stride, ++it) {
        const int64_t base = tile_base + (int64_t)threadIdx.x * V;
        const int nvalid = (int)max((int64_t)0, min((int64_t)V, p.env_end - base));
        // ---- stage the next tile of this thread while this one is computed ----
        const int64_t next_base = base + stride;
        const bool next_staged = kStaging && next_base + V <= p.env_end;
        ActVals<ACT, V> act_cur;
        if constexpr (!kActStaged) act_cur = act_next;
        const int sg = (kStages == 2) ? (it & 1) : 0;      // the stage holding this tile
        if constexpr (kStages == 2) {
            if (next_staged) {
                prefetch((it + 1) & 1, next_base);
                if constexpr (!kActStaged) act_next = fetch_actions<ACT, V>(act_ptr<ACT>(p.actions, next_base), V);
            }
            cp_async_commit();
            cp_async_wait<1>();  // everything but the newest group has landed: this tile's slices
        } else {
            cp_async_wait<0>();
        }
        if (staged) {
            if constexpr (kActStaged) act_cur = staged_actions(sg);
        } else {
            act_cur = fetch_actions<ACT, V>(act_ptr<ACT>(p.actions, base), nvalid);
        }
        uint32_t fw = 0u;
        Pack<R, V> vx, vy, vth, vgx, vgy, vwx, vwy;
        Pack<uint32_t, V> vtw, otw;
        if (staged) {
            vx = *reinterpret_cast<const Pack<R, V>*>(slot(sg, AQUA_PLANE_X));
            vy = *reinterpret_cast<const Pack<R, V>*>(slot(sg, AQUA_PLANE_Y));
            vth = *reinterpret_cast<const Pack<R, V>*>(slot(sg, AQUA_PLANE_THETA));
            vgx = *reinterpret_cast<const Pack<R, V>*>(slot(sg, AQUA_PLANE_GOAL_X));
            vgy = *reinterpret_cast<const Pack<R, V>*>(slot(sg, AQUA_PLANE_GOAL_Y));
            vwx = *reinterpret_cast<const Pack<R, V>*>(slot(sg, AQUA_PLANE_WAVE_X));
            vwy = *reinterpret_cast<const Pack<R, V>*>(slot(sg, AQUA_PLANE_WAVE_Y));
            vtw = *reinterpret_cast<const Pack<uint32_t, V>*>(slot(sg, AQUA_PLANE_TIME));
        } else if (nvalid > 0) {  // partial tail: the planes are padded to AQUA_TILE, full-width loads are in bounds
            vx = ld_pack<R, V>(pl0 + (int64_t)AQUA_PLANE_X * ld + base);
            vy = ld_pack<R, V>(pl0 + (int64_t)AQUA_PLANE_Y * ld + base);
            vth = ld_pack<R, V>(pl0 + (int64_t)AQUA_PLANE_THETA * ld + base);
            vgx = ld_pack<R, V>(pl0 + (int64_t)AQUA_PLANE_GOAL_X * ld + base);
            vgy = ld_pack<R, V>(pl0 + (int64_t)AQUA_PLANE_GOAL_Y * ld + base);
            vwx = ld_pack<R, V>(pl0 + (int64_t)AQUA_PLANE_WAVE_X * ld + base);
            vwy = ld_pack<R, V>(pl0 + (int64_t)AQUA_PLANE_WAVE_Y * ld + base);
            vtw = ld_pack<uint32_t, V>(ptw_ + base);
        }
        if constexpr (kStages == 1) {  // single buffer: refill it once this tile's slices are in registers
            if (next_staged) {
                prefetch(0, next_base);
                if constexpr (!kActStaged) act_next = fetch_actions<ACT, V>(act_ptr<ACT>(p.actions, next_base), V);
            }
            cp_async_commit();
        }
        if (nvalid > 0) {
            Drive<R> dr[V];
            decode_drives<R, ACT, V>(act_cur, dr, st.bad);

            EnvState<R> e[V];
            MoveTmp<R> m[V];
            R pxn[V], pyn[V], clr[V];
#pragma unroll
            for (int j = 0; j < V; ++j) {
                e[j].x = vx.v[j]; e[j].y = vy.v[j]; e[j].th = vth.v[j];
                e[j].wx = vwx.v[j]; e[j].wy = vwy.v[j];
                e[j].gx = vgx.v[j]; e[j].gy = vgy.v[j];
                e[j].tw = vtw.v[j];
                float nx = 0.f, ny = 0.f;
                if constexpr (INJECT) {
                    if (j < nvalid) {
                        nx = __ldg(p.noise_x + base + j);
                        ny = __ldg(p.noise_y + base + j);
                    }
                } else {
                    const uint32_t env_id = (uint32_t)(p.w.env_id_base + (uint64_t)(base + j));
                    if (waves) draw_wave_noise(p.w.keys, env_id, e[j].tw, nx, ny);
                }
                move_env<R>(e[j], dr[j], (R)nx, (R)ny, waves, m[j]);
                pxn[j] = e[j].x;
                pyn[j] = e[j].y;
                clr[j] = R(kClearFar);
            }
            obstacle_clearance_any<R, V>(tab, k_circ, k_obst, pxn, pyn, clr);

            Pack<R, V> vrew, ox, oy, oth, owx, owy;
            Pack<uint8_t, V> vfl;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                R reward;
                const uint32_t f = finish_env<R>(e[j], m[j], clr[j], reward);
                vrew.v[j] = reward;
                vfl.v[j] = (uint8_t)f;
                ox.v[j] = e[j].x; oy.v[j] = e[j].y; oth.v[j] = e[j].th;
                owx.v[j] = e[j].wx; owy.v[j] = e[j].wy; otw.v[j] = e[j].tw;
            }
            if (nvalid != V) {  // the one partial thread of a launch: padding envs count for nothing
#pragma unroll
                for (int j = 0; j < V; ++j)
                    if (j >= nvalid) { vfl.v[j] = 0; vrew.v[j] = R(0); }
            }
            fw = flags_word<V>(vfl);
            st.add_flags(fw, nvalid);
#pragma unroll
            for (int j = 0; j < V; ++j) st.add_env<R>(vrew.v[j]);
            if (!autoreset) {
#pragma unroll
                for (int j = 0; j < V; ++j) st.add_length(vfl.v[j], otw.v[j] & kTimeMask);
            }

            // ---- stores (the goal only changes at reset: phase B writes it) ----
            R* g = pl0 + base;  // plane x; running pointer
            st_pack<R, V>(g, ox);
            st_pack<R, V>(g += ld, oy);
            st_pack<R, V>(g += ld, oth);
            st_pack<R, V>(g += 3 * ld, owx);  // skip the two goal planes
            st_pack<R, V>(g += ld, owy);
            st_pack<uint32_t, V>(ptw_ + base, otw);
            R* const rew = reinterpret_cast<R*>(p.reward) + base;
            uint8_t* const fl = p.flags + base;
            if (nvalid == V) {
                st_pack<R, V>(rew, vrew);
                st_pack<uint8_t, V>(fl, vfl);
            } else {
#pragma unroll
                for (int j = 0; j < V; ++j)
                    if (j < nvalid) { rew[j] = vrew.v[j]; fl[j] = vfl.v[j]; }
            }
        }
        staged = next_staged;
        if (autoreset && __any_sync(0xffffffffu, fw != 0u)) {
            n_queued = reset_queue_push<V>(wq, n_queued, fw, (uint32_t)(base - p.env_begin), otw.v);
            if (n_queued >= kResetGroups)
                n_queued = reset_queue_drain<R, V>(p, tab, wq, n_queued, kResetGroups, st.length);
        }
    }
    cp_async_wait<0>();
    if (autoreset) reset_queue_drain<R, V>(p, tab, wq, n_queued, 1, st.length);  // flush the remainder
    __syncthreads();
    flush_stats(st, s_stats, p.stats);
}

// ------------------------------------------------------------------------------------------
// K1r: `steps` env-steps per launch under the synthetic random policy (ACTION stream); the
// state stays in registers between steps, per-step reward/flags are folded into the statistics
// ------------------------------------------------------------------------------------------
template <typename R, int V, bool CONTINUOUS>
__global__ void __launch_bounds__(kMaxThreads, V == 1 ? 4 : kMinBlocks) rollout_kernel(const __grid_constant__ StepArgs p) {
    __shared__ ObstacleBox<R> s_box[AQUA_MAX_OBSTACLES];
    __shared__ unsigned long long s_stats[AQUA_N_STATS];
    __shared__ WarpResetQueue<V> s_wq[kMaxThreads / 32];
    const ObstacleView<R> tab = stage_table<R>(s_box, s_stage, p.w.tab, p.w.k_obst, p.w.use_grid != 0);  // no staging buffers here
    if (threadIdx.x < AQUA_N_STATS) s_stats[threadIdx.x] = 0ull;
    __syncthreads();

    const int64_t ld = p.ld;
    const bool waves = p.w.waves != 0;
    const int k_obst = p.w.k_obst, k_circ = p.w.k_circ;
    const int64_t tile_envs = (int64_t)blockDim.x * V;
    const int64_t warp_off = (int64_t)(threadIdx.x & ~31u) * V;
    uint32_t* const ptw_ = plane_tw<R>(p.planes, ld);

    ThreadStats st;
    for (int64_t tile_base = p.env_begin + (int64_t)blockIdx.x * tile_envs; tile_base + warp_off < p.env_end;
         tile_base += (int64_t)gridDim.x * tile_envs) {
        const int64_t base = tile_base + (int64_t)threadIdx.x * V;
        const int nvalid = (int)max((int64_t)0, min((int64_t)V, p.env_end - base));
        EnvState<R> e[V];
        if (nvalid > 0) {
            const Pack<R, V> vx = ld_pack<R, V>(plane<R>(p.planes, ld, AQUA_PLANE_X) + base);
            const Pack<R, V> vy = ld_pack<R, V>(plane<R>(p.planes, ld, AQUA_PLANE_Y) + base);
            const Pack<R, V> vth = ld_pack<R, V>(plane<R>(p.planes, ld, AQUA_PLANE_THETA) + base);
            const Pack<R, V> vwx = ld_pack<R, V>(plane<R>(p.planes, ld, AQUA_PLANE_WAVE_X) + base);
            const Pack<R, V> vwy = ld_pack<R, V>(plane<R>(p.planes, ld, AQUA_PLANE_WAVE_Y) + base);
            const Pack<R, V> vgx = ld_pack<R, V>(plane<R>(p.planes, ld, AQUA_PLANE_GOAL_X) + base);
            const Pack<R, V> vgy = ld_pack<R, V>(plane<R>(p.planes, ld, AQUA_PLANE_GOAL_Y) + base);
            const Pack<uint32_t, V> vtw = ld_pack<uint32_t, V>(ptw_ + base);
#pragma unroll
            for (int j = 0; j < V; ++j) {
                e[j].x = vx.v[j]; e[j].y = vy.v[j]; e[j].th = vth.v[j];
                e[j].wx = vwx.v[j]; e[j].wy = vwy.v[j];
                e[j].gx = vgx.v[j]; e[j].gy = vgy.v[j];
                e[j].tw = vtw.v[j];
            }
        }
        for (int64_t k = 0; k < p.steps; ++k) {
            uint32_t fw = 0u;
            if (nvalid > 0) {
                const uint64_t step = p.step0 + (uint64_t)k;
                MoveTmp<R> m[V];
                R pxn[V], pyn[V], clr[V];
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const uint32_t env_id = (uint32_t)(p.w.env_id_base + (uint64_t)(base + j));
                    const uint2 ra = draw_action_words(p.w.keys, env_id, step);
                    Drive<R> dr;
                    if constexpr (CONTINUOUS) {
                        uint32_t bad = 0;
                        dr = drive_continuous<R>(draw_thrust(ra.x), draw_thrust(ra.y), bad);
                    } else {
                        dr = drive_from_index<R>(__umulhi(ra.x, 3u));
                    }
                    float nx = 0.f, ny = 0.f;
                    if (waves) draw_wave_noise(p.w.keys, env_id, e[j].tw, nx, ny);
                    move_env<R>(e[j], dr, (R)nx, (R)ny, waves, m[j]);
                    pxn[j] = e[j].x;
                    pyn[j] = e[j].y;
                    clr[j] = R(kClearFar);
                }
                obstacle_clearance_any<R, V>(tab, k_circ, k_obst, pxn, pyn, clr);
                Pack<uint8_t, V> vfl;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    R reward;
                    const uint32_t f = finish_env<R>(e[j], m[j], clr[j], reward);
                    vfl.v[j] = (uint8_t)((j < nvalid) ? f : 0u);
                    st.add_env<R>((j < nvalid) ? reward : R(0));
                }
                fw = flags_word<V>(vfl);
                st.add_flags(fw, nvalid);
            }
            if (__any_sync(0xffffffffu, fw != 0u)) {
                WarpResetQueue<V>& wq = s_wq[threadIdx.x >> 5];
                uint32_t tws[V];
#pragma unroll
                for (int j = 0; j < V; ++j) tws[j] = (nvalid > 0) ? e[j].tw : 0u;
                const int n_queued = reset_queue_push<V>(wq, 0, fw, (uint32_t)(base - p.env_begin), tws);
                reset_queue_drain<R, V>(p, tab, wq, n_queued, 1, st.length);  // the next step needs the fresh state
                if (fw != 0u) {  // pick the fresh episode up from the planes
#pragma unroll
                    for (int j = 0; j < V; ++j)
                        if ((fw >> (8 * j)) & 1u) {
                            const int64_t i = base + j;
                            e[j].x = plane<R>(p.planes, ld, AQUA_PLANE_X)[i];
                            e[j].y = plane<R>(p.planes, ld, AQUA_PLANE_Y)[i];
                            e[j].th = plane<R>(p.planes, ld, AQUA_PLANE_THETA)[i];
                            e[j].gx = plane<R>(p.planes, ld, AQUA_PLANE_GOAL_X)[i];
                            e[j].gy = plane<R>(p.planes, ld, AQUA_PLANE_GOAL_Y)[i];
                            e[j].wx = plane<R>(p.planes, ld, AQUA_PLANE_WAVE_X)[i];
                            e[j].wy = plane<R>(p.planes, ld, AQUA_PLANE_WAVE_Y)[i];
                            e[j].tw = ptw_[i];
                        }
                }
            }
        }
        if (nvalid > 0) {
            Pack<R, V> ox, oy, oth, owx, owy, ogx, ogy;
            Pack<uint32_t, V> otw;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                ox.v[j] = e[j].x; oy.v[j] = e[j].y; oth.v[j] = e[j].th;
                owx.v[j] = e[j].wx; owy.v[j] = e[j].wy;
                ogx.v[j] = e[j].gx; ogy.v[j] = e[j].gy; otw.v[j] = e[j].tw;
            }
            st_pack<R, V>(plane<R>(p.planes, ld, AQUA_PLANE_X) + base, ox);
            st_pack<R, V>(plane<R>(p.planes, ld, AQUA_PLANE_Y) + base, oy);
            st_pack<R, V>(plane<R>(p.planes, ld, AQUA_PLANE_THETA) + base, oth);
            st_pack<R, V>(plane<R>(p.planes, ld, AQUA_PLANE_WAVE_X) + base, owx);
            st_pack<R, V>(plane<R>(p.planes, ld, AQUA_PLANE_WAVE_Y) + base, owy);
            st_pack<R, V>(plane<R>(p.planes, ld, AQUA_PLANE_GOAL_X) + base, ogx);
            st_pack<R, V>(plane<R>(p.planes, ld, AQUA_PLANE_GOAL_Y) + base, ogy);
            st_pack<uint32_t, V>(ptw_ + base, otw);
        }
    }
    __syncthreads();
    flush_stats(st, s_stats, p.stats);
}

// ------------------------------------------------------------------------------------------
// K1s: the streaming form of K1 -- what large slabs run.  Same arithmetic as step_kernel (it
// inlines the same move_env / finish_env / obstacle_clearance, hence the same results bit for
// bit), without everything a full warp of full threads does not need:
//   * the launch covers whole warps only (n_stream is a multiple of 32 V; launch_step() gives the
//     few remaining envs to step_kernel), so there are no tail guards;
//   * one 32-bit env index, base pointers of the planes in the constant bank (one IMAD.WIDE per
//     address), 32-bit shared-memory addresses for the staging slots;
//   * `waves` is a template parameter, and the Philox2x32 chains of the thread's V envs are
//     advanced round by round side by side (V independent multiply chains in flight);
//   * the discrete drive is one 16-byte shared-memory read of a 3-entry table.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BYTES> __device__ __forceinline__ void cp_async_s(uint32_t d, const void* gsrc) {
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
    else if constexpr (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
    else {
        static_assert(BYTES == 4, "cp.async moves 4, 8 or 16 bytes");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
    }
}
template <int BYTES> __device__ __forceinline__ void lds_words(uint32_t a, uint32_t (&w)[BYTES / 4]) {
    if constexpr (BYTES == 16)
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(a));
    else if constexpr (BYTES == 8)
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(w[0]), "=r"(w[1]) : "r"(a));
    else {
        static_assert(BYTES == 4, "staging slots hold 4, 8 or 16 bytes");
        asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(w[0]) : "r"(a));
    }
}
template <typename T, int V> __device__ __forceinline__ Pack<T, V> lds_pack(uint32_t a) {
    uint32_t w[sizeof(T) * V / 4];
    lds_words<(int)sizeof(T) * V>(a, w);
    Pack<T, V> out;
    memcpy(&out, w, sizeof(T) * V);
    return out;
}

// wave increments of V consecutive envs (ids env_id .. env_id + V - 1): draw_wave_noise() for
// each, the ten rounds of the V counters interleaved
template <int V>
__device__ __forceinline__ void draw_wave_noise_v(const PhiloxKeys& keys, uint32_t env_id, const uint32_t (&tw)[V],
                                                  float (&nx)[V], float (&ny)[V]) {
    uint32_t c0[V], c1[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        c0[j] = env_id + (uint32_t)j;
        c1[j] = tw[j];
    }
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const unsigned long long pr = (unsigned long long)0xD256D193u * c0[j];
            const uint32_t n0 = (uint32_t)(pr >> 32) ^ c1[j] ^ keys.noise[r];
            c1[j] = (uint32_t)pr;
            c0[j] = n0;
        }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
        nx[j] = draw_noise(c0[j]);
        ny[j] = draw_noise(c1[j]);
    }
}

template <typename R> struct alignas(16) DriveSlot {
    R phi, h, chord, pad;
};

template <int ACT, int V> constexpr bool stream_supported() {
    constexpr int b = act_bytes<ACT, V>();
    return V > 1 && (b == 4 || b == 8 || b == 16 || b == 32);
}

template <typename R, int V, int ACT, bool WAVES>
__global__ void __launch_bounds__(kMaxThreads + kSchedThreads, kMinBlocks)
step_stream_kernel(const __grid_constant__ StepArgs p) {
    __shared__ ObstacleBox<R> s_box[AQUA_MAX_OBSTACLES];
    __shared__ unsigned long long s_stats[AQUA_N_STATS];
    __shared__ WarpResetQueue<V> s_wq[kMaxThreads / 32];
    __shared__ DriveSlot<R> s_drive[3];
#if AQUA_SCHED_WARP
    __shared__ uint32_t s_ring[kMaxThreads / 32][kRing];
    if (threadIdx.x < (kMaxThreads / 32) * kRing) (&s_ring[0][0])[threadIdx.x] = kRingEmpty;
#endif
    const ObstacleView<R> tab = stage_table<R>(s_box, s_stage + stage_bytes<R, V>(), p.w.tab, p.w.k_obst, p.w.use_grid != 0);
    if (threadIdx.x < AQUA_N_STATS) s_stats[threadIdx.x] = 0ull;
    if (ACT != AQUA_ACT_F32X2 && threadIdx.x < 3) {  // aqua.py:33-42
        const Drive<R> d = drive_from_index<R>(threadIdx.x);
        s_drive[threadIdx.x] = DriveSlot<R>{d.phi, d.h, d.chord, R(0)};
    }
    __syncthreads();

    constexpr int kSlot = stage_slot<R, V>(), kPitch = stage_pitch<R, V>();
    constexpr int kStage = kStageSlots * kPitch;
    constexpr int kActBytes = act_bytes<ACT, V>();
    constexpr int kActPiece = kActBytes < 16 ? kActBytes : 16;
    constexpr int kActPieces = kActBytes / kActPiece;
    constexpr int kActWords = kActBytes / 4;
    static_assert(stream_supported<ACT, V>() && (kSlot == 16 || kSlot == 8) && kStages == 2, "streaming kernel layout");

    const uint32_t n = p.n_stream;
    const bool autoreset = p.w.autoreset != 0;
    const int k_obst = p.w.k_obst, k_circ = p.w.k_circ;
    uint32_t s_cur = smem_u32(s_stage) + threadIdx.x * kSlot, s_nxt = s_cur + kStage;

    // the 8 plane slices + actions of the V envs starting at (relative) env index i
    auto prefetch = [&](uint32_t s, uint32_t i) {
#pragma unroll
        for (int pl = 0; pl < 7; ++pl) cp_async_s<kSlot>(s + pl * kPitch, p.pl0[pl] + (size_t)i * sizeof(R));
        cp_async_s<4 * V>(s + AQUA_PLANE_TIME * kPitch, p.pl0[AQUA_PLANE_TIME] + (size_t)i * 4);
#pragma unroll
        for (int q = 0; q < kActPieces; ++q)
            cp_async_s<kActPiece>(s + (AQUA_N_PLANES + q) * kPitch, p.act0 + (size_t)i * (kActBytes / V) + q * kActPiece);
    };

    ThreadStats st;
    WarpResetQueue<V>& wq = s_wq[(threadIdx.x >> 5) % (kMaxThreads / 32)];  // (the scheduler warp queues nothing)
    int n_queued = 0;  // warp-uniform

    // Work distribution: a warp tile is 32 V consecutive envs.  Every warp starts with four
    // statically assigned tiles and then claims further tiles from a global counter, one claim per
    // tile, issued two tiles ahead of its use (the round trip of an atomic through the loaded
    // memory system is longer than one tile): SMs do not run at the same speed (distance to the
    // memory partitions), and a static partition makes the kernel as slow as its slowest SM.
    // One counter cannot serve a claim per 128 envs (same-address atomics retire about one per
    // 1.5 ns), so warp w of every CTA draws from pool w mod n_pools: up to kTilePools counters on
    // separate lines, pool q owning the dynamic tiles congruent to q modulo n_pools.  Every pool has
    // the same number of warps on every CTA (warps per CTA: up to 8, or a multiple of 8), hence
    // the pools drain at the same rate.
    // Two staging buffers, prefetch distance two tiles: a buffer is refilled (with the tile after
    // next) as soon as its slices have been read into registers, so every copy has two tile
    // times to land.
    const uint32_t lane_off = (threadIdx.x & 31u) * V;
    constexpr uint32_t kWarpTile = 32u * V;
#if AQUA_SCHED_WARP
    // Scheduler warp: the last warp of the CTA does every tile claim, so that no worker ever waits
    // for the reply of a global atomic (measured: 21 % of a worker's time when it claimed its own
    // tiles).  Lane l feeds worker warp l through the ring s_ring[l]: it claims two tiles at a time
    // (atomicAdd 2 on the worker's pool) whenever two slots are free and writes kRingDone when the
    // pool is dry.  A worker pops one tile per tile it processes and always drains its ring up to
    // kRingDone, so no claimed tile is lost.  All spins are bounded (trap, not hang).
    const uint32_t n_workers = (blockDim.x >> 5) - 1u;
    const uint32_t n_warps = gridDim.x * n_workers;  // worker warps of the launch
    const uint32_t n_pools = min((uint32_t)kTilePools, n_workers);
    const uint32_t n_tiles = n / kWarpTile;
    const bool is_sched = (threadIdx.x >> 5) == n_workers;
    constexpr uint32_t kStatic = 3u;  // statically assigned tiles per worker
    if (is_sched) {
        const uint32_t l = threadIdx.x & 31u;
        if (l < n_workers) {
            volatile uint32_t* const ring = s_ring[l];
            const uint32_t q = l % n_pools;
            uint32_t* const ctr = p.tile_ctr + q * kTilePoolStride;
            uint32_t head = 0u;
            for (;;) {
                uint32_t spins = 0u;
                while (ring[head % kRing] != kRingEmpty || ring[(head + 1u) % kRing] != kRingEmpty) {
                    __nanosleep(100);
                    if (++spins > (1u << 24)) __trap();
                }
                const uint32_t c = atomicAdd(ctr, 2u);  // positions c, c + 1 of pool q's tile sequence
                const uint32_t t0 = kStatic * n_warps + c * n_pools + q, t1 = t0 + n_pools;
                if (t0 >= n_tiles) { ring[head % kRing] = kRingDone; break; }
                ring[head % kRing] = t0;
                if (t1 >= n_tiles) { ring[(head + 1u) % kRing] = kRingDone; break; }
                ring[(head + 1u) % kRing] = t1;
                head += 2u;
            }
        }
    }
    const uint32_t warp_g = blockIdx.x * n_workers + (threadIdx.x >> 5);
    uint32_t i = is_sched ? n : warp_g * kWarpTile + lane_off;       // this tile
    uint32_t i1 = is_sched ? n : i + n_warps * kWarpTile;            // the next one (in flight)
    uint32_t i2 = is_sched ? n : i1 + n_warps * kWarpTile;           // the one after (to be prefetched)
    volatile uint32_t* const my_ring = s_ring[is_sched ? 0u : (threadIdx.x >> 5)];
    uint32_t ring_head = 0u;
    bool ring_done = is_sched;  // warp-uniform
    // next tile of this worker (env index of this thread's slice), n once the pool is dry
    auto ring_pop = [&]() -> uint32_t {
        if (ring_done) return n;
        uint32_t t = 0u;
        if ((threadIdx.x & 31u) == 0u) {
            uint32_t spins = 0u;
            while ((t = my_ring[ring_head % kRing]) == kRingEmpty) {
                __nanosleep(50);
                if (++spins > (1u << 24)) __trap();
            }
            if (t != kRingDone) my_ring[ring_head % kRing] = kRingEmpty;  // hand the slot back
        }
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t == kRingDone) {
            ring_done = true;
            return n;
        }
        ++ring_head;
        return t * kWarpTile + lane_off;
    };
#else
    uint32_t lane_id;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_id));
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);  // of the launch
    uint32_t i = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kWarpTile + lane_off;  // this tile
    uint32_t i1 = i + n_warps * kWarpTile;   // the next one (in flight)
    uint32_t i2 = i1 + n_warps * kWarpTile;  // the one after (to be prefetched)
    uint32_t i3 = i2 + n_warps * kWarpTile;  // and its successor
    // pools in use = warps per CTA, at most kTilePools: every pool needs a drawer on every CTA
    const uint32_t n_pools = min((uint32_t)kTilePools, blockDim.x >> 5);
    const uint32_t pool = (threadIdx.x >> 5) % n_pools;
    uint32_t* const my_ctr = p.tile_ctr + pool * kTilePoolStride + (lane_id & p.tile_lane_mask);
    // ptxas lowers an atomic add on a warp-uniform address to the warp-aggregated form, whose
    // broadcast shuffle waits for the result on the spot; tile_lane_mask = 0 (a parameter) keeps
    // the address formally lane dependent, so the claim is a plain ATOMG consumed much later.
    auto claim_tile = [&](uint32_t& claim) {
        if ((threadIdx.x & 31u) == 0u)
            asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(claim) : "l"(my_ctr) : "memory");
    };
    uint32_t claim_old = 0u;      // issued during the previous tile: the successor of i3
    bool old_valid = i3 < n;      // warp-uniform
    if (old_valid) claim_tile(claim_old);
#endif
    if (i < n) prefetch(s_cur, i);
    cp_async_commit();
    if (i1 < n) prefetch(s_nxt, i1);
    cp_async_commit();
    for (;;) {
        const bool live = i < n;  // warp-uniform: n is a multiple of 32 V
        uint32_t fw = 0u;
        Pack<uint32_t, V> otw;
#if !AQUA_SCHED_WARP
        uint32_t claim_new = 0u;
#endif
        const uint32_t i_done = i;
        if (live) {
            cp_async_wait<1>();  // everything but the newest group has landed: this tile's slices

            const Pack<R, V> vx = lds_pack<R, V>(s_cur + AQUA_PLANE_X * kPitch);
            const Pack<R, V> vy = lds_pack<R, V>(s_cur + AQUA_PLANE_Y * kPitch);
            const Pack<R, V> vth = lds_pack<R, V>(s_cur + AQUA_PLANE_THETA * kPitch);
            const Pack<R, V> vgx = lds_pack<R, V>(s_cur + AQUA_PLANE_GOAL_X * kPitch);
            const Pack<R, V> vgy = lds_pack<R, V>(s_cur + AQUA_PLANE_GOAL_Y * kPitch);
            const Pack<R, V> vwx = lds_pack<R, V>(s_cur + AQUA_PLANE_WAVE_X * kPitch);
            const Pack<R, V> vwy = lds_pack<R, V>(s_cur + AQUA_PLANE_WAVE_Y * kPitch);
            const Pack<uint32_t, V> vtw = lds_pack<uint32_t, V>(s_cur + AQUA_PLANE_TIME * kPitch);
            uint32_t araw[kActWords];
#pragma unroll
            for (int q = 0; q < kActPieces; ++q) {
                uint32_t w[kActPiece / 4];
                lds_words<kActPiece>(s_cur + (AQUA_N_PLANES + q) * kPitch, w);
#pragma unroll
                for (int k = 0; k < kActPiece / 4; ++k) araw[q * (kActPiece / 4) + k] = w[k];
            }
            const ActVals<ACT, V> act = actions_from_raw<ACT, V, kActWords>(araw);
            if (i2 < n) prefetch(s_cur, i2);  // refill the buffer just read
            cp_async_commit();
#if !AQUA_SCHED_WARP
            const bool new_valid = old_valid && i3 < n;  // the pool had tiles left at the last look
            if (new_valid) claim_tile(claim_new);
#endif

#if AQUA_MEMBENCH
            // Measurement build only (make variant FLAGS=-DAQUA_MEMBENCH=1): the data movement of
            // this kernel -- staging pipeline, dynamic tiles, every load consumed, every store
            // issued -- without the step's arithmetic.  Flags stay 0 (no resets).  Timed next to the
            // real kernel on one box it gives the ceiling of this structure; never the shipped build.
            Pack<R, V> vrew, ox = vx, oy = vy, oth = vth, owx = vwx, owy = vwy;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                vrew.v[j] = vgx.v[j] + vgy.v[j];
                otw.v[j] = vtw.v[j] + act.w[j];
            }
            st.steps += V;
#else
            Drive<R> dr[V];
            if constexpr (ACT == AQUA_ACT_F32X2) {
                decode_drives<R, ACT, V>(act, dr, st.bad);
            } else {
                uint32_t worst = 0u;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    worst = max(worst, act.w[j]);
                    const DriveSlot<R> d = s_drive[min(act.w[j], 2u)];  // outside {0,1,2}: full throttle
                    dr[j].phi = d.phi;
                    dr[j].h = d.h;
                    dr[j].chord = d.chord;
                }
                if (worst > 2u) {  // rare
#pragma unroll
                    for (int j = 0; j < V; ++j) st.bad += (act.w[j] > 2u) ? 1u : 0u;
                }
            }

            float nx[V], ny[V];
            if constexpr (WAVES) {
                draw_wave_noise_v<V>(p.w.keys, p.env_id0 + i, vtw.v, nx, ny);
            } else {
#pragma unroll
                for (int j = 0; j < V; ++j) nx[j] = ny[j] = 0.f;
            }

            EnvState<R> e[V];
            MoveTmp<R> m[V];
            R pxn[V], pyn[V], clr[V];
#pragma unroll
            for (int j = 0; j < V; ++j) {
                e[j].x = vx.v[j]; e[j].y = vy.v[j]; e[j].th = vth.v[j];
                e[j].wx = vwx.v[j]; e[j].wy = vwy.v[j];
                e[j].gx = vgx.v[j]; e[j].gy = vgy.v[j];
                e[j].tw = vtw.v[j];
                move_env<R>(e[j], dr[j], (R)nx[j], (R)ny[j], WAVES, m[j]);
                pxn[j] = e[j].x;
                pyn[j] = e[j].y;
                clr[j] = R(kClearFar);
            }
            obstacle_clearance_any<R, V>(tab, k_circ, k_obst, pxn, pyn, clr);

            Pack<R, V> vrew, ox, oy, oth, owx, owy;
#pragma unroll
            for (int j = 0; j < V; ++j) {  // the flags word (byte j = env j) is built in 32-bit arithmetic
                R reward;
                fw |= finish_env<R>(e[j], m[j], clr[j], reward) << (8 * j);
                vrew.v[j] = reward;
                ox.v[j] = e[j].x; oy.v[j] = e[j].y; oth.v[j] = e[j].th;
                owx.v[j] = e[j].wx; owy.v[j] = e[j].wy; otw.v[j] = e[j].tw;
            }
            st.add_flags(fw, V);
#pragma unroll
            for (int j = 0; j < V; ++j) st.add_env<R>(vrew.v[j]);
            if (!autoreset) {
#pragma unroll
                for (int j = 0; j < V; ++j) st.add_length(fw >> (8 * j), otw.v[j] & kTimeMask);
            }

#endif
            // ---- stores (the goal only changes at reset: phase B writes it) ----
            const size_t off = (size_t)i * sizeof(R);
            st_pack_cs<R, V>(reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_X]) + off), ox);
            st_pack_cs<R, V>(reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_Y]) + off), oy);
            st_pack_cs<R, V>(reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_THETA]) + off), oth);
            st_pack_cs<R, V>(reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_WAVE_X]) + off), owx);
            st_pack_cs<R, V>(reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_WAVE_Y]) + off), owy);
            st_pack_cs<uint32_t, V>(reinterpret_cast<uint32_t*>(const_cast<char*>(p.pl0[AQUA_PLANE_TIME]) + (size_t)i * 4), otw);
            st_pack_cs<R, V>(reinterpret_cast<R*>(p.rew0 + off), vrew);
            if constexpr (V == 4) *reinterpret_cast<uint32_t*>(p.flg0 + i) = fw;
            else *reinterpret_cast<uint16_t*>(p.flg0 + i) = (uint16_t)fw;

            const uint32_t t = s_cur;
            s_cur = s_nxt;
            s_nxt = t;
#if AQUA_SCHED_WARP
            i = i1;
            i1 = i2;
            i2 = ring_pop();
#else
            // the claim of the previous tile is consumed here, after this tile's stores
            asm volatile("" : "+r"(claim_old) : : "memory");
            const uint32_t i4 =
                old_valid ? (4u * n_warps + __shfl_sync(0xffffffffu, claim_old, 0) * n_pools + pool) * kWarpTile + lane_off : n;
            i = i1;
            i1 = i2;
            i2 = i3;
            i3 = i4;
            claim_old = claim_new;
            old_valid = new_valid;
#endif
        }
        if (autoreset) {  // phase B, warp-collective
            if (live && __any_sync(0xffffffffu, fw != 0u)) n_queued = reset_queue_push<V>(wq, n_queued, fw, i_done, otw.v);
            const int need = live ? kResetGroups : 1;  // full passes only; the rest is flushed at the end
            if (n_queued >= need) n_queued = reset_queue_drain<R, V>(p, tab, wq, n_queued, need, st.length);
        }
        if (!live) break;
    }
    cp_async_wait<0>();
    __syncthreads();
    flush_stats(st, s_stats, p.stats);
    if (threadIdx.x == 0) {  // the last CTA to finish rearms the tile counter for the next launch
        __threadfence();
        if (atomicAdd(p.tile_ctr + kTilePools * kTilePoolStride, 1u) == gridDim.x - 1u) {
            for (int k = 0; k <= kTilePools; ++k) p.tile_ctr[k * kTilePoolStride] = 0u;
            __threadfence();
        }
    }
}

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
                                     WorldFlags{p.w.waves != 0, p.w.random_boat != 0, p.w.random_goal != 0});
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

// K3: fold the per-CTA slots into int64[AQUA_N_STATS]
__global__ void stats_reduce_kernel(unsigned long long* slots, int64_t* out, int clear) {
    __shared__ unsigned long long s[AQUA_N_STATS];
    if (threadIdx.x < AQUA_N_STATS) s[threadIdx.x] = 0ull;
    __syncthreads();
    unsigned long long acc[AQUA_N_STATS];
#pragma unroll
    for (int k = 0; k < AQUA_N_STATS; ++k) acc[k] = 0ull;
    for (int slot = threadIdx.x; slot < kStatSlots; slot += blockDim.x) {
#pragma unroll
        for (int k = 0; k < AQUA_N_STATS; ++k) {
            acc[k] += slots[slot * AQUA_N_STATS + k];
            if (clear) slots[slot * AQUA_N_STATS + k] = 0ull;
        }
    }
#pragma unroll
    for (int k = 0; k < AQUA_N_STATS; ++k) {
        unsigned long long v = acc[k];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s[k], v);
    }
    __syncthreads();
    if (threadIdx.x < AQUA_N_STATS) out[threadIdx.x] = (int64_t)s[threadIdx.x];
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
template <typename Op> static cudaError_t dispatch_rollout(int precision, int v, int continuous, const Op& op) {
    if (precision == AQUA_F64) {
        if (v == 1) return continuous ? op.template run<double, 1, true>() : op.template run<double, 1, false>();
        return continuous ? op.template run<double, Prec<double>::V, true>() : op.template run<double, Prec<double>::V, false>();
    }
    if (v == 1) return continuous ? op.template run<float, 1, true>() : op.template run<float, 1, false>();
    return continuous ? op.template run<float, Prec<float>::V, true>() : op.template run<float, Prec<float>::V, false>();
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
            const int threads = g.threads + kSchedThreads;
            if (a.w.waves) step_stream_kernel<R, V, ACT, true><<<g.blocks, threads, smem, s>>>(a);
            else step_stream_kernel<R, V, ACT, false><<<g.blocks, threads, smem, s>>>(a);
            return cudaGetLastError();
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

// Whole warps of wide threads go to the streaming kernel; what is left of the range (fewer than
// 32 V envs), noise injection and the narrow geometry go to the general kernel.
cudaError_t launch_step(int precision, int action_dtype, const StepArgs& a, LaunchGeom g, cudaStream_t s) {
    const int wide = envs_per_thread(precision);
    const int64_t n = a.env_end - a.env_begin;
    const int64_t warp_envs = 32 * wide;
    int64_t n_stream = 0;
    if (g.v == wide && wide > 1 && a.noise_x == nullptr && n <= ((int64_t)1 << 31) &&
        a.env_begin % wide == 0 && has_stream_kernel(precision, action_dtype))
        n_stream = n / warp_envs * warp_envs;
    if (n_stream == 0) return dispatch_step(precision, g.v, action_dtype, StepLaunchOp{a, g, s});

    const size_t es = precision == AQUA_F64 ? 8 : 4;
    size_t ab = 4;
    if (action_dtype == AQUA_ACT_I64 || action_dtype == AQUA_ACT_F32X2) ab = 8;
    else if (action_dtype == AQUA_ACT_U8) ab = 1;
    StepArgs b = a;
    b.env_end = a.env_begin + n_stream;
    for (int pl = 0; pl < AQUA_N_PLANES; ++pl)
        b.pl0[pl] = static_cast<const char*>(a.planes) + ((size_t)pl * (size_t)a.ld) * es +
                    (size_t)a.env_begin * (pl == AQUA_PLANE_TIME ? 4 : es);
    b.act0 = static_cast<const char*>(a.actions) + (size_t)a.env_begin * ab;
    b.rew0 = static_cast<char*>(a.reward) + (size_t)a.env_begin * es;
    b.flg0 = a.flags + a.env_begin;
    b.n_stream = (uint32_t)n_stream;
    b.tile_lane_mask = 0u;
    b.env_id0 = (uint32_t)(a.w.env_id_base + (uint64_t)a.env_begin);
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
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(step_stream_kernel<R, V, ACT, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(step_stream_kernel<R, V, ACT, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(step_stream_kernel<R, V, ACT, false>,
                                         cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(step_stream_kernel<R, V, ACT, true>,
                                         cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
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
                return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, step_stream_kernel<R, V, ACT, true>,
                                                                     threads + kSchedThreads, smem);
            return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, step_stream_kernel<R, V, ACT, false>,
                                                                 threads + kSchedThreads, smem);
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

struct RolloutLaunchOp {
    const StepArgs& a;
    LaunchGeom g;
    cudaStream_t s;
    template <typename R, int V, bool CONT> cudaError_t run() const {
        rollout_kernel<R, V, CONT><<<g.blocks, g.threads, cell_bytes(a.w.use_grid), s>>>(a);
        return cudaGetLastError();
    }
};
cudaError_t launch_rollout(int precision, int continuous, const StepArgs& a, LaunchGeom g, cudaStream_t s) {
    return dispatch_rollout(precision, g.v, continuous, RolloutLaunchOp{a, g, s});
}
struct RolloutOccupancyOp {
    int threads, use_grid;
    int* out;
    template <typename R, int V, bool CONT> cudaError_t run() const {
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, rollout_kernel<R, V, CONT>, threads, cell_bytes(use_grid));
    }
};
cudaError_t rollout_occupancy(int precision, int v, int continuous, int threads, int use_grid, int* blocks_per_sm) {
    return dispatch_rollout(precision, v, continuous, RolloutOccupancyOp{threads, use_grid, blocks_per_sm});
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

cudaError_t launch_stats_reduce(unsigned long long* slots, int64_t* out, int clear, cudaStream_t s) {
    stats_reduce_kernel<<<1, 256, 0, s>>>(slots, out, clear);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// host-side key schedule of the draw streams
// ------------------------------------------------------------------------------------------
static void philox4x32_10_host(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void make_philox_keys(uint64_t seed, PhiloxKeys* out) {
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int r = 0; r < 10; ++r) {
        out->rk4[2 * r] = key[0] + (uint32_t)r * 0x9E3779B9u;
        out->rk4[2 * r + 1] = key[1] + (uint32_t)r * 0xBB67AE85u;
    }
    uint32_t* streams[3] = {out->winit, out->noise, out->action};
    const uint32_t tags[3] = {kStreamWaveInit, kStreamWaveNoise, kStreamAction};
    for (int s = 0; s < 3; ++s) {
        const uint32_t ctr[4] = {tags[s], 0u, 0u, 0u};
        uint32_t o[4];
        philox4x32_10_host(ctr, key, o);
        for (int r = 0; r < 10; ++r) streams[s][r] = o[0] + (uint32_t)r * 0x9E3779B9u;
    }
}

// ------------------------------------------------------------------------------------------
// host-side obstacle tables (aqua.py:56-68, 373-390)
// ------------------------------------------------------------------------------------------
int envs_per_thread(int precision) { return precision == AQUA_F64 ? Prec<double>::V : Prec<float>::V; }

size_t obstacle_table_bytes(int precision) {
    return precision == AQUA_F64 ? sizeof(ObstacleTable<double>) : sizeof(ObstacleTable<float>);
}

// circles first (stable), then rectangles; returns the number of circles
static int sorted_rows(int k, const double* rows5, double* out5) {
    int n = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < k; ++i)
            if ((rows5[5 * i + 2] == 0.0) == (pass == 0)) {
                memcpy(out5 + 5 * n, rows5 + 5 * i, 5 * sizeof(double));
                ++n;
            }
    int k_circ = 0;
    for (int i = 0; i < k; ++i) k_circ += rows5[5 * i + 2] == 0.0;
    return k_circ;
}

// Classification of one grid cell (float64, 0.05 px safety margin that also covers the rounding of
// the cell index on the device).  *blocked: the whole cell lies inside one obstacle's clearance
// region (obstacle inflated by the radius 2.5) shrunk by the margin -- the distance to a convex
// obstacle is a convex function of the position, so its maximum over the cell is attained at a
// corner.  Otherwise the return value is the mask of the obstacles whose clearance region, grown
// by the margin, touches the cell.  (With 64 obstacles a genuine mask can equal kCellBlocked, hence
// the separate flag.)
static unsigned long long cell_mask(int k, const double* rows5, int cx, int cy, bool* blocked) {
    constexpr double kMargin = 0.05;
    const double cw = 100.0 / kGridDim;
    const double x0 = cx * cw, x1 = x0 + cw, y0 = cy * cw, y1 = y0 + cw;
    unsigned long long m = 0ull;
    for (int i = 0; i < k; ++i) {
        const double* o = rows5 + 5 * i;
        const bool circle = o[2] == 0.0;
        const double hx = circle ? 0.0 : o[3] / 2, hy = circle ? 0.0 : o[4] / 2;
        const double reach = circle ? o[3] + 2.5 : 2.5;
        // largest distance of a cell corner to the obstacle's box
        double far2 = 0.0;
        for (int c = 0; c < 4; ++c) {
            const double px = (c & 1) ? x1 : x0, py = (c & 2) ? y1 : y0;
            const double ex = std::max(std::fabs(px - o[0]) - hx, 0.0), ey = std::max(std::fabs(py - o[1]) - hy, 0.0);
            far2 = std::max(far2, ex * ex + ey * ey);
        }
        if (reach > kMargin && far2 <= (reach - kMargin) * (reach - kMargin)) {
            *blocked = true;
            return kCellBlocked;
        }
        // distance between the obstacle's box and the cell rectangle
        const double dx = std::max(std::max(x0 - (o[0] + hx), (o[0] - hx) - x1), 0.0);
        const double dy = std::max(std::max(y0 - (o[1] + hy), (o[1] - hy) - y1), 0.0);
        if (dx * dx + dy * dy <= (reach + kMargin) * (reach + kMargin)) m |= 1ull << i;
    }
    *blocked = false;
    return m;
}

// Returns whether the culling grid can be used: the table is dense enough to pay for it, and no
// cell that is not blocked carries a full mask (which would read as kCellBlocked; only possible
// with AQUA_MAX_OBSTACLES obstacles all reaching one 2 px cell -- such a table is tested by brute
// force).
template <typename R> static bool fill_table(int k, const double* rows5, ObstacleTable<R>* t) {
    memset(t, 0, sizeof(*t));
    bool grid_ok = k > kGridMinObstacles;
    for (int cy = 0; cy < kGridDim; ++cy)
        for (int cx = 0; cx < kGridDim; ++cx) {
            bool blocked = false;
            const unsigned long long m = cell_mask(k, rows5, cx, cy, &blocked);
            if (!blocked && m == kCellBlocked) grid_ok = false;  // a genuine full mask: ambiguous
            t->cell[cy * kGridDim + cx] = m;
        }
    for (int i = 0; i < k; ++i) {
        const double* o = rows5 + 5 * i;
        const bool circle = o[2] == 0.0;
        t->box[i].cx = (R)o[0];
        t->box[i].cy = (R)o[1];
        if (circle) {
            const double reach = o[3] + 2.5;  // boat / goal radius 2.5, aqua.py:72,75
            t->box[i].a = (R)(reach * reach);
            t->box[i].b = R(0);
        } else {
            t->box[i].a = (R)(o[3] / 2);
            t->box[i].b = (R)(o[4] / 2);
        }
    }
    return grid_ok;
}

int build_obstacle_tables(int precision, int k, const double* rows5, void* tab_host, double* tab64_host,
                          int* use_grid) {
    double sorted[AQUA_MAX_OBSTACLES * 5];
    const int k_circ = sorted_rows(k, rows5, sorted);
    bool grid_ok;
    if (precision == AQUA_F64) grid_ok = fill_table<double>(k, sorted, reinterpret_cast<ObstacleTable<double>*>(tab_host));
    else grid_ok = fill_table<float>(k, sorted, reinterpret_cast<ObstacleTable<float>*>(tab_host));
    *use_grid = grid_ok ? 1 : 0;
    for (int i = 0; i < k; ++i) {
        const double* o = sorted + 5 * i;
        const bool circle = o[2] == 0.0;
        double* d = tab64_host + 5 * i;
        d[0] = o[0];
        d[1] = o[1];
        d[2] = circle ? 0.0 : o[3] / 2;
        d[3] = circle ? 0.0 : o[4] / 2;
        const double reach = circle ? o[3] + 2.5 : 2.5;
        d[4] = reach * reach;
    }
    return k_circ;
}

void copy_table_cells(int precision, const void* tab_host, unsigned long long* out) {
    static_assert(kGridDim == AQUA_GRID_DIM && kCellBlocked == AQUA_CELL_BLOCKED, "header and kernels agree on the grid");
    const unsigned long long* cell = precision == AQUA_F64 ? reinterpret_cast<const ObstacleTable<double>*>(tab_host)->cell
                                                            : reinterpret_cast<const ObstacleTable<float>*>(tab_host)->cell;
    memcpy(out, cell, sizeof(unsigned long long) * kGridCells);
}

}  // namespace aqua
