// aqua_kernel_util.cuh -- what the step kernels share: 16-byte plane access, action decode, the
// cp.async staging helpers, per-thread episode statistics and their reduction, the obstacle table
// in shared memory, and the per-warp reset queue of the in-kernel auto-reset.
#pragma once

#include <cstring>

#include "aqua_device.cuh"
#include "aqua_packed.cuh"

namespace aqua {

// dynamic shared memory of the step kernels: the staging buffers, then the culling grid of a dense map
extern __shared__ __align__(16) unsigned char s_stage[];

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

// the three discrete drives (aqua.py:33-42) as a shared-memory table: one 16-byte read per env
template <typename R> struct alignas(16) DriveSlot {
    R phi, h, chord, pad;
};

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
    if (threadIdx.x < kSlotStats) {
        const unsigned long long v = s_stats[threadIdx.x];
        if (v) atomicAdd(&g_slots[(blockIdx.x % kStatSlots) * kSlotStats + threadIdx.x], v);
    }
}

// dynamic shared memory a launch needs for the culling grid (dense maps only), placed behind the
// staging buffers
constexpr size_t kCellBytes = sizeof(unsigned short) * kGridCells;
__host__ __device__ inline size_t cell_bytes(int use_grid) { return use_grid ? kCellBytes : 0; }
constexpr int kBoxSlots = AQUA_MAX_OBSTACLES + 1;  // s_box / s_nr2 entries: the table + the entry that never collides

// Copies the obstacle table of the launch (global ObstacleTable<R>) into shared memory and returns
// the view the kernels read: the brute-force entries into `s_box`, or (dense map) the unified
// entries into `s_box`, their reaches into `s_nr2` and the 16-bit cells into `s_cell`.  The caller
// synchronises the CTA.
template <typename R>
__device__ __forceinline__ ObstacleView<R> stage_table(ObstacleBox<R>* s_box, R* s_nr2, unsigned char* s_cell,
                                                       const void* g_tab_, int k_obst, bool use_grid) {
    const ObstacleTable<R>* g_tab = reinterpret_cast<const ObstacleTable<R>*>(g_tab_);
    constexpr int kWordsPerBox = sizeof(ObstacleBox<R>) / 4;
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_box);
    if (!use_grid) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(g_tab->box);
        for (int i = threadIdx.x; i < k_obst * kWordsPerBox; i += blockDim.x) dst[i] = __ldg(src + i);
        return ObstacleView<R>{s_box, s_nr2, nullptr, g_tab->cell};
    }
    const uint32_t* src = reinterpret_cast<const uint32_t*>(g_tab->ubox);
    for (int i = threadIdx.x; i < kBoxSlots * kWordsPerBox; i += blockDim.x) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < kBoxSlots; i += blockDim.x) s_nr2[i] = __ldg(g_tab->nr2 + i);
    const uint32_t* csrc = reinterpret_cast<const uint32_t*>(g_tab->cell16);
    uint32_t* cdst = reinterpret_cast<uint32_t*>(s_cell);
    for (int i = threadIdx.x; i < kGridCells / 2; i += blockDim.x) cdst[i] = __ldg(csrc + i);
    return ObstacleView<R>{s_box, s_nr2, reinterpret_cast<const unsigned short*>(s_cell), g_tab->cell};
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
// AQUA_LANE_PASS: queued envs a lane-per-env pass waits for (reset_queue_drain_lanes); 32 = only
// full warps.  The queue holds what is left below that plus one tile's worth of new entries.
#ifndef AQUA_LANE_PASS
#define AQUA_LANE_PASS 32
#endif
constexpr int kLanePass = AQUA_LANE_PASS;
template <int V> struct WarpResetQueue {
    uint32_t rel[32 * V + 32];  // env index relative to env_begin
    uint32_t tw[32 * V + 32];   // its terminal time word (episode, length)
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
                                                 p.w.reset_band, wf, valid, env_id, ep, p.w.forced);
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

// The same with one LANE per queued env (reset_env_lane): passes of up to 32 envs from the tail,
// while at least `min_batch` are queued (kLanePass: full passes; 1: flush).  What the wide kernels
// use: a warp collects 32 finished envs within a few tiles.
template <typename R, int V>
__device__ __forceinline__ int reset_queue_drain_lanes(const StepArgs& p, const ObstacleView<R>& tab,
                                                       const WarpResetQueue<V>& wq, int n, int min_batch,
                                                       uint32_t& length_sum) {
    const int lane = (int)(threadIdx.x & 31u);
    const WorldFlags wf{p.w.waves != 0, p.w.random_boat != 0, p.w.random_goal != 0};
    uint32_t* const ptw = plane_tw<R>(p.planes, p.ld);
    __syncwarp();  // queue entries are visible to the whole warp
    while (n >= min_batch && n > 0) {
        const int take = min(n, 32);
        const int first = n - take;
        if (lane < take) {
            const int64_t i = p.env_begin + (int64_t)wq.rel[first + lane];
            const uint32_t tw = wq.tw[first + lane];
            const uint32_t ep = ((tw >> AQUA_TIME_BITS) + 1u) & kEpisodeMask;
            const uint32_t env_id = (uint32_t)(p.w.env_id_base + (uint64_t)i);
            const CoopResetResult rr = reset_env_lane<R>(p.w.keys, tab, p.w.tab64, p.w.k_circ, p.w.k_obst,
                                                         p.w.reset_band, wf, env_id, ep, p.w.forced);
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

// Terminal observation of an env that finished (aqua.py:213 returns it together with done = True;
// under auto-reset the state planes move on to the next episode): x, y, angle, goal_x, goal_y into
// the 5 `term` planes, by the env's owner.  `i` = absolute env index.
template <typename R>
__device__ __forceinline__ void store_terminal(const StepArgs& p, int64_t i, const EnvState<R>& e) {
    R* const t = reinterpret_cast<R*>(p.term) + i;
    t[0] = e.x;
    t[p.ld] = e.y;
    t[2 * p.ld] = e.th;
    t[3 * p.ld] = e.gx;
    t[4 * p.ld] = e.gy;
}

// the V flag bytes as one word (byte j = env j)
template <int V> __device__ __forceinline__ uint32_t flags_word(const Pack<uint8_t, V>& f) {
    uint32_t w = 0;
    memcpy(&w, &f, V);
    return w;
}

}  // namespace aqua
