// aqua_step_many.cuh -- K1m (step_many_kernel): K consecutive env.step() calls in ONE launch.
//
// A single-step launch moves 65 bytes per env-step through HBM / L2 and cannot take less than a few
// microseconds, which is all a step of the BASELINE slab sizes (65,536 - 1M envs) is worth.  Here a
// thread keeps the state of its V envs in registers for K steps: per env-step it reads the action
// (4 bytes, from the caller's [K][row_stride] tape) and writes reward + flags (5 bytes); the state
// planes are read once before and written once after the K steps.  What bounds the kernel is the
// instruction issue rate, so the float32 arithmetic of env pairs is packed (aqua_packed.cuh).
// The same kernel with TAPE = false draws the actions from the Philox ACTION stream (the synthetic
// random policy of aqua_rollout_random) and writes no per-step outputs.
//
// Auto-reset without a stall: the state of an episode's first step depends only on (env id,
// episode index), never on how the previous episode ended, so it is drawn AHEAD OF TIME.  Every env
// owns a RING of kRingDepth 32-byte records in global memory: slot e mod kRingDepth holds the reset
// state of episode e, tagged e (the tags also sit in a plane of their own, 16 bytes per env).
//   * The record of each env's NEXT episode is in shared memory.  When an env finishes, its owner
//     swaps that record in (eight 32-bit shared-memory loads: see record_words) and starts an
//     asynchronous copy (cp.async) of the following slot from the ring into the same place.
//   * Used-up slots are drawn again, for the episode kRingDepth ahead, in ROUNDS of one try per lane
//     (ring_round): a lane carries one record through the reference's rejection loops
//     (aqua.py:100-126), one try per round, and takes the next queued record as soon as its own is
//     accepted, so the rejection loops of different envs never wait for the unluckiest one.
//   * Where the rounds run is a template parameter (INLOOP).  Dense maps and long launches: the owner
//     queues the slot when it swaps, and a round runs inside the step loop whenever 32 records are
//     queued or in the making (all lanes busy); between rounds the lanes' records sit in shared
//     memory.  Otherwise: before the first step of the next launch, which reads the tags of its envs
//     and queues every slot that does not hold one of the env's next kRingDepth episodes (that also
//     covers a fresh ring, and episodes moved on by a single-step launch or aqua_set_state).
// An env that finishes kRingDepth times before its used-up slots were drawn again finds a record with
// the wrong tag: everything queued is then drawn before the next step.  The ring is thus a cache:
// results do not depend on it (test_step_many_equals_stepwise,
// test_step_many_interleaved_with_single_steps).
#pragma once

#include <type_traits>

#include "aqua_kernel_util.cuh"

namespace aqua {

// Register budget: 128-thread CTAs, 5 of them per SM (96 registers) at 1 or 2 float envs per thread,
// 4 (128 registers) at 4 float envs and for double.  Measured on one box
// (profiles/r02k_step_many_occupancy_ab.txt): V = 4 at 128 instead of 146 registers (4 instead of 3
// CTAs per SM, no spills) +6 % at 2^20 envs, +9 % at 2^24, +14 % on the 64-obstacle map; V = 2 at 96
// instead of 107 (8 bytes spilled) +5 %.  One step further (V = 4 at 96) spills 200 bytes.
#ifndef AQUA_MANY_MINB
#define AQUA_MANY_MINB(V, F64) (((V) == 4 || (F64)) ? 4 : 5)
#endif
constexpr int kManyThreads = 128;  // launch bound; the launcher picks 32 .. 128 threads per CTA
constexpr uint32_t kStashEmpty = 0xffffffffu;

// the warp's queue of records to draw: what a round left behind (< 32) plus one entry per env of the warp
template <int V> struct RefillQueue {
    uint32_t rel[64 * V + 32];  // env index relative to env_begin
    uint32_t ep[64 * V + 32];   // episode index to draw the record for
};

// word: bit 0 of byte j set = queue env rel_base + j with episode ep[j]; returns the new length
template <int V>
__device__ __forceinline__ int refill_push(RefillQueue<V>& rq, int n, uint32_t word, uint32_t rel_base,
                                           const uint32_t (&ep)[V]) {
    const unsigned full = 0xffffffffu;
    const unsigned lt = (1u << (threadIdx.x & 31u)) - 1u;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const bool d = ((word >> (8 * j)) & 1u) != 0u;
        const unsigned mj = __ballot_sync(full, d);
        if (d) {
            const int at = n + __popc(mj & lt);
            rq.rel[at] = rel_base + (uint32_t)j;
            rq.ep[at] = ep[j];
        }
        n += __popc(mj);
    }
    return n;
}

__device__ __forceinline__ void stash_pack(const CoopResetResult& rr, uint32_t ep, uint4& a, uint4& b) {
    a = make_uint4(__float_as_uint(rr.x), __float_as_uint(rr.y), __float_as_uint(rr.th), __float_as_uint(rr.gx));
    b = make_uint4(__float_as_uint(rr.gy), __float_as_uint(rr.wx), __float_as_uint(rr.wy), ep);
}

// the record's reset state into the env, at time 0 of episode r1.w
template <typename R> __device__ __forceinline__ void stash_unpack(const uint4& r0, const uint4& r1, EnvState<R>& e) {
    e.x = (R)__uint_as_float(r0.x);
    e.y = (R)__uint_as_float(r0.y);
    e.th = (R)__uint_as_float(r0.z);
    e.gx = (R)__uint_as_float(r0.w);
    e.gy = (R)__uint_as_float(r1.x);
    e.wx = (R)__uint_as_float(r1.y);
    e.wy = (R)__uint_as_float(r1.z);
    e.tw = r1.w << AQUA_TIME_BITS;
}

// A record read from shared memory as eight separate 32-bit words.  A 16-byte load would tie x, y,
// heading and goal of ONE env to four consecutive registers, while the packed arithmetic of the
// step wants the same quantity of TWO envs side by side: the compiler then transposes the state
// with register moves in every step.  This path runs once per episode.
__device__ __forceinline__ void record_words(const uint4* rec, uint4& r0, uint4& r1) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(rec);
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r0.x) : "r"(a));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(r0.y) : "r"(a));
    asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(r0.z) : "r"(a));
    asm volatile("ld.shared.u32 %0, [%1+12];" : "=r"(r0.w) : "r"(a));
    asm volatile("ld.shared.u32 %0, [%1+16];" : "=r"(r1.x) : "r"(a));
    asm volatile("ld.shared.u32 %0, [%1+20];" : "=r"(r1.y) : "r"(a));
    asm volatile("ld.shared.u32 %0, [%1+24];" : "=r"(r1.z) : "r"(a));
    asm volatile("ld.shared.u32 %0, [%1+28];" : "=r"(r1.w) : "r"(a));
}
// the same from the ring in global memory (L2)
__device__ __forceinline__ void record_words_global(const uint4* rec, uint4& r0, uint4& r1) {
    const uint32_t* g = reinterpret_cast<const uint32_t*>(rec);
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(r0.x) : "l"(g));
    asm volatile("ld.global.cg.u32 %0, [%1+4];" : "=r"(r0.y) : "l"(g));
    asm volatile("ld.global.cg.u32 %0, [%1+8];" : "=r"(r0.z) : "l"(g));
    asm volatile("ld.global.cg.u32 %0, [%1+12];" : "=r"(r0.w) : "l"(g));
    asm volatile("ld.global.cg.u32 %0, [%1+16];" : "=r"(r1.x) : "l"(g));
    asm volatile("ld.global.cg.u32 %0, [%1+20];" : "=r"(r1.y) : "l"(g));
    asm volatile("ld.global.cg.u32 %0, [%1+24];" : "=r"(r1.z) : "l"(g));
    asm volatile("ld.global.cg.u32 %0, [%1+28];" : "=r"(r1.w) : "l"(g));
}

constexpr int kRingStartOnlySteps = 160;  // longer launches draw records inside the step loop on every map
constexpr int kRingDepth = 4;  // records per env; divides the episode range (2^21), so slot = episode & 3
constexpr uint32_t kNoWork = 0xffffffffu;

// the record a lane is drawing: env (relative to env_begin; kNoWork = idle), episode, try index
// (bit 31: the goal is accepted, the boat is being drawn) and the accepted goal
struct LaneWork {
    uint32_t rel, ep, n;
    float gx, gy;
};

// One round of the launch's record drawing (warp-collective): idle lanes take queued records from
// the tail of the queue, then every busy lane evaluates ONE try of its record's sampling sequence
// (try n = Philox(env, episode, n, RESET): the goal candidate while the goal is open, then the boat
// candidate -- exactly the sequence of reset_env_lane / aqua.py:100-126, accept decisions
// float64-exact).  A lane whose boat is accepted writes the record to the ring and is idle again.
// AQUA_RING_FULL_DRAIN=1: a launch on a sparse map draws every stale record before its first step;
// 0: only in full rounds of 32 -- what is left (fewer than 32 records, the ones farthest ahead: they
// are queued first and taken last) waits for the next launch
#ifndef AQUA_RING_FULL_DRAIN
#define AQUA_RING_FULL_DRAIN 1
#endif
#ifndef AQUA_RING_INLINE
#define AQUA_RING_INLINE 1
#endif
#if AQUA_RING_INLINE
#define AQUA_RING_FN __device__ __forceinline__
#else
#define AQUA_RING_FN __device__ __noinline__
#endif
// reset_blocked() through the small table (kernel parameters) where the launch has one: float only
template <typename R, int TAB>
__device__ __forceinline__ bool ring_blocked(const StepArgs& p, const ObstacleView<R>& tab, float x, float y) {
    if constexpr (TAB == 1 && sizeof(R) == 4) {
        const float px[1] = {x}, py[1] = {y};
        float clr[1] = {kClearFar};
        obstacle_clearance_small<1>(p.w.small, p.w.small_mode, px, py, clr);
        if (fabsf(clr[0]) <= p.w.reset_band) return hits_obstacle_g(p.w.tab64, p.w.k_obst, (double)x, (double)y);
        return clr[0] <= 0.f;
    } else {
        return reset_blocked<R, (TAB == 2 ? 1 : 0)>(tab, p.w.tab64, p.w.k_circ, p.w.k_obst, p.w.reset_band, x, y);
    }
}

template <typename R, int V, int TAB>
AQUA_RING_FN void ring_round(const StepArgs& p, const ObstacleView<R>& tab, const RefillQueue<V>& rq,
                                           int& n_q, int& n_act, LaneWork* sw) {
    const unsigned full = 0xffffffffu;
    LaneWork w = *sw;  // the lane's record lives in shared memory between rounds (rounds are rare)
    const unsigned lt = (1u << (threadIdx.x & 31u)) - 1u;
    const bool random_goal = p.w.random_goal != 0, random_boat = p.w.random_boat != 0;
    const unsigned mi = __ballot_sync(full, w.rel == kNoWork);
    if (n_q > 0 && mi != 0u) {  // warp-uniform
        const int rank = __popc(mi & lt);
        if (w.rel == kNoWork && rank < n_q) {
            w.rel = rq.rel[n_q - 1 - rank];
            w.ep = rq.ep[n_q - 1 - rank];
            w.n = random_goal ? 0u : 0x80000000u;
            w.gx = 25.f;  // aqua.py:107
            w.gy = 80.f;
        }
        n_q -= min(n_q, __popc(mi));
    }
    if (w.rel != kNoWork) {
        const uint32_t env_id = p.env_id0 + w.rel;
        const bool boat = (w.n >> 31) != 0u;
        const uint32_t nt = w.n & 0x7fffffffu;
        bool fin = false;
        float x = 85.f, y = 45.f, th = 0.f;  // aqua.py:117
        if (boat && !random_boat) {
            fin = true;
        } else {
            const uint4 r = philox_env(p.w.keys, env_id, w.ep, nt, kStreamReset);
            const float cx = draw_pos(boat ? r.x : r.z), cy = draw_pos(boat ? r.y : r.w);
            const bool force = nt + 1u >= kResetMaxTries;  // bounded loop: this try is accepted as is
            bool ok = true;
            if (force) {
                atomicAdd(p.w.forced, 1u);
            } else {
                ok = !(outside_border_f(cx, cy) || (boat && reset_reached<R>(w.gx, w.gy, cx, cy)) ||
                       ring_blocked<R, TAB>(p, tab, cx, cy));
            }
            w.n += 1u;
            if (ok && !boat) {  // aqua.py:102-105
                w.gx = cx;
                w.gy = cy;
                w.n |= 0x80000000u;
                fin = !random_boat;
            } else if (ok) {  // aqua.py:110-115
                fin = true;
                x = cx;
                y = cy;
                th = draw_angle(r.z);
            }
        }
        if (fin) {
            CoopResetResult rr;
            rr.x = x; rr.y = y; rr.th = th; rr.gx = w.gx; rr.gy = w.gy;
            rr.wx = rr.wy = 0.f;
            if (p.w.waves) draw_wave_init(p.w.keys, env_id, w.ep, rr.wx, rr.wy);  // aqua.py:124
            uint4 a, b;
            stash_pack(rr, w.ep, a, b);
            const size_t slot = (size_t)(p.env_begin + w.rel) * kRingDepth + (w.ep & (kRingDepth - 1));
            __stcg(p.stash + 2 * slot, a);
            __stcg(p.stash + 2 * slot + 1, b);
            __stcg(p.stash_tag + slot, w.ep);
            w.rel = kNoWork;
        }
    }
    *sw = w;
    n_act = __popc(__ballot_sync(full, w.rel != kNoWork));
}

// ten Philox2x32 rounds of N independent counters side by side
template <int N>
__device__ __forceinline__ void philox2x32_10_v(uint32_t (&c0)[N], uint32_t (&c1)[N], const uint32_t (&rk)[10],
                                                uint32_t key_add) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t k = rk[r] + key_add;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const unsigned long long pr = (unsigned long long)0xD256D193u * c0[j];
            const uint32_t n0 = (uint32_t)(pr >> 32) ^ c1[j] ^ k;
            c1[j] = (uint32_t)pr;
            c0[j] = n0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// The reference's Q-network (dqn.py:301-314: Dense 64 relu, Dense 64 relu, Dense 3 linear) evaluated by
// one lane for its own env.  `w` (shared memory) holds the Keras weights in their own layout:
// W1[5][64], b1[64], W2[64][64], b2[64], W3[64][3], b3[3] (kernels are [in][out]).  float32, every
// multiply-add an fma in ascending input order; the second hidden layer is produced 16 units at a
// time and folded into the three outputs at once, so only the first hidden layer is kept.
// Returns the first index of the largest output (numpy.argmax).
// ------------------------------------------------------------------------------------------
constexpr int kMlpIn = 5, kMlpHidden = 64, kMlpOut = 3;
constexpr int kMlpW1 = 0, kMlpB1 = kMlpW1 + kMlpIn * kMlpHidden, kMlpW2 = kMlpB1 + kMlpHidden,
              kMlpB2 = kMlpW2 + kMlpHidden * kMlpHidden, kMlpW3 = kMlpB2 + kMlpHidden,
              kMlpB3 = kMlpW3 + kMlpHidden * kMlpOut, kMlpWeights = kMlpB3 + kMlpOut;  // 4739
static_assert(kMlpWeights == AQUA_MLP_WEIGHTS, "header and kernel agree on the weight count");

__device__ __forceinline__ uint32_t mlp_greedy_action(const float* __restrict__ w, float x, float y, float th, float gx,
                                                      float gy) {
    // AquaStateNormalizer (utils.py:23-27) as observe_kernel computes it
    const float in[kMlpIn] = {x * (1.0f / 100.0f), y * (1.0f / 100.0f),
                              fmaf(th, 1.0f / (2.0f * 3.14159265358979323846f), 0.5f), gx * (1.0f / 100.0f),
                              gy * (1.0f / 100.0f)};
    float h1[kMlpHidden];
#pragma unroll
    for (int o = 0; o < kMlpHidden; o += 4) {
        float4 acc = *reinterpret_cast<const float4*>(w + kMlpB1 + o);
#pragma unroll
        for (int k = 0; k < kMlpIn; ++k) {
            const float4 ww = *reinterpret_cast<const float4*>(w + kMlpW1 + k * kMlpHidden + o);
            acc.x = fmaf(in[k], ww.x, acc.x);
            acc.y = fmaf(in[k], ww.y, acc.y);
            acc.z = fmaf(in[k], ww.z, acc.z);
            acc.w = fmaf(in[k], ww.w, acc.w);
        }
        h1[o] = fmaxf(acc.x, 0.f);
        h1[o + 1] = fmaxf(acc.y, 0.f);
        h1[o + 2] = fmaxf(acc.z, 0.f);
        h1[o + 3] = fmaxf(acc.w, 0.f);
    }
    float q[kMlpOut];
#pragma unroll
    for (int c = 0; c < kMlpOut; ++c) q[c] = w[kMlpB3 + c];
#pragma unroll 1
    for (int ob = 0; ob < kMlpHidden; ob += 16) {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            const float4 b = *reinterpret_cast<const float4*>(w + kMlpB2 + ob + i);
            acc[i] = b.x; acc[i + 1] = b.y; acc[i + 2] = b.z; acc[i + 3] = b.w;
        }
#pragma unroll
        for (int k = 0; k < kMlpHidden; ++k) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 ww = *reinterpret_cast<const float4*>(w + kMlpW2 + k * kMlpHidden + ob + i);
                acc[i] = fmaf(h1[k], ww.x, acc[i]);
                acc[i + 1] = fmaf(h1[k], ww.y, acc[i + 1]);
                acc[i + 2] = fmaf(h1[k], ww.z, acc[i + 2]);
                acc[i + 3] = fmaf(h1[k], ww.w, acc[i + 3]);
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float h2 = fmaxf(acc[i], 0.f);
#pragma unroll
            for (int c = 0; c < kMlpOut; ++c) q[c] = fmaf(h2, w[kMlpW3 + (ob + i) * kMlpOut + c], q[c]);
        }
    }
    uint32_t a = 0u;
    if (q[1] > q[0]) a = 1u;
    if (q[2] > q[a]) a = 2u;
    return a;
}

// the drives of two envs with discrete actions a, b in {0, 1, 2} from the pair table (three 8-byte loads)
template <typename R>
__device__ __forceinline__ void pair_drives(const R (*tab)[6], uint32_t a, uint32_t b, Drive<R>& da, Drive<R>& db) {
    const R* t = tab[3u * a + b];
    using P2 = typename std::conditional<sizeof(R) == 4, float2, double2>::type;
    const P2 phi = *reinterpret_cast<const P2*>(t), h = *reinterpret_cast<const P2*>(t + 2), chord = *reinterpret_cast<const P2*>(t + 4);
    da.phi = phi.x; db.phi = phi.y;
    da.h = h.x; db.h = h.y;
    da.chord = chord.x; db.chord = chord.y;
}

// starts the copy of the ring slot of episode `ep` of env `rel` into the env's shared-memory record
__device__ __forceinline__ void ring_prefetch(const StepArgs& p, uint32_t rel, uint32_t ep, uint4* dst) {
    const uint4* src = p.stash + 2 * ((size_t)(p.env_begin + rel) * kRingDepth + (ep & (kRingDepth - 1)));
    cp_async<16>(dst, src);
    cp_async<16>(dst + 1, src + 1);
    cp_async_commit();
}

// TAB: the obstacle test of the launch's table -- 0 brute force, 1 small table (World::small), 2 culling grid
// INLOOP: used-up ring records are drawn again inside the step loop (dense maps, long launches)
// MLP: the actions come from the reference's Q-network (dqn.py:301-314) evaluated by the env's own lane
// (aqua_rollout_policy); V = 1, discrete actions, no tape
template <typename R, int V, int ACT, bool TAPE, int TAB, bool INLOOP, bool MLP = false>
__global__ void __launch_bounds__(kManyThreads, MLP ? 3 : AQUA_MANY_MINB(V, sizeof(R) == 8)) step_many_kernel(const __grid_constant__ StepArgs p) {
    __shared__ ObstacleBox<R> s_box[kBoxSlots];
    __shared__ R s_nr2[kBoxSlots + 3];
    __shared__ unsigned long long s_stats[kSlotStats];
    __shared__ DriveSlot<R> s_drive[3];
    // the drives of an env PAIR (actions a, b): entry 3 a + b = {phi_a, phi_b}, {h_a, h_b}, {chord_a, chord_b},
    // each read with one 8-byte load straight into the register pair the packed arithmetic takes
    __shared__ __align__(16) R s_drive2[9][6];
    __shared__ __align__(16) float s_mlp[MLP ? kMlpWeights + 1 : 4];
    static_assert(!MLP || (V == 1 && !TAPE && ACT == AQUA_ACT_I32), "the in-kernel policy: one env per thread, discrete");
    // dynamic shared memory: the culling grid of a dense map, one refill queue per warp, the stash
    // records of the CTA's envs (2 x 16 bytes each)
    const ObstacleView<R> tab = stage_table<R>(s_box, s_nr2, s_stage, p.w.tab, p.w.k_obst, TAB == 2);
    RefillQueue<V>* const s_rq = reinterpret_cast<RefillQueue<V>*>(s_stage + cell_bytes(TAB == 2));
    uint4* const s_rec_cta = reinterpret_cast<uint4*>(s_rq + (blockDim.x >> 5));
    LaneWork* const sw = reinterpret_cast<LaneWork*>(s_rec_cta + 2 * (size_t)blockDim.x * V) + threadIdx.x;
    if (threadIdx.x < kSlotStats) s_stats[threadIdx.x] = 0ull;
    if (ACT != AQUA_ACT_F32X2 && threadIdx.x < 3) {  // aqua.py:33-42
        const Drive<R> d = drive_from_index<R>(threadIdx.x);
        s_drive[threadIdx.x] = DriveSlot<R>{d.phi, d.h, d.chord, R(0)};
    }
    if constexpr (MLP) {
        for (int i = threadIdx.x; i < kMlpWeights; i += blockDim.x) s_mlp[i] = __ldg(p.mlp_w + i);
    }
    if (ACT != AQUA_ACT_F32X2 && V > 1 && threadIdx.x < 9) {
        const Drive<R> da = drive_from_index<R>(threadIdx.x / 3), db = drive_from_index<R>(threadIdx.x % 3);
        R* t = s_drive2[threadIdx.x];
        t[0] = da.phi; t[1] = db.phi;
        t[2] = da.h; t[3] = db.h;
        t[4] = da.chord; t[5] = db.chord;
    }
    __syncthreads();
    // Programmatic dependent launch (the launcher sets the attribute): this CTA may have been made
    // resident, and have staged the obstacle table above, while the previous kernel of the stream was
    // still running; it waits here, before its first read of the state, until that kernel has
    // completed and flushed.  The trigger right behind lets the next launch do the same.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // Where the used-up records of the ring are drawn again.  INLOOP (dense maps: episodes of ~27 steps
    // on the 64-obstacle map; launches of more than kRingStartOnlySteps steps on any map): inside the step
    // loop, whenever they fill a round.  Otherwise: before the first step of the next launch, all at
    // once -- measured on one box (profiles/r03_ring_ab.txt) the step loop without that code runs 3 to
    // 12 % faster on the sparse maps, as long as a launch is short enough for four records per env.
    constexpr bool kInLoop = INLOOP;
    const unsigned full = 0xffffffffu;
    const uint32_t n = p.n_stream;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp_rel = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32u * V);  // the warp's first env
    const uint32_t rel = warp_rel + lane * V;                                                      // this thread's
    const int nvalid = rel >= n ? 0 : (int)min((uint32_t)V, n - rel);
    const bool waves = p.w.waves != 0, autoreset = p.w.autoreset != 0;
    const int k_obst = p.w.k_obst, k_circ = p.w.k_circ, small_mode = p.w.small_mode;
    const int steps = (int)p.steps;  // the launcher bounds it to 2^31 - 1
    uint32_t valid_mask = 0u;  // 0xff in byte j for the envs of this thread that exist
#pragma unroll
    for (int j = 0; j < V; ++j) valid_mask |= (j < nvalid ? 0xffu : 0u) << (8 * j);

    ThreadStats st;
    if (warp_rel < n) {  // warp-uniform
        RefillQueue<V>& rq = s_rq[threadIdx.x >> 5];
        int n_q = 0;           // queued records, warp-uniform
        int n_act = 0;         // lanes that are drawing a record, warp-uniform
        uint32_t pend = 0u;    // bit 0 of byte j: env j finished and waits for the refill of its record
        uint4* const s_rec = s_rec_cta + 2 * (size_t)(threadIdx.x & ~31u) * V;         // the warp's, shared
        uint4* const my_rec = s_rec + 2 * (size_t)lane * V;                            // this thread's, shared
        const uint32_t env_id = p.env_id0 + rel;

        // ---- state in: the planes are padded to AQUA_TILE, full-width loads are in bounds ----
        EnvState<R> e[V];
        {
            const size_t off = (size_t)rel * sizeof(R);
            const Pack<R, V> vx = ld_pack<R, V>(reinterpret_cast<const R*>(p.pl0[AQUA_PLANE_X] + off));
            const Pack<R, V> vy = ld_pack<R, V>(reinterpret_cast<const R*>(p.pl0[AQUA_PLANE_Y] + off));
            const Pack<R, V> vth = ld_pack<R, V>(reinterpret_cast<const R*>(p.pl0[AQUA_PLANE_THETA] + off));
            const Pack<R, V> vgx = ld_pack<R, V>(reinterpret_cast<const R*>(p.pl0[AQUA_PLANE_GOAL_X] + off));
            const Pack<R, V> vgy = ld_pack<R, V>(reinterpret_cast<const R*>(p.pl0[AQUA_PLANE_GOAL_Y] + off));
            const Pack<R, V> vwx = ld_pack<R, V>(reinterpret_cast<const R*>(p.pl0[AQUA_PLANE_WAVE_X] + off));
            const Pack<R, V> vwy = ld_pack<R, V>(reinterpret_cast<const R*>(p.pl0[AQUA_PLANE_WAVE_Y] + off));
            const Pack<uint32_t, V> vtw =
                ld_pack<uint32_t, V>(reinterpret_cast<const uint32_t*>(p.pl0[AQUA_PLANE_TIME] + (size_t)rel * 4));
#pragma unroll
            for (int j = 0; j < V; ++j) {
                e[j].x = vx.v[j]; e[j].y = vy.v[j]; e[j].th = vth.v[j];
                e[j].wx = vwx.v[j]; e[j].wy = vwy.v[j];
                e[j].gx = vgx.v[j]; e[j].gy = vgy.v[j];
                e[j].tw = vtw.v[j];
            }
        }

        // ---- the ring: queue the stale slots of this warp's envs and draw them, all lanes busy; then
        // the record of each env's next episode into shared memory ----
        if (autoreset) {
            sw->rel = kNoWork;
            uint32_t next_ep[V];
            uint4 tg[V];
#pragma unroll
            for (int j = 0; j < V; ++j) {
                next_ep[j] = ((e[j].tw >> AQUA_TIME_BITS) + 1u) & kEpisodeMask;
                tg[j] = make_uint4(0u, 0u, 0u, 0u);
                if (j < nvalid) tg[j] = __ldcg(reinterpret_cast<const uint4*>(p.stash_tag) + (p.env_begin + (int64_t)rel + j));
            }
#pragma unroll 1
            for (int d = kRingDepth - 1; d >= 0; --d) {  // the nearest episodes last: they are taken first
                uint32_t want[V], word = 0u;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    want[j] = (next_ep[j] + (uint32_t)d) & kEpisodeMask;
                    const uint32_t sl = want[j] & (kRingDepth - 1);
                    const uint32_t have = sl == 0u ? tg[j].x : (sl == 1u ? tg[j].y : (sl == 2u ? tg[j].z : tg[j].w));
                    word |= ((j < nvalid && have != want[j]) ? 1u : 0u) << (8 * j);
                }
                if (__any_sync(full, word != 0u)) n_q = refill_push<V>(rq, n_q, word, rel, want);
                __syncwarp();  // queue entries are visible to the whole warp
                while (n_q + n_act >= 32) ring_round<R, V, TAB>(p, tab, rq, n_q, n_act, sw);
            }
#if AQUA_RING_FULL_DRAIN
            // sparse maps: a launch draws all its records here and starts with a full ring
            if constexpr (!kInLoop)
                while (n_q + n_act > 0) ring_round<R, V, TAB>(p, tab, rq, n_q, n_act, sw);
#endif
            __syncwarp();  // orders the record stores of all lanes before the owners' loads
#pragma unroll
            for (int j = 0; j < V; ++j)
                if (j < nvalid) {
                    const uint4* src = p.stash + 2 * ((size_t)(p.env_begin + rel + j) * kRingDepth + (next_ep[j] & (kRingDepth - 1)));
                    my_rec[2 * j] = __ldcg(src);
                    my_rec[2 * j + 1] = __ldcg(src + 1);
                }
        }

        // ---- per-step I/O: row k of the tape is row_stride envs after row k - 1.  A thread without
        // envs (behind the last one of the launch) touches no row: the rows end at row_stride, which
        // only covers n rounded up to 16 ----
        constexpr int kActBytes = act_bytes<ACT, V>();
        const bool has = nvalid > 0;
        const char* act_ptr_k = TAPE ? p.act0 + (size_t)rel * (kActBytes / V) : nullptr;
        // (the row pointers advance every step whether or not the caller wants the output)
        const bool has_rew = p.reward != nullptr && has, has_flg = p.flags != nullptr && has;
        char* rew_ptr_k = p.rew0 + (size_t)rel * sizeof(R);
        uint8_t* flg_ptr_k = p.flg0 + rel;
        const size_t rew_row = (size_t)p.row_stride * sizeof(R), flg_row = (size_t)p.row_stride;
        // the actions of a step are loaded two steps ahead (a row of the tape comes from DRAM)
        const size_t act_row = (size_t)p.row_stride * (kActBytes / V);
        ActVals<ACT, V> act_next, act_next2;
#pragma unroll
        for (int j = 0; j < ActVals<ACT, V>::N; ++j) act_next.w[j] = act_next2.w[j] = 0u;
        if (TAPE && has) {
            act_next = fetch_actions<ACT, V>(act_ptr_k, V);
            if (steps > 1) act_next2 = fetch_actions<ACT, V>(act_ptr_k + act_row, V);
            act_ptr_k += 2 * act_row;  // the row of step k + 2
        }

        float mlp_eps = 0.f;
        uint64_t mlp_step0 = 0;
        if constexpr (MLP) {
            mlp_eps = p.mlp_eps_dev != nullptr ? __ldg(p.mlp_eps_dev) : p.mlp_eps;
            mlp_step0 = p.step0_dev != nullptr ? (uint64_t)__ldg(p.step0_dev) : p.step0;
        }
        // Two forms of the step loop around the same text (aqua_step_many_body.inc = one step,
        // aqua_step_many_pend.inc = the rare "record missing" path); which one is faster was measured on
        // one box (profiles/r03_ring_ab.txt, r04c - r04e).  Sparse maps: the step loop proper is an inner
        // do-while that holds nothing but the step; a warp with a missing record leaves it (+3 % over
        // the single loop).  Dense maps (records drawn in the loop): one loop with the rare path and the
        // rounds at its top (the nested form costs 2 to 4 % there).
        if constexpr (kInLoop) {
            for (int k = 0;; ++k) {
                if (autoreset) {
#include "aqua_step_many_pend.inc"
                    // the records used up so far are drawn again as soon as they fill a round
                    if (n_q + n_act >= 32) {
                        cp_async_wait<0>();
                        __syncwarp();
                        do ring_round<R, V, TAB>(p, tab, rq, n_q, n_act, sw);
                        while (n_q + n_act >= 32);
                        __syncwarp();  // orders the record stores before the copies the owners start later
                    }
                }
                if (k == steps) break;
#include "aqua_step_many_body.inc"
            }
        } else {
            int k = 0;
            bool ring_due = false;  // warp-uniform: an env waits for a record that was not in the ring
            for (;;) {
                if (ring_due) {
                    if (autoreset) {
#include "aqua_step_many_pend.inc"
                    }
                    ring_due = false;
                }
                if (k == steps) break;
                do {
#include "aqua_step_many_body.inc"
                    ++k;
                    if (autoreset) ring_due = __any_sync(full, pend != 0u);
                } while (k < steps && !ring_due);
            }
        }

        st.steps += (uint32_t)nvalid * (uint32_t)steps;

        cp_async_wait<0>();

        // ---- state out (full-width: the plane padding behind the last env is scratch) ----
        if (nvalid > 0) {
            Pack<R, V> ox, oy, oth, owx, owy, ogx, ogy;
            Pack<uint32_t, V> otw;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                ox.v[j] = e[j].x; oy.v[j] = e[j].y; oth.v[j] = e[j].th;
                owx.v[j] = e[j].wx; owy.v[j] = e[j].wy;
                ogx.v[j] = e[j].gx; ogy.v[j] = e[j].gy; otw.v[j] = e[j].tw;
            }
            const size_t off = (size_t)rel * sizeof(R);
            st_pack<R, V>(reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_X]) + off), ox);
            st_pack<R, V>(reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_Y]) + off), oy);
            st_pack<R, V>(reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_THETA]) + off), oth);
            st_pack<R, V>(reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_GOAL_X]) + off), ogx);
            st_pack<R, V>(reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_GOAL_Y]) + off), ogy);
            st_pack<R, V>(reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_WAVE_X]) + off), owx);
            st_pack<R, V>(reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_WAVE_Y]) + off), owy);
            st_pack<uint32_t, V>(reinterpret_cast<uint32_t*>(const_cast<char*>(p.pl0[AQUA_PLANE_TIME]) + (size_t)rel * 4), otw);
        }
    }
    __syncthreads();
    flush_stats(st, s_stats, p.stats);
}

}  // namespace aqua
