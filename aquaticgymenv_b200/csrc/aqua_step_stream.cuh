// aqua_step_stream.cuh -- K1s (step_stream_kernel): the streaming form of the env step, what
// large slabs run.
#pragma once

#include "aqua_kernel_util.cuh"

namespace aqua {

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

}  // namespace aqua
