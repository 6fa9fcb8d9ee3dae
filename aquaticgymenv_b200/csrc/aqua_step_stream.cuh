// aqua_step_stream.cuh -- K1s (step_stream_kernel): the streaming form of the env step, what
// large slabs run.
#pragma once

#include "aqua_kernel_util.cuh"

// tuning only (tools/tune.py variants): -DAQUA_STREAM_TAB=1 compiles the streaming kernel for small-table launches alone
#ifndef AQUA_STREAM_TAB
#define AQUA_STREAM_TAB -1
#endif

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

template <int ACT, int V> constexpr bool stream_supported() {
    constexpr int b = act_bytes<ACT, V>();
    return V > 1 && (b == 4 || b == 8 || b == 16 || b == 32);
}

// ------------------------------------------------------------------------------------------
// Bulk staging (AQUA_TMA=1, the default): the 8 plane segments + actions of a warp tile are
// contiguous runs of 32 x 16 bytes in global memory, so ONE lane of the warp moves each with a bulk
// asynchronous copy (cp.async.bulk.shared::cta.global, the 1-D form of the tensor memory
// accelerator: UBLKCP in SASS) that lands in the warp's 512-byte stretch of the staging slot and
// reports its bytes to the warp's mbarrier for that buffer.  Nine copy instructions per tile
// instead of 9 x 32 per-lane LDGSTS and their address arithmetic; the lanes wait on the mbarrier
// phase instead of a cp.async group.  A buffer is refilled after __syncwarp(): every lane has read
// its slices by then.  Both forms are compiled (template parameter TMA).  Measured on one box
// (profiles/r02a_variants_ab.txt): equal at 2^24 envs (7.89e10 both), per-lane cp.async 4.6 % faster at
// 2^20 envs, where a warp sees three tiles and the start-up of the copy pipeline counts.  launch_step
// therefore takes the bulk form from kTmaMinEnvs envs per launch (AQUA_TMA: 0 never, 1 by size, 2 always).
// ------------------------------------------------------------------------------------------
#ifndef AQUA_TMA
#define AQUA_TMA 1
#endif
constexpr int64_t kTmaMinEnvs = (int64_t)1 << 22;
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0u;
}
// bounded: a copy that never lands traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0u;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 26)) __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

template <typename R, int V, int ACT, bool WAVES, bool TMA>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks)
step_stream_kernel(const __grid_constant__ StepArgs p) {
    __shared__ ObstacleBox<R> s_box[kBoxSlots];
    __shared__ R s_nr2[kBoxSlots + 3];
    __shared__ unsigned long long s_stats[kSlotStats];
    __shared__ WarpResetQueue<V> s_wq[kMaxThreads / 32];
    __shared__ DriveSlot<R> s_drive[3];
    __shared__ __align__(8) unsigned long long s_bar[TMA ? kMaxThreads / 32 : 1][2];  // [warp][buffer]
    const ObstacleView<R> tab = stage_table<R>(s_box, s_nr2, s_stage + stage_bytes<R, V>(), p.w.tab, p.w.k_obst, p.w.use_grid != 0);
    if (threadIdx.x < kSlotStats) s_stats[threadIdx.x] = 0ull;
    if (ACT != AQUA_ACT_F32X2 && threadIdx.x < 3) {  // aqua.py:33-42
        const Drive<R> d = drive_from_index<R>(threadIdx.x);
        s_drive[threadIdx.x] = DriveSlot<R>{d.phi, d.h, d.chord, R(0)};
    }
    if constexpr (TMA) {
        if ((threadIdx.x & 31u) == 0u) {
            mbar_init(smem_u32(&s_bar[threadIdx.x >> 5][0]), 1u);
            mbar_init(smem_u32(&s_bar[threadIdx.x >> 5][1]), 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    __syncthreads();
    // Programmatic dependent launch (launch_step sets the attribute): the CTAs of this launch may
    // become resident, and stage the obstacle table above, while the previous kernel of the stream is
    // still running; they wait here, before their first access to the state, until it has completed.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

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
    const uint32_t lane_off = (threadIdx.x & 31u) * V;
    uint32_t s_cur = smem_u32(s_stage) + threadIdx.x * kSlot, s_nxt = s_cur + kStage;

    // The warp's stretch of a staging slot is the 32 slices of its lanes, kSlot bytes each.  With bulk
    // copies the actions are one linear run of kActBytes per thread instead (4 to 32 bytes: the two
    // action slots together hold kMaxThreads x 32), so thread t finds its actions act_off = t *
    // (kActBytes - kSlot) bytes from where a kSlot-byte slice would lie (modulo 2^32: the offset may be
    // negative); the same for the time plane, whose elements are 4 bytes whatever the state's.
    const uint32_t act_off = AQUA_N_PLANES * kPitch + (TMA ? threadIdx.x * (uint32_t)(kActBytes - kSlot) : 0u);
    const uint32_t tw_off = AQUA_PLANE_TIME * kPitch + (TMA ? threadIdx.x * (uint32_t)(4 * V - kSlot) : 0u);
    constexpr uint32_t kTxBytes = 32u * (7u * kSlot + 4u * V + kActBytes);
    uint32_t bar_cur = smem_u32(&s_bar[TMA ? threadIdx.x >> 5 : 0][0]), bar_nxt = bar_cur + 8u;
    uint32_t par_cur = 0u, par_nxt = 0u;  // phase parity to wait for on each buffer
    const bool issuer = (threadIdx.x & 31u) == 0u;
    // stages the tile whose first env (for this lane) is i into the buffer s (this lane's slice base)
    auto prefetch = [&](uint32_t s, uint32_t bar, uint32_t i) {
        if constexpr (TMA) {
            if (issuer) {  // lane 0: i is the tile's first env, s the warp's stretch
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the lanes' reads of s precede the refill
                mbar_expect_tx(bar, kTxBytes);
#pragma unroll
                for (int pl = 0; pl < 7; ++pl) bulk_g2s(s + pl * kPitch, p.pl0[pl] + (size_t)i * sizeof(R), 32u * kSlot, bar);
                bulk_g2s(s + tw_off, p.pl0[AQUA_PLANE_TIME] + (size_t)i * 4, 32u * 4u * V, bar);
                bulk_g2s(s + act_off, p.act0 + (size_t)i * (kActBytes / V), 32u * kActBytes, bar);
            }
        } else {  // the 8 plane slices + actions of the V envs starting at (relative) env index i, per lane
#pragma unroll
            for (int pl = 0; pl < 7; ++pl) cp_async_s<kSlot>(s + pl * kPitch, p.pl0[pl] + (size_t)i * sizeof(R));
            cp_async_s<4 * V>(s + AQUA_PLANE_TIME * kPitch, p.pl0[AQUA_PLANE_TIME] + (size_t)i * 4);
#pragma unroll
            for (int q = 0; q < kActPieces; ++q)
                cp_async_s<kActPiece>(s + (AQUA_N_PLANES + q) * kPitch, p.act0 + (size_t)i * (kActBytes / V) + q * kActPiece);
            cp_async_commit();
        }
    };

    ThreadStats st;
    WarpResetQueue<V>& wq = s_wq[threadIdx.x >> 5];
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
    constexpr uint32_t kWarpTile = 32u * V;
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
    // (the per-lane form commits one cp.async group per tile slot, empty when there is no tile)
    if (i < n) prefetch(s_cur, bar_cur, i);
    else if (!TMA) cp_async_commit();
    if (i1 < n) prefetch(s_nxt, bar_nxt, i1);
    else if (!TMA) cp_async_commit();
    for (;;) {
        const bool live = i < n;  // warp-uniform: n is a multiple of 32 V
        uint32_t fw = 0u;
        Pack<uint32_t, V> otw;
        uint32_t claim_new = 0u;
        const uint32_t i_done = i;
        if (live) {
            if constexpr (TMA) {
                mbar_wait(bar_cur, par_cur);  // this tile's nine copies have landed
                par_cur ^= 1u;
            } else {
                cp_async_wait<1>();  // everything but the newest group has landed: this tile's slices
            }

            const Pack<R, V> vx = lds_pack<R, V>(s_cur + AQUA_PLANE_X * kPitch);
            const Pack<R, V> vy = lds_pack<R, V>(s_cur + AQUA_PLANE_Y * kPitch);
            const Pack<R, V> vth = lds_pack<R, V>(s_cur + AQUA_PLANE_THETA * kPitch);
            const Pack<R, V> vgx = lds_pack<R, V>(s_cur + AQUA_PLANE_GOAL_X * kPitch);
            const Pack<R, V> vgy = lds_pack<R, V>(s_cur + AQUA_PLANE_GOAL_Y * kPitch);
            const Pack<R, V> vwx = lds_pack<R, V>(s_cur + AQUA_PLANE_WAVE_X * kPitch);
            const Pack<R, V> vwy = lds_pack<R, V>(s_cur + AQUA_PLANE_WAVE_Y * kPitch);
            const Pack<uint32_t, V> vtw = lds_pack<uint32_t, V>(s_cur + tw_off);
            uint32_t araw[kActWords];
#pragma unroll
            for (int q = 0; q < kActPieces; ++q) {
                uint32_t w[kActPiece / 4];
                lds_words<kActPiece>(s_cur + (TMA ? act_off + q * kActPiece : (AQUA_N_PLANES + q) * kPitch), w);
#pragma unroll
                for (int k = 0; k < kActPiece / 4; ++k) araw[q * (kActPiece / 4) + k] = w[k];
            }
            const ActVals<ACT, V> act = actions_from_raw<ACT, V, kActWords>(araw);
            if constexpr (TMA) __syncwarp();  // every lane has read its slices: the buffer may be refilled
            if (i2 < n) prefetch(s_cur, bar_cur, i2);  // refill the buffer just read
            else if (!TMA) cp_async_commit();
            const bool new_valid = old_valid && i3 < n;  // the pool had tiles left at the last look
            if (new_valid) claim_tile(claim_new);

            EnvState<R> e[V];
#pragma unroll
            for (int j = 0; j < V; ++j) {
                e[j].x = vx.v[j]; e[j].y = vy.v[j]; e[j].th = vth.v[j];
                e[j].wx = vwx.v[j]; e[j].wy = vwy.v[j];
                e[j].gx = vgx.v[j]; e[j].gy = vgy.v[j];
                e[j].tw = vtw.v[j];
            }
            Pack<R, V> vrew;
#if AQUA_MEMBENCH
            // Measurement build only (make variant FLAGS=-DAQUA_MEMBENCH=1): the data movement of
            // this kernel -- staging pipeline, dynamic tiles, every load consumed, every store
            // issued -- without the step's arithmetic.  Flags stay 0 (no resets).  Timed next to the
            // real kernel on one box it gives the ceiling of this structure; never the shipped build.
#pragma unroll
            for (int j = 0; j < V; ++j) {
                vrew.v[j] = vgx.v[j] + vgy.v[j];
                e[j].tw = vtw.v[j] + act.w[j];
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
            fw = step_envs<R, V, (sizeof(R) == 4 ? AQUA_STREAM_TAB : -1)>(e, dr, nx, ny, WAVES, tab, k_circ, k_obst, vrew.v, p.w.small, p.w.small_mode);
            st.add_flags(fw, V);
#pragma unroll
            for (int j = 0; j < V; ++j) st.add_env<R>(vrew.v[j]);
            if (!autoreset) {
#pragma unroll
                for (int j = 0; j < V; ++j) st.add_length(fw >> (8 * j), e[j].tw & kTimeMask);
            }
            if (p.term != nullptr && fw != 0u) {  // rare: keep the terminal observation (aqua.py:213)
#pragma unroll
                for (int j = 0; j < V; ++j)
                    if ((fw >> (8 * j)) & 1u) store_terminal<R>(p, p.env_begin + i + j, e[j]);
            }
#endif
            Pack<R, V> ox, oy, oth, owx, owy;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                ox.v[j] = e[j].x; oy.v[j] = e[j].y; oth.v[j] = e[j].th;
                owx.v[j] = e[j].wx; owy.v[j] = e[j].wy; otw.v[j] = e[j].tw;
            }
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
            if constexpr (TMA) {
                const uint32_t tb = bar_cur, tp = par_cur;
                bar_cur = bar_nxt;
                par_cur = par_nxt;
                bar_nxt = tb;
                par_nxt = tp;
            }
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
        }
        if (autoreset) {  // phase B, warp-collective
            if (live && __any_sync(0xffffffffu, fw != 0u)) n_queued = reset_queue_push<V>(wq, n_queued, fw, i_done, otw.v);
            const int need = live ? kLanePass : 1;  // full passes only; the rest is flushed at the end
            if (n_queued >= need) n_queued = reset_queue_drain_lanes<R, V>(p, tab, wq, n_queued, need, st.length);
        }
        if (!live) break;
    }
    if constexpr (!TMA) cp_async_wait<0>();
    __syncthreads();
    flush_stats(st, s_stats, p.stats);
    if (threadIdx.x == 0) {  // the last CTA to finish rearms the tile counter for the next launch
        __threadfence();
        if (atomicAdd(p.tile_ctr + kTilePools * kTilePoolStride, 1u) == gridDim.x - 1u) {
            __threadfence();  // every other CTA's claims precede the re-zeroing
            for (int k = 0; k <= kTilePools; ++k) p.tile_ctr[k * kTilePoolStride] = 0u;
            __threadfence();
        }
    }
}

}  // namespace aqua
