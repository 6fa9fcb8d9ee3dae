// aqua_step_general.cuh -- K1 (step_kernel): the general form of the env step.  It serves the tail
// of a range, narrow threads (V = 1) with injected wave noise; large slabs run step_stream_kernel
// (aqua_step_stream.cuh), which inlines the same arithmetic and agrees with it bit for bit.
#pragma once

#include "aqua_kernel_util.cuh"

namespace aqua {

// ------------------------------------------------------------------------------------------
// K1: env.step(action) for a range of envs -- aqua.py:135-213 (+ reset aqua.py:100-126 when an
// env finishes and autoreset is on)
// ------------------------------------------------------------------------------------------
template <typename R, int V, int ACT, bool INJECT>
__global__ void __launch_bounds__(kMaxThreads, V == 1 ? 4 : kMinBlocks) step_kernel(const __grid_constant__ StepArgs p) {
    __shared__ ObstacleBox<R> s_box[kBoxSlots];
    __shared__ R s_nr2[kBoxSlots + 3];
    __shared__ unsigned long long s_stats[kSlotStats];
    __shared__ WarpResetQueue<V> s_wq[kMaxThreads / 32];
    const ObstacleView<R> tab = stage_table<R>(s_box, s_nr2, s_stage + stage_bytes<R, V>(), p.w.tab, p.w.k_obst, p.w.use_grid != 0);
    if (threadIdx.x < kSlotStats) s_stats[threadIdx.x] = 0ull;
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
            float nx[V], ny[V];
#pragma unroll
            for (int j = 0; j < V; ++j) {
                e[j].x = vx.v[j]; e[j].y = vy.v[j]; e[j].th = vth.v[j];
                e[j].wx = vwx.v[j]; e[j].wy = vwy.v[j];
                e[j].gx = vgx.v[j]; e[j].gy = vgy.v[j];
                e[j].tw = vtw.v[j];
                nx[j] = ny[j] = 0.f;
                if constexpr (INJECT) {
                    if (j < nvalid) {
                        nx[j] = __ldg(p.noise_x + base + j);
                        ny[j] = __ldg(p.noise_y + base + j);
                    }
                } else {
                    const uint32_t env_id = (uint32_t)(p.w.env_id_base + (uint64_t)(base + j));
                    if (waves) draw_wave_noise(p.w.keys, env_id, e[j].tw, nx[j], ny[j]);
                }
            }
            Pack<R, V> vrew, ox, oy, oth, owx, owy;
            const uint32_t fw_all = step_envs<R, V>(e, dr, nx, ny, waves, tab, k_circ, k_obst, vrew.v, p.w.small, p.w.small_mode);
            Pack<uint8_t, V> vfl;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                vfl.v[j] = (uint8_t)(fw_all >> (8 * j));
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
            if (p.term != nullptr && fw != 0u) {  // rare: keep the terminal observation (aqua.py:213)
#pragma unroll
                for (int j = 0; j < V; ++j)
                    if (vfl.v[j] & 1u) store_terminal<R>(p, base + j, e[j]);
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
            if constexpr (V == 1) {
                if (n_queued >= kResetGroups) n_queued = reset_queue_drain<R, V>(p, tab, wq, n_queued, kResetGroups, st.length);
            } else {
                if (n_queued >= kLanePass) n_queued = reset_queue_drain_lanes<R, V>(p, tab, wq, n_queued, kLanePass, st.length);
            }
        }
    }
    cp_async_wait<0>();
    if (autoreset) {  // flush the remainder
        if constexpr (V == 1) reset_queue_drain<R, V>(p, tab, wq, n_queued, 1, st.length);
        else reset_queue_drain_lanes<R, V>(p, tab, wq, n_queued, 1, st.length);
    }
    __syncthreads();
    flush_stats(st, s_stats, p.stats);
}

}  // namespace aqua
