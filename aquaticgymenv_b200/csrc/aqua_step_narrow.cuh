// aqua_step_narrow.cuh -- K1n (step_narrow_kernel): the env step for small slabs, one env per
// thread.
//
// Below about 300k envs a launch cannot fill the GPU with wide threads and the step is bound by
// the latency of one thread's chain, not by bandwidth (tools/small_slab_probe.py).  This kernel is
// that chain and nothing else: no CTA barrier, no staging (the slab is L2-resident; every thread
// loads its 8 state words directly), the obstacle table read from global memory through L1 (all
// lanes read the same entry), one 32-bit env index, `waves` a template parameter, and statistics
// reduced per warp straight into the global slots.  A warp that finds finished envs resets them on
// the spot with the same cooperative routine as the other kernels.  Same arithmetic (move_env,
// obstacle_clearance, finish_env, coop_reset) as step_kernel and step_stream_kernel, hence the
// same results bit for bit; step_kernel keeps noise injection.
#pragma once

#include "aqua_kernel_util.cuh"

namespace aqua {

// Warp reduce (REDUX), then lanes 0..7 add one counter each to the warp's slot in global memory:
// no CTA barrier, no shared staging.  (For the streaming kernel the CTA-level flush_stats measured
// faster, profiles/r01Y_warp_flush_ab.txt; on small slabs, where a warp owns one tile, this one.)
__device__ __forceinline__ void flush_stats_warp(ThreadStats& t, unsigned long long* g_slots) {
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
    static_assert(AQUA_STAT_EPISODES == 0 && AQUA_STAT_COLLIDED == 1 && AQUA_STAT_TIMEOUT == 2 && AQUA_STAT_SUCCESS == 3 &&
                      AQUA_STAT_LENGTH_SUM == 4 && AQUA_STAT_REWARD_SUM == 5 && AQUA_STAT_STEPS == 6 &&
                      AQUA_STAT_BAD_ACTIONS == 7 && kSlotStats == 8,
                  "lane k carries counter k");
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long v = (unsigned long long)r;  // lane 5: the fixed-point reward sum
    v = lane == 0 ? d : lane == 1 ? c : lane == 2 ? o : lane == 3 ? s : lane == 4 ? l : lane == 6 ? n : lane == 7 ? b : v;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (lane < kSlotStats && v != 0ull) atomicAdd(&g_slots[(warp % kStatSlots) * kSlotStats + lane], v);
}

// AQUA_PDL=1: consecutive narrow launches overlap their launch latency (see the kernel's first lines)
#ifndef AQUA_PDL
#define AQUA_PDL 1
#endif

// launch: ceil(n_stream / blockDim.x) CTAs of kNarrowThreads threads; pl0 / act0 / rew0 / flg0 /
// n_stream / env_id0 of StepArgs filled as for the streaming kernel (any n_stream, no multiple)
template <typename R, int ACT, bool WAVES>
__global__ void __launch_bounds__(kNarrowThreads, 8) step_narrow_kernel(const __grid_constant__ StepArgs p) {
    __shared__ WarpResetQueue<1> s_wq[kNarrowThreads / 32];
    const ObstacleTable<R>* const g_tab = reinterpret_cast<const ObstacleTable<R>*>(p.w.tab);
    const ObstacleView<R> tab{p.w.use_grid ? g_tab->ubox : g_tab->box, g_tab->nr2, p.w.use_grid ? g_tab->cell16 : nullptr, g_tab->cell};

#if AQUA_PDL
    // Programmatic dependent launch (launch_step sets the attribute): the CTAs of this launch may
    // become resident while the previous kernel of the stream is still running, and wait here,
    // before their first read, until it has completed and its writes are visible.  The trigger right
    // behind lets the NEXT launch do the same with respect to this one.  Without the launch
    // attribute both instructions do nothing.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
    const uint32_t n = p.n_stream;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // env index relative to env_begin
    const bool autoreset = p.w.autoreset != 0;
    ThreadStats st;
    uint32_t fw = 0u;
    uint32_t tw_out[1] = {0u};
    if (i < n) {
        const size_t off = (size_t)i * sizeof(R);
        EnvState<R> e[1];
        e[0].x = *reinterpret_cast<const R*>(p.pl0[AQUA_PLANE_X] + off);
        e[0].y = *reinterpret_cast<const R*>(p.pl0[AQUA_PLANE_Y] + off);
        e[0].th = *reinterpret_cast<const R*>(p.pl0[AQUA_PLANE_THETA] + off);
        e[0].gx = *reinterpret_cast<const R*>(p.pl0[AQUA_PLANE_GOAL_X] + off);
        e[0].gy = *reinterpret_cast<const R*>(p.pl0[AQUA_PLANE_GOAL_Y] + off);
        e[0].wx = *reinterpret_cast<const R*>(p.pl0[AQUA_PLANE_WAVE_X] + off);
        e[0].wy = *reinterpret_cast<const R*>(p.pl0[AQUA_PLANE_WAVE_Y] + off);
        e[0].tw = *reinterpret_cast<const uint32_t*>(p.pl0[AQUA_PLANE_TIME] + (size_t)i * 4);
        const ActVals<ACT, 1> act = fetch_actions<ACT, 1>(p.act0 + (size_t)i * act_bytes<ACT, 1>(), 1);

        Drive<R> dr[1];
        decode_drives<R, ACT, 1>(act, dr, st.bad);
        float nx = 0.f, ny = 0.f;
        if constexpr (WAVES) draw_wave_noise(p.w.keys, p.env_id0 + i, e[0].tw, nx, ny);
        MoveTmp<R> m;
        move_env<R>(e[0], dr[0], (R)nx, (R)ny, WAVES, m);
        const R pxn[1] = {e[0].x}, pyn[1] = {e[0].y};
        R clr[1] = {R(kClearFar)};
        if constexpr (sizeof(R) == 4) {
            if (p.w.small_mode != 0) obstacle_clearance_small<1>(p.w.small, p.w.small_mode, pxn, pyn, clr);
            else obstacle_clearance_any<R, 1>(tab, p.w.k_circ, p.w.k_obst, pxn, pyn, clr);
        } else {
            obstacle_clearance_any<R, 1>(tab, p.w.k_circ, p.w.k_obst, pxn, pyn, clr);
        }
        R reward;
        fw = finish_env<R>(e[0], m, clr[0], reward);
        st.add_flags(fw, 1);
        st.add_env<R>(reward);
        if (!autoreset) st.add_length(fw, e[0].tw & kTimeMask);
        tw_out[0] = e[0].tw;
        if (p.term != nullptr && (fw & 1u)) store_terminal<R>(p, p.env_begin + i, e[0]);  // aqua.py:213

        // ---- stores (the goal only changes at reset: phase B writes it) ----
        *reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_X]) + off) = e[0].x;
        *reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_Y]) + off) = e[0].y;
        *reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_THETA]) + off) = e[0].th;
        *reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_WAVE_X]) + off) = e[0].wx;
        *reinterpret_cast<R*>(const_cast<char*>(p.pl0[AQUA_PLANE_WAVE_Y]) + off) = e[0].wy;
        *reinterpret_cast<uint32_t*>(const_cast<char*>(p.pl0[AQUA_PLANE_TIME]) + (size_t)i * 4) = e[0].tw;
        *reinterpret_cast<R*>(p.rew0 + off) = reward;
        p.flg0[i] = (uint8_t)fw;
    }
    // phase B, warp-collective (lanes beyond n carry fw = 0): finished envs are reset at once
    if (autoreset && __any_sync(0xffffffffu, fw != 0u)) {
        WarpResetQueue<1>& wq = s_wq[threadIdx.x >> 5];
        const int n_queued = reset_queue_push<1>(wq, 0, fw, i, tw_out);
        reset_queue_drain<R, 1>(p, tab, wq, n_queued, 1, st.length);
    }
    flush_stats_warp(st, p.stats);
}

}  // namespace aqua
