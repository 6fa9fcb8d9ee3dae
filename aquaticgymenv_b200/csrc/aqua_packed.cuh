// aqua_packed.cuh -- one env step for the V envs of a thread, with the float32 arithmetic of env
// PAIRS issued as packed FP32 instructions (Blackwell: add / mul / fma .rn.f32x2 -> FADD2 / FMUL2 /
// FFMA2, two IEEE-rounded results per instruction).
//
// step_envs<R, V>() is what every step kernel calls.  For float with V even it runs the pair form
// below; otherwise (double, V = 1) the scalar functions of aqua_device.cuh.  The pair form issues
// exactly the operations of move_env<float> / obstacle_clearance<float> / finish_env<float>, in
// the same order, each rounded once to nearest per lane -- the library is compiled with
// -fmad=false and every fused multiply-add is written out in both forms -- so the two agree bit
// for bit (tests: test_wide_and_narrow_kernels_agree_bitwise, test_step_many_equals_stepwise).
// What has no packed form stays scalar per half: min / max / abs, comparisons and selects, the
// integer side of the quadrant fix-up, MUFU sqrt / rcp, conversions.
#pragma once

#include "aqua_device.cuh"

#ifndef AQUA_PACKED
#define AQUA_PACKED 1
#endif

namespace aqua {

// two floats in one 64-bit register pair (lo = env 2q, hi = env 2q + 1)
struct F2 {
    unsigned long long u;
};
__device__ __forceinline__ F2 f2(float lo, float hi) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.u) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ F2 f2(float both) { return f2(both, both); }
__device__ __forceinline__ float lo(F2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v.u));
    return a;
}
__device__ __forceinline__ float hi(F2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v.u));
    return b;
}
__device__ __forceinline__ F2 operator+(F2 a, F2 b) {
    F2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u));
    return r;
}
__device__ __forceinline__ F2 operator-(F2 a, F2 b) {
    F2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u));
    return r;
}
__device__ __forceinline__ F2 operator*(F2 a, F2 b) {
    F2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u));
    return r;
}
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
    F2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.u) : "l"(a.u), "l"(b.u), "l"(c.u));
    return r;
}

// quadrant fix-up of Prec<float>::sincos_ for one half: q = bits of the rounded quadrant count
__device__ __forceinline__ void quadrant_fix(uint32_t q, float sr, float cr, float& s, float& c) {
    const bool swap = (q & 1u) != 0u;
    const float ss = swap ? cr : sr, cc = swap ? sr : cr;
    s = __uint_as_float(__float_as_uint(ss) ^ ((q & 2u) << 30));
    c = __uint_as_float(__float_as_uint(cc) ^ (((q + 1u) & 2u) << 30));
}

// move_env<float> for two envs (aqua.py:159-191)
__device__ __forceinline__ void move_pair(EnvState<float>& A, EnvState<float>& B, const Drive<float>& da,
                                          const Drive<float>& db, float nxa, float nya, float nxb, float nyb,
                                          bool waves, MoveTmp<float>& ma, MoveTmp<float>& mb) {
    const F2 th = f2(A.th, B.th);
    const F2 a = th + f2(da.h, db.h);
    // Prec<float>::sincos_(a)
    const F2 t = fma2(a, f2(0.636619772f), f2(12582912.0f));
    const F2 fq = t + f2(-12582912.0f);  // t - 12582912.0f
    F2 r = fma2(fq, f2(-1.57079637f), a);
    r = fma2(fq, f2(4.37113883e-8f), r);
    const F2 r2 = r * r;
    const F2 ps = fma2(r2, fma2(r2, f2(-1.9515295891e-4f), f2(8.3321608736e-3f)), f2(-1.6666654611e-1f));
    const F2 sr = fma2(r * r2, ps, r);
    const F2 pc = fma2(r2, fma2(r2, f2(2.443315711809948e-5f), f2(-1.388731625493765e-3f)), f2(4.166664568298827e-2f));
    const F2 cr = fma2(r2 * r2, pc, fma2(r2, f2(-0.5f), f2(1.0f)));
    float sna, csa, snb, csb;
    quadrant_fix(__float_as_uint(lo(t)), lo(sr), lo(cr), sna, csa);
    quadrant_fix(__float_as_uint(hi(t)), hi(sr), hi(cr), snb, csb);

    const F2 wx = f2(A.wx, B.wx), wy = f2(A.wy, B.wy);
    const F2 dx = fma2(f2(-da.chord, -db.chord), f2(sna, snb), wx);  // heading 0 = +y; aqua.py:173-181
    const F2 dy = fma2(f2(da.chord, db.chord), f2(csa, csb), wy);
    const F2 px = f2(A.x, B.x), py = f2(A.y, B.y);
    const F2 xn = px + dx, yn = py + dy;
    ma.dx = lo(dx); mb.dx = hi(dx);
    ma.dy = lo(dy); mb.dy = hi(dy);
    ma.px = A.x; mb.px = B.x;
    ma.py = A.y; mb.py = B.y;
    A.x = lo(xn); B.x = hi(xn);
    A.y = lo(yn); B.y = hi(yn);
    const F2 tn = th + f2(da.phi, db.phi);  // aqua.py:183
    A.th = wrap_angle(lo(tn));
    B.th = wrap_angle(hi(tn));
    {  // aqua.py:188-191; without waves the clip range is [0, 0] (aqua.py:23-25): selects, no branch
        const F2 wxn = wx + f2(nxa, nxb), wyn = wy + f2(nya, nyb);
        A.wx = waves ? fminf(fmaxf(lo(wxn), -0.05f), 0.05f) : 0.f;
        B.wx = waves ? fminf(fmaxf(hi(wxn), -0.05f), 0.05f) : 0.f;
        A.wy = waves ? fminf(fmaxf(lo(wyn), -0.05f), 0.05f) : 0.f;
        B.wy = waves ? fminf(fmaxf(hi(wyn), -0.05f), 0.05f) : 0.f;
    }
    A.tw += ((A.tw & kTimeMask) != kTimeMask) ? 1u : 0u;  // aqua.py:141
    B.tw += ((B.tw & kTimeMask) != kTimeMask) ? 1u : 0u;
}

// obstacle_clearance<float, V> with the V positions taken in pairs (brute force over the table)
template <int V>
__device__ __forceinline__ void obstacle_clearance_pairs(const ObstacleView<float>& tab, int k_circ, int k_obst,
                                                         const float (&px)[V], const float (&py)[V], float (&clr)[V]) {
    F2 x2[V / 2], y2[V / 2];
#pragma unroll
    for (int q = 0; q < V / 2; ++q) {
        x2[q] = f2(px[2 * q], px[2 * q + 1]);
        y2[q] = f2(py[2 * q], py[2 * q + 1]);
    }
    // four circles / two rectangles per trip, clamped index (see obstacle_clearance)
#pragma unroll 1
    for (int k = 0; k < k_circ; k += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const ObstacleBox<float> o = tab.box[min(k + u, k_circ - 1)];
#pragma unroll
            for (int q = 0; q < V / 2; ++q) {
                const F2 ddx = x2[q] - f2(o.cx), ddy = y2[q] - f2(o.cy);
                const F2 d = fma2(ddx, ddx, fma2(ddy, ddy, f2(-o.a)));
                clr[2 * q] = fminf(clr[2 * q], lo(d));
                clr[2 * q + 1] = fminf(clr[2 * q + 1], hi(d));
            }
        }
    }
#pragma unroll 1
    for (int k = k_circ; k < k_obst; k += 2) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const ObstacleBox<float> o = tab.box[min(k + u, k_obst - 1)];
#pragma unroll
            for (int q = 0; q < V / 2; ++q) {
                const F2 ex = x2[q] - f2(o.cx), ey = y2[q] - f2(o.cy);
                const F2 ux = f2(fabsf(lo(ex)), fabsf(hi(ex))) - f2(o.a);
                const F2 uy = f2(fabsf(lo(ey)), fabsf(hi(ey))) - f2(o.b);
                const F2 ddx = f2(fmaxf(lo(ux), 0.f), fmaxf(hi(ux), 0.f));
                const F2 ddy = f2(fmaxf(lo(uy), 0.f), fmaxf(hi(uy), 0.f));
                const F2 d = fma2(ddx, ddx, fma2(ddy, ddy, f2(-(float)kRectR2)));
                clr[2 * q] = fminf(clr[2 * q], lo(d));
                clr[2 * q + 1] = fminf(clr[2 * q + 1], hi(d));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Small tables (World::small, at most 4 circles + 4 rectangles): the entries are kernel parameters,
// i.e. constant-bank operands, and the tests are straight-line code: no loads, no loop, every test
// independent of the others.  Same operations per obstacle as obstacle_clearance(); the unused
// entries never collide and the minimum does not depend on the order, so the decision is the same.
// mode 1: the table has at most 2 rectangles, mode 2: up to 4.
// ------------------------------------------------------------------------------------------
template <int V>
__device__ __forceinline__ void obstacle_clearance_small(const SmallTable& t, int mode, const float (&px)[V],
                                                         const float (&py)[V], float (&clr)[V]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float ddx = px[j] - t.c[u][0], ddy = py[j] - t.c[u][1];
            clr[j] = fminf(clr[j], fmaf(ddx, ddx, fmaf(ddy, ddy, t.c[u][2])));
        }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        if (half == 1 && mode != 2) break;  // kernel-uniform
#pragma unroll
        for (int u = 2 * half; u < 2 * half + 2; ++u) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float ddx = fmaxf(fabsf(px[j] - t.r[u][0]) - t.r[u][2], 0.f);
                const float ddy = fmaxf(fabsf(py[j] - t.r[u][1]) - t.r[u][3], 0.f);
                clr[j] = fminf(clr[j], fmaf(ddx, ddx, fmaf(ddy, ddy, -(float)kRectR2)));
            }
        }
    }
}

// the same with the V positions taken in pairs (packed arithmetic)
template <int V>
__device__ __forceinline__ void obstacle_clearance_small_pairs(const SmallTable& t, int mode, const float (&px)[V],
                                                               const float (&py)[V], float (&clr)[V]) {
    F2 x2[V / 2], y2[V / 2];
#pragma unroll
    for (int q = 0; q < V / 2; ++q) {
        x2[q] = f2(px[2 * q], px[2 * q + 1]);
        y2[q] = f2(py[2 * q], py[2 * q + 1]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int q = 0; q < V / 2; ++q) {
            const F2 ddx = x2[q] - f2(t.c[u][0]), ddy = y2[q] - f2(t.c[u][1]);
            const F2 d = fma2(ddx, ddx, fma2(ddy, ddy, f2(t.c[u][2])));
            clr[2 * q] = fminf(clr[2 * q], lo(d));
            clr[2 * q + 1] = fminf(clr[2 * q + 1], hi(d));
        }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        if (half == 1 && mode != 2) break;  // kernel-uniform
#pragma unroll
        for (int u = 2 * half; u < 2 * half + 2; ++u) {
#pragma unroll
            for (int q = 0; q < V / 2; ++q) {
                const F2 ex = x2[q] - f2(t.r[u][0]), ey = y2[q] - f2(t.r[u][1]);
                const F2 ux = f2(fabsf(lo(ex)), fabsf(hi(ex))) - f2(t.r[u][2]);
                const F2 uy = f2(fabsf(lo(ey)), fabsf(hi(ey))) - f2(t.r[u][3]);
                const F2 ddx = f2(fmaxf(lo(ux), 0.f), fmaxf(hi(ux), 0.f));
                const F2 ddy = f2(fmaxf(lo(uy), 0.f), fmaxf(hi(uy), 0.f));
                const F2 d = fma2(ddx, ddx, fma2(ddy, ddy, f2(-(float)kRectR2)));
                clr[2 * q] = fminf(clr[2 * q], lo(d));
                clr[2 * q + 1] = fminf(clr[2 * q + 1], hi(d));
            }
        }
    }
}

// the termination ladder of finish_env<float> for one env (aqua.py:194-211)
__device__ __forceinline__ uint32_t ladder_f32(const EnvState<float>& e, float clearance, float dc2, float step_reward,
                                               float& reward) {
    const float xn = e.x, yn = e.y;
    bool collided = !((fminf(xn, yn) >= 2.5f) & (fmaxf(xn, yn) <= 97.5f));
    collided |= clearance <= 0.f;
    const bool reached = dc2 <= 25.f;
    const bool late = (e.tw & kTimeMask) > 1000u;
    const uint32_t f_goal = reached ? (AQUA_FLAG_DONE | AQUA_FLAG_SUCCESS) : 0u;
    const uint32_t f_time = late ? (AQUA_FLAG_DONE | AQUA_FLAG_TIME) : f_goal;
    const float r_goal = reached ? 10.f : step_reward;
    reward = (collided | late) ? -10.f : r_goal;
    return collided ? (AQUA_FLAG_DONE | AQUA_FLAG_COLLIDED) : f_time;
}

// finish_env<float> for two envs; returns flags(A) | flags(B) << 8
__device__ __forceinline__ uint32_t finish_pair(const EnvState<float>& A, const EnvState<float>& B,
                                                const MoveTmp<float>& ma, const MoveTmp<float>& mb, float clra,
                                                float clrb, float& rewa, float& rewb) {
    using P = Prec<float>;
    const F2 gx = f2(A.gx, B.gx), gy = f2(A.gy, B.gy);
    const F2 gcx = f2(A.x, B.x) - gx, gcy = f2(A.y, B.y) - gy;
    const F2 dc2 = fma2(gcx, gcx, gcy * gcy);
    const F2 gpx = f2(ma.px, mb.px) - gx, gpy = f2(ma.py, mb.py) - gy;
    const F2 dp2 = fma2(gpx, gpx, gpy * gpy);
    const F2 dx = f2(ma.dx, mb.dx), dy = f2(ma.dy, mb.dy);
    // -num of finish_env: fma(dx, fma(2, gpx, dx), dy * fma(2, gpy, dy))
    const F2 neg_num = fma2(dx, fma2(f2(2.f), gpx, dx), dy * fma2(f2(2.f), gpy, dy));
    const F2 den = f2(P::sqrt_(lo(dp2)), P::sqrt_(hi(dp2))) + f2(P::sqrt_(lo(dc2)), P::sqrt_(hi(dc2)));
    // (0.7 * num) * rcp(den) with num = -neg_num: 0.7 * (-x) and (-0.7) * x round alike
    const F2 sr = (f2(-0.7f) * neg_num) * f2(P::rcp_(lo(den)), P::rcp_(hi(den)));
    const uint32_t fa = ladder_f32(A, clra, lo(dc2), lo(sr), rewa);
    const uint32_t fb = ladder_f32(B, clrb, hi(dc2), hi(sr), rewb);
    return fa | (fb << 8);
}

// wave increments of V consecutive envs (ids env_id .. env_id + V - 1): draw_wave_noise() for
// each, the ten rounds of the V counters interleaved (V independent multiply chains in flight)
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
#if AQUA_PACKED
        // draw_noise() of both words in one FFMA2
        const F2 d = fma2(f2(u24(c0[j]), u24(c1[j])), f2(0.002f * 0x1p-24f), f2(-0.001f));
        nx[j] = lo(d);
        ny[j] = hi(d);
#else
        nx[j] = draw_noise(c0[j]);
        ny[j] = draw_noise(c1[j]);
#endif
    }
}

// One env.step() (aqua.py:135-211) for the V envs of a thread: drives and wave increments in,
// new state in place, rewards out; returns the flags word (byte j = env j).
// `small` / `small_mode`: World::small and World::small_mode of the launch (float only; 0 = use `tab`).
// TAB: the obstacle test the kernel is compiled for -- 0 brute force over `tab`, 1 the small table,
// 2 the culling grid (step_many_kernel: one instantiation per kind, so that its step loop holds one
// test only) -- or -1: chosen per launch.
template <typename R, int V, int TAB = -1>
__device__ __forceinline__ uint32_t step_envs(EnvState<R> (&e)[V], const Drive<R> (&dr)[V], const float (&nx)[V],
                                              const float (&ny)[V], bool waves, const ObstacleView<R>& tab,
                                              int k_circ, int k_obst, R (&reward)[V], const SmallTable& small,
                                              int small_mode) {
    MoveTmp<R> m[V];
    R pxn[V], pyn[V], clr[V];
    uint32_t fw = 0u;
    if constexpr (AQUA_PACKED && sizeof(R) == 4 && V % 2 == 0) {
#pragma unroll
        for (int q = 0; q < V / 2; ++q)
            move_pair(e[2 * q], e[2 * q + 1], dr[2 * q], dr[2 * q + 1], nx[2 * q], ny[2 * q], nx[2 * q + 1], ny[2 * q + 1],
                      waves, m[2 * q], m[2 * q + 1]);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            pxn[j] = e[j].x;
            pyn[j] = e[j].y;
            clr[j] = kClearFar;
        }
        if constexpr (TAB == 1) {
            obstacle_clearance_small_pairs<V>(small, small_mode, pxn, pyn, clr);
        } else if constexpr (TAB == 2) {
            obstacle_clearance_grid<R, V>(tab, k_circ, pxn, pyn, clr);
        } else if constexpr (TAB == 0) {
            obstacle_clearance_pairs<V>(tab, k_circ, k_obst, pxn, pyn, clr);
        } else {
            if (small_mode != 0) obstacle_clearance_small_pairs<V>(small, small_mode, pxn, pyn, clr);
            else if (tab.cell16 != nullptr) obstacle_clearance_grid<R, V>(tab, k_circ, pxn, pyn, clr);
            else obstacle_clearance_pairs<V>(tab, k_circ, k_obst, pxn, pyn, clr);
        }
#pragma unroll
        for (int q = 0; q < V / 2; ++q)
            fw |= finish_pair(e[2 * q], e[2 * q + 1], m[2 * q], m[2 * q + 1], clr[2 * q], clr[2 * q + 1], reward[2 * q],
                              reward[2 * q + 1])
                  << (16 * q);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            move_env<R>(e[j], dr[j], (R)nx[j], (R)ny[j], waves, m[j]);
            pxn[j] = e[j].x;
            pyn[j] = e[j].y;
            clr[j] = R(kClearFar);
        }
        if constexpr (sizeof(R) == 4 && TAB == 1) {
            obstacle_clearance_small<V>(small, small_mode, pxn, pyn, clr);
        } else if constexpr (TAB == 2) {
            obstacle_clearance_any<R, V, 1>(tab, k_circ, k_obst, pxn, pyn, clr);
        } else if constexpr (TAB == 0) {
            obstacle_clearance_any<R, V, 0>(tab, k_circ, k_obst, pxn, pyn, clr);
        } else if constexpr (sizeof(R) == 4) {
            if (small_mode != 0) obstacle_clearance_small<V>(small, small_mode, pxn, pyn, clr);
            else obstacle_clearance_any<R, V>(tab, k_circ, k_obst, pxn, pyn, clr);
        } else {
            obstacle_clearance_any<R, V>(tab, k_circ, k_obst, pxn, pyn, clr);
        }
#pragma unroll
        for (int j = 0; j < V; ++j) fw |= finish_env<R>(e[j], m[j], clr[j], reward[j]) << (8 * j);
    }
    return fw;
}

}  // namespace aqua
