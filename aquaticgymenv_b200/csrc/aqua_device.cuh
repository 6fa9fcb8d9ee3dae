// aqua_device.cuh -- device-side primitives of the batched AquaGym step (sm_100a).
//
// Compiled with -fmad=false: every fused multiply-add below is written out as fma(), so the
// arithmetic (and hence every trajectory) is the same in all kernels that inline these functions
// and does not depend on the compiler's contraction choices.
//
// One env instance is advanced by one call of step_env<R>() with its whole state in
// registers.  The kinematics use the closed form of the reference's ICC rotation
// (gym_aqua/envs/aqua.py:159-183; derivation in DESIGN.md section 3) instead of the reference's
// "rotate about a centre 1.25e8 px away" formulation, which is unusable in float32.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/aquagym.h"
#include "aqua_kernels.cuh"

namespace aqua {

constexpr uint32_t kTimeMask = (1u << AQUA_TIME_BITS) - 1u;  // 2047: time saturates here
constexpr uint32_t kEpisodeMask = AQUA_EPISODE_MASK;
constexpr uint32_t kResetMaxTries = 4096u;

enum : uint32_t { kStreamReset = 1u, kStreamWaveInit = 2u, kStreamWaveNoise = 3u, kStreamAction = 4u };

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  key = seed, counter = (env id, c1, c2, stream tag).
// The ten round keys (k0 + r*W0, k1 + r*W1) are computed once on the host and travel in the
// kernel parameter block, so that they are constant-bank operands of the round's XORs.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               const PhiloxKeys& k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k.rk4[2 * r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k.rk4[2 * r + 1];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint4 philox_env(const PhiloxKeys& k, uint32_t env_id, uint32_t c1, uint32_t c2,
                                            uint32_t tag) {
    return philox4x32_10(env_id, c1, c2, tag, k);
}

// Philox2x32-10: the per-step draws need two words only (half the multiplies of the 4x32 form).
// `rk` = the stream's ten round keys; `key_add` shifts the key (ACTION stream: high step word).
__device__ __forceinline__ uint2 philox2x32_10(uint32_t c0, uint32_t c1, const uint32_t (&rk)[10],
                                               uint32_t key_add = 0u) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p = (unsigned long long)0xD256D193u * c0;
        const uint32_t n0 = (uint32_t)(p >> 32) ^ c1 ^ (rk[r] + key_add);
        c1 = (uint32_t)p;
        c0 = n0;
    }
    return make_uint2(c0, c1);
}

// 24 high bits -> float in [0,1): u01 = (u >> 8) * 2^-24, exact.  The draw maps below fold the
// power-of-two factor into their constant, which is bit-identical (the product is the same real
// number, rounded once) and saves a multiply per draw.
__device__ __forceinline__ float u24(uint32_t u) { return __uint2float_rn(u >> 8); }

// The draws are float32 by definition: every instantiation and the CPU
// oracle consume bit-identical values.
constexpr float kPiDnF = 3.1415925f;     // largest float below pi
constexpr float kTwoPiDnF = 6.283185f;   // 2 * kPiDnF exactly
__device__ __forceinline__ float draw_pos(uint32_t u) { return __fmul_rn(u24(u), 100.0f * 0x1p-24f); }
__device__ __forceinline__ float draw_angle(uint32_t u) { return __fmaf_rn(u24(u), kTwoPiDnF * 0x1p-24f, -kPiDnF); }
__device__ __forceinline__ float draw_wave0(uint32_t u) { return __fmaf_rn(u24(u), 0.1f * 0x1p-24f, -0.05f); }
__device__ __forceinline__ float draw_noise(uint32_t u) { return __fmaf_rn(u24(u), 0.002f * 0x1p-24f, -0.001f); }
__device__ __forceinline__ float draw_thrust(uint32_t u) { return __fmaf_rn(u24(u), 0.3f * 0x1p-24f, 0.2f); }

// wave increment of the step taken at time word `tw` = (episode << 11) | t -- aqua.py:188
__device__ __forceinline__ void draw_wave_noise(const PhiloxKeys& keys, uint32_t env_id, uint32_t tw,
                                                float& nx, float& ny) {
    const uint2 r = philox2x32_10(env_id, tw, keys.noise);
    nx = draw_noise(r.x);
    ny = draw_noise(r.y);
}
// wave speed of a fresh episode -- aqua.py:124
__device__ __forceinline__ void draw_wave_init(const PhiloxKeys& keys, uint32_t env_id, uint32_t ep,
                                               float& wx, float& wy) {
    const uint2 r = philox2x32_10(env_id, ep, keys.winit);
    wx = draw_wave0(r.x);
    wy = draw_wave0(r.y);
}
// synthetic random policy (test_random.py:19-20)
__device__ __forceinline__ uint2 draw_action_words(const PhiloxKeys& keys, uint32_t env_id, uint64_t step) {
    return philox2x32_10(env_id, (uint32_t)step, keys.action, (uint32_t)(step >> 32));
}

// ------------------------------------------------------------------------------------------
// precision traits
// ------------------------------------------------------------------------------------------
#ifndef AQUA_VF32
#define AQUA_VF32 4
#endif
template <typename R> struct Prec;
template <> struct Prec<float> {
    static constexpr int V = AQUA_VF32;  // envs per thread: 16-byte plane loads at V = 4
    // sin and cos of a heading: |a| <= pi + 0.06, so two quadrant-reduction terms suffice and no
    // large-argument path is needed.  Cephes single-precision minimax kernels on [-pi/4, pi/4]
    // (about 1 ulp); quadrant fix-up on the sign bits.
    static __device__ __forceinline__ void sincos_(float a, float* s, float* c) {
        const float t = fmaf(a, 0.636619772f, 12582912.0f);      // 1.5 * 2^23: round to nearest integer
        const uint32_t q = __float_as_uint(t);                   // low bits = quadrant (two's complement)
        const float fq = t - 12582912.0f;
        float r = fmaf(fq, -1.57079637f, a);                     // pi/2 = 1.57079637 - 4.37113883e-8
        r = fmaf(fq, 4.37113883e-8f, r);
        const float r2 = r * r;
        const float ps = fmaf(r2, fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f);
        const float sr = fmaf(r * r2, ps, r);
        const float pc = fmaf(r2, fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f), 4.166664568298827e-2f);
        const float cr = fmaf(r2 * r2, pc, fmaf(r2, -0.5f, 1.0f));
        const bool swap = (q & 1u) != 0u;
        const float ss = swap ? cr : sr, cc = swap ? sr : cr;
        *s = __uint_as_float(__float_as_uint(ss) ^ ((q & 2u) << 30));
        *c = __uint_as_float(__float_as_uint(cc) ^ (((q + 1u) & 2u) << 30));
    }
    static __device__ __forceinline__ float sqrt_(float v) {
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
        return r;
    }
    static __device__ __forceinline__ float rcp_(float v) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
        return r;
    }
};
template <> struct Prec<double> {
    static constexpr int V = 2;
    static __device__ __forceinline__ void sincos_(double a, double* s, double* c) { sincos(a, s, c); }
    static __device__ __forceinline__ double sqrt_(double v) { return sqrt(v); }
    static __device__ __forceinline__ double rcp_(double v) { return 1.0 / v; }
};

// Obstacle table in shared memory (aqua.py:56-68, 373-390).  Circles come first (k < k_circ):
// entry = {cx, cy, r2, -} with r2 = (radius + 2.5)^2; then rectangles: {cx, cy, hx, hy} (half
// extents), whose clearance is the boat / goal radius 2.5 alone (kRectR2).  The distance of a
// point to an axis-aligned box is the norm of max(|p - c| - h, 0); a circle is the box with zero
// extents, so both tests are the reference's `distance <= 0` written on squared distances.
// All lanes read the same entry (one 16-byte broadcast LDS per obstacle, conflict free).
template <typename R> struct alignas(16) ObstacleBox {
    R cx, cy, a, b;
};
// Dense maps (more than kGridMinObstacles obstacles) are culled through a uniform grid over the
// world.  A cell lists the obstacles whose clearance region (obstacle inflated by the radius 2.5,
// plus a safety margin) touches it, so only those are tested exactly; a cell that lies entirely
// inside one obstacle's clearance region (shrunk by the margin) is marked blocked: every point of
// it collides, no test needed.  Both classifications are made on the host in float64 with a
// 0.05 px margin, hence conservative: the collision decision is unchanged.
//
// What the kernels read per position is one 16-bit word: the first two listed obstacles (7 bits
// each, kCellNone = none), the blocked bit and an overflow bit.  Both listed obstacles are always
// tested, in one unified form (a circle is a box of zero extents), kCellNone being a table entry
// that never collides: straight-line code, no divergence between the lanes of a warp, whichever
// cells they sit in.  With 96 x 96 cells (1.04 px) fewer than one position in a thousand of the
// 64-obstacle stress map lies in a cell that lists more than two obstacles (overflow bit): those
// lanes walk the cell's full 64-bit mask from global memory.
constexpr int kGridDim = 96;
constexpr int kGridCells = kGridDim * kGridDim;
constexpr int kGridMinObstacles = 12;
constexpr unsigned long long kCellBlocked = ~0ull;
constexpr uint32_t kCellNone = AQUA_MAX_OBSTACLES;  // index of the entry that never collides
constexpr uint32_t kCell16Blocked = 0x4000u, kCell16Overflow = 0x8000u;
// host / global-memory layout
template <typename R> struct ObstacleTable {
    ObstacleBox<R> box[AQUA_MAX_OBSTACLES];        // brute-force layout (see ObstacleBox)
    ObstacleBox<R> ubox[AQUA_MAX_OBSTACLES + 1];   // unified {cx, cy, half_x, half_y}; [kCellNone] never collides
    R nr2[AQUA_MAX_OBSTACLES + 4];                 // -(reach^2) of the unified entries
    unsigned short cell16[kGridCells];             // id0 | id1 << 7 | blocked | overflow
    unsigned long long cell[kGridCells];           // full masks (kCellBlocked = blocked); global memory only
};
// What the kernels read.  Brute force (cell16 == NULL): `box` = the table's box[] in shared memory.
// Culling grid: `box` = the unified entries, `nr2`, `cell16` in shared memory (cell16 at the end of
// the dynamic shared memory, see cell_bytes()), `cell` in global memory.
template <typename R> struct ObstacleView {
    const ObstacleBox<R>* box;
    const R* nr2;
    const unsigned short* cell16;
    const unsigned long long* cell;
};
constexpr double kRectR2 = 6.25;  // (boat or goal radius 2.5)^2, aqua.py:72,75

// Signed squared clearance of V positions against the whole table, min-accumulated: an env
// collides iff the minimum is <= 0.  The table entry is read once per obstacle.
template <typename R, int V>
__device__ __forceinline__ void obstacle_clearance(const ObstacleView<R>& tab, int k_circ, int k_obst,
                                                   const R (&px)[V], const R (&py)[V], R (&clr)[V]) {
    // Four circles / two rectangles per trip.  The counts are runtime values; a trip that runs past the
    // last entry tests that entry again instead (clamped index): min() is idempotent, so the result
    // is the same bit for bit, and no remainder loop is needed.
#pragma unroll 1
    for (int k = 0; k < k_circ; k += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const ObstacleBox<R> o = tab.box[min(k + u, k_circ - 1)];
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const R ddx = px[j] - o.cx, ddy = py[j] - o.cy;
                clr[j] = fmin(clr[j], fma(ddx, ddx, fma(ddy, ddy, -o.a)));
            }
        }
    }
#pragma unroll 1
    for (int k = k_circ; k < k_obst; k += 2) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const ObstacleBox<R> o = tab.box[min(k + u, k_obst - 1)];
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const R ddx = fmax(fabs(px[j] - o.cx) - o.a, R(0));
                const R ddy = fmax(fabs(py[j] - o.cy) - o.b, R(0));
                clr[j] = fmin(clr[j], fma(ddx, ddx, fma(ddy, ddy, R(-kRectR2))));
            }
        }
    }
}
constexpr float kClearFar = 1e30f;  // initial clearance: "no obstacle near"

// squared clearance of one position against one unified entry
template <typename R>
__device__ __forceinline__ R unified_clearance(const ObstacleBox<R>& o, R nr2, R px, R py) {
    const R ddx = fmax(fabs(px - o.cx) - o.a, R(0));
    const R ddy = fmax(fabs(py - o.cy) - o.b, R(0));
    return fma(ddx, ddx, fma(ddy, ddy, nr2));
}

// same result (sign of the minimum clearance) through the culling grid
template <typename R, int V>
__device__ __forceinline__ void obstacle_clearance_grid(const ObstacleView<R>& tab, int k_circ,
                                                        const R (&px)[V], const R (&py)[V], R (&clr)[V]) {
    (void)k_circ;
    uint32_t over = 0u;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int cx = min(max((int)((float)px[j] * (kGridDim / 100.0f)), 0), kGridDim - 1);
        const int cy = min(max((int)((float)py[j] * (kGridDim / 100.0f)), 0), kGridDim - 1);
        const uint32_t c = tab.cell16[cy * kGridDim + cx];
        const uint32_t i0 = c & 127u, i1 = (c >> 7) & 127u;
        const R d0 = unified_clearance<R>(tab.box[i0], tab.nr2[i0], px[j], py[j]);
        const R d1 = unified_clearance<R>(tab.box[i1], tab.nr2[i1], px[j], py[j]);
        const R d = fmin(clr[j], fmin(d0, d1));
        clr[j] = (c & kCell16Blocked) ? R(-1) : d;  // blocked: inside an obstacle's clearance region for sure
        over |= (c & kCell16Overflow) ? (1u << j) : 0u;
    }
    if (over != 0u) {  // rare: a cell that lists more than two obstacles
#pragma unroll
        for (int j = 0; j < V; ++j) {
            if (!((over >> j) & 1u)) continue;
            const int cx = min(max((int)((float)px[j] * (kGridDim / 100.0f)), 0), kGridDim - 1);
            const int cy = min(max((int)((float)py[j] * (kGridDim / 100.0f)), 0), kGridDim - 1);
            unsigned long long mask = __ldg(tab.cell + cy * kGridDim + cx);
            while (mask != 0ull) {
                const int k = __ffsll((long long)mask) - 1;
                mask &= mask - 1ull;
                clr[j] = fmin(clr[j], unified_clearance<R>(tab.box[k], tab.nr2[k], px[j], py[j]));
            }
        }
    }
}

// brute force for the small tables of the reference, grid for dense maps.  GRID: 1 / 0 = the kernel is
// compiled for launches with / without the culling grid (step_many_kernel), -1 = decided per launch
template <typename R, int V, int GRID = -1>
__device__ __forceinline__ void obstacle_clearance_any(const ObstacleView<R>& tab, int k_circ, int k_obst,
                                                       const R (&px)[V], const R (&py)[V], R (&clr)[V]) {
    if constexpr (GRID == 1) {
        obstacle_clearance_grid<R, V>(tab, k_circ, px, py, clr);
    } else if constexpr (GRID == 0) {
        obstacle_clearance<R, V>(tab, k_circ, k_obst, px, py, clr);
    } else {
        if (tab.cell16 != nullptr) obstacle_clearance_grid<R, V>(tab, k_circ, px, py, clr);
        else obstacle_clearance<R, V>(tab, k_circ, k_obst, px, py, clr);
    }
}

// same test against the float64 table in global memory (reset path: rare, always float64 so
// that accept/reject decisions are those of the float64 reference)
__device__ __forceinline__ bool hits_obstacle_g(const double* __restrict__ tab64, int k_obst, double px,
                                                double py) {
    for (int k = 0; k < k_obst; ++k) {
        const double* o = tab64 + 5 * k;
        const double ddx = fmax(fabs(px - o[0]) - o[2], 0.0);
        const double ddy = fmax(fabs(py - o[1]) - o[3], 0.0);
        if (fma(ddx, ddx, fma(ddy, ddy, -o[4])) <= 0.0) return true;
    }
    return false;
}

// aqua.py:424-427 / 453-455 (strict inequalities; boat and goal radius are both 2.5)
__device__ __forceinline__ bool outside_border(double px, double py) {
    return (px - 2.5 < 0.0) || (py - 2.5 < 0.0) || (px + 2.5 > 100.0) || (py + 2.5 > 100.0);
}

struct WorldFlags {
    bool waves, random_boat, random_goal;
};

struct ResetResult {
    double x, y, th, wx, wy, gx, gy;
};

// env.reset() for one env: aqua.py:100-126.  `ep` is the NEW episode index.  Try n of the
// reference's `observation_space.sample()` sequence is Philox(env, ep, n, RESET).
// A rejection loop that reaches try kResetMaxTries - 1 accepts it untested (the reference would loop
// on: aqua.py:102-115 have no bound); every such accept is counted in *forced (AQUA_STAT_FORCED_RESETS).
inline __device__ __noinline__ ResetResult reset_env(const PhiloxKeys& keys, uint32_t env_id, uint32_t ep,
                                              const double* __restrict__ tab64, int k_obst,
                                              WorldFlags wf, unsigned int* forced) {
    ResetResult o;
    uint32_t n = 0;
    if (wf.random_goal) {  // aqua.py:102-105
        for (;;) {
            const uint4 r = philox_env(keys, env_id, ep, n++, kStreamReset);
            o.gx = (double)draw_pos(r.z);
            o.gy = (double)draw_pos(r.w);
            if (n >= kResetMaxTries) { atomicAdd(forced, 1u); break; }
            if (!(outside_border(o.gx, o.gy) || hits_obstacle_g(tab64, k_obst, o.gx, o.gy))) break;
        }
    } else {  // aqua.py:107
        o.gx = 25.0;
        o.gy = 80.0;
    }
    if (wf.random_boat) {  // aqua.py:110-115
        for (;;) {
            const uint4 r = philox_env(keys, env_id, ep, n++, kStreamReset);
            o.x = (double)draw_pos(r.x);
            o.y = (double)draw_pos(r.y);
            o.th = (double)draw_angle(r.z);
            if (n >= kResetMaxTries) { atomicAdd(forced, 1u); break; }
            const double dgx = o.gx - o.x, dgy = o.gy - o.y;
            const bool reached = sqrt(fma(dgx, dgx, dgy * dgy)) - 5.0 <= 0.0;  // aqua.py:421-422
            if (!(reached || outside_border(o.x, o.y) || hits_obstacle_g(tab64, k_obst, o.x, o.y))) break;
        }
    } else {  // aqua.py:117
        o.x = 85.0;
        o.y = 45.0;
        o.th = 0.0;
    }
    if (wf.waves) {  // aqua.py:124
        float wx, wy;
        draw_wave_init(keys, env_id, ep, wx, wy);
        o.wx = (double)wx;
        o.wy = (double)wy;
    } else {
        o.wx = 0.0;
        o.wy = 0.0;
    }
    return o;
}

// ------------------------------------------------------------------------------------------
// Cooperative reset (in-kernel auto-reset).  Finished envs are rare (about 1 in 70 env-steps) and
// a rejection loop run by the owning lane alone would serialise the whole warp behind it.  The
// step kernels therefore queue finished envs per CTA and hand each one to a group of
// kCoopLanes lanes that evaluates kCoopLanes consecutive tries of the reference's sampling
// sequence at once: lane l of round q tests try n = 4q + l (one Philox call yields both the goal
// candidate, words z/w, and the boat candidate, words x/y/z of `observation_space.sample()`).
// The accepted goal is the first passing try g, the accepted boat the first passing try > g
// (aqua.py:102-115): exactly the sequential result, found in the first round about 19 times in 20.
//
// Accept/reject decisions are those of float64 arithmetic on the float32 draws: the float table
// is used with a guard band and anything inside the band is re-evaluated in float64.
// ------------------------------------------------------------------------------------------
constexpr int kCoopLanes = 4;

__device__ __forceinline__ bool outside_border_f(float x, float y) {
    // (double)x - 2.5 and (double)x + 2.5 are exact for a float x in [0, 128): same predicate
    return !((x >= 2.5f) & (x <= 97.5f) & (y >= 2.5f) & (y <= 97.5f));
}

// `band` = guard band of the float32 clearance (1e-4 x the largest squared reach of the table,
// far above its rounding error): outside the band the float32 sign is the float64 sign.
template <typename R, int GRID = -1>
__device__ __forceinline__ bool reset_blocked(const ObstacleView<R>& tab, const double* __restrict__ tab64,
                                              int k_circ, int k_obst, float band, float x, float y) {
    const R px[1] = {(R)x}, py[1] = {(R)y};
    R clr[1] = {R(kClearFar)};
    obstacle_clearance_any<R, 1, GRID>(tab, k_circ, k_obst, px, py, clr);
    if constexpr (sizeof(R) == 4) {
        if (fabsf(clr[0]) <= band) return hits_obstacle_g(tab64, k_obst, (double)x, (double)y);
    }
    return clr[0] <= R(0);
}

// aqua.py:421-422 at reset: goal and boat radius 2.5 each
template <typename R> __device__ __forceinline__ bool reset_reached(float gx, float gy, float x, float y) {
    if constexpr (sizeof(R) == 4) {
        const float dgx = gx - x, dgy = gy - y;
        const float sd = fmaf(dgx, dgx, dgy * dgy) - 25.f;
        if (fabsf(sd) > 2.5e-3f) return sd <= 0.f;
    }
    const double dgx = (double)gx - (double)x, dgy = (double)gy - (double)y;
    return sqrt(fma(dgx, dgx, dgy * dgy)) - 5.0 <= 0.0;
}

struct CoopResetResult {
    float x, y, th, wx, wy, gx, gy;  // float32 draws: exact in either precision
};

// Must be called by all 32 lanes of a warp (it uses full-mask ballots and shuffles).  Lanes
// [4s, 4s+4) serve one env: `valid`, `env_id`, `ep` (the NEW episode index) are group-uniform.
// The result is returned to every lane of the group.
template <typename R>
__device__ __forceinline__ CoopResetResult coop_reset(const PhiloxKeys& keys, const ObstacleView<R>& tab,
                                                      const double* __restrict__ tab64, int k_circ, int k_obst,
                                                      float band, WorldFlags wf, bool valid, uint32_t env_id,
                                                      uint32_t ep, unsigned int* forced) {
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned sub = lane & ~(unsigned)(kCoopLanes - 1);  // first lane of this group
    const unsigned l = lane - sub;
    CoopResetResult o;
    o.gx = 25.f; o.gy = 80.f;            // aqua.py:107
    o.x = 85.f; o.y = 45.f; o.th = 0.f;  // aqua.py:117
    bool goal_done = !wf.random_goal || !valid;
    bool boat_done = !wf.random_boat || !valid;
    uint32_t n0 = 0;
    while (__any_sync(full, !(goal_done && boat_done))) {
        const uint32_t n = n0 + l;
        const uint4 r = philox_env(keys, env_id, ep, n, kStreamReset);
        const bool force = n >= kResetMaxTries - 1u;  // bounded loop: this try is accepted as is
        // goal candidate of try n: aqua.py:103-105
        const float cgx = draw_pos(r.z), cgy = draw_pos(r.w);
        bool ok_g = false;
        if (!goal_done)
            ok_g = force || !(outside_border_f(cgx, cgy) || reset_blocked<R>(tab, tab64, k_circ, k_obst, band, cgx, cgy));
        const unsigned mg = (__ballot_sync(full, ok_g) >> sub) & ((1u << kCoopLanes) - 1u);
        unsigned first_boat = goal_done ? 0u : (unsigned)kCoopLanes;  // first lane whose try follows the goal's
        const unsigned lg = mg ? (unsigned)(__ffs(mg) - 1) : 0u;
        const float sgx = __shfl_sync(full, cgx, sub + lg), sgy = __shfl_sync(full, cgy, sub + lg);
        if (!goal_done && mg) {
            if (force && l == lg) atomicAdd(forced, 1u);  // the forced try is the accepted one: no try passed
            o.gx = sgx;
            o.gy = sgy;
            goal_done = true;
            first_boat = lg + 1u;
        }
        // boat candidate of try n: aqua.py:110-115
        const float cbx = draw_pos(r.x), cby = draw_pos(r.y), cbt = draw_angle(r.z);
        bool ok_b = false;
        if (goal_done && !boat_done && l >= first_boat)
            ok_b = force || !(reset_reached<R>(o.gx, o.gy, cbx, cby) || outside_border_f(cbx, cby) ||
                              reset_blocked<R>(tab, tab64, k_circ, k_obst, band, cbx, cby));
        const unsigned mb = (__ballot_sync(full, ok_b) >> sub) & ((1u << kCoopLanes) - 1u);
        const unsigned lb = mb ? (unsigned)(__ffs(mb) - 1) : 0u;
        const float sbx = __shfl_sync(full, cbx, sub + lb), sby = __shfl_sync(full, cby, sub + lb);
        const float sbt = __shfl_sync(full, cbt, sub + lb);
        if (mb) {
            if (force && l == lb && !boat_done) atomicAdd(forced, 1u);
            o.x = sbx;
            o.y = sby;
            o.th = sbt;
            boat_done = true;
        }
        n0 += kCoopLanes;
    }
    if (wf.waves) {  // aqua.py:124
        draw_wave_init(keys, env_id, ep, o.wx, o.wy);
    } else {
        o.wx = 0.f;
        o.wy = 0.f;
    }
    return o;
}

// ------------------------------------------------------------------------------------------
// One lane resets one env: the reference's two rejection loops in sequence (aqua.py:100-126), try n
// = Philox(env, ep, n, RESET), exactly the result of reset_env() and coop_reset().  No warp
// collectives: the lanes of a warp each serve another env and leave their loops when their own
// tries pass, so a pass over 32 queued envs costs as many rounds as its unluckiest lane (about
// 5 on the default map) where 32 envs take four coop_reset() passes of 8.  Used wherever finished
// envs can be collected 32 at a time: the streaming / general step kernels and the stash refill of
// the multi-step kernel.  Accept/reject decisions are float64-exact (guard band, see above).
// ------------------------------------------------------------------------------------------
template <typename R, int GRID = -1>
__device__ __forceinline__ CoopResetResult reset_env_lane(const PhiloxKeys& keys, const ObstacleView<R>& tab,
                                                          const double* __restrict__ tab64, int k_circ, int k_obst,
                                                          float band, WorldFlags wf, uint32_t env_id, uint32_t ep,
                                                          unsigned int* forced) {
    CoopResetResult o;
    o.gx = 25.f; o.gy = 80.f;            // aqua.py:107
    o.x = 85.f; o.y = 45.f; o.th = 0.f;  // aqua.py:117
    uint32_t n = 0;
    if (wf.random_goal) {  // aqua.py:102-105
        for (;;) {
            const uint4 r = philox_env(keys, env_id, ep, n++, kStreamReset);
            o.gx = draw_pos(r.z);
            o.gy = draw_pos(r.w);
            if (n >= kResetMaxTries) { atomicAdd(forced, 1u); break; }
            if (!(outside_border_f(o.gx, o.gy) || reset_blocked<R, GRID>(tab, tab64, k_circ, k_obst, band, o.gx, o.gy))) break;
        }
    }
    if (wf.random_boat) {  // aqua.py:110-115
        for (;;) {
            const uint4 r = philox_env(keys, env_id, ep, n++, kStreamReset);
            o.x = draw_pos(r.x);
            o.y = draw_pos(r.y);
            o.th = draw_angle(r.z);
            if (n >= kResetMaxTries) { atomicAdd(forced, 1u); break; }
            if (!(reset_reached<R>(o.gx, o.gy, o.x, o.y) || outside_border_f(o.x, o.y) ||
                  reset_blocked<R, GRID>(tab, tab64, k_circ, k_obst, band, o.x, o.y)))
                break;
        }
    }
    if (wf.waves) {  // aqua.py:124
        draw_wave_init(keys, env_id, ep, o.wx, o.wy);
    } else {
        o.wx = 0.f;
        o.wy = 0.f;
    }
    return o;
}

// ------------------------------------------------------------------------------------------
// one env step, state in registers
// ------------------------------------------------------------------------------------------
template <typename R> struct EnvState {
    R x, y, th, wx, wy, gx, gy;
    uint32_t tw;  // (episode << 11) | time
};

// heading wrap to [-pi, pi): aqua.py:128-133
__device__ __forceinline__ float wrap_angle(float a) {
    // |a| <= pi + 0.12.  kPiUp = 3.14159274f is the float next above pi and 3.1415925f the one next
    // below, so "a >= pi" reads a >= kPiUp and "a < -pi" reads a <= -kPiUp (-kPiUp itself lies below
    // -pi and wraps to just under +pi, as the reference's [-pi, pi) range wants); 2*pi is
    // subtracted as hi + lo.
    constexpr float kPiUp = 3.14159274f, kTwoPiHi = 6.28318548f, kTwoPiLo = -1.74845553e-7f;
    if (a >= kPiUp) a = (a - kTwoPiHi) - kTwoPiLo;
    else if (a <= -kPiUp) a = (a + kTwoPiHi) + kTwoPiLo;
    return a;
}
__device__ __forceinline__ double wrap_angle(double a) {
    // the reference's own expression, same constants and order
    constexpr double kLow = -3.141592653589793, kHigh = 3.141592653589793;
    const double width = kHigh - kLow;
    const double off = a - kLow;
    return (off - (floor(off / width) * width)) + kLow;
}

template <typename R> struct MoveTmp {
    R px, py, dx, dy;  // position before the move, displacement
};

// The part of the move that depends on the thrusts only: turn angle phi = w * tau, half angle h
// and chord length v * tau * sin(h)/h of the arc (aqua.py:159-170 in closed form).
template <typename R> struct Drive {
    R phi, h, chord;
};

template <typename R> __device__ __forceinline__ Drive<R> drive_from_thrusts(R tl, R tr) {
    constexpr bool kF64 = sizeof(R) == 8;
    R d = tr - tl;
    if (fabs(d) < R(1e-8)) d = copysign(R(1e-8), d);  // aqua.py:159-160 (epsilon sentinel)
    Drive<R> o;
    R sinc;
    if constexpr (kF64) o.phi = d / R(2.5);           // aqua.py:168
    else o.phi = d * R(0.4);
    const R v = (tr + tl) * R(0.5);                   // |w * r|,  aqua.py:166-170
    o.h = o.phi * R(0.5);
    const R h2 = o.h * o.h;
    // sin(h)/h for |h| <= 0.06
    if constexpr (kF64)
        sinc = fma(h2, fma(h2, fma(h2, fma(h2, R(1.0 / 362880), R(-1.0 / 5040)), R(1.0 / 120)), R(-1.0 / 6)), R(1));
    else
        sinc = fma(h2, fma(h2, R(1.0 / 120), R(-1.0 / 6)), R(1));
    o.chord = v * sinc;
    return o;
}

// Discrete table aqua.py:33-42: 0 -> (0.2, 0.5) rotate left, 1 -> (0.5, 0.2) rotate right (the
// exact mirror image of action 0: same chord, opposite angles), 2 -> (0.5, 0.5) full throttle.
// The three drives are compile-time constants of the same arithmetic as drive_from_thrusts().
// An index outside {0,1,2} selects full throttle (and is counted by the caller).
template <typename R> __device__ __forceinline__ Drive<R> drive_from_index(uint32_t a) {
    const Drive<R> turn = drive_from_thrusts<R>(R(0.2), R(0.5));
    const Drive<R> ahead = drive_from_thrusts<R>(R(0.5), R(0.5));
    Drive<R> o;
    o.phi = (a == 0u) ? turn.phi : ((a == 1u) ? -turn.phi : ahead.phi);
    o.h = (a == 0u) ? turn.h : ((a == 1u) ? -turn.h : ahead.h);
    o.chord = (a < 2u) ? turn.chord : ahead.chord;
    return o;
}

// Phase 1 of step(): drive -> new pose, wave update, time.  (nx, ny) is the wave increment of
// this step.
template <typename R>
__device__ __forceinline__ void move_env(EnvState<R>& e, const Drive<R>& dr, R nx, R ny, bool waves,
                                         MoveTmp<R>& m) {
    using P = Prec<R>;
    const R phi = dr.phi, h = dr.h, chord = dr.chord;
    R sn, cs;
    P::sincos_(e.th + h, &sn, &cs);
    m.dx = fma(-chord, sn, e.wx);  // heading 0 = +y; aqua.py:173-181, tau = 1
    m.dy = fma(chord, cs, e.wy);
    m.px = e.x;
    m.py = e.y;
    e.x = m.px + m.dx;
    e.y = m.py + m.dy;
    e.th = wrap_angle(e.th + phi);  // aqua.py:183

    // ---- waves, aqua.py:188-191 (after the move) ----
    {  // has_waves = 0 clips to [0, 0] (aqua.py:23-25,189-191): selects, no branch
        const R wmax = R(0.05);
        const R wxn = fmin(fmax(e.wx + nx, -wmax), wmax), wyn = fmin(fmax(e.wy + ny, -wmax), wmax);
        e.wx = waves ? wxn : R(0);
        e.wy = waves ? wyn : R(0);
    }

    // ---- time, aqua.py:141 ----
    e.tw += ((e.tw & kTimeMask) != kTimeMask) ? 1u : 0u;  // saturates at 2047
}

// Phase 2: termination ladder aqua.py:194-211.  `clearance` comes from obstacle_clearance().
// Returns the AQUA_FLAG_* byte; `reward` receives the step reward.
template <typename R>
__device__ __forceinline__ uint32_t finish_env(const EnvState<R>& e, const MoveTmp<R>& m, R clearance,
                                               R& reward) {
    using P = Prec<R>;
    constexpr bool kF64 = sizeof(R) == 8;
    const R xn = e.x, yn = e.y;
    bool collided;
    if constexpr (kF64)
        collided = (xn - R(2.5) < R(0)) | (yn - R(2.5) < R(0)) | (xn + R(2.5) > R(100)) | (yn + R(2.5) > R(100));
    else  // float: p - 2.5 and p + 2.5 are exact near the borders, so this is the same predicate
        collided = !((fmin(xn, yn) >= R(2.5)) & (fmax(xn, yn) <= R(97.5)));
    collided |= clearance <= R(0);

    const R gcx = xn - e.gx, gcy = yn - e.gy;
    const R dc2 = fma(gcx, gcx, gcy * gcy);
    const R gpx = m.px - e.gx, gpy = m.py - e.gy;
    const R dp2 = fma(gpx, gpx, gpy * gpy);
    R step_reward;
    bool reached;
    if constexpr (kF64) {  // the reference's expressions: aqua.py:89-90, 392-402, 421-422
        const R dcur = P::sqrt_(dc2) - R(5), dprev = P::sqrt_(dp2) - R(5);
        reached = dcur <= R(0);
        step_reward = (dprev - dcur) * R(0.7);
    } else {
        // d_prev - d_cur = (d_prev^2 - d_cur^2) / (d_prev + d_cur), the numerator taken from the
        // displacement itself: no cancellation of two ~50 px norms in float32
        reached = dc2 <= R(25);
        const R num = -fma(m.dx, fma(R(2), gpx, m.dx), m.dy * fma(R(2), gpy, m.dy));
        const R den = P::sqrt_(dp2) + P::sqrt_(dc2);
        step_reward = (R(0.7) * num) * P::rcp_(den);
    }
    // collision > time > goal > running, written as selects (no divergent branches)
    const bool late = (e.tw & kTimeMask) > 1000u;
    const uint32_t f_goal = reached ? (AQUA_FLAG_DONE | AQUA_FLAG_SUCCESS) : 0u;
    const uint32_t f_time = late ? (AQUA_FLAG_DONE | AQUA_FLAG_TIME) : f_goal;
    const R r_goal = reached ? R(10) : step_reward;
    reward = (collided | late) ? R(-10) : r_goal;
    return collided ? (AQUA_FLAG_DONE | AQUA_FLAG_COLLIDED) : f_time;
}

// action decode ---------------------------------------------------------------------------
// continuous aqua.py:143-151: values outside [0.2, 0.5] are clipped (and counted)
template <typename R>
__device__ __forceinline__ Drive<R> drive_continuous(float al, float ar, uint32_t& bad) {
    const R l = (R)al, r = (R)ar;
    const R tl = fmin(fmax(l, R(0.2)), R(0.5));
    const R tr = fmin(fmax(r, R(0.2)), R(0.5));
    bad += (tl != l || tr != r) ? 1u : 0u;
    return drive_from_thrusts<R>(tl, tr);
}

}  // namespace aqua
