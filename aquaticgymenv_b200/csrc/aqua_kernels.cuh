// aqua_kernels.cuh -- launch interface between the C ABI (aqua_abi.cu) and the kernels
// (aqua_kernels.cu).  Plain structs and launcher prototypes; no CUDA device code here.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/aquagym.h"

namespace aqua {

// CTA shape of the wide kernels (V envs per thread): kWideThreads threads, compiled for kMinBlocks
// resident CTAs per SM.  Narrow launches (V = 1) use 128-thread CTAs of the same kernels.
// Measured on one box (profiles/r01F_cta_shape_ab.txt): 256 x 2 (112 registers, 16 warps per SM)
// 8.27e10 env-steps/s; 192 x 3 (96 registers, 18 warps per SM) 8.15e10, and 11 % slower on the
// 64-obstacle map: the registers are worth more than the two extra warps.
#ifndef AQUA_WIDE_THREADS
#define AQUA_WIDE_THREADS 256
#endif
#ifndef AQUA_MIN_BLOCKS
#define AQUA_MIN_BLOCKS 2
#endif
constexpr int kWideThreads = AQUA_WIDE_THREADS;
constexpr int kNarrowThreads = 128;
constexpr int kMaxThreads = kWideThreads > kNarrowThreads ? kWideThreads : kNarrowThreads;  // launch bounds
constexpr int kMinBlocks = AQUA_MIN_BLOCKS;
// dynamic tile claims of the streaming kernel: kTilePools counters + the finished-CTA count, one
// 128-byte line each
constexpr int kTilePools = 8;
constexpr int kTilePoolStride = 32;                                   // uint32 words
constexpr int kTileCounterWords = (kTilePools + 1) * kTilePoolStride;
constexpr int kStatSlots = 1024;   // per-CTA statistics slots (CTA b adds into slot b % kStatSlots)
constexpr int kSlotStats = 8;      // counters per slot: AQUA_STAT_EPISODES .. AQUA_STAT_BAD_ACTIONS; the ninth
                                   // (AQUA_STAT_FORCED_RESETS) is one global counter, World::forced

// Round keys of the draw streams, computed once on the host so that in the kernels they are
// constant-bank operands of the rounds' XORs:
//   rk4[2r], rk4[2r+1] = k0 + r*0x9E3779B9, k1 + r*0xBB67AE85  (Philox4x32-10, key = seed:
//                        RESET stream, 4 words per sampling try)
//   noise/winit/action[r] = key32(tag) + r*0x9E3779B9          (Philox2x32-10, per-step streams;
//                        key32(tag) = Philox4x32-10(counter (tag,0,0,0), key seed).x)
struct PhiloxKeys {
    uint32_t rk4[20];
    uint32_t noise[10];
    uint32_t winit[10];
    uint32_t action[10];
};
void make_philox_keys(uint64_t seed, PhiloxKeys* out);

// A small obstacle table (at most 4 circles and 4 rectangles: every layout the reference ships) as
// kernel parameters: its entries are constant-bank operands of the clearance arithmetic, which then
// is straight-line code without loads (step_envs, aqua_packed.cuh).  Unused entries never collide.
struct SmallTable {
    float c[4][4];  // circles {cx, cy, -(radius + 2.5)^2, 0}
    float r[4][4];  // rectangles {cx, cy, half_x, half_y}
};

// Everything a kernel needs to know about the world; travels by value in the parameter block.
struct World {
    PhiloxKeys keys;
    uint64_t env_id_base;   // global id of env 0 of this slab
    const void* tab;        // device ObstacleTable<R> (R = float or double per precision)
    const double* tab64;    // device double[K][5] = {cx, cy, hx, hy, r2}: float64 table (reset path)
    int32_t k_obst;
    int32_t k_circ;         // obstacles [0, k_circ) are circles, [k_circ, k_obst) rectangles
    float reset_band;       // guard band of the float32 reset test, see reset_blocked()
    uint8_t waves, random_boat, random_goal, autoreset;
    uint8_t use_grid;       // dense map: obstacles are culled through the cell grid of the table
    uint8_t small_mode;     // float32 tables of <= 4 circles and <= 4 rectangles: 1 (<= 2 rectangles) or 2; else 0
    unsigned int* forced;   // device counter of forced reset accepts (AQUA_STAT_FORCED_RESETS)
    SmallTable small;       // valid when small_mode != 0
};

struct StepArgs {
    World w;
    void* planes;           // plane 0; plane p starts p * ld elements later
    int64_t ld;
    int64_t env_begin;      // first env of this launch (multiple of AQUA_TILE)
    int64_t env_end;        // one past the last env of this launch
    const void* actions;    // indexed by absolute env index
    void* reward;
    uint8_t* flags;
    const float* noise_x;   // injected wave increments (or NULL): indexed by absolute env index
    const float* noise_y;
    unsigned long long* stats;  // [kStatSlots][kSlotStats]
    int64_t steps;          // rollout kernel: steps per launch
    uint64_t step0;         // rollout kernel: ACTION-stream index of the first step
    // streaming step kernel (whole warps only): everything below is filled by launch_step() so
    // that the kernel addresses with one 32-bit index and constant-bank base pointers
    const char* pl0[AQUA_N_PLANES];  // plane p at env_begin
    const char* act0;                // actions / reward / flags at env_begin
    char* rew0;
    uint8_t* flg0;
    uint32_t* tile_ctr;              // one counter set (kTileCounterWords), zero between launches
    uint32_t n_stream;               // envs of the launch, a multiple of 32 * (envs per thread)
    uint32_t env_id0;                // global id of env_begin
    uint32_t tile_lane_mask;         // 0 (see the tile claim in step_stream_kernel)
    // terminal observations (aqua.py:213 returns the terminal obs with done = True; with auto-reset
    // the planes already hold the next episode): 5 planes [ld] of the state's element type, written
    // by the env's owner thread when an env finishes; NULL = not kept
    void* term;
    // multi-step kernel (step_many_kernel): rows of the action tape / reward / flags outputs are
    // row_stride envs apart; stash = one 32-byte record per env with the pre-drawn next episode
    int64_t row_stride;
    uint4* stash;          // the ring: [env][kRingDepth] records of 32 bytes
    uint32_t* stash_tag;   // [env][kRingDepth]: the episode each slot holds
    // in-kernel policy (aqua_rollout_policy): the Q-network's AQUA_MLP_WEIGHTS floats, the exploration
    // rate, and (or NULL) the chosen actions [steps][row_stride]
    const float* mlp_w;
    float mlp_eps;
    int32_t* act_out;
    // the same two values read on the device (aqua_rollout_policy_dev: a captured CUDA graph replays
    // the launch with a new exploration rate and step index each time), or NULL
    const float* mlp_eps_dev;
    const unsigned long long* step0_dev;
};

struct LaunchGeom {
    int threads;  // per CTA
    int blocks;
    int v;        // envs per thread: envs_per_thread(precision) (wide) or 1 (narrow, small slabs)
};

// every launcher returns the cudaError_t of the launch (cudaGetLastError)
cudaError_t launch_step(int precision, int action_dtype, const StepArgs& a, LaunchGeom g, cudaStream_t s);
// K env-steps per launch (aqua_many.cu).  tape != 0: actions[K][row_stride] in, reward / flags
// [K][row_stride] out (either may be NULL); tape == 0: the synthetic random policy (ACTION stream)
cudaError_t launch_step_many(int precision, int action_dtype, int tape, const StepArgs& a, LaunchGeom g, cudaStream_t s);
// K env-steps per launch with the actions chosen in the kernel by the Q-network a.mlp_w (V = 1)
cudaError_t launch_step_policy(int precision, const StepArgs& a, int threads, cudaStream_t s);
cudaError_t many_occupancy(int precision, int v, int action_dtype, int tape, int threads, int tab, int* blocks_per_sm);
cudaError_t launch_stash_init(void* ring, size_t bytes, cudaStream_t s);
int stash_ring_depth();
cudaError_t launch_reset(int precision, const StepArgs& a, const uint8_t* mask, cudaStream_t s);
cudaError_t launch_init_state(int precision, void* planes, int64_t ld, cudaStream_t s);
cudaError_t launch_random_actions(const World& w, int64_t n, uint64_t step_index, void* actions,
                                  int action_dtype, cudaStream_t s);
cudaError_t launch_get_state(int precision, const void* planes, int64_t ld, int64_t n, double* out,
                             cudaStream_t s);
cudaError_t launch_set_state(int precision, void* planes, int64_t ld, int64_t n, const double* in,
                             cudaStream_t s);
cudaError_t launch_observe(int precision, const void* planes, int64_t ld, int64_t n, float* out, int normalized,
                           cudaStream_t s);
// planes 0..4 of envs [0, n) into out[5][n], then *done = seq (both mapped host memory)
cudaError_t launch_copy_obs_planes(int precision, const void* planes, int64_t ld, int64_t n, void* out, uint32_t* done,
                                   uint32_t seq, cudaStream_t s);
// writes the goals of the finished envs of [lo, lo + len) into the caller's (device-mapped) host planes
cudaError_t launch_goal_updates(int precision, const void* planes, int64_t ld, const uint8_t* flags, int64_t lo,
                                int64_t len, void* h_gx, void* h_gy, cudaStream_t s);
// x, y, angle (+ goals, or the finished envs' goals only), reward, flags of envs [lo, lo + len) into MAPPED host buffers
cudaError_t launch_pack_host(int precision, const void* planes, int64_t ld, const void* reward, const uint8_t* flags,
                             int64_t lo, int64_t len, int64_t n, void* h_obs, void* h_rew, uint8_t* h_flg, int sparse_goals,
                             cudaStream_t s);
cudaError_t launch_stats_reduce(unsigned long long* slots, unsigned int* forced, int64_t* out, int clear, cudaStream_t s);
// per-device, once: opt the step kernels in to their dynamic shared memory (staging buffers)
cudaError_t configure_kernels(int precision);
// occupancy of the step kernel for (precision, action dtype) at `threads` per CTA; `use_grid`:
// the launch carries the culling grid of a dense map in its dynamic shared memory
cudaError_t step_occupancy(int precision, int v, int action_dtype, int threads, int use_grid, int* blocks_per_sm);
// same for the streaming kernel; *blocks_per_sm = 0 when (precision, action dtype) has none
cudaError_t stream_occupancy(int precision, int action_dtype, int waves, int threads, int use_grid, int* blocks_per_sm);

int envs_per_thread(int precision);

// host-side construction of the device obstacle tables; `out` sized for the precision
size_t obstacle_table_bytes(int precision);
// (circles first, then rectangles); returns the number of circles.  *use_grid: the table is dense
// enough for the culling grid and its cell masks are unambiguous (see kCellBlocked)
int build_obstacle_tables(int precision, int k, const double* rows5, void* tab_host, double* tab64_host,
                          int* use_grid);
// World::small / small_mode from the caller's rows {cx, cy, type, a, b} (same float32 values as the table)
void fill_small_table(int precision, int k, const double* rows5, int use_grid, World* w);
// the cell masks of a host table built by build_obstacle_tables (aqua_obstacle_grid)
void copy_table_cells(int precision, const void* tab_host, unsigned long long* out);
void copy_table_cells16(int precision, const void* tab_host, unsigned short* out);

}  // namespace aqua
