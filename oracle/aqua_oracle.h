/*
 * aqua_oracle.h -- CPU oracle for the AquaGym environment step.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C, float64 restatement of the
 * reference environment `gym_aqua/envs/aqua.py` (ilVecc/AquaticGymEnv).  It is
 * the checker the CUDA path is compared against; it is never the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it.
 *
 * Pinning: the oracle is pinned against golden vectors produced by EXECUTING the
 * unmodified reference (`oracle/gen_golden.py`, fixtures in tests/golden/), see
 * tests/test_oracle_golden.py.  The reference itself ships no tests.
 *
 * Every function cites the reference lines (relative to /root/reference) it
 * follows.  The random draws (which the reference takes from two unseeded NumPy
 * generators, aqua.py:103-115,124,188) are replaced by a counter-based
 * Philox4x32-10 stream; the draw layout is documented in DESIGN.md section 4 and
 * is the same layout the CUDA kernels use, so both sides see identical inputs.
 */
#ifndef AQUA_ORACLE_H
#define AQUA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AQO_MAX_OBSTACLES 64

/* flags byte written by step (same bit layout as the CUDA path) */
#define AQO_FLAG_DONE 1u
#define AQO_FLAG_COLLIDED 2u
#define AQO_FLAG_TIME 4u
#define AQO_FLAG_SUCCESS 8u

/* Philox stream tags (counter word 3) */
#define AQO_STREAM_RESET 1u      /* reset try n -> (x, y, theta | gx, gy) */
#define AQO_STREAM_WAVE_INIT 2u  /* wave speed at reset (Philox2x32) */
#define AQO_STREAM_WAVE_NOISE 3u /* wave-speed increment of one step (Philox2x32) */
#define AQO_STREAM_ACTION 4u     /* synthetic random actions (bench / rollout; Philox2x32) */

#define AQO_RESET_MAX_TRIES 4096u
/* the kernels pack (episode << 11 | time) in one 32-bit word: episode is 21 bits and
 * starts at AQO_EPISODE_NONE so that the first reset opens episode 0 */
#define AQO_EPISODE_MASK 0x1FFFFFu
#define AQO_EPISODE_NONE 0x1FFFFFu

typedef struct {
    int32_t continuous;  /* aqua.py:11,458-459 */
    int32_t waves;       /* aqua.py:15 */
    int32_t random_boat; /* aqua.py:16 */
    int32_t random_goal; /* aqua.py:17 */
    int32_t n_obstacles;
    uint32_t key_winit;  /* derived stream keys: aqo_world_init_keys() */
    uint32_t key_noise;
    uint32_t key_action;
    uint64_t seed;
    uint64_t env_id_base;
    /* obstacle k: {cx, cy, type (0 = circle, 1 = rectangle), a, b};
     * circle: a = radius; rectangle: a = width, b = height (centre anchored),
     * aqua.py:56-68, 373-390 */
    double obstacles[AQO_MAX_OBSTACLES * 5];
} aqo_world;

/* structure-of-arrays state, n entries each */
typedef struct {
    double *x, *y, *theta; /* boat pose, aqua.py:110-117 */
    double *wx, *wy;       /* wave speed, aqua.py:124 */
    double *gx, *gy;       /* goal, aqua.py:103-107 */
    int32_t *time;         /* aqua.py:125,141 */
    uint32_t *episode;     /* episode counter (key of the draw streams) */
} aqo_state;

/* fills key_noise / key_action from seed; call once after setting the fields */
void aqo_world_init_keys(aqo_world* w);

/* ---- draw streams (shared spec with the CUDA kernels) ---- */
void aqo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void aqo_philox2x32_10(const uint32_t ctr[2], uint32_t key, uint32_t out[2]);
uint32_t aqo_key32(uint64_t seed, uint32_t tag);
float aqo_u01(uint32_t u);
void aqo_draw_reset_sample(uint64_t seed, uint32_t env_id, uint32_t episode, uint32_t n,
                           double out5[5]);
void aqo_draw_wave_init(uint64_t seed, uint32_t env_id, uint32_t episode, int waves,
                        double out2[2]);
void aqo_draw_wave_noise(uint64_t seed, uint32_t env_id, uint32_t episode, uint32_t t, int waves,
                         double out2[2]);
uint32_t aqo_draw_action_discrete(uint64_t seed, uint32_t env_id, uint64_t step);
void aqo_draw_action_continuous(uint64_t seed, uint32_t env_id, uint64_t step, float out2[2]);

/* ---- single-env primitives (reference restatement) ---- */
double aqo_normalize_angle(double value);
double aqo_distance_circle_circle(double ax, double ay, double ar, double bx, double by, double br);
double aqo_distance_rectangle_circle(double rx, double ry, double rw, double rh, double cx,
                                     double cy, double cr);
int aqo_collided_border(double px, double py, double radius);
int aqo_collided_obstacle(const aqo_world* w, double px, double py, double radius);
double aqo_nearest_obstacle_distance(const aqo_world* w, double px, double py, double radius);

/* reset env i (advances episode[i]); returns the number of reset samples drawn */
int aqo_reset_env(const aqo_world* w, aqo_state* s, int64_t i);

/* one step of env i.  action: discrete index (as double) or thrust pair.
 * noise: NULL -> Philox stream, else the injected wave increment {nx, ny}.
 * margin (nullable): the reference's decision margin after the move =
 * min(|border slack|, |nearest obstacle distance|, |goal distance|).
 * returns the flags byte; *reward receives the reward. */
uint32_t aqo_step_env(const aqo_world* w, aqo_state* s, int64_t i, const double* action,
                      const double* noise, double* reward, double* margin);

/* ---- batched drivers ---- */
void aqo_reset_batch(const aqo_world* w, aqo_state* s, int64_t n, const uint8_t* mask);
/* actions: discrete int32[n] (actions_i32) or continuous float[n*2] (actions_f32);
 * noise: nullable double[2*n] laid out {nx[n], ny[n]};
 * autoreset != 0: envs whose step returned done are reset in place afterwards. */
void aqo_step_batch(const aqo_world* w, aqo_state* s, int64_t n, const int32_t* actions_i32,
                    const float* actions_f32, const double* noise, double* reward, uint8_t* flags,
                    double* margin, int autoreset, int nthreads);
/* K steps of synthetic random actions drawn from AQO_STREAM_ACTION (bench helper);
 * returns the number of finished episodes. step0 = index of the first step. */
int64_t aqo_rollout_random(const aqo_world* w, aqo_state* s, int64_t n, int64_t steps,
                           uint64_t step0, int nthreads, double* reward_sum);

#ifdef __cplusplus
}
#endif
#endif
