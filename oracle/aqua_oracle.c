/*
 * aqua_oracle.c -- CPU oracle for the AquaGym environment step (float64, plain C).
 *
 * TEST INFRASTRUCTURE ONLY -- see aqua_oracle.h.  The arithmetic below follows
 * the reference `gym_aqua/envs/aqua.py` operation by operation (same formulation,
 * same evaluation order, float64), NOT the closed form the CUDA kernels use, so
 * that agreement between the two is a real check.
 *
 * Build: make -C oracle      (gcc -O2 -shared, pthreads; no -ffast-math, no FMA
 * contraction, so that the arithmetic is the plain IEEE sequence written here)
 */
#define _GNU_SOURCE
#include "aqua_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <string.h>

/* ---------------- world constants (aqua.py:20-25, 72-81, 86-94) ------------- */
static const double WORLD_SIZE = 100.0;         /* aqua.py:20 */
static const double MOTOR_MIN_THRUST = 0.2;     /* aqua.py:21 */
static const double MOTOR_MAX_THRUST = 0.5;     /* aqua.py:22 */
static const double WAVE_MAX_SPEED = 0.05;      /* aqua.py:23-24 (times has_waves) */
static const double WAVE_SPEED_VARIANCE = 0.001;/* aqua.py:25 (times has_waves) */
static const double GOAL_RADIUS = 2.5;          /* aqua.py:72 */
static const double BOAT_RADIUS = 2.5;          /* aqua.py:75 */
static const double AXLE_LENGTH = 2.5;          /* aqua.py:76 (= boat radius) */
static const double THRUST_EPSILON = 1e-8;      /* aqua.py:81 */
static const double REWARD_GOAL = 10.0;         /* aqua.py:86 */
static const double REWARD_MAX_STEPS = -10.0;   /* aqua.py:87 */
static const double REWARD_COLLISION = -10.0;   /* aqua.py:88 */
static const double REWARD_STEP_SCALE = 0.7;    /* aqua.py:90 */
static const int32_t TIME_LIMIT = 1000;         /* aqua.py:91 */
static const double TAU = 1.0;                  /* aqua.py:94 */
/* observation_space bounds, aqua.py:46-49 */
static const double OBS_LOW_ANGLE = -M_PI;
static const double OBS_HIGH_ANGLE = M_PI;

/* ---------------- Philox4x32-10 (Salmon et al., SC'11; Random123) ------------
 * Not part of the reference: it replaces the reference's unseeded NumPy draws.
 * Known-answer vectors are checked in tests/test_philox.py. */
void aqo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Philox2x32-10 (same paper): the per-step draws need two words only */
void aqo_philox2x32_10(const uint32_t ctr[2], uint32_t key, uint32_t out[2]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], k = key;
    for (int r = 0; r < 10; ++r) {
        uint64_t p = (uint64_t)0xD256D193u * c0;
        uint32_t n0 = (uint32_t)(p >> 32) ^ k ^ c1;
        c1 = (uint32_t)p;
        c0 = n0;
        k += 0x9E3779B9u;
    }
    out[0] = c0; out[1] = c1;
}

/* 32-bit key of a per-step stream, derived from the 64-bit seed */
uint32_t aqo_key32(uint64_t seed, uint32_t tag) {
    uint32_t ctr[4] = {tag, 0u, 0u, 0u};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    aqo_philox4x32_10(ctr, key, out);
    return out[0];
}

/* 24 high bits -> float32 in [0, 1); exactly representable */
float aqo_u01(uint32_t u) { return (float)(u >> 8) * 0x1p-24f; }

static void philox_env(uint64_t seed, uint32_t env_id, uint32_t c1, uint32_t c2, uint32_t tag,
                       uint32_t out[4]) {
    uint32_t ctr[4] = {env_id, c1, c2, tag};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    aqo_philox4x32_10(ctr, key, out);
}

/* All draws are computed in float32 exactly as the kernel does and then widened,
 * so that the float32 and float64 instantiations of the kernel and this oracle
 * see bit-identical inputs. */

/* largest float32 below pi and its double: theta ~ U[-PI_DN, PI_DN) is inside
 * the reference's [-pi, pi) (aqua.py:46-49, Box.sample) */
#define AQO_PI_DN_F 3.1415925f
#define AQO_TWO_PI_DN_F 6.283185f /* = 2 * PI_DN_F exactly */

/* One call of `observation_space.sample()` (aqua.py:103,105,110,115): a 5-vector
 * of which the goal loop uses [3:] and the boat loop uses [:3].  Components
 * [2] and [3] share a word; no caller uses both. */
void aqo_draw_reset_sample(uint64_t seed, uint32_t env_id, uint32_t episode, uint32_t n,
                           double out5[5]) {
    uint32_t r[4];
    philox_env(seed, env_id, episode, n, AQO_STREAM_RESET, r);
    out5[0] = (double)(aqo_u01(r[0]) * 100.0f);
    out5[1] = (double)(aqo_u01(r[1]) * 100.0f);
    out5[2] = (double)(fmaf(aqo_u01(r[2]), AQO_TWO_PI_DN_F, -AQO_PI_DN_F));
    out5[3] = (double)(aqo_u01(r[2]) * 100.0f);
    out5[4] = (double)(aqo_u01(r[3]) * 100.0f);
}

/* `wave_space.sample()` (aqua.py:124): U[-0.05, 0.05)^2 * has_waves */
static void wave_init_k(uint32_t key, uint32_t env_id, uint32_t episode, double out2[2]) {
    uint32_t ctr[2] = {env_id, episode};
    uint32_t r[2];
    aqo_philox2x32_10(ctr, key, r);
    out2[0] = (double)fmaf(aqo_u01(r[0]), 0.1f, -0.05f);
    out2[1] = (double)fmaf(aqo_u01(r[1]), 0.1f, -0.05f);
}
void aqo_draw_wave_init(uint64_t seed, uint32_t env_id, uint32_t episode, int waves,
                        double out2[2]) {
    if (!waves) { out2[0] = 0.0; out2[1] = 0.0; return; }
    wave_init_k(aqo_key32(seed, AQO_STREAM_WAVE_INIT), env_id, episode, out2);
}

/* `np.random.uniform(-var, var, 2)` (aqua.py:188); t = env time BEFORE the step's
 * increment (0 for the first step of an episode): Philox2x32 keyed by the packed
 * (episode << 11 | t) word */
static void wave_noise_k(uint32_t key, uint32_t env_id, uint32_t episode, uint32_t t, double out2[2]) {
    if (t > 2047u) t = 2047u;
    uint32_t ctr[2] = {env_id, (episode << 11) | t};
    uint32_t r[2];
    aqo_philox2x32_10(ctr, key, r);
    out2[0] = (double)fmaf(aqo_u01(r[0]), 0.002f, -0.001f);
    out2[1] = (double)fmaf(aqo_u01(r[1]), 0.002f, -0.001f);
}
void aqo_draw_wave_noise(uint64_t seed, uint32_t env_id, uint32_t episode, uint32_t t, int waves,
                         double out2[2]) {
    if (!waves) { out2[0] = 0.0; out2[1] = 0.0; return; }
    wave_noise_k(aqo_key32(seed, AQO_STREAM_WAVE_NOISE), env_id, episode, t, out2);
}

/* derived keys of the per-step streams; call after setting w->seed */
void aqo_world_init_keys(aqo_world* w) {
    w->key_winit = aqo_key32(w->seed, AQO_STREAM_WAVE_INIT);
    w->key_noise = aqo_key32(w->seed, AQO_STREAM_WAVE_NOISE);
    w->key_action = aqo_key32(w->seed, AQO_STREAM_ACTION);
}

/* synthetic random policy: `action_space.sample()` stand-in (test_random.py:20) */
static void action_words_k(uint32_t key_action, uint32_t env_id, uint64_t step, uint32_t r[2]) {
    uint32_t ctr[2] = {env_id, (uint32_t)step};
    aqo_philox2x32_10(ctr, key_action + (uint32_t)(step >> 32), r);
}
static uint32_t action_discrete_k(uint32_t key_action, uint32_t env_id, uint64_t step) {
    uint32_t r[2];
    action_words_k(key_action, env_id, step, r);
    return (uint32_t)(((uint64_t)r[0] * 3u) >> 32);
}
static void action_continuous_k(uint32_t key_action, uint32_t env_id, uint64_t step, float out2[2]) {
    uint32_t r[2];
    action_words_k(key_action, env_id, step, r);
    out2[0] = fmaf(aqo_u01(r[0]), 0.3f, 0.2f);
    out2[1] = fmaf(aqo_u01(r[1]), 0.3f, 0.2f);
}
uint32_t aqo_draw_action_discrete(uint64_t seed, uint32_t env_id, uint64_t step) {
    return action_discrete_k(aqo_key32(seed, AQO_STREAM_ACTION), env_id, step);
}
void aqo_draw_action_continuous(uint64_t seed, uint32_t env_id, uint64_t step, float out2[2]) {
    action_continuous_k(aqo_key32(seed, AQO_STREAM_ACTION), env_id, step, out2);
}

/* ---------------- reference primitives ---------------- */

/* aqua.py:128-133 */
double aqo_normalize_angle(double value) {
    double range_start = OBS_LOW_ANGLE, range_end = OBS_HIGH_ANGLE;
    double width = range_end - range_start;
    double offset_value = value - range_start;
    return (offset_value - (floor(offset_value / width) * width)) + range_start;
}

/* np.linalg.norm of a 2-vector */
static double norm2(double dx, double dy) { return sqrt(dx * dx + dy * dy); }

/* aqua.py:373-377 */
double aqo_distance_circle_circle(double ax, double ay, double ar, double bx, double by,
                                  double br) {
    double centers_dist = norm2(ax - bx, ay - by);
    double total_radius_length = ar + br;
    return centers_dist - total_radius_length;
}

static double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* aqua.py:379-390 */
double aqo_distance_rectangle_circle(double rx, double ry, double rw, double rh, double cx,
                                     double cy, double cr) {
    double rect_l = rx - rw / 2, rect_r = rx + rw / 2;
    double rect_b = ry - rh / 2, rect_t = ry + rh / 2;
    double dx = cx - clipd(cx, rect_l, rect_r);
    double dy = cy - clipd(cy, rect_b, rect_t);
    return norm2(dx, dy) - cr;
}

static double obstacle_distance(const double* o, double px, double py, double radius) {
    /* aqua.py:431-436: 'c' -> circle, anything else -> rectangle */
    if (o[2] == 0.0) return aqo_distance_circle_circle(o[0], o[1], o[3], px, py, radius);
    return aqo_distance_rectangle_circle(o[0], o[1], o[3], o[4], px, py, radius);
}

/* aqua.py:424-427 (boat) and aqua.py:453-455 (goal): strict inequalities */
int aqo_collided_border(double px, double py, double radius) {
    return (px - radius < 0.0) || (py - radius < 0.0) || (px + radius > WORLD_SIZE) ||
           (py + radius > WORLD_SIZE);
}

/* aqua.py:429-439 (boat) and aqua.py:442-451 (goal) */
int aqo_collided_obstacle(const aqo_world* w, double px, double py, double radius) {
    for (int k = 0; k < w->n_obstacles; ++k)
        if (obstacle_distance(&w->obstacles[k * 5], px, py, radius) <= 0.0) return 1;
    return 0;
}

/* aqua.py:409-418 (unused by the reference's step; margin oracle only) */
double aqo_nearest_obstacle_distance(const aqo_world* w, double px, double py, double radius) {
    double min_dist = INFINITY;
    for (int k = 0; k < w->n_obstacles; ++k) {
        double d = obstacle_distance(&w->obstacles[k * 5], px, py, radius);
        if (d < min_dist) min_dist = d;
    }
    return min_dist;
}

/* aqua.py:392-395 */
static double distance_state_from_goal(const aqo_state* s, int64_t i, double px, double py) {
    return aqo_distance_circle_circle(s->gx[i], s->gy[i], GOAL_RADIUS, px, py, BOAT_RADIUS);
}

/* ---------------- reset (aqua.py:100-126) ---------------- */
int aqo_reset_env(const aqo_world* w, aqo_state* s, int64_t i) {
    uint32_t env_id = (uint32_t)(w->env_id_base + (uint64_t)i);
    uint32_t ep = (s->episode[i] + 1u) & AQO_EPISODE_MASK; /* new episode index */
    uint32_t n = 0;
    double v[5];
    /* goal: aqua.py:102-107 */
    if (w->random_goal) {
        for (;;) {
            aqo_draw_reset_sample(w->seed, env_id, ep, n++, v);
            s->gx[i] = v[3];
            s->gy[i] = v[4];
            if (n >= AQO_RESET_MAX_TRIES) break; /* not in the reference: bounded loop */
            if (!(aqo_collided_border(s->gx[i], s->gy[i], GOAL_RADIUS) ||
                  aqo_collided_obstacle(w, s->gx[i], s->gy[i], GOAL_RADIUS)))
                break;
        }
    } else {
        s->gx[i] = 25.0;
        s->gy[i] = 80.0;
    }
    /* boat: aqua.py:109-117 */
    if (w->random_boat) {
        for (;;) {
            aqo_draw_reset_sample(w->seed, env_id, ep, n++, v);
            s->x[i] = v[0];
            s->y[i] = v[1];
            s->theta[i] = v[2];
            if (n >= AQO_RESET_MAX_TRIES) break;
            if (!(distance_state_from_goal(s, i, s->x[i], s->y[i]) <= 0.0 ||
                  aqo_collided_border(s->x[i], s->y[i], BOAT_RADIUS) ||
                  aqo_collided_obstacle(w, s->x[i], s->y[i], BOAT_RADIUS)))
                break;
        }
    } else {
        s->x[i] = 85.0;
        s->y[i] = 45.0;
        s->theta[i] = 0.0;
    }
    /* velocities: aqua.py:120-125 */
    double wv[2];
    if (w->waves) wave_init_k(w->key_winit, env_id, ep, wv);
    else { wv[0] = 0.0; wv[1] = 0.0; }
    s->wx[i] = wv[0];
    s->wy[i] = wv[1];
    s->time[i] = 0;
    s->episode[i] = ep;
    return (int)n;
}

/* ---------------- step (aqua.py:135-213) ---------------- */
uint32_t aqo_step_env(const aqo_world* w, aqo_state* s, int64_t i, const double* action,
                      const double* noise, double* reward, double* margin) {
    uint32_t env_id = (uint32_t)(w->env_id_base + (uint64_t)i);
    const double has_waves = (double)(w->waves ? 1 : 0);
    /* aqua.py:140-141 */
    double prev_x = s->x[i], prev_y = s->y[i];
    int32_t t_before = s->time[i];
    s->time[i] += 1;

    /* thrusts: aqua.py:143-154 */
    double thrust_left, thrust_right;
    if (w->continuous) {
        double a0 = action[0], a1 = action[1];
        /* Box.contains -> else np.clip (aqua.py:145-150); clip is the identity in bounds */
        a0 = clipd(a0, MOTOR_MIN_THRUST, MOTOR_MAX_THRUST);
        a1 = clipd(a1, MOTOR_MIN_THRUST, MOTOR_MAX_THRUST);
        thrust_left = a0;
        thrust_right = a1;
    } else {
        /* action table aqua.py:33-42 */
        int a = (int)action[0];
        static const double table[3][2] = {{0.2, 0.5}, {0.5, 0.2}, {0.5, 0.5}};
        thrust_left = table[a][0];
        thrust_right = table[a][1];
    }

    /* aqua.py:159-160 */
    double thrust_diff = thrust_right - thrust_left;
    thrust_diff = copysign(fmax(fabs(thrust_diff), THRUST_EPSILON), thrust_diff);

    /* aqua.py:166-168 */
    double r = AXLE_LENGTH / 2 * (thrust_right + thrust_left) / thrust_diff;
    double wv = thrust_diff / AXLE_LENGTH;

    /* aqua.py:173-174 */
    double angle = M_PI / 2 + s->theta[i];
    double icc_x = s->x[i] + r * (-sin(angle));
    double icc_y = s->y[i] + r * cos(angle);
    /* aqua.py:176-177 */
    double c = cos(wv * TAU), sn = sin(wv * TAU);
    /* aqua.py:180-181: rot.dot(p - icc) + icc + wave * tau */
    double vx = s->x[i] - icc_x, vy = s->y[i] - icc_y;
    double rx = c * vx + (-sn) * vy;
    double ry = sn * vx + c * vy;
    s->x[i] = rx + icc_x + s->wx[i] * TAU;
    s->y[i] = ry + icc_y + s->wy[i] * TAU;
    /* aqua.py:183 */
    s->theta[i] = aqo_normalize_angle(s->theta[i] + wv * TAU);

    /* waves: aqua.py:188-191 (after the move) */
    double nz[2];
    if (noise) { nz[0] = noise[0]; nz[1] = noise[1]; }
    else if (!w->waves) { nz[0] = 0.0; nz[1] = 0.0; }
    else wave_noise_k(w->key_noise, env_id, s->episode[i], (uint32_t)t_before, nz);
    (void)WAVE_SPEED_VARIANCE;
    double wmin = -WAVE_MAX_SPEED * has_waves, wmax = WAVE_MAX_SPEED * has_waves;
    s->wx[i] = clipd(s->wx[i] + nz[0], wmin, wmax);
    s->wy[i] = clipd(s->wy[i] + nz[1], wmin, wmax);

    /* termination ladder: aqua.py:194-211 */
    uint32_t flags;
    double px = s->x[i], py = s->y[i];
    double d_cur = distance_state_from_goal(s, i, px, py);
    if (aqo_collided_obstacle(w, px, py, BOAT_RADIUS) || aqo_collided_border(px, py, BOAT_RADIUS)) {
        flags = AQO_FLAG_DONE | AQO_FLAG_COLLIDED;
        *reward = REWARD_COLLISION;
    } else if (s->time[i] > TIME_LIMIT) {
        flags = AQO_FLAG_DONE | AQO_FLAG_TIME;
        *reward = REWARD_MAX_STEPS;
    } else if (d_cur <= 0.0) { /* aqua.py:421-422 */
        flags = AQO_FLAG_DONE | AQO_FLAG_SUCCESS;
        *reward = REWARD_GOAL;
    } else {
        flags = 0;
        double d_prev = distance_state_from_goal(s, i, prev_x, prev_y);
        *reward = (d_prev - d_cur) * REWARD_STEP_SCALE; /* aqua.py:89-90 */
    }

    if (margin) {
        /* aqua.py:404-407 and :409-418, plus the goal distance */
        double bl = fmin(px - BOAT_RADIUS - 0.0, py - BOAT_RADIUS - 0.0);
        double tr = fmin(WORLD_SIZE - (px + BOAT_RADIUS), WORLD_SIZE - (py + BOAT_RADIUS));
        double m = fabs(fmin(bl, tr));
        double od = fabs(aqo_nearest_obstacle_distance(w, px, py, BOAT_RADIUS));
        if (od < m) m = od;
        if (fabs(d_cur) < m) m = fabs(d_cur);
        *margin = m;
    }
    return flags;
}

/* ---------------- batched drivers ---------------- */
void aqo_reset_batch(const aqo_world* w, aqo_state* s, int64_t n, const uint8_t* mask) {
    for (int64_t i = 0; i < n; ++i)
        if (!mask || mask[i]) aqo_reset_env(w, s, i);
}

/* plain pthread fan-out over contiguous env ranges (envs are independent) */
typedef struct {
    void (*fn)(void* arg, int64_t lo, int64_t hi, int tid);
    void* arg;
    int64_t lo, hi;
    int tid;
} aqo_job;

static void* job_main(void* p) {
    aqo_job* j = (aqo_job*)p;
    j->fn(j->arg, j->lo, j->hi, j->tid);
    return NULL;
}

#define AQO_MAX_THREADS 512
static void parallel_ranges(int64_t n, int nthreads,
                            void (*fn)(void*, int64_t, int64_t, int), void* arg) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > AQO_MAX_THREADS) nthreads = AQO_MAX_THREADS;
    if ((int64_t)nthreads > n) nthreads = n > 0 ? (int)n : 1;
    if (nthreads == 1) { fn(arg, 0, n, 0); return; }
    pthread_t th[AQO_MAX_THREADS];
    aqo_job jobs[AQO_MAX_THREADS];
    int started[AQO_MAX_THREADS];
    int64_t chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        int64_t lo = (int64_t)t * chunk, hi = lo + chunk;
        if (lo > n) lo = n;
        if (hi > n) hi = n;
        jobs[t].fn = fn; jobs[t].arg = arg; jobs[t].lo = lo; jobs[t].hi = hi; jobs[t].tid = t;
        started[t] = (pthread_create(&th[t], NULL, job_main, &jobs[t]) == 0);
        if (!started[t]) fn(arg, lo, hi, t); /* could not spawn: run inline */
    }
    for (int t = 0; t < nthreads; ++t)
        if (started[t]) pthread_join(th[t], NULL);
}

typedef struct {
    const aqo_world* w; aqo_state* s; int64_t n;
    const int32_t* actions_i32; const float* actions_f32; const double* noise;
    double* reward; uint8_t* flags; double* margin; int autoreset;
} step_args;

static void step_range(void* p, int64_t lo, int64_t hi, int tid) {
    (void)tid;
    step_args* a = (step_args*)p;
    const aqo_world* w = a->w;
    for (int64_t i = lo; i < hi; ++i) {
        double act[2], nz[2], rew = 0.0, mg = 0.0;
        if (w->continuous) {
            act[0] = (double)a->actions_f32[2 * i];
            act[1] = (double)a->actions_f32[2 * i + 1];
        } else {
            act[0] = (double)a->actions_i32[i];
            act[1] = 0.0;
        }
        if (a->noise) { nz[0] = a->noise[i]; nz[1] = a->noise[a->n + i]; }
        uint32_t f = aqo_step_env(w, a->s, i, act, a->noise ? nz : NULL, &rew,
                                  a->margin ? &mg : NULL);
        a->reward[i] = rew;
        a->flags[i] = (uint8_t)f;
        if (a->margin) a->margin[i] = mg;
        if (a->autoreset && (f & AQO_FLAG_DONE)) aqo_reset_env(w, a->s, i);
    }
}

void aqo_step_batch(const aqo_world* w, aqo_state* s, int64_t n, const int32_t* actions_i32,
                    const float* actions_f32, const double* noise, double* reward, uint8_t* flags,
                    double* margin, int autoreset, int nthreads) {
    step_args a = {w, s, n, actions_i32, actions_f32, noise, reward, flags, margin, autoreset};
    parallel_ranges(n, nthreads, step_range, &a);
}

typedef struct {
    const aqo_world* w; aqo_state* s; int64_t steps; uint64_t step0;
    int64_t episodes[AQO_MAX_THREADS]; double rsum[AQO_MAX_THREADS];
} rollout_args;

static void rollout_range(void* p, int64_t lo, int64_t hi, int tid) {
    rollout_args* a = (rollout_args*)p;
    const aqo_world* w = a->w;
    int64_t episodes = 0;
    double rsum = 0.0;
    for (int64_t i = lo; i < hi; ++i) {
        uint32_t env_id = (uint32_t)(w->env_id_base + (uint64_t)i);
        for (int64_t k = 0; k < a->steps; ++k) {
            double act[2], rew;
            if (w->continuous) {
                float t2[2];
                action_continuous_k(w->key_action, env_id, a->step0 + (uint64_t)k, t2);
                act[0] = t2[0];
                act[1] = t2[1];
            } else {
                act[0] = (double)action_discrete_k(w->key_action, env_id, a->step0 + (uint64_t)k);
                act[1] = 0.0;
            }
            uint32_t f = aqo_step_env(w, a->s, i, act, NULL, &rew, NULL);
            rsum += rew;
            if (f & AQO_FLAG_DONE) {
                ++episodes;
                aqo_reset_env(w, a->s, i);
            }
        }
    }
    a->episodes[tid] += episodes;
    a->rsum[tid] += rsum;
}

int64_t aqo_rollout_random(const aqo_world* w, aqo_state* s, int64_t n, int64_t steps,
                           uint64_t step0, int nthreads, double* reward_sum) {
    static rollout_args a; /* large per-thread arrays: keep off the stack; not re-entrant */
    memset(&a, 0, sizeof a);
    a.w = w; a.s = s; a.steps = steps; a.step0 = step0;
    parallel_ranges(n, nthreads, rollout_range, &a);
    int64_t episodes = 0;
    double rsum = 0.0;
    for (int t = 0; t < AQO_MAX_THREADS; ++t) { episodes += a.episodes[t]; rsum += a.rsum[t]; }
    if (reward_sum) *reward_sum = rsum;
    return episodes;
}
