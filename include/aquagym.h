/*
 * aquagym.h -- C ABI of libaquagym.so: the B200-native batched AquaGym environment step.
 *
 * This is the drop-in boundary for ONE hot path of ilVecc/AquaticGymEnv: the environment
 * `reset()` / `step()` of `gym_aqua/envs/aqua.py` (discrete `AquaEnv` and continuous
 * `AquaContinuousEnv`), batched over N independent env instances that live in GPU memory.
 * The reference has no native interface (it is pure Python); every entry point below cites
 * the reference method it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - Plain C: opaque handle, raw device pointers, sizes.  No torch / C++ types.
 *   - Every function returns AQUA_OK (0) or a negative aqua_status; the message of the last
 *     failure on the calling thread is returned by aqua_last_error().
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All launches
 *     are asynchronous on that stream; only aqua_get_stats() and aqua_step_host() synchronise.
 *   - A handle is owned by one host thread at a time (calls on it are not thread safe); distinct
 *     handles are independent.  At most 64 step launches of one handle may be in flight at once on
 *     DIFFERENT streams (each takes one of 64 tile-counter sets, handed out round robin; launches on
 *     one stream serialise and never collide); a CUDA graph that captured steps of a handle must not
 *     be replayed concurrently with other steps of that handle.
 *   - There is no CPU fallback: without a CUDA device aqua_create() fails.
 *
 * State layout (structure of arrays, one "plane" per quantity, `ld` elements apart):
 *     plane 0..6 : x, y, theta, goal_x, goal_y, wave_x, wave_y   (float, or double if
 *                  precision == AQUA_F64)
 *     plane 7    : uint32 word  (episode << 11) | time           (time saturates at 2047)
 *   The observation of the reference (aqua.py:126,213: concat(boat_state[3], goal[2])) is planes
 *   0..4, i.e. the first 5 * ld elements viewed as [5][ld]; no separate observation copy is made.
 */
#ifndef AQUAGYM_H
#define AQUAGYM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AQUAGYM_ABI_VERSION 3

typedef struct aqua_env aqua_env;

typedef enum {
    AQUA_OK = 0,
    AQUA_ERR_INVALID = -1,   /* bad argument / configuration */
    AQUA_ERR_CUDA = -2,      /* a CUDA runtime call failed (message has the CUDA error string) */
    AQUA_ERR_NO_DEVICE = -3, /* no usable CUDA device: there is no CPU fallback */
    AQUA_ERR_ALLOC = -4
} aqua_status;

enum { AQUA_F32 = 0, AQUA_F64 = 1 };

/* Kernel family: wide threads own 4 (float) / 2 (double) consecutive envs (16-byte plane accesses,
 * the bandwidth layout); narrow threads own one env (small, latency-bound slabs).  AUTO picks
 * narrow below about 680k envs per launch.  Results are identical either way. */
enum { AQUA_GEOM_AUTO = 0, AQUA_GEOM_WIDE = 1, AQUA_GEOM_NARROW = 2 };

/* element type of the `actions` buffer of aqua_step */
enum {
    AQUA_ACT_I32 = 0,   /* discrete: int32[N]  action index, aqua.py:33-43,154 */
    AQUA_ACT_I64 = 1,   /* discrete: int64[N]  (torch argmax / randint default) */
    AQUA_ACT_U8 = 2,    /* discrete: uint8[N] */
    AQUA_ACT_F32X2 = 3  /* continuous: float[N][2] = (thrust_left, thrust_right), aqua.py:143-151 */
};

/* bits of the per-env flags byte written by aqua_step (aqua.py:194-211) */
#define AQUA_FLAG_DONE 1u      /* `done` */
#define AQUA_FLAG_COLLIDED 2u  /* info['Termination.collided'] */
#define AQUA_FLAG_TIME 4u      /* info['Termination.time'] */
#define AQUA_FLAG_SUCCESS 8u   /* info['Termination.success'] */

#define AQUA_MAX_OBSTACLES 64
#define AQUA_TIME_BITS 11
#define AQUA_EPISODE_MASK 0x1FFFFFu
#define AQUA_N_PLANES 8
enum {
    AQUA_PLANE_X = 0, AQUA_PLANE_Y = 1, AQUA_PLANE_THETA = 2, AQUA_PLANE_GOAL_X = 3,
    AQUA_PLANE_GOAL_Y = 4, AQUA_PLANE_WAVE_X = 5, AQUA_PLANE_WAVE_Y = 6, AQUA_PLANE_TIME = 7
};
#define AQUA_TILE 1024 /* `ld` must be a multiple of this */

/* index of the counters returned by aqua_get_stats / aqua_stats_reduce */
enum {
    AQUA_STAT_EPISODES = 0,   /* finished episodes (steps with done) */
    AQUA_STAT_COLLIDED = 1,
    AQUA_STAT_TIMEOUT = 2,
    AQUA_STAT_SUCCESS = 3,
    AQUA_STAT_LENGTH_SUM = 4, /* sum of episode lengths of the finished episodes */
    AQUA_STAT_REWARD_SUM = 5, /* sum of ALL step rewards, fixed point: value * 2^20 */
    AQUA_STAT_STEPS = 6,      /* env-steps executed */
    AQUA_STAT_BAD_ACTIONS = 7, /* discrete index outside {0,1,2} or thrust outside [0.2,0.5] */
    AQUA_STAT_FORCED_RESETS = 8, /* reset rejection loops (aqua.py:102-115) cut off after 4096 tries: the
                                    last try was accepted untested.  The reference would loop on; a
                                    non-zero count means the map leaves (almost) no free space */
    AQUA_N_STATS = 9
};
#define AQUA_REWARD_FIXED_POINT 1048576.0

/* Constructor arguments.  Replaces `AquaEnv.__init__(obstacles, waves, random_boat,
 * random_goal)` (aqua.py:13-98) plus the `continuous` class attribute (aqua.py:11,458-459). */
typedef struct {
    uint32_t struct_size;  /* = sizeof(aqua_config), for ABI evolution */
    int32_t device;        /* CUDA device ordinal */
    int64_t n_envs;        /* N >= 1 */
    int32_t continuous;    /* 0: AquaEnv, 1: AquaContinuousEnv */
    int32_t waves;         /* aqua.py:15   has_waves */
    int32_t random_boat;   /* aqua.py:16; 0 -> boat starts at (85, 45, 0), aqua.py:117 */
    int32_t random_goal;   /* aqua.py:17; 0 -> goal at (25, 80), aqua.py:107 */
    int32_t precision;     /* AQUA_F32 (throughput layout) or AQUA_F64 (parity-grade pose) */
    int32_t autoreset;     /* 1: a done env is reset inside the step kernel (vector-env
                              semantics: reward/flags are the terminal step's, the state planes
                              hold the first observation of the next episode); 0: the terminal
                              state stays until aqua_reset (the reference's single-env flow) */
    int32_t n_obstacles;   /* 0..AQUA_MAX_OBSTACLES */
    int32_t geometry;      /* AQUA_GEOM_AUTO (0), or force one kernel family (tests, tuning) */
    /* HOST array [n_obstacles][5] = {cx, cy, type (0 circle / 1 rectangle), a, b}: circle
     * a = radius; rectangle a = width, b = height, centre anchored (aqua.py:56-68,373-390) */
    const double* obstacles;
    uint64_t seed;         /* key of the Philox4x32-10 draw streams */
    uint64_t env_id_base;  /* global id of env 0 of this slab (rank * n_envs when sharded) */
    /* Optional caller-owned DEVICE state buffer of AQUA_N_PLANES * ld elements (element =
     * float or double per `precision`); NULL -> the library allocates it. */
    void* state;
    int64_t ld;            /* plane stride in elements, multiple of AQUA_TILE, >= n_envs;
                              0 -> n_envs rounded up to AQUA_TILE */
    int32_t keep_terminal; /* 1: the observation a finished env returned with done = True (aqua.py:213)
                              is kept in 5 extra planes (aqua_state_view.terminal) -- under autoreset
                              the state planes already hold the next episode's first observation */
    int32_t _reserved;
    void* terminal;        /* optional caller-owned DEVICE buffer for them, 5 * ld elements of the
                              state's element type; NULL -> the library allocates it */
} aqua_config;

typedef struct {
    void* planes[AQUA_N_PLANES]; /* device pointers, see "State layout" */
    int64_t ld;
    int64_t n_envs;
    int32_t precision;
    int32_t _pad;
    /* keep_terminal: x, y, theta, goal_x, goal_y of the last finished episode of each env, `ld` apart,
     * same element type as the state; valid for env i once it has reported done; else NULL */
    void* terminal[5];
} aqua_state_view;

/* ---- life cycle ---------------------------------------------------------------------- */
int aqua_abi_version(void);
const char* aqua_last_error(void);

/* gym.make(id, **kwargs) -> AquaEnv(...)  (gym_aqua/__init__.py:4-41, aqua.py:13).  The state
 * is undefined until the first aqua_reset. */
int aqua_create(aqua_env** out, const aqua_config* cfg);
/* env.close() (aqua.py:367-371) + release of device memory */
int aqua_destroy(aqua_env* env);
int aqua_state(aqua_env* env, aqua_state_view* out);

/* ---- the hot path -------------------------------------------------------------------- */
/* env.reset() (aqua.py:100-126).  mask: device uint8[N] (non-zero = reset that env) or NULL
 * for all.  Opens the next episode of each selected env and draws goal, boat pose and wave
 * speed by the reference's rejection loops from the Philox RESET / WAVE_INIT streams. */
int aqua_reset(aqua_env* env, void* stream, const uint8_t* mask);

/* env.step(action) (aqua.py:135-213) for all N envs.
 *   actions : device buffer, element type `action_dtype`
 *   reward  : device float[N] (double[N] if precision == AQUA_F64)  -- aqua.py:86-90
 *   flags   : device uint8[N], AQUA_FLAG_* bits                      -- aqua.py:194-208
 * The observation is the state planes (see aqua_state).  Out-of-range actions are clamped
 * (aqua.py:145-150 prints and clips; a bad discrete index raises IndexError in the reference,
 * here it selects "full throttle") and counted in AQUA_STAT_BAD_ACTIONS. */
int aqua_step(aqua_env* env, void* stream, const void* actions, int action_dtype, void* reward,
              uint8_t* flags);

/* Same, with the wave-speed increment of aqua.py:188 injected instead of drawn:
 * noise = device float[2][N] {nx[N], ny[N]} (NULL = Philox stream).  Parity/debug. */
int aqua_step_ex(aqua_env* env, void* stream, const void* actions, int action_dtype, void* reward,
                 uint8_t* flags, const float* noise);

/* Synthetic random policy (`action_space.sample()`, test_random.py:19-20) from the Philox
 * ACTION stream keyed by (global env id, step_index): writes actions of `action_dtype`. */
int aqua_random_actions(aqua_env* env, void* stream, uint64_t step_index, void* actions,
                        int action_dtype);

/* `steps` consecutive env.step() calls (aqua.py:135-213) in ONE launch: the state stays in registers
 * between the steps, per env-step only the action is read and reward + flags are written.
 *   actions : device [steps][row_stride] of `action_dtype` (continuous: float[steps][row_stride][2]);
 *             row k holds the actions of step k, env i at column i
 *   reward  : device float / double [steps][row_stride], or NULL
 *   flags   : device uint8 [steps][row_stride], or NULL
 *   row_stride : columns per row, >= n_envs and a multiple of 16 (columns >= n_envs are padding: never
 *             read as actions that count, but the outputs' padding may be written)
 * Identical, bit for bit, to `steps` calls of aqua_step on the rows.  With autoreset the next
 * episode of a finished env starts in the same launch (its first observation is NOT materialised
 * per step: the state planes hold the observation after the last step; keep_terminal keeps the
 * terminal one).  The first call allocates 144 bytes per env for the ring of pre-drawn next episodes
 * (four records + their tags; so it cannot be captured in a CUDA graph; later calls can).  Launches
 * of up to 160 steps draw the records an env used up before the first step of the next call; longer
 * launches, and dense maps, draw them inside the step loop. */
int aqua_step_many(aqua_env* env, void* stream, int64_t steps, const void* actions, int action_dtype,
                   void* reward, uint8_t* flags, int64_t row_stride);

/* The same under the synthetic random policy above (aqua_random_actions keyed by first_step_index,
 * first_step_index + 1, ...): no tape, no per-step outputs, the statistics count everything. */
int aqua_rollout_random(aqua_env* env, void* stream, int64_t steps, uint64_t first_step_index);

/* `steps` env.step() calls in ONE launch with the actions chosen IN the kernel by the reference's
 * Q-network (main/impl/dqn.py:301-314: Dense 64 relu, Dense 64 relu, Dense 3 linear) on the normalised
 * observation (main/impl/utils.py:23-27), greedy (test_dqn.py:14-22) or epsilon-greedy
 * (dqn.py:212-228): what an evaluation loop, or the acting side of the DQN loop, calls instead of
 * observe -> network -> argmax -> step per env-step.  Discrete-action handles only.
 *   weights : DEVICE float[AQUA_MLP_WEIGHTS] in Keras layout: W1[5][64], b1[64], W2[64][64], b2[64],
 *             W3[64][3], b3[3] (kernels are [in][out])
 *   epsilon : with this probability a step takes a uniform random action instead (0 = greedy); the
 *             draws are those of aqua_random_actions keyed by first_step_index + k
 *   actions_out : device int32 [steps][row_stride] or NULL: the actions taken
 *   reward / flags / row_stride : as in aqua_step_many (each may be NULL; row_stride may be 0 then)
 * The network is evaluated in float32 with fused multiply-adds in ascending input order; the action is
 * the first index of the largest output. */
#define AQUA_MLP_WEIGHTS 4739
int aqua_rollout_policy(aqua_env* env, void* stream, int64_t steps, const float* weights, float epsilon,
                        uint64_t first_step_index, int32_t* actions_out, void* reward, uint8_t* flags,
                        int64_t row_stride);

/* The same with the exploration rate and the first step index read ON THE DEVICE when the launch runs
 * (DEVICE float[1] and uint64[1]): a CUDA graph that captured the call replays it with the values
 * of the moment -- the acting side of a DQN loop captured as one graph per vector step. */
int aqua_rollout_policy_dev(aqua_env* env, void* stream, int64_t steps, const float* weights, const float* epsilon_dev,
                            const uint64_t* first_step_index_dev, int32_t* actions_out, void* reward, uint8_t* flags,
                            int64_t row_stride);

/* Host-buffer step: the call a host-side gym user makes.  Copies `actions` (HOST, preferably
 * pinned) to the device, steps, and copies obs (float[5][N] planes x,y,theta,gx,gy; double if
 * F64), reward and flags back to HOST buffers, chunked and pipelined over internal streams;
 * returns after everything has landed.  obs/reward/flags may be NULL to skip that copy.
 * `stream` is the stream earlier asynchronous calls on this handle were issued on: the internal
 * streams wait for it first.  Handles of at most 8192 envs (the single-env drop-in) take another
 * route: the kernels read the actions from and write the outputs to mapped pinned memory of the
 * library, on `stream` itself, and the call polls a completion word (two launches per step, no
 * copy-engine transfer; AQUA_HOST_NO_MAPPED in the environment disables it). */
int aqua_step_host(aqua_env* env, void* stream, const void* actions, int action_dtype, void* obs,
                   void* reward, uint8_t* flags);

/* Same; persistent_obs != 0 promises that `obs` is the very buffer of the previous aqua_step_host(_ex)
 * call on this handle and has not been written since.  If it is pinned memory (cudaHostAlloc /
 * cudaHostRegister / torch pin_memory), its goal planes (2 of the 5) are then brought up to date by
 * per-env writes for the envs whose episode changed instead of being copied whole: 17 instead of 25
 * bytes per env-step come back over PCIe (float32).  Any other call that
 * changes the state (aqua_step, aqua_step_many, aqua_reset, aqua_set_state, ...) makes the next host
 * step send the planes whole again, whatever the promise. */
int aqua_step_host_ex(aqua_env* env, void* stream, const void* actions, int action_dtype, void* obs,
                      void* reward, uint8_t* flags, int persistent_obs);

/* Observation rows for a policy network: device float[N][5] = (x, y, angle, goal_x, goal_y)
 * gathered from the state planes (aqua.py:126,213).  normalized != 0 applies the scaling of
 * AquaStateNormalizer.__call__ (main/impl/utils.py:23-27): obs / (high - low), angle + 0.5, so
 * that every component lies in [0, 1].  Always float32, whatever the env precision. */
int aqua_observe(aqua_env* env, void* stream, float* out, int normalized);

/* ---- state access (parity teacher-forcing, checkpoint/restore) ------------------------ */
/* canonical DEVICE double[9][N]: x, y, theta, wave_x, wave_y, goal_x, goal_y, time, episode */
int aqua_get_state(aqua_env* env, void* stream, double* out);
int aqua_set_state(aqua_env* env, void* stream, const double* in);

/* ---- episode statistics --------------------------------------------------------------- */
/* Fold the per-block counter slots into DEVICE int64[AQUA_N_STATS] (the vector that is
 * all-reduced across ranks with NCCL); clear != 0 zeroes the internal counters afterwards. */
int aqua_stats_reduce(aqua_env* env, void* stream, int64_t* out_dev, int clear);
/* Same into a HOST int64[AQUA_N_STATS]; synchronises `stream`. */
int aqua_get_stats(aqua_env* env, void* stream, int64_t* out_host, int clear);

/* ---- obstacle table inspection (host only) --------------------------------------------- */
/* The obstacle tests of aqua.py:429-451 are exact for every table; tables of more than 12
 * obstacles are first culled through a grid of AQUA_GRID_DIM x AQUA_GRID_DIM cells over the
 * 100 x 100 world.  This returns that grid as the kernels would use it, for tests of its
 * classification.  Pure host computation: needs no CUDA device and no handle.
 *   obstacles  : HOST [n_obstacles][5], as in aqua_config
 *   rows_out   : HOST double[n_obstacles][5] = {cx, cy, half_x, half_y, reach^2} in the kernels'
 *                order (circles first; a circle has zero half extents): a boat or goal centre p
 *                collides with row i iff max(|px-cx|-half_x,0)^2 + max(|py-cy|-half_y,0)^2 <= reach^2
 *   cells_out  : HOST uint64[AQUA_GRID_DIM * AQUA_GRID_DIM], cell = cy * AQUA_GRID_DIM + cx with
 *                cx = (int)(float(px) * AQUA_GRID_DIM / 100) clamped: bit i set = row i may be
 *                reached from that cell; AQUA_CELL_BLOCKED = every point of the cell collides
 *   use_grid_out : 1 if the kernels cull this table through the grid, 0 if they test all rows */
#define AQUA_GRID_DIM 96
#define AQUA_CELL_BLOCKED 0xFFFFFFFFFFFFFFFFull
int aqua_obstacle_grid(const double* obstacles, int n_obstacles, double* rows_out, uint64_t* cells_out,
                       int* use_grid_out);
/* The same plus the 16-bit word the kernels actually read per position (ABI version 3):
 *   cells16_out : HOST uint16[AQUA_GRID_DIM * AQUA_GRID_DIM] or NULL: bits 0-6 and 7-13 = the first
 *                 two rows listed in the cell's mask (AQUA_CELL16_NONE = none; both are always
 *                 tested), AQUA_CELL16_BLOCKED = every point of the cell collides,
 *                 AQUA_CELL16_OVERFLOW = the mask lists more than two rows (the kernel then walks
 *                 the full mask for that position) */
#define AQUA_CELL16_NONE 64
#define AQUA_CELL16_BLOCKED 0x4000
#define AQUA_CELL16_OVERFLOW 0x8000
int aqua_obstacle_grid_ex(const double* obstacles, int n_obstacles, double* rows_out, uint64_t* cells_out,
                          uint16_t* cells16_out, int* use_grid_out);

#ifdef __cplusplus
}
#endif
#endif /* AQUAGYM_H */
