// aqua_abi.cu -- the C ABI of libaquagym.so (include/aquagym.h): handle life cycle, argument
// validation, launch geometry and the host-buffer step pipeline.  No exceptions cross the ABI;
// every CUDA failure becomes a status + thread-local message.  There is no CPU fallback.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "aqua_kernels.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return status;
}

#define AQ_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(AQUA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                   \
    } while (0)

// restores the caller's current device on scope exit
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = (prev == dev) || (cudaSetDevice(dev) == cudaSuccess);
        if (prev == dev) prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

constexpr int kHostStreams = 4;       // host-step pipeline depth
constexpr int kTileCounters = 64;     // launches that may be in flight at once on one handle
constexpr int64_t kMappedMaxEnvs = 8192;  // host steps of at most this many envs go through mapped host memory
constexpr int64_t kHostChunkMin = 1024 * 1024;  // every copy costs several us to issue and start: few, large
                                                // ones (tools/e2e_probe.py: 1 chunk at 2^20 envs, 16 at 2^24)

size_t elem_size(int precision) { return precision == AQUA_F64 ? 8 : 4; }

size_t action_bytes(int dtype) {
    switch (dtype) {
        case AQUA_ACT_I32: return 4;
        case AQUA_ACT_I64: return 8;
        case AQUA_ACT_U8: return 1;
        case AQUA_ACT_F32X2: return 8;
        default: return 0;
    }
}

}  // namespace

struct aqua_env {
    aqua_config cfg;
    int device = 0;
    int sm_count = 0;
    void* state = nullptr;
    bool owns_state = false;
    int64_t ld = 0;
    void* d_tab = nullptr;
    double* d_tab64 = nullptr;
    unsigned long long* d_slots = nullptr;
    int64_t* d_stats_out = nullptr;
    uint32_t* d_tile_ctr = nullptr;  // kTileCounters counter sets, handed out round robin per launch
    unsigned int* d_forced = nullptr;  // AQUA_STAT_FORCED_RESETS
    uint4* d_stash = nullptr;          // step_many: ring of pre-drawn episodes, 4 x 32 bytes per env + tags (lazy)
    uint32_t* d_stash_tag = nullptr;
    void* d_term = nullptr;            // keep_terminal: 5 planes [ld]
    bool owns_term = false;
    unsigned launch_seq = 0;
    aqua::World world;
    // resident CTAs per SM, queried once at create: [action dtype][narrow (V = 1, 128 threads) /
    // wide (256 threads)]
    int occ_step[4][2] = {};
    // host-step pipeline (lazily created)
    cudaStream_t hs[kHostStreams] = {};
    cudaEvent_t ev_in = nullptr;
    bool host_ready = false;
    void* d_actions = nullptr;
    size_t d_actions_bytes = 0;
    void* d_reward = nullptr;
    uint8_t* d_flags = nullptr;
    // small slabs (aqua_step_host with at most kMappedMaxEnvs envs): one block of mapped pinned host
    // memory that the kernels read the actions from and write reward, flags and observation to
    void* h_map = nullptr;
    void* d_map = nullptr;  // the same memory as the device addresses it
    uint32_t map_seq = 0;   // sequence number of the last mapped step (the completion word the host polls)
    // goal planes of the caller's host observation buffer: sent whole once, then only the entries of
    // envs whose episode changed (see aqua_step_host)
    const void* host_obs_last = nullptr;  // the obs buffer whose goal planes are current, or NULL
};

namespace {

aqua::StepArgs base_args(aqua_env* env) {
    aqua::StepArgs a;
    memset(&a, 0, sizeof a);
    a.tile_ctr = env->d_tile_ctr + aqua::kTileCounterWords * (env->launch_seq++ % kTileCounters);
    a.w = env->world;
    a.planes = env->state;
    a.ld = env->ld;
    a.env_begin = 0;
    a.env_end = env->cfg.n_envs;
    a.stats = env->d_slots;
    a.term = env->d_term;
    a.stash = env->d_stash;
    a.stash_tag = env->d_stash_tag;
    return a;
}

// Launch geometry for `n` envs.  Large slabs: wide threads (16-byte plane accesses), 256-thread
// CTAs, a persistent grid of (resident CTAs per SM) x (SM count).  Smaller slabs are latency-bound:
// they use narrow threads (one env each) in 128-thread CTAs, which quarters the dependent work per
// thread and spreads the slab over all SMs.  The crossover was measured on one box
// (profiles/r03z_single_geom.txt): narrow 3.99e10 against wide 3.78e10 env-steps/s at 2^19 envs,
// 4.61e10 against 5.35e10 at 2^20 -- about 4.5 wide tiles per resident warp.
int pick_geom(const aqua_env* env, int64_t n, bool rollout, int action_dtype, aqua::LaunchGeom* g) {
    const int wide = aqua::envs_per_thread(env->cfg.precision);
    const bool narrow = env->cfg.geometry == AQUA_GEOM_NARROW ||
                        (env->cfg.geometry == AQUA_GEOM_AUTO && n < (int64_t)256 * wide * env->sm_count * 9 / 2);  // ~680k envs
    const int V = narrow ? 1 : wide;
    const int threads = narrow ? aqua::kNarrowThreads : aqua::kWideThreads;
    const int ti = narrow ? 0 : 1;
    (void)rollout;
    const int per_sm = env->occ_step[action_dtype][ti];
    const int64_t tiles = (n + (int64_t)threads * V - 1) / ((int64_t)threads * V);
    const int64_t resident = (int64_t)per_sm * env->sm_count;
    g->threads = threads;
    g->v = V;
    g->blocks = (int)std::max<int64_t>(1, std::min(tiles, resident));
    return AQUA_OK;
}

int check_env(const aqua_env* env) {
    if (env == nullptr) return fail(AQUA_ERR_INVALID, "env handle is NULL");
    return AQUA_OK;
}

int check_action_dtype(const aqua_env* env, int dtype) {
    if (action_bytes(dtype) == 0) return fail(AQUA_ERR_INVALID, "unknown action dtype %d", dtype);
    if (env->cfg.continuous && dtype != AQUA_ACT_F32X2)
        return fail(AQUA_ERR_INVALID, "AquaContinuousEnv takes AQUA_ACT_F32X2 actions (aqua.py:143-151), got dtype %d", dtype);
    if (!env->cfg.continuous && dtype == AQUA_ACT_F32X2)
        return fail(AQUA_ERR_INVALID, "AquaEnv takes a discrete action index (aqua.py:154), got AQUA_ACT_F32X2");
    return AQUA_OK;
}

// Host step of a small slab (the single-env drop-in above all: aqua.py:135-213 once per call).  Copies
// through the copy engines cost several microseconds each to issue and to complete, four of them
// per step; here the step kernel reads the actions from, and writes reward and flags to, a block of
// MAPPED pinned host memory, a second small kernel writes the five observation planes there, and
// the call waits once: two launches and one synchronisation per step.
int step_host_mapped(aqua_env* env, cudaStream_t s, const void* actions, int action_dtype, void* obs, void* reward,
                     uint8_t* flags) {
    const int64_t n = env->cfg.n_envs;
    const size_t es = elem_size(env->cfg.precision);
    const size_t ab = action_bytes(action_dtype);
    const size_t cap = (size_t)kMappedMaxEnvs;
    // layout: actions [cap x 8] | observation [5][n] (cap x 5 x 8) | reward [cap x 8] | flags [cap]
    //         | completion word
    const size_t off_obs = cap * 8, off_rew = off_obs + cap * 40, off_flg = off_rew + cap * 8, off_done = off_flg + cap,
                 total = off_done + 64;
    if (env->h_map == nullptr) {
        AQ_CUDA(cudaHostAlloc(&env->h_map, total, cudaHostAllocMapped));
        AQ_CUDA(cudaHostGetDevicePointer(&env->d_map, env->h_map, 0));
        memset(env->h_map, 0, total);
    }
    char* h = static_cast<char*>(env->h_map);
    char* d = static_cast<char*>(env->d_map);
    memcpy(h, actions, (size_t)n * ab);
    aqua::StepArgs a = base_args(env);
    a.actions = d;
    a.reward = d + off_rew;
    a.flags = reinterpret_cast<uint8_t*>(d + off_flg);
    aqua::LaunchGeom g;
    pick_geom(env, n, false, action_dtype, &g);
    AQ_CUDA(aqua::launch_step(env->cfg.precision, action_dtype, a, g, s));
    // The observation kernel runs after the step kernel (stream order) and sets the completion word
    // last; the host polls that word -- a synchronisation through the driver costs several microseconds
    // more -- and falls back to the stream if it does not arrive (a failed launch shows up there).
    const uint32_t seq = ++env->map_seq;
    AQ_CUDA(aqua::launch_copy_obs_planes(env->cfg.precision, env->state, env->ld, n, d + off_obs,
                                         reinterpret_cast<uint32_t*>(d + off_done), seq, s));
    volatile uint32_t* done = reinterpret_cast<volatile uint32_t*>(h + off_done);
    bool arrived = false;
    for (int spin = 0; spin < 4000000; ++spin) {
        if (*done == seq) {
            arrived = true;
            break;
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    if (!arrived) AQ_CUDA(cudaStreamSynchronize(s));
    std::atomic_thread_fence(std::memory_order_acquire);
    if (obs != nullptr) memcpy(obs, h + off_obs, (size_t)5 * (size_t)n * es);
    if (reward != nullptr) memcpy(reward, h + off_rew, (size_t)n * es);
    if (flags != nullptr) memcpy(flags, h + off_flg, (size_t)n);
    env->host_obs_last = nullptr;
    return AQUA_OK;
}

int ensure_host_pipeline(aqua_env* env, int action_dtype) {
    if (!env->host_ready) {
        for (int i = 0; i < kHostStreams; ++i) AQ_CUDA(cudaStreamCreateWithFlags(&env->hs[i], cudaStreamNonBlocking));
        AQ_CUDA(cudaEventCreateWithFlags(&env->ev_in, cudaEventDisableTiming));
        AQ_CUDA(cudaMalloc(&env->d_reward, (size_t)env->ld * elem_size(env->cfg.precision)));
        AQ_CUDA(cudaMalloc(&env->d_flags, (size_t)env->ld));
        env->host_ready = true;
    }
    const size_t need = (size_t)env->ld * action_bytes(action_dtype);
    if (env->d_actions_bytes < need) {
        if (env->d_actions) AQ_CUDA(cudaFree(env->d_actions));
        env->d_actions = nullptr;
        env->d_actions_bytes = 0;
        AQ_CUDA(cudaMalloc(&env->d_actions, need));
        env->d_actions_bytes = need;
    }
    return AQUA_OK;
}

}  // namespace

namespace {
int rollout_policy_impl(aqua_env* env, void* stream, int64_t steps, const float* weights, float epsilon,
                        uint64_t first_step_index, const float* epsilon_dev, const uint64_t* first_step_index_dev,
                        int32_t* actions_out, void* reward, uint8_t* flags, int64_t row_stride);
}

extern "C" {

int aqua_abi_version(void) { return AQUAGYM_ABI_VERSION; }

const char* aqua_last_error(void) { return g_err; }

int aqua_create(aqua_env** out, const aqua_config* cfg) {
    if (out == nullptr) return fail(AQUA_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (cfg == nullptr) return fail(AQUA_ERR_INVALID, "cfg is NULL");
    if (cfg->struct_size != sizeof(aqua_config))
        return fail(AQUA_ERR_INVALID, "aqua_config.struct_size = %u, this library expects %zu (ABI %d)",
                    cfg->struct_size, sizeof(aqua_config), AQUAGYM_ABI_VERSION);
    if (cfg->n_envs < 1) return fail(AQUA_ERR_INVALID, "n_envs must be >= 1, got %lld", (long long)cfg->n_envs);
    if (cfg->precision != AQUA_F32 && cfg->precision != AQUA_F64)
        return fail(AQUA_ERR_INVALID, "precision must be AQUA_F32 or AQUA_F64, got %d", cfg->precision);
    if (cfg->geometry < AQUA_GEOM_AUTO || cfg->geometry > AQUA_GEOM_NARROW)
        return fail(AQUA_ERR_INVALID, "geometry must be AQUA_GEOM_AUTO, _WIDE or _NARROW, got %d", cfg->geometry);
    if (cfg->n_obstacles < 0 || cfg->n_obstacles > AQUA_MAX_OBSTACLES)
        return fail(AQUA_ERR_INVALID, "n_obstacles must be in [0, %d], got %d", AQUA_MAX_OBSTACLES, cfg->n_obstacles);
    if (cfg->n_obstacles > 0 && cfg->obstacles == nullptr)
        return fail(AQUA_ERR_INVALID, "n_obstacles = %d but obstacles is NULL", cfg->n_obstacles);
    for (int k = 0; k < cfg->n_obstacles; ++k) {
        const double* o = cfg->obstacles + 5 * k;
        if (!(o[2] == 0.0 || o[2] == 1.0))
            return fail(AQUA_ERR_INVALID, "obstacle %d: type must be 0 (circle) or 1 (rectangle), got %g", k, o[2]);
        if (!(o[3] >= 0.0) || !(o[2] == 0.0 || o[4] >= 0.0))
            return fail(AQUA_ERR_INVALID, "obstacle %d: negative or NaN size", k);
    }
    if (cfg->env_id_base + (uint64_t)cfg->n_envs > (1ull << 32))
        return fail(AQUA_ERR_INVALID, "global env ids must fit 32 bits (env_id_base + n_envs <= 2^32)");
    int64_t ld = cfg->ld;
    if (ld == 0) ld = (cfg->n_envs + AQUA_TILE - 1) / AQUA_TILE * AQUA_TILE;
    if (ld < cfg->n_envs || ld % AQUA_TILE != 0)
        return fail(AQUA_ERR_INVALID, "ld = %lld must be a multiple of %d and >= n_envs", (long long)ld, AQUA_TILE);
    if (cfg->state != nullptr && (reinterpret_cast<uintptr_t>(cfg->state) & 15u))
        return fail(AQUA_ERR_INVALID, "state buffer must be 16-byte aligned");
    if (cfg->terminal != nullptr && (reinterpret_cast<uintptr_t>(cfg->terminal) & 15u))
        return fail(AQUA_ERR_INVALID, "terminal buffer must be 16-byte aligned");

    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(AQUA_ERR_NO_DEVICE, "no CUDA device (%s): libaquagym has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= n_dev)
        return fail(AQUA_ERR_INVALID, "device %d out of range [0, %d)", cfg->device, n_dev);

    DeviceGuard guard(cfg->device);
    if (!guard.ok) return fail(AQUA_ERR_CUDA, "cudaSetDevice(%d) failed", cfg->device);

    aqua_env* env = new (std::nothrow) aqua_env();
    if (env == nullptr) return fail(AQUA_ERR_ALLOC, "out of host memory");
    env->cfg = *cfg;
    env->cfg.obstacles = nullptr;
    env->cfg.ld = ld;
    env->device = cfg->device;
    env->ld = ld;

#define AQ_CREATE(call)                                                                            \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            const int rc_ = fail(AQUA_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));   \
            aqua_destroy(env);                                                                     \
            /* aqua_destroy succeeded or not, report the original failure */                       \
            fail(AQUA_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                   \
            return rc_;                                                                            \
        }                                                                                          \
    } while (0)

    AQ_CREATE(cudaDeviceGetAttribute(&env->sm_count, cudaDevAttrMultiProcessorCount, env->device));
    const size_t es = elem_size(cfg->precision);
    if (cfg->state != nullptr) {
        env->state = cfg->state;
    } else {
        AQ_CREATE(cudaMalloc(&env->state, (size_t)AQUA_N_PLANES * (size_t)ld * es));
        env->owns_state = true;
    }
    // obstacle tables
    const size_t tab_bytes = aqua::obstacle_table_bytes(cfg->precision);
    int k_circ = 0, use_grid = 0;
    {
        void* h_tab = malloc(tab_bytes);
        double h_tab64[AQUA_MAX_OBSTACLES * 5] = {0};
        if (h_tab == nullptr) {
            aqua_destroy(env);
            return fail(AQUA_ERR_ALLOC, "out of host memory");
        }
        k_circ = aqua::build_obstacle_tables(cfg->precision, cfg->n_obstacles, cfg->obstacles, h_tab, h_tab64, &use_grid);
        cudaError_t e1 = cudaMalloc(&env->d_tab, tab_bytes);
        if (e1 == cudaSuccess) e1 = cudaMemcpy(env->d_tab, h_tab, tab_bytes, cudaMemcpyHostToDevice);
        free(h_tab);
        AQ_CREATE(e1);
        AQ_CREATE(cudaMalloc(&env->d_tab64, sizeof h_tab64));
        AQ_CREATE(cudaMemcpy(env->d_tab64, h_tab64, sizeof h_tab64, cudaMemcpyHostToDevice));
    }
    AQ_CREATE(cudaMalloc(&env->d_slots, sizeof(unsigned long long) * aqua::kStatSlots * aqua::kSlotStats));
    AQ_CREATE(cudaMemset(env->d_slots, 0, sizeof(unsigned long long) * aqua::kStatSlots * aqua::kSlotStats));
    AQ_CREATE(cudaMalloc(&env->d_forced, sizeof(unsigned int)));
    AQ_CREATE(cudaMemset(env->d_forced, 0, sizeof(unsigned int)));
    if (cfg->keep_terminal && cfg->terminal != nullptr) {
        env->d_term = cfg->terminal;
    } else if (cfg->keep_terminal) {
        AQ_CREATE(cudaMalloc(&env->d_term, (size_t)5 * (size_t)ld * es));
        env->owns_term = true;
        AQ_CREATE(cudaMemset(env->d_term, 0, (size_t)5 * (size_t)ld * es));
    }
    AQ_CREATE(cudaMalloc(&env->d_stats_out, sizeof(int64_t) * AQUA_N_STATS));
    AQ_CREATE(cudaMalloc(&env->d_tile_ctr, sizeof(uint32_t) * aqua::kTileCounterWords * kTileCounters));
    AQ_CREATE(cudaMemset(env->d_tile_ctr, 0, sizeof(uint32_t) * aqua::kTileCounterWords * kTileCounters));

    aqua::World& w = env->world;
    memset(&w, 0, sizeof w);
    aqua::make_philox_keys(cfg->seed, &w.keys);
    w.env_id_base = cfg->env_id_base;
    w.tab = env->d_tab;
    w.tab64 = env->d_tab64;
    w.k_obst = cfg->n_obstacles;
    w.k_circ = k_circ;
    {
        double reach2 = 6.25;  // rectangles: (2.5)^2
        for (int k = 0; k < cfg->n_obstacles; ++k)
            if (cfg->obstacles[5 * k + 2] == 0.0) {
                const double r = cfg->obstacles[5 * k + 3] + 2.5;
                reach2 = std::max(reach2, r * r);
            }
        w.reset_band = (float)(1e-4 * reach2);
    }
    w.waves = cfg->waves != 0;
    w.random_boat = cfg->random_boat != 0;
    w.random_goal = cfg->random_goal != 0;
    w.autoreset = cfg->autoreset != 0;
    w.use_grid = use_grid != 0;
    w.forced = env->d_forced;
    aqua::fill_small_table(cfg->precision, cfg->n_obstacles, cfg->obstacles, use_grid, &w);
    if (const char* s = getenv("AQUA_SMALL_TABLE"))  // tests / tuning: 0 = the shared-memory table for every layout
        if (atoi(s) == 0) w.small_mode = 0;

    AQ_CREATE(aqua::configure_kernels(cfg->precision));
    for (int ti = 0; ti < 2; ++ti) {
        const int threads = ti == 0 ? aqua::kNarrowThreads : aqua::kWideThreads;
        const int v = ti == 0 ? 1 : aqua::envs_per_thread(cfg->precision);
        for (int act = 0; act < 4; ++act) {
            AQ_CREATE(aqua::step_occupancy(cfg->precision, v, act, threads, use_grid, &env->occ_step[act][ti]));
            env->occ_step[act][ti] = std::max(env->occ_step[act][ti], 1);
            if (ti == 1) {  // wide slabs run the streaming kernel where there is one: size the grid for it
                int occ = 0;
                AQ_CREATE(aqua::stream_occupancy(cfg->precision, act, cfg->waves != 0, threads, use_grid, &occ));
                if (occ > 0) env->occ_step[act][ti] = occ;
            }
        }
    }
    AQ_CREATE(aqua::launch_init_state(cfg->precision, env->state, ld, nullptr));
    AQ_CREATE(cudaStreamSynchronize(nullptr));
#undef AQ_CREATE
    *out = env;
    return AQUA_OK;
}

int aqua_destroy(aqua_env* env) {
    if (env == nullptr) return AQUA_OK;
    DeviceGuard guard(env->device);
    cudaError_t first = cudaSuccess;
    auto keep = [&](cudaError_t e) {
        if (first == cudaSuccess && e != cudaSuccess) first = e;
    };
    if (env->host_ready) {
        for (int i = 0; i < kHostStreams; ++i)
            if (env->hs[i]) {
                keep(cudaStreamSynchronize(env->hs[i]));
                keep(cudaStreamDestroy(env->hs[i]));
            }
        if (env->ev_in) keep(cudaEventDestroy(env->ev_in));
    }
    if (env->h_map) keep(cudaFreeHost(env->h_map));
    if (env->d_actions) keep(cudaFree(env->d_actions));
    if (env->d_reward) keep(cudaFree(env->d_reward));
    if (env->d_flags) keep(cudaFree(env->d_flags));
    if (env->owns_state && env->state) keep(cudaFree(env->state));
    if (env->d_tab) keep(cudaFree(env->d_tab));
    if (env->d_tab64) keep(cudaFree(env->d_tab64));
    if (env->d_slots) keep(cudaFree(env->d_slots));
    if (env->d_stats_out) keep(cudaFree(env->d_stats_out));
    if (env->d_tile_ctr) keep(cudaFree(env->d_tile_ctr));
    if (env->d_forced) keep(cudaFree(env->d_forced));
    if (env->d_stash) keep(cudaFree(env->d_stash));
    if (env->owns_term && env->d_term) keep(cudaFree(env->d_term));
    delete env;
    if (first != cudaSuccess) return fail(AQUA_ERR_CUDA, "aqua_destroy: %s", cudaGetErrorString(first));
    return AQUA_OK;
}

int aqua_state(aqua_env* env, aqua_state_view* out) {
    if (int rc = check_env(env)) return rc;
    if (out == nullptr) return fail(AQUA_ERR_INVALID, "out is NULL");
    const size_t es = elem_size(env->cfg.precision);
    for (int p = 0; p < AQUA_N_PLANES; ++p)
        out->planes[p] = static_cast<char*>(env->state) + (size_t)p * (size_t)env->ld * es;
    out->ld = env->ld;
    out->n_envs = env->cfg.n_envs;
    out->precision = env->cfg.precision;
    out->_pad = 0;
    for (int k = 0; k < 5; ++k)
        out->terminal[k] = env->d_term ? static_cast<char*>(env->d_term) + (size_t)k * (size_t)env->ld * es : nullptr;
    return AQUA_OK;
}

int aqua_reset(aqua_env* env, void* stream, const uint8_t* mask) {
    if (int rc = check_env(env)) return rc;
    env->host_obs_last = nullptr;
    DeviceGuard guard(env->device);
    const aqua::StepArgs a = base_args(env);
    AQ_CUDA(aqua::launch_reset(env->cfg.precision, a, mask, static_cast<cudaStream_t>(stream)));
    return AQUA_OK;
}

int aqua_step_ex(aqua_env* env, void* stream, const void* actions, int action_dtype, void* reward,
                 uint8_t* flags, const float* noise) {
    if (int rc = check_env(env)) return rc;
    if (int rc = check_action_dtype(env, action_dtype)) return rc;
    if (actions == nullptr || reward == nullptr || flags == nullptr)
        return fail(AQUA_ERR_INVALID, "actions, reward and flags must be device pointers, got NULL");
    env->host_obs_last = nullptr;
    if ((reinterpret_cast<uintptr_t>(actions) & 15u) || (reinterpret_cast<uintptr_t>(reward) & 15u) ||
        (reinterpret_cast<uintptr_t>(flags) & 3u))
        return fail(AQUA_ERR_INVALID, "actions and reward must be 16-byte aligned, flags 4-byte aligned");
    DeviceGuard guard(env->device);
    aqua::StepArgs a = base_args(env);
    a.actions = actions;
    a.reward = reward;
    a.flags = flags;
    if (noise != nullptr) {
        a.noise_x = noise;
        a.noise_y = noise + env->cfg.n_envs;
    }
    aqua::LaunchGeom g;
    if (int rc = pick_geom(env, env->cfg.n_envs, false, action_dtype, &g)) return rc;
    AQ_CUDA(aqua::launch_step(env->cfg.precision, action_dtype, a, g, static_cast<cudaStream_t>(stream)));
    return AQUA_OK;
}

int aqua_step(aqua_env* env, void* stream, const void* actions, int action_dtype, void* reward,
              uint8_t* flags) {
    return aqua_step_ex(env, stream, actions, action_dtype, reward, flags, nullptr);
}

int aqua_random_actions(aqua_env* env, void* stream, uint64_t step_index, void* actions, int action_dtype) {
    if (int rc = check_env(env)) return rc;
    if (int rc = check_action_dtype(env, action_dtype)) return rc;
    if (actions == nullptr) return fail(AQUA_ERR_INVALID, "actions is NULL");
    DeviceGuard guard(env->device);
    AQ_CUDA(aqua::launch_random_actions(env->world, env->cfg.n_envs, step_index, actions, action_dtype,
                                        static_cast<cudaStream_t>(stream)));
    return AQUA_OK;
}

namespace {

// the episode ring of step_many_kernel, allocated on first use (144 bytes per env; not capturable)
int ensure_stash(aqua_env* env, cudaStream_t s) {
    if (env->d_stash != nullptr) return AQUA_OK;
    const size_t depth = (size_t)aqua::stash_ring_depth();
    void* ptr = nullptr;
    AQ_CUDA(cudaMalloc(&ptr, (size_t)env->ld * depth * 36));  // records, then the tag plane
    env->d_stash = static_cast<uint4*>(ptr);
    env->d_stash_tag = reinterpret_cast<uint32_t*>(static_cast<char*>(ptr) + (size_t)env->ld * depth * 32);
    AQ_CUDA(aqua::launch_stash_init(ptr, (size_t)env->ld * depth * 36, s));
    return AQUA_OK;
}

// Geometry of a multi-step launch: every env belongs to one thread for the whole launch, so the
// grid is not persistent.  More envs per thread give a thread more independent work and cost fewer
// instructions per env-step, but leave fewer warps.  Measured over slab sizes on one box, launches
// back to back (profiles/r04b_step_many_v_by_size_*.txt, three maps): one env per thread up to ~330
// envs per SM (49,152 envs on a B200), two envs per thread above (at 65,536 envs the two are equal on the
// default map, two is 27 % ahead with continuous actions); four per thread (128 registers) no longer
// wins anywhere.  128-thread CTAs on dense maps from ~1300 envs per SM (the culling grid is staged
// once per CTA), 64-thread CTAs otherwise.
// AQUA_MANY_V / AQUA_MANY_THREADS (environment) override the choice: tuning only.
int pick_many_geom(const aqua_env* env, int64_t n, aqua::LaunchGeom* g) {
    const int wide = aqua::envs_per_thread(env->cfg.precision);
    const int64_t per_sm = n / std::max(env->sm_count, 1);
    const bool grid = env->world.use_grid != 0;
    int threads = 64, v = 2;
    if (per_sm <= 332) v = 1;
    if (grid && per_sm >= 1328) threads = 128;
    if (env->cfg.geometry == AQUA_GEOM_WIDE) v = wide;
    if (env->cfg.geometry == AQUA_GEOM_NARROW) v = 1;
    if (const char* s = getenv("AQUA_MANY_V")) {
        const int x = atoi(s);
        if (x == 1 || x == 2 || (x == 4 && wide == 4)) v = x;
    }
    if (const char* s = getenv("AQUA_MANY_THREADS")) {
        const int x = atoi(s);
        if (x == 32 || x == 64 || x == 128) threads = x;  // the kernel is bounded to 128 threads
    }
    g->v = v;
    g->threads = threads;
    g->blocks = 0;  // launch_step_many derives it
    return AQUA_OK;
}

}  // namespace

int aqua_step_many(aqua_env* env, void* stream, int64_t steps, const void* actions, int action_dtype,
                   void* reward, uint8_t* flags, int64_t row_stride) {
    if (int rc = check_env(env)) return rc;
    if (int rc = check_action_dtype(env, action_dtype)) return rc;
    if (steps < 0 || steps > 0x7fffffff) return fail(AQUA_ERR_INVALID, "steps must be in [0, 2^31)");
    if (actions == nullptr) return fail(AQUA_ERR_INVALID, "actions is NULL");
    if (row_stride < env->cfg.n_envs || row_stride % 16 != 0)
        return fail(AQUA_ERR_INVALID, "row_stride = %lld must be a multiple of 16 and >= n_envs = %lld",
                    (long long)row_stride, (long long)env->cfg.n_envs);
    if ((reinterpret_cast<uintptr_t>(actions) & 15u) || (reinterpret_cast<uintptr_t>(reward) & 15u) ||
        (reinterpret_cast<uintptr_t>(flags) & 15u))
        return fail(AQUA_ERR_INVALID, "actions, reward and flags must be 16-byte aligned");
    if (steps == 0) return AQUA_OK;
    env->host_obs_last = nullptr;
    DeviceGuard guard(env->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int rc = ensure_stash(env, s)) return rc;
    aqua::StepArgs a = base_args(env);
    a.actions = actions;
    a.reward = reward;
    a.flags = flags;
    a.row_stride = row_stride;
    a.steps = steps;
    aqua::LaunchGeom g;
    if (int rc = pick_many_geom(env, env->cfg.n_envs, &g)) return rc;
    AQ_CUDA(aqua::launch_step_many(env->cfg.precision, action_dtype, 1, a, g, s));
    return AQUA_OK;
}

int aqua_rollout_random(aqua_env* env, void* stream, int64_t steps, uint64_t first_step_index) {
    if (int rc = check_env(env)) return rc;
    if (steps < 0 || steps > 0x7fffffff) return fail(AQUA_ERR_INVALID, "steps must be in [0, 2^31)");
    if (steps == 0) return AQUA_OK;
    env->host_obs_last = nullptr;
    DeviceGuard guard(env->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int rc = ensure_stash(env, s)) return rc;
    aqua::StepArgs a = base_args(env);
    a.steps = steps;
    a.step0 = first_step_index;
    a.row_stride = 0;
    aqua::LaunchGeom g;
    if (int rc = pick_many_geom(env, env->cfg.n_envs, &g)) return rc;
    AQ_CUDA(aqua::launch_step_many(env->cfg.precision, env->cfg.continuous ? AQUA_ACT_F32X2 : AQUA_ACT_I32, 0, a, g, s));
    return AQUA_OK;
}

int aqua_rollout_policy(aqua_env* env, void* stream, int64_t steps, const float* weights, float epsilon,
                        uint64_t first_step_index, int32_t* actions_out, void* reward, uint8_t* flags,
                        int64_t row_stride) {
    return rollout_policy_impl(env, stream, steps, weights, epsilon, first_step_index, nullptr, nullptr, actions_out, reward,
                               flags, row_stride);
}

int aqua_rollout_policy_dev(aqua_env* env, void* stream, int64_t steps, const float* weights, const float* epsilon_dev,
                            const uint64_t* first_step_index_dev, int32_t* actions_out, void* reward, uint8_t* flags,
                            int64_t row_stride) {
    if (epsilon_dev == nullptr || first_step_index_dev == nullptr)
        return fail(AQUA_ERR_INVALID, "epsilon_dev and first_step_index_dev must not be NULL");
    return rollout_policy_impl(env, stream, steps, weights, 0.f, 0, epsilon_dev, first_step_index_dev, actions_out, reward,
                               flags, row_stride);
}

}  // extern "C"

namespace {
int rollout_policy_impl(aqua_env* env, void* stream, int64_t steps, const float* weights, float epsilon,
                        uint64_t first_step_index, const float* epsilon_dev, const uint64_t* first_step_index_dev,
                        int32_t* actions_out, void* reward, uint8_t* flags, int64_t row_stride) {
    if (int rc = check_env(env)) return rc;
    if (env->cfg.continuous) return fail(AQUA_ERR_INVALID, "aqua_rollout_policy: the Q-network chooses among the discrete actions");
    if (steps < 0 || steps > 0x7fffffff) return fail(AQUA_ERR_INVALID, "steps must be in [0, 2^31)");
    if (weights == nullptr) return fail(AQUA_ERR_INVALID, "weights is NULL");
    if (!(epsilon >= 0.f && epsilon <= 1.f)) return fail(AQUA_ERR_INVALID, "epsilon must be in [0, 1]");
    const bool outputs = actions_out != nullptr || reward != nullptr || flags != nullptr;
    if (outputs && (row_stride < env->cfg.n_envs || row_stride % 16 != 0))
        return fail(AQUA_ERR_INVALID, "row_stride = %lld must be a multiple of 16 and >= n_envs = %lld",
                    (long long)row_stride, (long long)env->cfg.n_envs);
    if ((reinterpret_cast<uintptr_t>(actions_out) & 15u) || (reinterpret_cast<uintptr_t>(reward) & 15u) ||
        (reinterpret_cast<uintptr_t>(flags) & 15u) || (reinterpret_cast<uintptr_t>(weights) & 15u))
        return fail(AQUA_ERR_INVALID, "weights, actions_out, reward and flags must be 16-byte aligned");
    if (steps == 0) return AQUA_OK;
    env->host_obs_last = nullptr;
    DeviceGuard guard(env->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int rc = ensure_stash(env, s)) return rc;
    aqua::StepArgs a = base_args(env);
    a.reward = reward;
    a.flags = flags;
    a.row_stride = row_stride;
    a.steps = steps;
    a.step0 = first_step_index;
    a.mlp_w = weights;
    a.mlp_eps = epsilon;
    a.act_out = actions_out;
    a.mlp_eps_dev = epsilon_dev;
    a.step0_dev = reinterpret_cast<const unsigned long long*>(first_step_index_dev);
    AQ_CUDA(aqua::launch_step_policy(env->cfg.precision, a, 128, s));
    return AQUA_OK;
}
}  // namespace

extern "C" {

int aqua_step_host(aqua_env* env, void* stream, const void* actions, int action_dtype, void* obs,
                   void* reward, uint8_t* flags) {
    return aqua_step_host_ex(env, stream, actions, action_dtype, obs, reward, flags, 0);
}

int aqua_step_host_ex(aqua_env* env, void* stream, const void* actions, int action_dtype, void* obs,
                      void* reward, uint8_t* flags, int persistent_obs) {
    if (int rc = check_env(env)) return rc;
    if (int rc = check_action_dtype(env, action_dtype)) return rc;
    if (actions == nullptr) return fail(AQUA_ERR_INVALID, "actions is NULL");
    DeviceGuard guard(env->device);
    if (env->cfg.n_envs <= kMappedMaxEnvs && getenv("AQUA_HOST_NO_MAPPED") == nullptr)
        return step_host_mapped(env, static_cast<cudaStream_t>(stream), actions, action_dtype, obs, reward, flags);
    if (int rc = ensure_host_pipeline(env, action_dtype)) return rc;

    const int64_t n = env->cfg.n_envs;
    const size_t es = elem_size(env->cfg.precision);
    const size_t ab = action_bytes(action_dtype);
    // chunking: whole AQUA_TILEs, at least kHostChunkMin envs, at most 4 chunks per stream
    int64_t chunk = (n + kHostStreams * 4 - 1) / (kHostStreams * 4);
    chunk = std::max(chunk, kHostChunkMin);
    if (const char* e = getenv("AQUA_HOST_CHUNKS")) {  // tuning only: that many chunks (1 .. 16)
        const int64_t k = std::min<int64_t>(std::max<int64_t>(atoll(e), 1), kHostStreams * 4);
        chunk = (n + k - 1) / k;
    }
    chunk = (chunk + AQUA_TILE - 1) / AQUA_TILE * AQUA_TILE;

    AQ_CUDA(cudaEventRecord(env->ev_in, static_cast<cudaStream_t>(stream)));
    for (int i = 0; i < kHostStreams; ++i) AQ_CUDA(cudaStreamWaitEvent(env->hs[i], env->ev_in, 0));

    static_assert(AQUA_PLANE_X == 0 && AQUA_PLANE_Y == 1 && AQUA_PLANE_THETA == 2 && AQUA_PLANE_GOAL_X == 3 &&
                      AQUA_PLANE_GOAL_Y == 4, "the observation (aqua.py:213) is planes 0..4 in order");
    // a failure inside the loop must not return while copies into the caller's buffers are still
    // in flight: remember it, stop issuing, wait for every internal stream, then report
    cudaError_t err = cudaSuccess;
    const char* what = "";
#define AQ_TRY(call)                                  \
    if (err == cudaSuccess) {                         \
        err = (call);                                 \
        if (err != cudaSuccess) what = #call;         \
    }
    // The goal planes (2 of the 5 observation planes) only change for envs whose episode changed:
    // they travel whole when the caller's obs buffer is new to this handle (or anything else touched
    // the state in between, or the buffer is pageable) and as per-env writes of a small kernel straight
    // into the caller's pinned planes otherwise -- 17 instead of 25 bytes per env-step come back.
    char* obs_dev = nullptr;  // the caller's obs buffer as the device sees it (pinned memory only)
    if (persistent_obs != 0 && obs != nullptr && env->host_obs_last == obs) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, obs) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer != nullptr)
            obs_dev = static_cast<char*>(pa.devicePointer);
        else
            cudaGetLastError();  // pageable memory: not an error, just no shortcut
    }
    const bool sparse_goals = obs_dev != nullptr;
    const int n_obs_planes = sparse_goals ? 3 : 5;
    // All the caller's output buffers pinned and mapped: ONE kernel per chunk writes them straight
    // from the SMs (launch_pack_host) instead of three copy-engine transfers and the goal kernel.
    // (AQUA_HOST_NO_PACK in the environment keeps the copy-engine route: tuning / tests.)
    auto mapped = [](const void* h) -> char* {
        if (h == nullptr) return nullptr;
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, h) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer != nullptr)
            return static_cast<char*>(pa.devicePointer);
        cudaGetLastError();
        return nullptr;
    };
    char* const m_obs = mapped(obs);
    char* const m_rew = mapped(reward);
    char* const m_flg = mapped(flags);
    const bool pack = getenv("AQUA_HOST_NO_PACK") == nullptr && (obs == nullptr || m_obs != nullptr) &&
                      (reward == nullptr || m_rew != nullptr) && flags != nullptr && m_flg != nullptr;
    if (pack && getenv("AQUA_HOST_CHUNKS") == nullptr) {  // a few chunks on alternating streams: a chunk's writes overlap the next one's step
        // (measured, profiles/r03v_e2e_*.jsonl: four chunks at 2^20 envs, chunks of 2^20 at 2^24)
        chunk = std::min<int64_t>(std::max<int64_t>((n + 3) / 4, 256 * 1024), 1024 * 1024);
        chunk = (chunk + AQUA_TILE - 1) / AQUA_TILE * AQUA_TILE;
    }
    int c = 0;
    for (int64_t lo = 0; lo < n && err == cudaSuccess; lo += chunk, ++c) {
        const int64_t hi = std::min(n, lo + chunk);
        const int64_t len = hi - lo;
        cudaStream_t s = env->hs[c % kHostStreams];
        AQ_TRY(cudaMemcpyAsync(static_cast<char*>(env->d_actions) + (size_t)lo * ab,
                               static_cast<const char*>(actions) + (size_t)lo * ab, (size_t)len * ab,
                               cudaMemcpyHostToDevice, s));
        aqua::StepArgs a = base_args(env);
        a.env_begin = lo;
        a.env_end = hi;
        a.actions = env->d_actions;
        a.reward = env->d_reward;
        a.flags = env->d_flags;
        aqua::LaunchGeom g;
        pick_geom(env, len, false, action_dtype, &g);
        AQ_TRY(aqua::launch_step(env->cfg.precision, action_dtype, a, g, s));
        if (pack) {
            AQ_TRY(aqua::launch_pack_host(env->cfg.precision, env->state, env->ld, env->d_reward, env->d_flags, lo, len, n,
                                          m_obs, m_rew, reinterpret_cast<uint8_t*>(m_flg), sparse_goals ? 1 : 0, s));
            continue;
        }
        if (flags != nullptr)
            AQ_TRY(cudaMemcpyAsync(flags + lo, env->d_flags + lo, (size_t)len, cudaMemcpyDeviceToHost, s));
        if (reward != nullptr)
            AQ_TRY(cudaMemcpyAsync(static_cast<char*>(reward) + (size_t)lo * es,
                                   static_cast<char*>(env->d_reward) + (size_t)lo * es, (size_t)len * es,
                                   cudaMemcpyDeviceToHost, s));
        if (obs != nullptr)  // planes x, y, angle (, goal_x, goal_y) are rows 0.. of both sides: ONE strided copy
            AQ_TRY(cudaMemcpy2DAsync(static_cast<char*>(obs) + (size_t)lo * es, (size_t)n * es,
                                     static_cast<char*>(env->state) + (size_t)lo * es, (size_t)env->ld * es,
                                     (size_t)len * es, (size_t)n_obs_planes, cudaMemcpyDeviceToHost, s));
        if (sparse_goals)
            AQ_TRY(aqua::launch_goal_updates(env->cfg.precision, env->state, env->ld, env->d_flags, lo, len,
                                             obs_dev + (size_t)3 * (size_t)n * es, obs_dev + (size_t)4 * (size_t)n * es, s));
    }
    for (int i = 0; i < kHostStreams; ++i) {
        const cudaError_t e = cudaStreamSynchronize(env->hs[i]);
        if (err == cudaSuccess && e != cudaSuccess) {
            err = e;
            what = "cudaStreamSynchronize(host pipeline)";
        }
    }
#undef AQ_TRY
    if (err != cudaSuccess) {
        env->host_obs_last = nullptr;
        return fail(AQUA_ERR_CUDA, "aqua_step_host: %s failed: %s", what, cudaGetErrorString(err));
    }
    env->host_obs_last = obs;  // its goal planes are current now (NULL: nothing kept)
    return AQUA_OK;
}

int aqua_observe(aqua_env* env, void* stream, float* out, int normalized) {
    if (int rc = check_env(env)) return rc;
    if (out == nullptr) return fail(AQUA_ERR_INVALID, "out is NULL");
    DeviceGuard guard(env->device);
    AQ_CUDA(aqua::launch_observe(env->cfg.precision, env->state, env->ld, env->cfg.n_envs, out, normalized,
                                 static_cast<cudaStream_t>(stream)));
    return AQUA_OK;
}

int aqua_get_state(aqua_env* env, void* stream, double* out) {
    if (int rc = check_env(env)) return rc;
    if (out == nullptr) return fail(AQUA_ERR_INVALID, "out is NULL");
    DeviceGuard guard(env->device);
    AQ_CUDA(aqua::launch_get_state(env->cfg.precision, env->state, env->ld, env->cfg.n_envs, out,
                                   static_cast<cudaStream_t>(stream)));
    return AQUA_OK;
}

int aqua_set_state(aqua_env* env, void* stream, const double* in) {
    if (int rc = check_env(env)) return rc;
    env->host_obs_last = nullptr;
    if (in == nullptr) return fail(AQUA_ERR_INVALID, "in is NULL");
    DeviceGuard guard(env->device);
    AQ_CUDA(aqua::launch_set_state(env->cfg.precision, env->state, env->ld, env->cfg.n_envs, in,
                                   static_cast<cudaStream_t>(stream)));
    return AQUA_OK;
}

int aqua_stats_reduce(aqua_env* env, void* stream, int64_t* out_dev, int clear) {
    if (int rc = check_env(env)) return rc;
    if (out_dev == nullptr) return fail(AQUA_ERR_INVALID, "out_dev is NULL");
    DeviceGuard guard(env->device);
    AQ_CUDA(aqua::launch_stats_reduce(env->d_slots, env->d_forced, out_dev, clear, static_cast<cudaStream_t>(stream)));
    return AQUA_OK;
}

int aqua_get_stats(aqua_env* env, void* stream, int64_t* out_host, int clear) {
    if (int rc = check_env(env)) return rc;
    if (out_host == nullptr) return fail(AQUA_ERR_INVALID, "out_host is NULL");
    DeviceGuard guard(env->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    AQ_CUDA(aqua::launch_stats_reduce(env->d_slots, env->d_forced, env->d_stats_out, clear, s));
    AQ_CUDA(cudaMemcpyAsync(out_host, env->d_stats_out, sizeof(int64_t) * AQUA_N_STATS, cudaMemcpyDeviceToHost, s));
    AQ_CUDA(cudaStreamSynchronize(s));
    return AQUA_OK;
}

int aqua_obstacle_grid(const double* obstacles, int n_obstacles, double* rows_out, uint64_t* cells_out,
                       int* use_grid_out) {
    return aqua_obstacle_grid_ex(obstacles, n_obstacles, rows_out, cells_out, nullptr, use_grid_out);
}

int aqua_obstacle_grid_ex(const double* obstacles, int n_obstacles, double* rows_out, uint64_t* cells_out,
                          uint16_t* cells16_out, int* use_grid_out) {
    if (n_obstacles < 0 || n_obstacles > AQUA_MAX_OBSTACLES)
        return fail(AQUA_ERR_INVALID, "n_obstacles must be in [0, %d], got %d", AQUA_MAX_OBSTACLES, n_obstacles);
    if ((n_obstacles > 0 && (obstacles == nullptr || rows_out == nullptr)) || cells_out == nullptr || use_grid_out == nullptr)
        return fail(AQUA_ERR_INVALID, "obstacles, rows_out, cells_out and use_grid_out must not be NULL");
    void* h_tab = malloc(aqua::obstacle_table_bytes(AQUA_F64));
    if (h_tab == nullptr) return fail(AQUA_ERR_ALLOC, "out of host memory");
    double h_tab64[AQUA_MAX_OBSTACLES * 5] = {0};
    aqua::build_obstacle_tables(AQUA_F64, n_obstacles, obstacles, h_tab, h_tab64, use_grid_out);
    aqua::copy_table_cells(AQUA_F64, h_tab, reinterpret_cast<unsigned long long*>(cells_out));
    if (cells16_out != nullptr) aqua::copy_table_cells16(AQUA_F64, h_tab, cells16_out);
    free(h_tab);
    if (n_obstacles > 0) memcpy(rows_out, h_tab64, sizeof(double) * 5 * (size_t)n_obstacles);
    return AQUA_OK;
}

}  // extern "C"
