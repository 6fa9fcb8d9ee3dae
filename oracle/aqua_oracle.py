"""ctypes front of the C oracle (`libaqua_oracle.so`, built by `make -C oracle`).

TEST INFRASTRUCTURE ONLY: imported by tests/, `__graft_entry__.smoke()` and
bench.py's cpu_baseline / `--impl reference` legs, never by the product package
`aquaticgymenv_b200`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libaqua_oracle.so")

MAX_OBSTACLES = 64
FLAG_DONE, FLAG_COLLIDED, FLAG_TIME, FLAG_SUCCESS = 1, 2, 4, 8
EPISODE_NONE = 0x1FFFFF


class _World(C.Structure):
    _fields_ = [("continuous", C.c_int32), ("waves", C.c_int32), ("random_boat", C.c_int32),
                ("random_goal", C.c_int32), ("n_obstacles", C.c_int32), ("key_winit", C.c_uint32),
                ("key_noise", C.c_uint32), ("key_action", C.c_uint32), ("seed", C.c_uint64), ("env_id_base", C.c_uint64),
                ("obstacles", C.c_double * (MAX_OBSTACLES * 5))]


class _State(C.Structure):
    _fields_ = [(k, C.POINTER(C.c_double)) for k in ("x", "y", "theta", "wx", "wy", "gx", "gy")] + \
               [("time", C.POINTER(C.c_int32)), ("episode", C.POINTER(C.c_uint32))]


def build(force=False):
    """Compile the oracle with plain gcc (idempotent)."""
    src = os.path.join(_HERE, "aqua_oracle.c")
    hdr = os.path.join(_HERE, "aqua_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libaqua_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        dp, u8p, i32p, f32p = (C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int32),
                               C.POINTER(C.c_float))
        L.aqo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.aqo_philox2x32_10.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32)]
        L.aqo_key32.restype = C.c_uint32
        L.aqo_key32.argtypes = [C.c_uint64, C.c_uint32]
        L.aqo_world_init_keys.argtypes = [C.POINTER(_World)]
        L.aqo_u01.restype = C.c_float
        L.aqo_u01.argtypes = [C.c_uint32]
        L.aqo_draw_reset_sample.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, dp]
        L.aqo_draw_wave_init.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, dp]
        L.aqo_draw_wave_noise.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, dp]
        L.aqo_draw_action_discrete.restype = C.c_uint32
        L.aqo_draw_action_discrete.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64]
        L.aqo_draw_action_continuous.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, f32p]
        L.aqo_normalize_angle.restype = C.c_double
        L.aqo_normalize_angle.argtypes = [C.c_double]
        L.aqo_distance_circle_circle.restype = C.c_double
        L.aqo_distance_circle_circle.argtypes = [C.c_double] * 6
        L.aqo_distance_rectangle_circle.restype = C.c_double
        L.aqo_distance_rectangle_circle.argtypes = [C.c_double] * 7
        L.aqo_collided_border.argtypes = [C.c_double] * 3
        L.aqo_collided_obstacle.argtypes = [C.POINTER(_World), C.c_double, C.c_double, C.c_double]
        L.aqo_reset_env.argtypes = [C.POINTER(_World), C.POINTER(_State), C.c_int64]
        L.aqo_reset_batch.argtypes = [C.POINTER(_World), C.POINTER(_State), C.c_int64, u8p]
        L.aqo_step_batch.argtypes = [C.POINTER(_World), C.POINTER(_State), C.c_int64, i32p, f32p,
                                     dp, dp, u8p, dp, C.c_int, C.c_int]
        L.aqo_rollout_random.restype = C.c_int64
        L.aqo_rollout_random.argtypes = [C.POINTER(_World), C.POINTER(_State), C.c_int64,
                                         C.c_int64, C.c_uint64, C.c_int, dp]
        _lib = L
    return _lib


def obstacle_table(obstacles):
    """Reference obstacle list [(pos, 'c'|'r', radius|(w, h)), ...] (aqua.py:56-68) or an
    array [K, 5] -> float64 [K, 5] rows {cx, cy, type, a, b}."""
    if isinstance(obstacles, np.ndarray):
        return np.ascontiguousarray(obstacles, dtype=np.float64).reshape(-1, 5)
    rows = []
    for pos, kind, dims in obstacles:
        if kind == 'c':
            rows.append([float(pos[0]), float(pos[1]), 0.0, float(dims), 0.0])
        else:  # the reference treats every non-'c' type as a rectangle (aqua.py:434-435)
            rows.append([float(pos[0]), float(pos[1]), 1.0, float(dims[0]), float(dims[1])])
    return np.array(rows, dtype=np.float64).reshape(-1, 5)


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


class OracleVecEnv(object):
    """A batch of independent reference environments stepped by the C oracle."""

    def __init__(self, n, continuous=False, obstacles=(), waves=True, random_boat=True,
                 random_goal=True, seed=0, env_id_base=0):
        self.n = int(n)
        self.L = lib()
        tab = obstacle_table(list(obstacles) if not isinstance(obstacles, np.ndarray) else obstacles)
        assert tab.shape[0] <= MAX_OBSTACLES
        self.table = tab
        w = _World()
        w.continuous, w.waves = int(continuous), int(waves)
        w.random_boat, w.random_goal = int(random_boat), int(random_goal)
        w.n_obstacles = tab.shape[0]
        w.seed, w.env_id_base = int(seed), int(env_id_base)
        self.L.aqo_world_init_keys(C.byref(w))
        for i, v in enumerate(tab.reshape(-1)):
            w.obstacles[i] = v
        self.world = w
        self.continuous = bool(continuous)
        z = lambda: np.zeros(self.n, dtype=np.float64)
        self.x, self.y, self.theta, self.wx, self.wy, self.gx, self.gy = (z() for _ in range(7))
        self.time = np.zeros(self.n, dtype=np.int32)
        self.episode = np.full(self.n, EPISODE_NONE, dtype=np.uint32)
        self._bind()

    def _bind(self):
        s = _State()
        for k in ("x", "y", "theta", "wx", "wy", "gx", "gy"):
            setattr(s, k, _ptr(getattr(self, k), C.c_double))
        s.time = _ptr(self.time, C.c_int32)
        s.episode = _ptr(self.episode, C.c_uint32)
        self.state = s

    # -- state access (canonical [9, n] float64: x y theta wx wy gx gy time episode) --
    def get_state(self):
        return np.stack([self.x, self.y, self.theta, self.wx, self.wy, self.gx, self.gy,
                         self.time.astype(np.float64), self.episode.astype(np.float64)])

    def set_state(self, st):
        st = np.asarray(st, dtype=np.float64)
        for k, name in enumerate(("x", "y", "theta", "wx", "wy", "gx", "gy")):
            getattr(self, name)[:] = st[k]
        self.time[:] = st[7].astype(np.int32)
        self.episode[:] = st[8].astype(np.uint32)

    def obs(self):
        return np.stack([self.x, self.y, self.theta, self.gx, self.gy], axis=1)

    def reset(self, mask=None):
        m = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            m = _ptr(mask, C.c_uint8)
        self.L.aqo_reset_batch(C.byref(self.world), C.byref(self.state), self.n, m)
        return self.obs()

    def step(self, actions, noise=None, autoreset=True, want_margin=False, nthreads=1):
        """-> (obs [n,5] after auto-reset, reward [n], flags [n] uint8, margin [n] | None)."""
        reward = np.empty(self.n, dtype=np.float64)
        flags = np.empty(self.n, dtype=np.uint8)
        margin = np.empty(self.n, dtype=np.float64) if want_margin else None
        a_i32 = a_f32 = None
        if self.continuous:
            a_f32 = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.n, 2)
        else:
            a_i32 = np.ascontiguousarray(actions, dtype=np.int32).reshape(self.n)
        nz = None
        if noise is not None:
            nz = np.ascontiguousarray(noise, dtype=np.float64).reshape(2, self.n)
        self.L.aqo_step_batch(
            C.byref(self.world), C.byref(self.state), self.n,
            None if a_i32 is None else _ptr(a_i32, C.c_int32),
            None if a_f32 is None else _ptr(a_f32, C.c_float),
            None if nz is None else _ptr(nz, C.c_double),
            _ptr(reward, C.c_double), _ptr(flags, C.c_uint8),
            None if margin is None else _ptr(margin, C.c_double), int(autoreset), int(nthreads))
        return self.obs(), reward, flags, margin

    def rollout_random(self, steps, step0=0, nthreads=1):
        """K steps of the synthetic random policy; -> (episodes finished, reward sum)."""
        rs = C.c_double(0.0)
        ep = self.L.aqo_rollout_random(C.byref(self.world), C.byref(self.state), self.n,
                                       int(steps), int(step0), int(nthreads), C.byref(rs))
        return int(ep), float(rs.value)
