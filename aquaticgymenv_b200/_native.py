"""ctypes binding of ``libaquagym.so`` (C ABI: ``include/aquagym.h``).

The shared library is built in-tree by ``aquaticgymenv_b200.build.build()`` (nvcc, sm_100a).
There is no CPU fallback: if the library is missing ``load()`` raises, and without a CUDA device
``aqua_create`` returns ``AQUA_ERR_NO_DEVICE`` which surfaces as :class:`AquaError`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AQUAGYM_LIB selects another build of the same library (tuning variants, tools/tune.py)
LIB_PATH = os.environ.get("AQUAGYM_LIB") or os.path.join(_HERE, "libaquagym.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "aquagym.h")

ABI_VERSION = 3
OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_ALLOC = 0, -1, -2, -3, -4
F32, F64 = 0, 1
GEOM_AUTO, GEOM_WIDE, GEOM_NARROW = 0, 1, 2
ACT_I32, ACT_I64, ACT_U8, ACT_F32X2 = 0, 1, 2, 3
FLAG_DONE, FLAG_COLLIDED, FLAG_TIME, FLAG_SUCCESS = 1, 2, 4, 8
MAX_OBSTACLES = 64
TIME_BITS = 11
EPISODE_MASK = 0x1FFFFF
N_PLANES = 8
TILE = 1024
N_STATS = 9
MLP_WEIGHTS = 4739  # AQUA_MLP_WEIGHTS
STAT_NAMES = ("episodes", "collided", "timeout", "success", "length_sum", "reward_sum_fx", "steps",
              "bad_actions", "forced_resets")
REWARD_FIXED_POINT = 1048576.0


class AquaError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("libaquagym: %s (status %d)" % (message, status))
        self.status = status


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("device", C.c_int32), ("n_envs", C.c_int64),
        ("continuous", C.c_int32), ("waves", C.c_int32), ("random_boat", C.c_int32),
        ("random_goal", C.c_int32), ("precision", C.c_int32), ("autoreset", C.c_int32),
        ("n_obstacles", C.c_int32), ("geometry", C.c_int32),
        ("obstacles", C.POINTER(C.c_double)), ("seed", C.c_uint64), ("env_id_base", C.c_uint64),
        ("state", C.c_void_p), ("ld", C.c_int64), ("keep_terminal", C.c_int32), ("_reserved", C.c_int32),
        ("terminal", C.c_void_p)]


class StateView(C.Structure):
    _fields_ = [("planes", C.c_void_p * N_PLANES), ("ld", C.c_int64), ("n_envs", C.c_int64),
                ("precision", C.c_int32), ("_pad", C.c_int32), ("terminal", C.c_void_p * 5)]


_vp, _u64, _i64, _int = C.c_void_p, C.c_uint64, C.c_int64, C.c_int
# name -> (restype, argtypes); must list every symbol include/aquagym.h declares
SIGNATURES = {
    "aqua_abi_version": (_int, []),
    "aqua_last_error": (C.c_char_p, []),
    "aqua_create": (_int, [C.POINTER(_vp), C.POINTER(Config)]),
    "aqua_destroy": (_int, [_vp]),
    "aqua_state": (_int, [_vp, C.POINTER(StateView)]),
    "aqua_reset": (_int, [_vp, _vp, _vp]),
    "aqua_step": (_int, [_vp, _vp, _vp, _int, _vp, _vp]),
    "aqua_step_ex": (_int, [_vp, _vp, _vp, _int, _vp, _vp, _vp]),
    "aqua_random_actions": (_int, [_vp, _vp, _u64, _vp, _int]),
    "aqua_step_many": (_int, [_vp, _vp, _i64, _vp, _int, _vp, _vp, _i64]),
    "aqua_rollout_random": (_int, [_vp, _vp, _i64, _u64]),
    "aqua_rollout_policy": (_int, [_vp, _vp, _i64, _vp, C.c_float, _u64, _vp, _vp, _vp, _i64]),
    "aqua_rollout_policy_dev": (_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64]),
    "aqua_step_host": (_int, [_vp, _vp, _vp, _int, _vp, _vp, _vp]),
    "aqua_step_host_ex": (_int, [_vp, _vp, _vp, _int, _vp, _vp, _vp, _int]),
    "aqua_observe": (_int, [_vp, _vp, _vp, _int]),
    "aqua_get_state": (_int, [_vp, _vp, _vp]),
    "aqua_set_state": (_int, [_vp, _vp, _vp]),
    "aqua_stats_reduce": (_int, [_vp, _vp, _vp, _int]),
    "aqua_get_stats": (_int, [_vp, _vp, C.POINTER(C.c_int64), _int]),
    "aqua_obstacle_grid": (_int, [_vp, _int, _vp, _vp, C.POINTER(_int)]),
    "aqua_obstacle_grid_ex": (_int, [_vp, _int, _vp, _vp, _vp, C.POINTER(_int)]),
}
GRID_DIM = 96                      # AQUA_GRID_DIM
CELL_BLOCKED = 0xFFFFFFFFFFFFFFFF  # AQUA_CELL_BLOCKED
CELL16_NONE, CELL16_BLOCKED, CELL16_OVERFLOW = 64, 0x4000, 0x8000

_lib = None


def load():
    """Load libaquagym.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C aquaticgymenv_b200/csrc` (nvcc, sm_100a). aquaticgymenv_b200 has no CPU "
            "fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lenient = bool(os.environ.get("AQUAGYM_LIB_ANY_ABI"))  # tools/many_ab.py: time an older build
    for name, (res, args) in SIGNATURES.items():
        if lenient and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    got = lib.aqua_abi_version()
    if got != ABI_VERSION and not lenient:
        raise ImportError("libaquagym ABI %d, binding expects %d: rebuild" % (got, ABI_VERSION))
    _lib = lib
    return lib


def check(status):
    if status != OK:
        raise AquaError(status, load().aqua_last_error().decode("utf-8", "replace"))


def obstacle_grid(obstacles):
    """aqua_obstacle_grid: the culling grid the kernels use for an obstacle table [K, 5]
    (host-only, no CUDA device needed).  Returns (rows [K, 5] = {cx, cy, half_x, half_y, reach^2} in
    the kernels' order, cells uint64 [GRID_DIM, GRID_DIM] indexed [cy, cx], use_grid)."""
    import numpy as np
    tab = np.ascontiguousarray(np.asarray(obstacles, np.float64).reshape(-1, 5))
    k = tab.shape[0]
    rows = np.zeros((max(k, 1), 5), np.float64)
    cells = np.zeros(GRID_DIM * GRID_DIM, np.uint64)
    use_grid = C.c_int(0)
    check(load().aqua_obstacle_grid(tab.ctypes.data if k else None, k, rows.ctypes.data, cells.ctypes.data,
                                    C.byref(use_grid)))
    return rows[:k], cells.reshape(GRID_DIM, GRID_DIM), bool(use_grid.value)


def obstacle_grid16(obstacles):
    """aqua_obstacle_grid_ex: as obstacle_grid plus the 16-bit cell words the kernels read
    (uint16 [GRID_DIM, GRID_DIM]).  Returns (rows, cells, cells16, use_grid)."""
    import numpy as np
    tab = np.ascontiguousarray(np.asarray(obstacles, np.float64).reshape(-1, 5))
    k = tab.shape[0]
    rows = np.zeros((max(k, 1), 5), np.float64)
    cells = np.zeros(GRID_DIM * GRID_DIM, np.uint64)
    cells16 = np.zeros(GRID_DIM * GRID_DIM, np.uint16)
    use_grid = C.c_int(0)
    check(load().aqua_obstacle_grid_ex(tab.ctypes.data if k else None, k, rows.ctypes.data, cells.ctypes.data,
                                       cells16.ctypes.data, C.byref(use_grid)))
    return rows[:k], cells.reshape(GRID_DIM, GRID_DIM), cells16.reshape(GRID_DIM, GRID_DIM), bool(use_grid.value)

