"""Batched AquaGym environment on one GPU: the host side of ``libaquagym.so``.

``AquaVecEnv`` keeps the reference's ``reset()`` / ``step(action)`` contract
(``gym_aqua/envs/aqua.py:100-126, 135-213``) for ``num_envs`` independent instances whose state
lives in GPU memory as structure-of-arrays planes.  Everything that touches env state runs in the
hand-written CUDA kernels behind the C ABI (``include/aquagym.h``); PyTorch only owns the device
buffers and the stream.  There is no CPU implementation here.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N

# default obstacle layout, aqua.py:59-66 ("example_with_obs"); rows {cx, cy, type, a, b}
DEFAULT_OBSTACLES = np.array([
    [15.0, 75.0, 0.0, 5.0, 0.0],
    [20.0, 35.0, 0.0, 10.0, 0.0],
    [65.0, 85.0, 1.0, 5.0, 5.0],
    [85.0, 20.0, 0.0, 10.0, 0.0],
    [85.0, 75.0, 0.0, 5.0, 0.0]], dtype=np.float64)

# "difficult" layout of the v2 ids, gym_aqua/__init__.py:24-31
DIFFICULT_OBSTACLES = np.array([
    [15.0, 70.0, 0.0, 5.0, 0.0],
    [25.0, 40.0, 1.0, 10.0, 10.0],
    [40.0, 80.0, 1.0, 10.0, 10.0],
    [55.0, 20.0, 0.0, 10.0, 0.0],
    [60.0, 55.0, 1.0, 20.0, 20.0],
    [85.0, 75.0, 0.0, 5.0, 0.0]], dtype=np.float64)


def obstacle_table(obstacles):
    """``obstacles`` as the reference constructor takes it (aqua.py:13,56-68): ``False`` -> none,
    ``True`` -> the default layout, a list of ``(pos, 'c'|'r', radius|(w, h))`` tuples, or an array
    ``[K, 5]`` of rows ``{cx, cy, type (0 circle / 1 rectangle), a, b}`` -> float64 ``[K, 5]``."""
    if isinstance(obstacles, np.ndarray):
        tab = np.ascontiguousarray(obstacles, dtype=np.float64).reshape(-1, 5)
    elif isinstance(obstacles, (list, tuple)):
        rows = []
        for pos, kind, dims in obstacles:
            if kind == "c":
                rows.append([float(pos[0]), float(pos[1]), 0.0, float(dims), 0.0])
            else:  # any non-'c' type is a rectangle in the reference's collision code (aqua.py:434-435)
                rows.append([float(pos[0]), float(pos[1]), 1.0, float(dims[0]), float(dims[1])])
        tab = np.array(rows, dtype=np.float64).reshape(-1, 5)
    elif obstacles:
        tab = DEFAULT_OBSTACLES.copy()
    else:
        tab = np.zeros((0, 5), dtype=np.float64)
    if tab.shape[0] > N.MAX_OBSTACLES:
        raise ValueError("at most %d obstacles, got %d" % (N.MAX_OBSTACLES, tab.shape[0]))
    return tab


def table_to_reference_list(tab):
    """Inverse of :func:`obstacle_table` (the ``env.obstacles`` attribute, aqua.py:56-68)."""
    out = []
    for cx, cy, typ, a, b in np.asarray(tab, dtype=np.float64).reshape(-1, 5):
        out.append((np.array([cx, cy]), "c", a) if typ == 0.0 else (np.array([cx, cy]), "r", (a, b)))
    return out


class LazyInfo(dict):
    """``info`` of a batched step: the reference's three keys (aqua.py:194-198) as bool tensors,
    materialised from the flags byte on first access.  With ``keep_terminal`` it also carries
    ``terminal_observation`` ``[N, 5]``: for the envs that finished in this step, the observation the
    reference returns together with ``done=True`` (aqua.py:213) -- the ``obs`` of an auto-resetting
    env already shows the next episode (rows of envs that did not finish hold their last terminal
    observation)."""
    _BITS = {"Termination.collided": N.FLAG_COLLIDED, "Termination.time": N.FLAG_TIME,
             "Termination.success": N.FLAG_SUCCESS}

    def __init__(self, flags, terminal=None):
        super().__init__(flags=flags)
        if terminal is not None:
            self["terminal_observation"] = terminal

    def __missing__(self, key):
        bit = self._BITS[key]
        val = (self["flags"] & bit) != 0
        self[key] = val
        return val


_ACT_DTYPES = {torch.int32: N.ACT_I32, torch.int64: N.ACT_I64, torch.uint8: N.ACT_U8}
_NP_ACT_DTYPES = {np.dtype(np.int32): N.ACT_I32, np.dtype(np.int64): N.ACT_I64,
                  np.dtype(np.uint8): N.ACT_U8}


class AquaVecEnv(object):
    """``num_envs`` AquaGym environments advanced together by one kernel launch per step.

    Constructor arguments mirror ``AquaEnv.__init__`` (aqua.py:13) plus the batching knobs:

    * ``continuous``: ``AquaContinuousEnv`` (thrust pairs) instead of ``AquaEnv`` (3 actions);
    * ``seed`` / ``env_id_base``: key and first global env id of the Philox draw streams, so
      that a slab of a sharded job reproduces the same per-env trajectories as a single-GPU job;
    * ``precision``: ``"float32"`` (throughput layout, 65 B per env-step) or ``"float64"``
      (parity-grade pose);
    * ``autoreset``: finished envs are reset inside the step kernel (vector-env semantics);
    * ``geometry``: ``"auto"`` (default), or force the ``"wide"`` / ``"narrow"`` kernel family
      (tests and tuning; results are identical);
    * ``keep_terminal``: keep the observation a finished env returned with ``done=True``
      (``terminal_obs`` ``[N, 5]``, ``info["terminal_observation"]``).
    """

    def __init__(self, num_envs, continuous=False, obstacles=False, waves=True, random_boat=True,
                 random_goal=True, seed=0, env_id_base=0, device=None, precision="float32",
                 autoreset=True, geometry="auto", keep_terminal=False):
        self._h = None
        self.lib = N.load()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        device = torch.device(device)
        if device.type != "cuda":
            raise ValueError("AquaVecEnv runs on a CUDA device only (no CPU fallback), got %r" % (device,))
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        if precision in ("float32", torch.float32, np.float32):
            self.precision, self.dtype = N.F32, torch.float32
        elif precision in ("float64", torch.float64, np.float64):
            self.precision, self.dtype = N.F64, torch.float64
        else:
            raise ValueError("precision must be 'float32' or 'float64'")
        self.device = device
        self.num_envs = int(num_envs)
        self.continuous = bool(continuous)
        self.has_waves = int(bool(waves))
        self.random_boat, self.random_goal = bool(random_boat), bool(random_goal)
        self.autoreset = bool(autoreset)
        self.seed, self.env_id_base = int(seed), int(env_id_base)
        self.table = obstacle_table(obstacles)
        self.obstacles = table_to_reference_list(self.table)
        self.ld = (self.num_envs + N.TILE - 1) // N.TILE * N.TILE

        cfg = N.Config()
        cfg.struct_size = C.sizeof(N.Config)
        cfg.device = device.index
        cfg.n_envs = self.num_envs
        cfg.continuous, cfg.waves = int(self.continuous), self.has_waves
        cfg.random_boat, cfg.random_goal = int(self.random_boat), int(self.random_goal)
        cfg.precision, cfg.autoreset = self.precision, int(self.autoreset)
        cfg.n_obstacles = self.table.shape[0]
        cfg.geometry = {"auto": N.GEOM_AUTO, "wide": N.GEOM_WIDE, "narrow": N.GEOM_NARROW}[geometry]
        tab = np.ascontiguousarray(self.table)
        cfg.obstacles = tab.ctypes.data_as(C.POINTER(C.c_double)) if tab.size else None
        cfg.seed, cfg.env_id_base = self.seed, self.env_id_base
        cfg.keep_terminal = int(bool(keep_terminal))
        self._terminal = None
        if self.num_envs >= 1 and torch.cuda.is_available():
            # caller-owned state: the planes are a torch tensor, observations are views of it
            self._state = torch.zeros(N.N_PLANES * self.ld, dtype=self.dtype, device=device)
            torch.cuda.current_stream(device).synchronize()  # aqua_create initialises it on stream 0
            cfg.state = self._state.data_ptr()
            cfg.ld = self.ld
            if keep_terminal:  # caller-owned as well: the terminal observations are views of it
                self._terminal = torch.zeros(5 * self.ld, dtype=self.dtype, device=device)
                cfg.terminal = self._terminal.data_ptr()
        else:
            self._state = None
        h = C.c_void_p()
        N.check(self.lib.aqua_create(C.byref(h), C.byref(cfg)))
        self._h = h
        planes = self._state.view(N.N_PLANES, self.ld)
        self.planes = planes
        self.obs_planes = planes[:5, :self.num_envs]          # [5, N] zero-copy: x, y, theta, gx, gy
        self.obs = self.obs_planes.t()                         # [N, 5] strided view of the same memory
        self.wave = planes[5:7, :self.num_envs]
        self.time_word = planes[7].view(torch.int32)[:self.ld][:self.num_envs]
        self.reward = torch.zeros(self.ld, dtype=self.dtype, device=device)[:self.num_envs]
        self.flags = torch.zeros(self.ld, dtype=torch.uint8, device=device)[:self.num_envs]
        self._stats_dev = torch.zeros(N.N_STATS, dtype=torch.int64, device=device)
        self.terminal_obs_planes = self.terminal_obs = None
        if keep_terminal:
            self.terminal_obs_planes = self._terminal.view(5, self.ld)[:, :self.num_envs]  # [5, N]
            self.terminal_obs = self.terminal_obs_planes.t()                                # [N, 5]
        self._step_index = 0
        self._host_obs_prev = None
        # constants of the per-step call (the wrapper used to cost more than the launch itself on
        # small slabs: 11 us per call against 5 us through the bare ABI, tools/small_slab_probe.py)
        self._dev_index = torch.device(device).index or 0
        self._reward_ptr, self._flags_ptr = self.reward.data_ptr(), self.flags.data_ptr()
        self._act_numel = self.num_envs * (2 if self.continuous else 1)
        self._raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)

    # ---------------------------------------------------------------- plumbing
    def _stream(self):
        if self._raw_stream is not None:  # the handle itself, without building a torch Stream object
            return self._raw_stream(self._dev_index)
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            h, self._h = self._h, None
            N.check(self.lib.aqua_destroy(h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check_open(self):
        if self._h is None:
            raise RuntimeError("environment is closed")

    # ---------------------------------------------------------------- the reference API
    def reset(self, mask=None):
        """``env.reset()`` (aqua.py:100-126) for all envs, or those where ``mask`` is non-zero.
        Returns the observation view ``[N, 5]``."""
        self._check_open()
        ptr = None
        if mask is not None:
            mask = torch.as_tensor(mask, device=self.device)
            mask = (mask != 0).to(torch.uint8).contiguous()
            if mask.numel() != self.num_envs:
                raise ValueError("mask must have %d entries" % self.num_envs)
            ptr = C.c_void_p(mask.data_ptr())
        N.check(self.lib.aqua_reset(self._h, self._stream(), ptr))
        return self.obs

    def _action_arg(self, actions):
        actions = torch.as_tensor(actions, device=self.device)
        if self.continuous:
            actions = actions.to(torch.float32).contiguous()
            if actions.numel() != 2 * self.num_envs:
                raise ValueError("continuous actions must have shape [%d, 2]" % self.num_envs)
            return actions, N.ACT_F32X2
        if actions.dtype not in _ACT_DTYPES:
            if actions.dtype.is_floating_point or actions.dtype == torch.bool:
                raise TypeError("discrete actions must be an integer tensor, got %s" % actions.dtype)
            actions = actions.to(torch.int64)
        actions = actions.contiguous()
        if actions.numel() != self.num_envs:
            raise ValueError("discrete actions must have %d entries" % self.num_envs)
        return actions, _ACT_DTYPES[actions.dtype]

    def step(self, actions, noise=None):
        """``env.step(action)`` (aqua.py:135-213) for every env in one launch.

        ``actions``: integer tensor ``[N]`` (discrete) or float ``[N, 2]`` (continuous) on the env's
        device.  ``noise`` (parity/debug): float32 ``[2, N]`` wave increments replacing the Philox
        draw of aqua.py:188.  Returns ``(obs [N,5] view, reward [N], done [N] bool, info)``; with
        ``autoreset`` the observation of a finished env is the first one of its next episode.
        The returned tensors are reused by the next call."""
        flags = self.step_flags(actions, noise)
        return self.obs, self.reward, flags != 0, LazyInfo(flags, self.terminal_obs)

    def step_flags(self, actions, noise=None):
        """The launch only: updates ``obs``/``reward``/``flags`` in place and returns ``flags``
        (uint8 ``AQUA_FLAG_*`` bits) without any extra torch op."""
        if self._h is None:
            raise RuntimeError("environment is closed")
        # fast path: a tensor that already is what the kernel reads (the common case in a loop)
        adt = None
        if type(actions) is torch.Tensor and actions.is_cuda and actions.numel() == self._act_numel \
                and actions.is_contiguous() and actions.device.index == self._dev_index:
            adt = N.ACT_F32X2 if (self.continuous and actions.dtype == torch.float32) else \
                (None if self.continuous else _ACT_DTYPES.get(actions.dtype))
        if adt is None:
            actions, adt = self._action_arg(actions)
        nz = None
        if noise is not None:
            noise = torch.as_tensor(noise, device=self.device).to(torch.float32).contiguous()
            if noise.shape != (2, self.num_envs):
                raise ValueError("noise must have shape [2, %d]" % self.num_envs)
            nz = noise.data_ptr()
        rc = self.lib.aqua_step_ex(self._h, self._stream(), actions.data_ptr(), adt, self._reward_ptr,
                                   self._flags_ptr, nz)
        if rc:
            N.check(rc)
        self._step_index += 1
        return self.flags

    def step_host(self, actions, obs_out=None, reward_out=None, flags_out=None):
        """Host-buffer step (``aqua_step_host``): NumPy in, NumPy out, copies pipelined inside.
        ``obs_out`` is ``[5, N]`` (planes x, y, theta, goal_x, goal_y).  Passing the SAME ``obs_out``
        array call after call (and leaving it untouched in between) lets the library send only the
        goals that changed: 17 instead of 25 bytes per env-step come back."""
        self._check_open()
        np_real = np.float64 if self.precision == N.F64 else np.float32
        if self.continuous:
            actions = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.num_envs, 2)
            adt = N.ACT_F32X2
        else:
            actions = np.ascontiguousarray(actions)
            if actions.dtype not in _NP_ACT_DTYPES:
                actions = actions.astype(np.int64)
            adt = _NP_ACT_DTYPES[actions.dtype]
            if actions.size != self.num_envs:
                raise ValueError("discrete actions must have %d entries" % self.num_envs)
        if obs_out is None:
            obs_out = np.empty((5, self.num_envs), dtype=np_real)
        if reward_out is None:
            reward_out = np.empty(self.num_envs, dtype=np_real)
        if flags_out is None:
            flags_out = np.empty(self.num_envs, dtype=np.uint8)
        for name, a, dt in (("obs_out", obs_out, np_real), ("reward_out", reward_out, np_real),
                            ("flags_out", flags_out, np.uint8)):
            if a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("%s must be C-contiguous %s" % (name, np.dtype(dt)))
        # the array object of the previous call (a reference is kept, so its memory cannot have been
        # recycled): the library may update its goal planes incrementally
        persistent = int(obs_out is self._host_obs_prev)
        N.check(self.lib.aqua_step_host_ex(self._h, self._stream(), C.c_void_p(actions.ctypes.data), adt,
                                           C.c_void_p(obs_out.ctypes.data),
                                           C.c_void_p(reward_out.ctypes.data),
                                           C.c_void_p(flags_out.ctypes.data), persistent))
        self._host_obs_prev = obs_out
        self._step_index += 1
        return obs_out, reward_out, flags_out

    def step_host_raw(self, actions_ptr, action_dtype, obs_ptr, reward_ptr, flags_ptr):
        """``aqua_step_host`` on raw host addresses (``ctypes.c_void_p``) the caller keeps alive and has
        validated once: what the single-env drop-in calls every step (no per-call conversions)."""
        N.check(self.lib.aqua_step_host_ex(self._h, self._stream(), actions_ptr, action_dtype, obs_ptr, reward_ptr,
                                           flags_ptr, 0))
        self._host_obs_prev = None
        self._step_index += 1

    def host_d2h_bytes_per_env(self):
        """Bytes per env-step that :meth:`step_host` copies device -> host when all three outputs are
        requested (obs planes that changed, reward, flags)."""
        es = 8 if self.precision == N.F64 else 4
        return 3 * es + es + 1  # x, y, angle + reward + flags; plus a list entry per finished env (~1.5 %)

    def render(self, mode="human"):
        """aqua.py:215-365 draws a pyglet window; the batched env has no viewer (out of scope)."""
        return None

    # ---------------------------------------------------------------- synthetic policy / rollouts
    def random_actions(self, step_index=None, out=None):
        """``action_space.sample()`` for every env (test_random.py:19-20) from the Philox ACTION
        stream keyed by (global env id, step index)."""
        self._check_open()
        if step_index is None:
            step_index = self._step_index
        if out is None:
            out = (torch.empty(self.num_envs, 2, dtype=torch.float32, device=self.device) if self.continuous
                   else torch.empty(self.num_envs, dtype=torch.int32, device=self.device))
        adt = N.ACT_F32X2 if self.continuous else _ACT_DTYPES[out.dtype]
        N.check(self.lib.aqua_random_actions(self._h, self._stream(), int(step_index),
                                             C.c_void_p(out.data_ptr()), adt))
        return out

    def step_many(self, actions, reward_out=None, flags_out=None):
        """``K`` consecutive ``env.step`` calls in ONE launch (``aqua_step_many``): ``actions`` is the
        tape ``[K, N]`` (discrete, integer dtype) or ``[K, N, 2]`` (continuous float32) on the env's
        device.  Returns ``(reward [K, N], flags [K, N] uint8)``; bit for bit what ``K`` calls of
        :meth:`step` give.  The state planes (``obs``) hold the observation after the last step.
        Rows must be 16 envs aligned: for ``N % 16 != 0`` pass (and get) tensors ``[K, N16]`` with
        ``N16 = N`` rounded up to 16 and use the first ``N`` columns."""
        self._check_open()
        if not (type(actions) is torch.Tensor and actions.is_cuda and actions.is_contiguous()):
            raise TypeError("step_many takes a contiguous CUDA tensor [K, N] (or [K, N, 2])")
        if self.continuous:
            if actions.dtype != torch.float32 or actions.dim() != 3 or actions.shape[2] != 2:
                raise ValueError("continuous tape must be float32 [K, N, 2]")
            adt = N.ACT_F32X2
        else:
            if actions.dim() != 2 or actions.dtype not in _ACT_DTYPES:
                raise ValueError("discrete tape must be int32 / int64 / uint8 [K, N]")
            adt = _ACT_DTYPES[actions.dtype]
        K, stride = int(actions.shape[0]), int(actions.shape[1])
        if stride < self.num_envs or stride % 16:
            raise ValueError("tape rows must have >= %d columns, a multiple of 16; got %d" % (self.num_envs, stride))
        if reward_out is None:
            reward_out = torch.empty(K, stride, dtype=self.dtype, device=self.device)
        if flags_out is None:
            flags_out = torch.empty(K, stride, dtype=torch.uint8, device=self.device)
        for name, t, dt in (("reward_out", reward_out, self.dtype), ("flags_out", flags_out, torch.uint8)):
            if t is not False and (t.shape != (K, stride) or t.dtype != dt or not t.is_contiguous()):
                raise ValueError("%s must be a contiguous %s [%d, %d] tensor" % (name, dt, K, stride))
        N.check(self.lib.aqua_step_many(
            self._h, self._stream(), K, C.c_void_p(actions.data_ptr()), adt,
            None if reward_out is False else C.c_void_p(reward_out.data_ptr()),
            None if flags_out is False else C.c_void_p(flags_out.data_ptr()), stride))
        self._step_index += K
        return reward_out, flags_out

    def rollout_random(self, steps, first_step_index=None):
        """``steps`` env-steps in one launch under the synthetic random policy (always auto-resets);
        only the state and the statistics are updated."""
        self._check_open()
        if first_step_index is None:
            first_step_index = self._step_index
        N.check(self.lib.aqua_rollout_random(self._h, self._stream(), int(steps), int(first_step_index)))
        self._step_index = int(first_step_index) + int(steps)

    @staticmethod
    def pack_q_network(q_function):
        """The weights of the reference's Q-network (dqn.py:301-314; a torch Sequential with three Linear
        layers 5-64-64-3, e.g. ``savedmodel.load_q_network`` or ``dqn.create_model``) as the flat float32
        vector ``aqua_rollout_policy`` takes: Keras layout W1[5][64], b1, W2[64][64], b2, W3[64][3], b3."""
        linears = [m for m in q_function if isinstance(m, torch.nn.Linear)]
        shapes = [tuple(m.weight.shape) for m in linears]
        if shapes != [(64, 5), (64, 64), (3, 64)]:
            raise ValueError("rollout_policy evaluates the reference's 5-64-64-3 network, got layers %r" % (shapes,))
        parts = []
        for m in linears:
            parts += [m.weight.detach().t().contiguous().reshape(-1), m.bias.detach().reshape(-1)]  # torch keeps [out, in]
        w = torch.cat(parts).to(torch.float32).contiguous()
        assert w.numel() == N.MLP_WEIGHTS
        return w

    def rollout_policy(self, weights, steps, epsilon=0.0, first_step_index=None, want_actions=False, want_reward=False,
                       want_flags=False):
        """``steps`` env-steps in ONE launch with the actions chosen in the kernel by the Q-network
        ``weights`` (:meth:`pack_q_network`), greedy or epsilon-greedy (``aqua_rollout_policy``).
        Returns ``(actions, reward, flags)``: ``[steps, row_stride]`` tensors (columns >= num_envs are
        padding) for the outputs asked for, ``None`` otherwise; the statistics count everything."""
        self._check_open()
        if first_step_index is None:
            first_step_index = self._step_index
        w = weights.to(self.device, torch.float32).contiguous()
        if w.numel() != N.MLP_WEIGHTS:
            raise ValueError("weights must hold %d floats (pack_q_network)" % N.MLP_WEIGHTS)
        stride = (self.num_envs + 15) // 16 * 16
        acts = torch.empty(steps, stride, dtype=torch.int32, device=self.device) if want_actions else None
        rew = torch.empty(steps, stride, dtype=self.dtype, device=self.device) if want_reward else None
        flg = torch.empty(steps, stride, dtype=torch.uint8, device=self.device) if want_flags else None
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        N.check(self.lib.aqua_rollout_policy(self._h, self._stream(), int(steps), ptr(w), float(epsilon),
                                             int(first_step_index), ptr(acts), ptr(rew), ptr(flg), stride))
        self._step_index = int(first_step_index) + int(steps)
        return acts, rew, flg

    def rollout_policy_dev(self, weights, steps, epsilon_dev, step_index_dev, actions_out=None, reward_out=None,
                           flags_out=None):
        """:meth:`rollout_policy` with the exploration rate (float32 tensor of one element) and the
        first step index (int64 tensor of one element, >= 0) read on the device when the launch
        runs, into caller-owned ``[steps, stride]`` outputs (``stride`` = num_envs rounded up to 16):
        nothing here depends on a host value, so the call can be captured in a CUDA graph
        (``aqua_rollout_policy_dev``).  The caller advances ``step_index_dev``."""
        self._check_open()
        stride = (self.num_envs + 15) // 16 * 16
        for t, dt in ((actions_out, torch.int32), (reward_out, self.dtype), (flags_out, torch.uint8)):
            if t is not None and (t.dtype != dt or tuple(t.shape) != (steps, stride) or not t.is_contiguous()):
                raise ValueError("outputs must be contiguous [%d, %d] tensors of int32 / %s / uint8" % (steps, stride, self.dtype))
        if weights.numel() != N.MLP_WEIGHTS or weights.dtype != torch.float32 or not weights.is_contiguous():
            raise ValueError("weights: a contiguous float32 tensor of %d elements" % N.MLP_WEIGHTS)
        if epsilon_dev.dtype != torch.float32 or step_index_dev.dtype != torch.int64:
            raise ValueError("epsilon_dev: float32, step_index_dev: int64 (one element each, on the device)")
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        N.check(self.lib.aqua_rollout_policy_dev(self._h, self._stream(), int(steps), ptr(weights), ptr(epsilon_dev),
                                                 ptr(step_index_dev), ptr(actions_out), ptr(reward_out), ptr(flags_out),
                                                 stride))

    def observe(self, normalized=False, out=None):
        """Observation rows ``[N, 5]`` float32, contiguous (network input).  ``normalized`` applies
        ``AquaStateNormalizer.__call__`` (main/impl/utils.py:23-27) in the same kernel."""
        self._check_open()
        if out is None:
            out = torch.empty(self.num_envs, 5, dtype=torch.float32, device=self.device)
        elif out.shape != (self.num_envs, 5) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 [%d, 5] tensor" % self.num_envs)
        N.check(self.lib.aqua_observe(self._h, self._stream(), C.c_void_p(out.data_ptr()),
                                      int(bool(normalized))))
        return out

    # ---------------------------------------------------------------- state / statistics
    def get_state(self):
        """Canonical float64 ``[9, N]`` device tensor: x, y, theta, wave_x, wave_y, goal_x, goal_y,
        time, episode."""
        self._check_open()
        out = torch.empty(9, self.num_envs, dtype=torch.float64, device=self.device)
        N.check(self.lib.aqua_get_state(self._h, self._stream(), C.c_void_p(out.data_ptr())))
        return out

    def set_state(self, state):
        self._check_open()
        state = torch.as_tensor(state, dtype=torch.float64, device=self.device).contiguous()
        if state.shape != (9, self.num_envs):
            raise ValueError("state must have shape [9, %d]" % self.num_envs)
        N.check(self.lib.aqua_set_state(self._h, self._stream(), C.c_void_p(state.data_ptr())))

    def stats_tensor(self, clear=False):
        """Device int64 ``[8]`` counters (see ``AQUA_STAT_*``), asynchronous on the current stream:
        the vector that is all-reduced across ranks."""
        self._check_open()
        N.check(self.lib.aqua_stats_reduce(self._h, self._stream(),
                                           C.c_void_p(self._stats_dev.data_ptr()), int(bool(clear))))
        return self._stats_dev

    def stats(self, clear=False):
        """Host dict of the counters (synchronises the current stream)."""
        self._check_open()
        out = (C.c_int64 * N.N_STATS)()
        N.check(self.lib.aqua_get_stats(self._h, self._stream(), out, int(bool(clear))))
        return stats_dict(list(out))


def stats_dict(values):
    d = {k: int(v) for k, v in zip(N.STAT_NAMES, values)}
    d["reward_sum"] = d["reward_sum_fx"] / N.REWARD_FIXED_POINT
    ep = max(d["episodes"], 1)
    d["success_rate"] = d["success"] / ep
    d["mean_episode_length"] = d["length_sum"] / ep
    d["mean_episode_return"] = d["reward_sum"] / ep
    return d
