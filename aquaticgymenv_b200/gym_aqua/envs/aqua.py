"""Drop-in ``AquaEnv`` / ``AquaContinuousEnv`` (reference: ``gym_aqua/envs/aqua.py``).

The classes keep the reference's constructor, attributes and ``reset()`` / ``step(action)`` /
``render()`` / ``close()`` contract for ONE environment, so that ``main/impl/dqn.py``,
``main/testing/__init__.py`` and the test scripts run unchanged; the dynamics run in the CUDA
kernels of ``libaquagym.so`` (float64 instantiation, no auto-reset: the reference's episodic
flow).  For throughput use ``make(id, num_envs=N)`` / :class:`AquaVecEnv` instead.
"""
import numpy as np

from .. import _base
from ... import spaces
from ... import _native as N
from ...vec_env import AquaVecEnv


class AquaEnv(_base.Env):
    metadata = {'render.modes': ['human']}
    continuous = False

    def __init__(self, obstacles=False, waves=True, random_boat=True, random_goal=True, seed=None,
                 device=None):
        """``seed=None`` (default) draws a fresh 64-bit key for the Philox streams, so that two envs
        -- e.g. the train and test envs of dqn_train_policy.py / test_dqn.py -- see independent
        episodes, like the unseeded reference; pass an int (or call :meth:`seed`) to replay."""
        # aqua.py:15-25
        self.has_waves = int(waves)
        self.random_boat = random_boat
        self.random_goal = random_goal
        self.world_size = 100
        self.motor_min_thrust = +0.2
        self.motor_max_thrust = +0.5
        self.wave_min_speed = -0.05 * self.has_waves
        self.wave_max_speed = +0.05 * self.has_waves
        self.wave_speed_variance = 0.001 * self.has_waves
        # aqua.py:28-53
        if self.continuous:
            self.action_space = spaces.Box(self.motor_min_thrust, self.motor_max_thrust, shape=[2],
                                           dtype=np.float64)
        else:
            self.actions = [(self.motor_min_thrust, self.motor_max_thrust),
                            (self.motor_max_thrust, self.motor_min_thrust),
                            (self.motor_max_thrust, self.motor_max_thrust)]
            self.action_space = spaces.Discrete(len(self.actions))
        self.observation_space = spaces.Box(
            np.array([0, 0, -np.pi, 0, 0]),
            np.array([self.world_size, self.world_size, np.pi, self.world_size, self.world_size]),
            dtype=np.float64)
        self.wave_space = spaces.Box(self.wave_min_speed, self.wave_max_speed, shape=[2], dtype=np.float64)
        # aqua.py:86-94: the rewards are callables in the reference
        self.reward_goal = lambda: +10
        self.reward_max_steps = lambda: -10
        self.reward_collision = lambda: -10
        self.reward_step = lambda: (self._distance_from_goal(self._prev_obs) - self._distance_from_goal(self._last_obs)) * 0.7
        self.time_limit = 1000
        self.tau = 1
        self._ctor = dict(continuous=self.continuous, obstacles=obstacles, waves=waves, random_boat=random_boat,
                          random_goal=random_goal, device=device, precision="float64", autoreset=False)
        self._vec = None
        self.seed(seed)
        self.obstacles = self._vec.obstacles
        self.time = 0
        # host buffers of the one-call step (aqua_step_host: action in, obs + reward + flags out)
        self._act_h = np.zeros((1, 2), np.float32) if self.continuous else np.zeros(1, np.int32)
        self._obs_h = np.zeros((5, 1), np.float64)
        self._rew_h = np.zeros(1, np.float64)
        self._flg_h = np.zeros(1, np.uint8)
        self._prev_obs = self._last_obs = None
        import ctypes
        self._ptrs = tuple(ctypes.c_void_p(a.ctypes.data) for a in (self._act_h, self._obs_h, self._rew_h, self._flg_h))
        self._adt = N.ACT_F32X2 if self.continuous else N.ACT_I32

    def seed(self, seed=None):
        """Key of the env's draw streams (reset samples, wave noise): ``None`` = a fresh one from
        the OS.  Re-creates the device env; call :meth:`reset` afterwards."""
        if seed is None:
            import os
            seed = int.from_bytes(os.urandom(8), "little")
        if self._vec is not None:
            self._vec.close()
        self._seed = int(seed) & ((1 << 64) - 1)
        self._vec = AquaVecEnv(1, seed=self._seed, **self._ctor)
        return [self._seed]

    @staticmethod
    def _distance_from_goal(obs):
        """aqua.py:392-402 without the radii (they cancel in reward_step)"""
        return float(np.hypot(obs[3] - obs[0], obs[4] - obs[1]))

    def reset(self):
        """aqua.py:100-126 -> float64[5] = (x, y, angle, goal_x, goal_y)."""
        self._vec.reset()
        self.time = 0
        self._last_obs = self._vec.obs_planes[:, 0].cpu().numpy().astype(np.float64)
        self._prev_obs = self._last_obs
        return self._last_obs.copy()

    def step(self, action):
        """aqua.py:135-213 -> (obs float64[5], reward, done bool, info dict of 3 bools).  One call
        into the library per step: action up, observation + reward + flags down, one synchronisation."""
        if self.continuous:
            # aqua.py:137-138.  The device step takes float32 thrusts (AQUA_ACT_F32X2): parity with the
            # reference's float64 arithmetic holds for float32-representable thrusts
            self._act_h[0] = np.asarray(action, dtype=np.float64).reshape(2)
        else:
            idx = int(action)
            self.actions[idx]  # IndexError on a bad index, like aqua.py:154
            self._act_h[0] = idx % len(self.actions)
        self._vec.step_host_raw(self._ptrs[0], self._adt, self._ptrs[1], self._ptrs[2], self._ptrs[3])
        flags = int(self._flg_h[0])
        self.time += 1
        self._prev_obs, self._last_obs = self._last_obs, self._obs_h[:, 0].copy()
        info = {'Termination.collided': bool(flags & 2), 'Termination.time': bool(flags & 4),
                'Termination.success': bool(flags & 8)}
        reward = self._rew_h[0]
        # aqua.py:199-211: terminal rewards are the Python ints of the reward_* callables
        reward = int(reward) if flags & 1 else float(reward)
        return self._last_obs.copy(), reward, bool(flags & 1), info

    def render(self, mode='human'):
        """aqua.py:215-365 opens a pyglet window; not part of the accelerated path: no-op."""
        return None

    def close(self):
        """aqua.py:367-371"""
        if getattr(self, "_vec", None) is not None:
            self._vec.close()

    def normalize_angle(self, value):
        """aqua.py:128-133"""
        lo, hi = self.observation_space.low[2], self.observation_space.high[2]
        width = hi - lo
        off = value - lo
        return (off - (np.floor(off / width) * width)) + lo


class AquaContinuousEnv(AquaEnv):
    continuous = True  # aqua.py:458-459
