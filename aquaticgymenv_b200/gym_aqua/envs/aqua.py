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
from ...vec_env import AquaVecEnv


class AquaEnv(_base.Env):
    metadata = {'render.modes': ['human']}
    continuous = False

    def __init__(self, obstacles=False, waves=True, random_boat=True, random_goal=True, seed=0,
                 device=None):
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
        # aqua.py:86-94
        self.reward_goal, self.reward_max_steps, self.reward_collision = +10, -10, -10
        self.time_limit = 1000
        self.tau = 1
        self._vec = AquaVecEnv(1, continuous=self.continuous, obstacles=obstacles, waves=waves,
                               random_boat=random_boat, random_goal=random_goal, seed=seed,
                               device=device, precision="float64", autoreset=False)
        self.obstacles = self._vec.obstacles
        self.time = 0

    def _obs(self):
        return self._vec.obs_planes[:, 0].cpu().numpy().astype(np.float64)

    def reset(self):
        """aqua.py:100-126 -> float64[5] = (x, y, angle, goal_x, goal_y)."""
        self._vec.reset()
        self.time = 0
        return self._obs()

    def step(self, action):
        """aqua.py:135-213 -> (obs float64[5], reward, done bool, info dict of 3 bools)."""
        import torch
        if self.continuous:
            a = np.asarray(action, dtype=np.float64).reshape(2)  # aqua.py:137-138
            act = torch.as_tensor(a.astype(np.float32)).reshape(1, 2)
        else:
            idx = int(action)
            self.actions[idx]  # IndexError on a bad index, like aqua.py:154
            act = torch.tensor([idx % len(self.actions)], dtype=torch.int32)
        flags = int(self._vec.step_flags(act.to(self._vec.device)).cpu()[0])
        reward = float(self._vec.reward.cpu()[0])
        self.time += 1
        info = {'Termination.collided': bool(flags & 2), 'Termination.time': bool(flags & 4),
                'Termination.success': bool(flags & 8)}
        return self._obs(), reward, bool(flags & 1), info

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
