"""Batched evaluation harness and the reference's hand-written policies
(main/testing/__init__.py:17-36 rollout loop, test_random.py:19-20, test_optimal.py:8-28,
dqn_test_plotter.py:23-35 statistics), running on the device against an :class:`AquaVecEnv`."""
import math

import torch


class Policy(object):
    """``get_action(obs)`` maps observation rows ``[N, 5]`` to actions for every env."""

    def get_action(self, state):
        raise NotImplementedError("no action selection strategy has been implemented")


class RandomAquaPolicy(Policy):
    """test_random.py:19-20 (`action_space.sample()`), from the env's Philox ACTION stream."""

    def __init__(self, env):
        self.env = env

    def get_action(self, state):
        return self.env.random_actions()


class OptimalAquaPolicy(Policy):
    """test_optimal.py:8-28: steer towards the goal bearing, threshold 8 degrees.  Meant for
    discrete, obstacle-free setups (test_optimal.py:31)."""
    ANGLE_THRESHOLD = 8 / 180 * math.pi

    @staticmethod
    def _normalize_angle(angle):
        return torch.remainder(angle + 2 * math.pi, 2 * math.pi)

    def get_action(self, state):
        boat, goal = state[:, :2], state[:, 3:5]
        boat_angle_wrt_zero = self._normalize_angle(state[:, 2] + math.pi / 2)
        d = goal - boat
        goal_angle_wrt_boat = self._normalize_angle(torch.atan2(d[:, 1], d[:, 0]))
        angle_diff = goal_angle_wrt_boat - boat_angle_wrt_zero
        turn = torch.where(angle_diff > 0, 0, 1)
        return torch.where(angle_diff.abs() > self.ANGLE_THRESHOLD, turn, 2).to(torch.int32)


class GreedyQPolicy(Policy):
    """test_dqn.py:14-22: argmax of a Q-network on the normalised observation."""

    def __init__(self, env, q_function):
        self.env, self.q = env, q_function

    @torch.no_grad()
    def get_action(self, state):
        return self.q(self.env.observe(normalized=True)).argmax(dim=1).to(torch.int32)


@torch.no_grad()
def evaluate(env, policy, episodes, max_steps=None):
    """The statistics of dqn_test_plotter.py:23-35 (success rate, mean and standard deviation of
    the episode reward, mean length, termination counts) over at least ``episodes`` episodes, all
    envs in parallel with auto-reset.

    Which episodes count is fixed up front: the FIRST ``ceil(episodes / num_envs)`` episodes of
    every env.  (Counting whichever episodes finish first would favour short ones -- early
    collisions -- over long and timed-out ones.)  Every episode ends within 1001 steps
    (aqua.py:203), so the default step limit lets every env complete its quota: no truncation.
    With an explicit smaller ``max_steps`` the unfinished quota is reported in ``truncated``.
    The loop does not wait for the device: counters live there and are read every 64 steps."""
    if not env.autoreset:
        raise ValueError("evaluate() needs an auto-resetting AquaVecEnv")
    env.reset()
    n, dev = env.num_envs, env.device
    quota = -(-int(episodes) // n)
    limit = max_steps if max_steps is not None else 1001 * quota + 1
    ret = torch.zeros(n, dtype=torch.float64, device=dev)
    length = torch.zeros(n, dtype=torch.int64, device=dev)
    count = torch.zeros(n, dtype=torch.int64, device=dev)
    acc = torch.zeros(7, dtype=torch.float64, device=dev)  # episodes, collided, time, success, sum r, sum r^2, sum len
    t = 0
    while t < limit:
        flags = env.step_flags(policy.get_action(env.obs))  # env.obs: after auto-reset the next episode's first obs
        ret += env.reward.double()
        length += 1
        done = (flags & 1) != 0
        counted = done & (count < quota)
        c = counted.double()
        acc += torch.stack([c.sum(), (c * ((flags & 2) != 0)).sum(), (c * ((flags & 4) != 0)).sum(),
                            (c * ((flags & 8) != 0)).sum(), (c * ret).sum(), (c * ret * ret).sum(),
                            (c * length).sum()])
        count += counted
        ret.masked_fill_(done, 0.0)
        length.masked_fill_(done, 0)
        t += 1
        if t % 64 == 0 and bool((count >= quota).all()):
            break
    k, coll, tim, succ, sr, sr2, sl = acc.tolist()
    k = int(k)
    mean = sr / max(k, 1)
    var = max(sr2 - k * mean * mean, 0.0) / max(k - 1, 1)
    return {"episodes": k, "success_rate": succ / max(k, 1), "collided": int(coll), "time": int(tim),
            "success": int(succ), "reward_mean": mean, "reward_std": var ** 0.5,
            "length_mean": sl / max(k, 1), "steps": t, "episodes_per_env": quota,
            "truncated": int(n * quota - k)}
