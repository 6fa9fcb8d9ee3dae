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
    """Runs the policy until at least ``episodes`` episodes have finished (all envs in parallel,
    auto-reset) and returns the statistics of dqn_test_plotter.py:23-35: success rate, mean and
    standard deviation of the episode reward, mean length, and the termination counts."""
    if not env.autoreset:
        raise ValueError("evaluate() needs an auto-resetting AquaVecEnv")
    env.reset()
    n = env.num_envs
    ret = torch.zeros(n, dtype=torch.float64, device=env.device)
    rewards, lengths = [], []
    counts = torch.zeros(3, dtype=torch.int64, device=env.device)  # collided, time, success
    steps_in_ep = torch.zeros(n, dtype=torch.int64, device=env.device)
    done_total, t = 0, 0
    limit = max_steps if max_steps is not None else 1002 * (episodes // n + 2)
    while done_total < episodes and t < limit:
        obs = env.obs  # [N, 5] view: after auto-reset it already holds the next episode's first obs
        flags = env.step_flags(policy.get_action(obs))
        ret += env.reward.double()
        steps_in_ep += 1
        done = (flags & 1) != 0
        if bool(done.any()):
            rewards.append(ret[done].clone())
            lengths.append(steps_in_ep[done].clone())
            counts += torch.stack([((flags & 2) != 0).sum(), ((flags & 4) != 0).sum(), ((flags & 8) != 0).sum()])
            done_total += int(done.sum())
            ret[done] = 0.0
            steps_in_ep[done] = 0
        t += 1
    rewards = torch.cat(rewards) if rewards else torch.zeros(0, dtype=torch.float64)
    lengths = torch.cat(lengths) if lengths else torch.zeros(0, dtype=torch.int64)
    c = counts.tolist()
    total = max(len(rewards), 1)
    return {"episodes": len(rewards), "success_rate": c[2] / total, "collided": c[0], "time": c[1],
            "success": c[2], "reward_mean": float(rewards.mean()) if len(rewards) else 0.0,
            "reward_std": float(rewards.std()) if len(rewards) > 1 else 0.0,
            "length_mean": float(lengths.double().mean()) if len(lengths) else 0.0, "steps": t}
