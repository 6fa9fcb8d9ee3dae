"""Host-side helpers of the reference's agent code (main/impl/utils.py), kept so that the
reference's scripts find the same names.  The batched, on-device equivalent of the normaliser is
``AquaVecEnv.observe(normalized=True)`` (one kernel, no host round trip)."""
import numpy as np


def aqua_term_string(info, episode_total_reward, steps):
    """main/impl/utils.py:4-12"""
    if info['Termination.time']:
        msg = "reached max steps                "
    elif info['Termination.collided']:
        msg = "collided ({:3d} steps)             ".format(steps)
    else:
        msg = ">>> REACHED GOAL ({: 3d} steps) <<<".format(steps)
    return "{}  [REW: {: 7.2f}]".format(msg, episode_total_reward)


class Normalizer(object):
    """main/impl/dqn.py:47-52"""

    def __call__(self, *args, **kwargs):
        pass

    def inverse(self, norm_state):
        pass


class AquaStateNormalizer(Normalizer):
    """main/impl/utils.py:15-33: obs / (high - low), angle + 0.5, add a batch dimension."""

    def __init__(self, env):
        self.env = env
        self.observation_range = env.observation_space.high - env.observation_space.low

    def __call__(self, state):
        norm_state = np.array(state, dtype=np.float64)
        norm_state /= self.observation_range
        norm_state[2] += 0.5
        return norm_state[None, :]

    def inverse(self, norm_state):
        state = np.array(norm_state, dtype=np.float64)
        state[0, 2] -= 0.5
        state *= self.observation_range
        return state[0, :]
