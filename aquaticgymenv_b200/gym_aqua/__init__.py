"""``gym_aqua`` drop-in: the six environment ids of the reference (gym_aqua/__init__.py:4-41).

    from aquaticgymenv_b200 import gym_aqua
    env = gym_aqua.make("AquaEnv-v0", obstacles=True)            # one env, reference API
    venv = gym_aqua.make("AquaEnv-v1", num_envs=65536)           # batched GPU env (AquaVecEnv)

If the third-party ``gym`` package is installed the ids are also registered with it, so
``gym.make("AquaEnv-v0", ...)`` of dqn_train_policy.py:19 resolves to these classes.
"""
import numpy as np

from ..vec_env import DEFAULT_OBSTACLES, DIFFICULT_OBSTACLES, AquaVecEnv, table_to_reference_list
from .envs import AquaContinuousEnv, AquaEnv

# gym_aqua/__init__.py:24-31
difficult_obstacles = table_to_reference_list(DIFFICULT_OBSTACLES)

# id -> (class, continuous, default kwargs)   gym_aqua/__init__.py:4-41
REGISTRY = {
    'AquaEnv-v0': (AquaEnv, False, {}),
    'AquaContinuousEnv-v0': (AquaContinuousEnv, True, {}),
    'AquaEnv-v1': (AquaEnv, False, {'obstacles': True}),
    'AquaContinuousEnv-v1': (AquaContinuousEnv, True, {'obstacles': True}),
    'AquaEnv-v2': (AquaEnv, False, {'obstacles': difficult_obstacles}),
    'AquaContinuousEnv-v2': (AquaContinuousEnv, True, {'obstacles': difficult_obstacles}),
}


def make(env_id, num_envs=None, **kwargs):
    """``gym.make(id, **kwargs)``: registered kwargs overridden by the caller's.  With
    ``num_envs`` the batched :class:`AquaVecEnv` is returned instead of the single-env class."""
    if env_id not in REGISTRY:
        raise KeyError("No registered env with id: %s" % env_id)
    cls, continuous, defaults = REGISTRY[env_id]
    kw = dict(defaults)
    kw.update(kwargs)
    if num_envs is None:
        return cls(**kw)
    return AquaVecEnv(num_envs, continuous=continuous, **kw)


def _register_with_gym():  # pragma: no cover - only when gym is installed
    try:
        from gym.envs.registration import register
    except Exception:
        return
    for env_id, (cls, _, defaults) in REGISTRY.items():
        try:
            register(id=env_id, entry_point='aquaticgymenv_b200.gym_aqua.envs:%s' % cls.__name__,
                     kwargs=defaults)
        except Exception:
            pass


_register_with_gym()
