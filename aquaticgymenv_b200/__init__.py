"""aquaticgymenv_b200 -- B200-native batched AquaGym environment step.

Public surface:
  * ``AquaVecEnv``            batched env on one GPU (torch tensors in/out)
  * ``gym_aqua.make(id, ..)`` the reference's six env ids; single-env drop-in classes
  * ``sharding``              one env slab per rank + all-reduce of the episode statistics
  * ``dqn`` / ``policies``    batched torch DQN trainer, evaluation harness and baseline policies
  * ``savedmodel``            reader of the reference's shipped TF SavedModel policies
The compute path is ``libaquagym.so`` (hand-written sm_100a CUDA behind ``include/aquagym.h``);
importing this package without the built library raises.
"""
from . import _native

_native.load()  # fail loudly if the CUDA library has not been built: there is no fallback

from .vec_env import (AquaVecEnv, DEFAULT_OBSTACLES, DIFFICULT_OBSTACLES, LazyInfo,  # noqa: E402,F401
                      obstacle_table, stats_dict, table_to_reference_list)
from . import gym_aqua  # noqa: E402,F401
from . import sharding  # noqa: E402,F401

__all__ = ["AquaVecEnv", "gym_aqua", "sharding", "obstacle_table", "DEFAULT_OBSTACLES",
           "DIFFICULT_OBSTACLES"]
