"""Minimal stand-in for the third-party `gym` package (absent from this image).

TEST INFRASTRUCTURE ONLY: it exists so that `oracle/gen_golden.py` can import and
EXECUTE the unmodified reference `gym_aqua` package from /root/reference to
generate golden vectors.  It implements exactly the surface the reference touches
(SURVEY.md section 8c): `gym.Env`, `gym.spaces.Box/Discrete`,
`gym.envs.registration.register`, `gym.make`.  Semantics follow gym 0.17/0.18
(the versions contemporaneous with the reference): 4-tuple `step`, obs-only
`reset`, no TimeLimit wrapper unless `max_episode_steps` is registered.
"""
from gym import spaces  # noqa: F401
from gym.core import Env  # noqa: F401
from gym.envs.registration import make, register  # noqa: F401

__version__ = "0.17.standin"
