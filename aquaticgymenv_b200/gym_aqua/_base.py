"""``gym.Env`` stand-in base (aqua.py:9 subclasses ``gym.Env``; gym is third party and may be
absent).  Uses the real class when gym is importable so isinstance checks of user code hold."""
try:  # pragma: no cover - depends on the user's environment
    import gym as _gym
    Env = _gym.Env
except Exception:  # gym not installed
    class Env(object):
        metadata = {}
        reward_range = (-float("inf"), float("inf"))
        action_space = None
        observation_space = None

        def seed(self, seed=None):
            return [seed]

        @property
        def unwrapped(self):
            return self

        def __enter__(self):
            return self

        def __exit__(self, *a):
            self.close()
            return False
