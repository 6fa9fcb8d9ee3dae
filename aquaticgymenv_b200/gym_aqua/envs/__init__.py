from .aqua import AquaEnv, AquaContinuousEnv  # noqa: F401  (gym_aqua/envs/__init__.py:1)
