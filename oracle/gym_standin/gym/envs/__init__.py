from gym.envs.registration import make, register, registry  # noqa: F401
