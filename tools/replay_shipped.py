#!/usr/bin/env python
"""Replays the two policies the reference ships (tests/golden/policy_*.npz = the weights of
example_policies/*/models/*, read by aquaticgymenv_b200/savedmodel.py) greedily on the GPU env and
prints the statistics of dqn_test_plotter.py:23-35 next to the published ones.
usage (GPU box): python tools/replay_shipped.py [episodes]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aquaticgymenv_b200 import gym_aqua, policies, savedmodel  # noqa: E402

episodes = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
for name, env_id, pub_s, pub_r in (("no_obs", "AquaEnv-v0", 0.938, 38.83), ("with_obs", "AquaEnv-v1", 0.667, 25.52)):
    for precision in ("float32", "float64"):
        env = gym_aqua.make(env_id, num_envs=16384, seed=11, precision=precision)
        q = savedmodel.load_q_network(os.path.join(ROOT, "tests", "golden", "policy_%s.npz" % name), device=env.device)
        res = policies.evaluate(env, policies.GreedyQPolicy(env, q), episodes=episodes)
        res.update(policy=name, env=env_id, precision=precision, published_success=pub_s, published_reward_mean=pub_r)
        print(json.dumps(res), flush=True)
        env.close()
