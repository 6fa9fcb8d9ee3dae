#!/usr/bin/env python
"""BASELINE config 5: the batched DQN loop driving the GPU env with on-device epsilon-greedy
actions.  Prints one JSON line with env-steps/s and learner-steps/s.
usage: python tools/dqn_bench.py [num_envs] [vector_steps] [graph|eager]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aquaticgymenv_b200 import dqn, gym_aqua, policies  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 600
env = gym_aqua.make("AquaEnv-v0", obstacles=True, num_envs=n, seed=0)   # dqn_train_policy.py:19
graph = (sys.argv[3] if len(sys.argv) > 3 else "graph") == "graph"   # "eager": one launch per torch op
agent = dqn.BatchedDQN(env, hyperparams={"EPSILON_DECAY": 40000, "BUFFER_SIZE": 2_000_000,
                                         "MINIBATCH_SIZE": 4096, "UPDATES_PER_STEP": 1}, cuda_graph=graph)
agent.train(env_steps=20)
torch.cuda.synchronize()
t0 = time.perf_counter()
agent.train(env_steps=steps)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
res = policies.evaluate(env, policies.GreedyQPolicy(env, agent.q_function), episodes=20000)
print(json.dumps({"config": "DQN loop (dqn.py hyper-parameters, batch 4096) on AquaEnv-v0 obstacles=True",
                  "mode": "one CUDA graph per vector step" if graph else "eager",
                  "num_envs": n, "vector_steps": steps, "env_steps_per_s": n * steps / dt,
                  "learner_steps_per_s": steps / dt, "episodes_seen": len(agent.reward_list),
                  "greedy_success_rate_after": res["success_rate"], "greedy_reward_mean": res["reward_mean"]}))
