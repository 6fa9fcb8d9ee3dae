#!/usr/bin/env python
"""Hyper-parameter probe for the batched DQN trainer on the CPU: the trainer of
aquaticgymenv_b200/dqn.py driving the C ORACLE env through a small adapter (there is no GPU in the
build container).  Development tool only -- the product trainer runs on the CUDA env.
usage: python tools/dqn_cpu_probe.py [n_envs] [vector_steps] [batch] [updates_per_step] [eps_decay_episodes] [obstacles 0/1]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import aqua_oracle as ao  # noqa: E402
from aquaticgymenv_b200 import dqn, policies  # noqa: E402


class OracleTorchEnv(object):
    continuous = False
    autoreset = True

    def __init__(self, n, obstacles, seed=0):
        self.num_envs, self.device = n, torch.device("cpu")
        self.o = ao.OracleVecEnv(n, obstacles=obstacles, seed=seed)
        self.reward = torch.zeros(n)
        self._step_index = 0
        self._scale = np.array([100, 100, 2 * np.pi, 100, 100.0])

    def reset(self):
        self.o.reset()

    @property
    def obs(self):
        return torch.as_tensor(self.o.obs())

    def observe(self, normalized=True):
        x = self.o.obs() / self._scale
        x[:, 2] += 0.5
        return torch.as_tensor(x, dtype=torch.float32)

    def step_flags(self, a):
        _, r, f, _ = self.o.step(a.numpy().astype(np.int32), autoreset=True, nthreads=4)
        self.reward = torch.as_tensor(r, dtype=torch.float32)
        self._step_index += 1
        return torch.as_tensor(f)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    ups = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    decay = int(sys.argv[5]) if len(sys.argv) > 5 else 20000
    obst = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    lr = float(sys.argv[7]) if len(sys.argv) > 7 else 1e-3
    buf = int(sys.argv[8]) if len(sys.argv) > 8 else 200000
    table = np.load(os.path.join(ROOT, "tests", "golden", "v1_table.npy")) if obst else np.zeros((0, 5))
    torch.set_num_threads(4)
    env = OracleTorchEnv(n, table, seed=1)
    agent = dqn.BatchedDQN(env, hyperparams={"EPSILON_DECAY": decay, "BUFFER_SIZE": buf, "MINIBATCH_SIZE": batch,
                                             "UPDATES_PER_STEP": ups, "LEARNING_RATE": lr}, seed=0)
    t0 = time.time()
    done = 0
    chunk = max(steps // 8, 1)
    while done < steps:
        agent.train(env_steps=chunk)
        done += chunk
        ev = OracleTorchEnv(1024, table, seed=99)
        res = policies.evaluate(ev, policies.GreedyQPolicy(ev, agent.q_function), episodes=1024)
        print("vector steps %6d  env-steps %.3g  learner steps %d  episodes %d  eps %.3f  train success(last 2000) %.3f  "
              "greedy success %.3f  reward %.2f  len %.1f  [%.0f s]" % (
                  done, agent.env_steps, agent.learner_steps, agent.episode,
                  float(agent._epsilon(agent._decay)), float(np.mean(agent.success_list[-2000:])) if agent.success_list else 0,
                  res["success_rate"], res["reward_mean"], res["length_mean"], time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
