#!/usr/bin/env python
"""BASELINE config 5 as a LEARNING run: the batched DQN trainer (aquaticgymenv_b200/dqn.py, the
reference's hyper-parameters dqn.py:17-30 with the batch / buffer scaled to a vector env) driving
the GPU env until the greedy policy matches the reference's published success (0.938 without
obstacles, 0.667 with the default ones; example_policies/test_results.pickle).  One JSON line per
setting with the learning curve, env-steps/s and learner-steps/s OF THAT RUN, and the greedy
success over 8192 test episodes.
usage (GPU box): python tools/dqn_train.py [vector_steps] [num_envs] [batch]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aquaticgymenv_b200 import dqn, gym_aqua, policies  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 48000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
HP = {"EPSILON_DECAY": 15000, "BUFFER_SIZE": 300000, "MINIBATCH_SIZE": batch, "UPDATES_PER_STEP": 1,
      "LEARNING_RATE": 1e-3}   # GAMMA 0.98, TAU 0.005, EPSILON 1.0 -> 0.05, double DQN: the reference's

only = os.environ.get("DQN_ONLY")   # "0" / "1": just that obstacle setting
for obstacles, published in ((False, 0.938), (True, 0.667)):
    if only is not None and int(only) != int(obstacles):
        continue
    HP["UPDATES_PER_STEP"] = int(os.environ.get("DQN_UPS", "1"))
    env = gym_aqua.make("AquaEnv-v0", obstacles=obstacles, num_envs=n, seed=1)   # dqn_train_policy.py:19
    test_env = gym_aqua.make("AquaEnv-v0", obstacles=obstacles, num_envs=2048, seed=99)
    in_kernel = os.environ.get("DQN_IN_KERNEL", "0") == "1"   # acting through aqua_rollout_policy_dev
    agent = dqn.BatchedDQN(env, hyperparams=HP, seed=0, cuda_graph=True, in_kernel_acting=in_kernel)
    curve, busy = [], 0.0
    chunk = max(steps // 8, 1)
    done = 0
    while done < steps:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        agent.train(env_steps=chunk)
        torch.cuda.synchronize()
        busy += time.perf_counter() - t0
        done += chunk
        res = policies.evaluate(test_env, policies.GreedyQPolicy(test_env, agent.q_function), episodes=8192)
        curve.append({"vector_steps": done, "episodes": agent.episode,
                      "train_success_last_2000": float(sum(agent.success_list[-2000:]) / max(len(agent.success_list[-2000:]), 1)),
                      "greedy_success": res["success_rate"], "greedy_reward_mean": res["reward_mean"]})
    print(json.dumps({
        "config": "BASELINE config 5: batched double-DQN (5-64-64-3, Adam 1e-3, gamma 0.98, tau 0.005, eps 1.0->0.05 over "
                  "15000 episodes) on AquaEnv-v0 obstacles=%s, %d envs, batch %d, 1 learner step per vector step, "
                  "one CUDA graph per vector step%s" % (obstacles, n, batch, ", acting inside the step kernel" if in_kernel else ""),
        "hyperparams": HP, "vector_steps": steps, "env_steps": n * steps, "learner_steps": agent.learner_steps,
        "env_steps_per_s": n * steps / busy, "learner_steps_per_s": agent.learner_steps / busy,
        "episodes_seen": agent.episode, "greedy_success_rate": curve[-1]["greedy_success"],
        "best_checkpoint_success": max(c["greedy_success"] for c in curve),
        "greedy_reward_mean": curve[-1]["greedy_reward_mean"], "reference_published_success": published,
        "curve": curve}), flush=True)
    env.close()
    test_env.close()
