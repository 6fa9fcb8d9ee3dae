#!/usr/bin/env python
"""Steps-per-launch sweep of aqua_step_many at one slab size: 8 launches back to back per point
(launch overhead overlapped, as in a real rollout loop), CUDA events around the 8.
usage (GPU box): python tools/many_ksweep.py [envs] [v] [threads]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
os.environ["AQUA_MANY_V"] = sys.argv[2] if len(sys.argv) > 2 else "1"
os.environ["AQUA_MANY_THREADS"] = sys.argv[3] if len(sys.argv) > 3 else "64"
from aquaticgymenv_b200 import gym_aqua  # noqa: E402

env = gym_aqua.make("AquaEnv-v1", num_envs=n, seed=0)
env.reset()
n16 = (n + 15) // 16 * 16
for K in (4, 8, 16, 32, 64, 128, 256):
    tape = torch.randint(0, 3, (K, n16), dtype=torch.int32, device="cuda")
    rew = torch.empty(K, n16, device="cuda")
    flg = torch.empty(K, n16, dtype=torch.uint8, device="cuda")
    for _ in range(4):
        env.step_many(tape, rew, flg)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(8):
            env.step_many(tape, rew, flg)
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e-3 / 8)
    print(json.dumps({"envs": n, "K": K, "us_per_launch": best * 1e6, "us_per_step": best * 1e6 / K,
                      "env_steps_per_s": n * K / best}), flush=True)
