#!/usr/bin/env python
"""One configuration of aqua_step_many for profiling (ncu): python tools/many_one.py envs K v threads [env_id] [dense64]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
n, K, v, threads = (int(a) for a in sys.argv[1:5])
env_id = sys.argv[5] if len(sys.argv) > 5 else "AquaEnv-v1"
os.environ["AQUA_MANY_V"], os.environ["AQUA_MANY_THREADS"] = str(v), str(threads)
from aquaticgymenv_b200 import gym_aqua  # noqa: E402
kw = {}
if len(sys.argv) > 6 and sys.argv[6] == "dense64":
    kw["obstacles"] = np.load(os.path.join(ROOT, "tests", "golden", "dense64_table.npy"))
env = gym_aqua.make(env_id, num_envs=n, seed=0, **kw)
env.reset()
n16 = (n + 15) // 16 * 16
tape = torch.rand(K, n16, 2, device="cuda") * 0.3 + 0.2 if env.continuous else torch.randint(0, 3, (K, n16), dtype=torch.int32, device="cuda")
rew = torch.empty(K, n16, device="cuda")
flg = torch.empty(K, n16, dtype=torch.uint8, device="cuda")
for _ in range(5):
    env.step_many(tape, rew, flg)
torch.cuda.synchronize()
print(env.stats())
