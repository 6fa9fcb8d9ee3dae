#!/usr/bin/env python
"""One launch per step (aqua_step): narrow (one env per thread) against wide (streaming) geometry by
slab size, eager launches back to back.  usage (GPU box): python tools/single_geom_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aquaticgymenv_b200 import gym_aqua  # noqa: E402

print("# AquaEnv-v1, aqua_step, env-steps/s by slab size and geometry (400 eager launches back to back)")
for n in (65536, 131072, 262144, 524288, 1 << 20, 1 << 21, 1 << 22):
    row = []
    for geom in ("narrow", "wide", "auto"):
        env = gym_aqua.make("AquaEnv-v1", num_envs=n, seed=0, geometry=geom)
        env.reset()
        pool = [env.random_actions(k) for k in range(4)]
        for k in range(30):
            env.step_flags(pool[k % 4])
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 400 if n <= (1 << 20) else 100
        s.record()
        for k in range(steps):
            env.step_flags(pool[k % 4])
        e.record()
        torch.cuda.synchronize()
        row.append("%s %.3g (%.1f us)" % (geom, n * steps / (s.elapsed_time(e) * 1e-3), s.elapsed_time(e) * 1e3 / steps))
        env.close()
    print("%8d  " % n + "  ".join(row), flush=True)
