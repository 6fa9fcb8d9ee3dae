#!/usr/bin/env python
"""Time ONE configuration of the batched step (single-step launches, actions resident in HBM) and
print its episode statistics -- the method of bench.py's `other_configs`, for profiling a config
on its own (e.g. under ncu).
usage (GPU box): python tools/config_rate.py [env_id] [envs] [float32|float64] [dense64|-] [steps]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

env_id = sys.argv[1] if len(sys.argv) > 1 else "AquaEnv-v0"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
precision = sys.argv[3] if len(sys.argv) > 3 else "float32"
table = sys.argv[4] if len(sys.argv) > 4 else "dense64"
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 100

import torch  # noqa: E402
from aquaticgymenv_b200 import gym_aqua  # noqa: E402

kw = {}
if table == "dense64":
    kw["obstacles"] = np.load(os.path.join(ROOT, "tests", "golden", "dense64_table.npy"))
env = gym_aqua.make(env_id, num_envs=n, seed=0, precision=precision, **kw)
env.reset()
pool = [env.random_actions(k) for k in range(4)]
for k in range(30):  # past the first episodes: the reset rate of the steady state
    env.step_flags(pool[k % 4])
env.stats(clear=True)
torch.cuda.synchronize()
start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
start.record()
for k in range(steps):
    env.step_flags(pool[k % 4])
end.record()
torch.cuda.synchronize()
dt = start.elapsed_time(end) * 1e-3
st = env.stats()
env.close()
print(json.dumps({"env": env_id, "table": table, "envs": n, "precision": precision, "steps": steps,
                  "env_steps_per_s": n * steps / dt, "us_per_step": 1e6 * dt / steps,
                  "resets_per_env_step": st["episodes"] / max(st["steps"], 1),
                  "mean_episode_length": st.get("mean_episode_length"),
                  "success_rate": st.get("success_rate")}))
