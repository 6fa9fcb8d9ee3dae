#!/usr/bin/env python
"""Host-buffer step (aqua_step_host) rate against the number of pipeline chunks, with a plain
pinned device-to-host copy of the same bytes beside it.  usage (GPU box): python tools/e2e_probe.py [envs]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
if len(sys.argv) > 2:  # child: one chunk count
    import numpy as np
    import torch
    from aquaticgymenv_b200 import gym_aqua
    env = gym_aqua.make("AquaEnv-v1", num_envs=n, seed=0)
    env.reset()
    a = torch.empty(n, dtype=torch.int32).pin_memory()
    a.copy_(env.random_actions(0).cpu())
    o = torch.empty(5, n).pin_memory()
    r = torch.empty(n).pin_memory()
    f = torch.empty(n, dtype=torch.uint8).pin_memory()
    an, on, rn, fn = a.numpy(), o.numpy(), r.numpy(), f.numpy()
    for _ in range(3):
        env.step_host(an, on, rn, fn)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        env.step_host(an, on, rn, fn)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print(json.dumps({"envs": n, "chunks": os.environ.get("AQUA_HOST_CHUNKS", "default"), "us_per_step": dt * 1e6,
                      "env_steps_per_s": n / dt, "d2h_gbs": 17 * n / dt / 1e9}))
else:
    for route in ("pack kernel (SM writes to the mapped buffers)", "copy engines"):
        for k in ("default", "1", "2", "4", "8", "16"):
            env = dict(os.environ)
            if k != "default":
                env["AQUA_HOST_CHUNKS"] = k
            if route == "copy engines":
                env["AQUA_HOST_NO_PACK"] = "1"
            out = subprocess.run([sys.executable, __file__, str(n), "child"], env=env, capture_output=True, text=True)
            print(route + ": " + (out.stdout.strip() or out.stderr[-400:]), flush=True)
