#!/usr/bin/env python
"""Which envs-per-thread instantiation of aqua_step_many wins at which slab size (K = 64, 8 launches
back to back).  usage (GPU box): python tools/many_vsweep.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, ROOT)
    from aquaticgymenv_b200 import gym_aqua
    n, K = int(sys.argv[1]), 64
    env = gym_aqua.make("AquaEnv-v1", num_envs=n, seed=0)
    env.reset()
    n16 = (n + 15) // 16 * 16
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
    print(json.dumps({"envs": n, "K": K, "v": os.environ["AQUA_MANY_V"], "threads": os.environ["AQUA_MANY_THREADS"],
                      "env_steps_per_s": n * K / best}))
else:
    for n in (65536, 98304, 131072, 196608, 262144, 393216, 524288, 786432, 1048576, 2097152):
        row = []
        for v, thr in ((1, 64), (1, 128), (2, 64), (2, 128), (4, 64), (4, 128)):
            env = dict(os.environ, AQUA_MANY_V=str(v), AQUA_MANY_THREADS=str(thr))
            out = subprocess.run([sys.executable, __file__, str(n)], env=env, capture_output=True, text=True)
            try:
                row.append("V%d/%d %.3g" % (v, thr, json.loads(out.stdout.strip().splitlines()[-1])["env_steps_per_s"]))
            except Exception:
                row.append("V%d/%d FAIL" % (v, thr))
        print("%8d  " % n + "  ".join(row), flush=True)
