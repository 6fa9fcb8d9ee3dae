#!/usr/bin/env python
"""Times aqua_step_many (K env-steps per launch) over slab sizes, envs per thread and CTA sizes on
ONE box.  usage (GPU box): python tools/many_probe.py [quick]
Prints one JSON line per point: env-steps/s with the int32 action tape resident, reward + flags
written; `frac65` = that rate x 65 B / the measured copy bandwidth (the single-step roofline
convention of SURVEY.md section 8d; the kernel itself moves 9 + 64/K bytes per env-step)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aquaticgymenv_b200 import gym_aqua  # noqa: E402

PEAK = 6549.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def rate(env_id, n, K, v, threads, obstacles=None, reps=5, mode="tape", precision="float32", b2b=1):
    """b2b: launches back to back per timing (as bench.py times them); v = 0: the library's own geometry"""
    os.environ["AQUA_MANY_V"] = str(v)
    os.environ["AQUA_MANY_THREADS"] = str(threads)
    kw = {}
    if obstacles == "dense64":
        kw["obstacles"] = np.load(os.path.join(ROOT, "tests", "golden", "dense64_table.npy"))
    env = gym_aqua.make(env_id, num_envs=n, seed=0, precision=precision, **kw)
    env.reset()
    n16 = (n + 15) // 16 * 16
    if mode == "tape":
        if env.continuous:
            tape = torch.rand(K, n16, 2, device="cuda") * 0.3 + 0.2
        else:
            tape = torch.randint(0, 3, (K, n16), dtype=torch.int32, device="cuda")
        rew = torch.empty(K, n16, dtype=env.dtype, device="cuda")
        flg = torch.empty(K, n16, dtype=torch.uint8, device="cuda")
        run = lambda: env.step_many(tape, rew, flg)
    else:
        run = lambda: env.rollout_random(K)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(b2b):
            run()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e-3 / b2b)
    st = env.stats()
    env.close()
    r = n * K / best
    return {"env": env_id + ("+" + obstacles if obstacles else ""), "mode": mode, "envs": n, "K": K, "v": v,
            "threads": threads, "precision": precision, "env_steps_per_s": r, "us_per_launch": best * 1e6,
            "frac65": r * 65 / 1e9 / PEAK, "episodes_per_step": st["episodes"] / max(st["steps"], 1)}


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    sizes = [65536, 131072, 524288, 1 << 20] if not quick else [65536, 1 << 20]
    for n in sizes:
        for K in ((32,) if quick else (16, 64)):
            for v, threads in ((1, 32), (1, 64), (1, 128), (2, 64), (2, 128), (4, 64), (4, 128)):
                print(json.dumps(rate("AquaEnv-v1", n, K, v, threads)), flush=True)
    for n, v, threads in ((524288, 2, 128), (524288, 4, 128), (524288, 1, 128), (1 << 22, 4, 128), (1 << 22, 2, 128)):
        print(json.dumps(rate("AquaEnv-v0", n, 32, v, threads, obstacles="dense64")), flush=True)
    for n, v, threads in ((131072, 1, 64), (131072, 2, 128), (1 << 20, 2, 128), (1 << 20, 4, 128)):
        print(json.dumps(rate("AquaContinuousEnv-v1", n, 32, v, threads)), flush=True)
    for n, v, threads in ((65536, 1, 64), (1 << 20, 2, 128), (1 << 20, 4, 128), (1 << 24, 4, 128), (1 << 24, 2, 128)):
        print(json.dumps(rate("AquaEnv-v1", n, 32, v, threads, mode="random")), flush=True)
    print(json.dumps(rate("AquaEnv-v1", 1 << 20, 32, 2, 128, precision="float64")), flush=True)
    print(json.dumps(rate("AquaEnv-v1", 1 << 24, 32, 4, 128)), flush=True)


if __name__ == "__main__":
    main()
