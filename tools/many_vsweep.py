#!/usr/bin/env python
"""Which envs-per-thread instantiation of aqua_step_many wins at which slab size (K = 64, launches
back to back; one process, the geometry is chosen per launch through AQUA_MANY_V / AQUA_MANY_THREADS).
usage (GPU box): python tools/many_vsweep.py [env_id] [dense64|-]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aquaticgymenv_b200 import gym_aqua  # noqa: E402

env_id = sys.argv[1] if len(sys.argv) > 1 else "AquaEnv-v1"
kw = {}
if len(sys.argv) > 2 and sys.argv[2] == "dense64":
    kw["obstacles"] = np.load(os.path.join(ROOT, "tests", "golden", "dense64_table.npy"))
K = 64
print("# %s %s, aqua_step_many K = %d: env-steps/s by slab size (rows) and envs per thread / threads per CTA; "
      "'auto' = the library's own choice" % (env_id, sys.argv[2] if len(sys.argv) > 2 else "", K))
for n in (16384, 32768, 65536, 98304, 131072, 196608, 262144, 393216, 524288, 1048576, 2097152, 4194304):
    env = gym_aqua.make(env_id, num_envs=n, seed=0, **kw)
    env.reset()
    n16 = (n + 15) // 16 * 16
    if env.continuous:
        tape = torch.rand(K, n16, 2, device="cuda") * 0.3 + 0.2
    else:
        tape = torch.randint(0, 3, (K, n16), dtype=torch.int32, device="cuda")
    rew = torch.empty(K, n16, device="cuda")
    flg = torch.empty(K, n16, dtype=torch.uint8, device="cuda")
    row = []
    for v, thr in ((1, 64), (2, 64), (4, 64), (2, 128), (4, 128), (0, 0)):
        os.environ.pop("AQUA_MANY_V", None)
        os.environ.pop("AQUA_MANY_THREADS", None)
        if v:
            os.environ["AQUA_MANY_V"], os.environ["AQUA_MANY_THREADS"] = str(v), str(thr)
        for _ in range(3):
            env.step_many(tape, rew, flg)
        torch.cuda.synchronize()
        best = 1e9
        reps = 4 if n <= (1 << 20) else 2
        for _ in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(reps):
                env.step_many(tape, rew, flg)
            e.record()
            torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e) * 1e-3 / reps)
        row.append(("V%d/%d" % (v, thr) if v else "auto") + " %.3g" % (n * K / best))
    env.close()
    del tape, rew, flg
    print("%8d  " % n + "  ".join(row), flush=True)
