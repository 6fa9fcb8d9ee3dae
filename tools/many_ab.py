#!/usr/bin/env python
"""Same-box A/B of libaquagym.so builds (launches timed four at a time, back to back) (aquaticgymenv_b200/variants/*.so) on the aqua_step_many
workloads of BASELINE.json: one subprocess per (build, workload), the builds interleaved, twice.
usage (GPU box): python tools/many_ab.py [K]      worker: python tools/many_ab.py --one <workload index> K"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# (label, env id, envs, obstacles, V, threads); V = 0: the library's own geometry
WORK = [("c2 65536", "AquaEnv-v1", 65536, None, 0, 0), ("c3 131072 cont", "AquaContinuousEnv-v1", 131072, None, 0, 0),
        ("head 2^20", "AquaEnv-v1", 1 << 20, None, 0, 0), ("c4 dense 2^19", "AquaEnv-v0", 1 << 19, "dense64", 0, 0),
        ("c4 dense 2^22", "AquaEnv-v0", 1 << 22, "dense64", 0, 0), ("98304", "AquaEnv-v1", 98304, None, 0, 0),
        ("c2 65536 v1", "AquaEnv-v1", 65536, None, 1, 64)]

if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    from tools.many_probe import rate
    w = WORK[int(sys.argv[2])]
    print(json.dumps(rate(w[1], w[2], int(sys.argv[3]), w[4], w[5], obstacles=w[3], reps=5, b2b=4)))
    sys.exit(0)

K = sys.argv[1] if len(sys.argv) > 1 else "64"
libs = sorted(glob.glob(os.path.join(ROOT, "aquaticgymenv_b200", "variants", "*.so")))
print("# K = %s; env-steps/s per build, two rounds" % K)
for wi, w in enumerate(WORK):
    row = []
    for rep in range(2):
        for lib in libs:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(wi), K],
                                 env=dict(os.environ, AQUAGYM_LIB=lib, AQUAGYM_LIB_ANY_ABI="1"), capture_output=True, text=True)
            try:
                row.append("%s %.4g" % (os.path.basename(lib)[11:-3], json.loads(out.stdout.strip().splitlines()[-1])["env_steps_per_s"]))
            except Exception:
                row.append("%s FAILED %s" % (os.path.basename(lib)[11:-3], out.stderr[-200:].replace("\n", " | ")))
    print("%-20s %s" % (w[0], "  ".join(row)), flush=True)
