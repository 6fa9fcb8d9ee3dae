#!/usr/bin/env python
"""Time the build variants of libaquagym.so (make -C aquaticgymenv_b200/csrc variant NAME=.. FLAGS=..)
on ONE box, one subprocess per variant and workload (tools/config_rate.py: single-step launches,
actions resident).  Boxes differ by a few per cent, so variants are only comparable within one
call.  usage (GPU box): python tools/tune.py"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKLOADS = [("AquaEnv-v1", 1 << 24, "float32", "-", 200), ("AquaEnv-v1", 1 << 20, "float32", "-", 400),
             ("AquaEnv-v0", 1 << 22, "float32", "dense64", 100), ("AquaContinuousEnv-v1", 1 << 24, "float32", "-", 100),
             ("AquaEnv-v1", 1 << 23, "float64", "-", 100)]
if os.environ.get("TUNE_QUICK"):  # the two default-map float32 sizes only
    WORKLOADS = WORKLOADS[:2]
libs = sorted(glob.glob(os.path.join(ROOT, "aquaticgymenv_b200", "variants", "*.so")))
for rep in range(2):  # twice, interleaved: shows the run-to-run spread on this box
    for lib in libs:
        env = dict(os.environ, AQUAGYM_LIB=lib, AQUAGYM_LIB_ANY_ABI="1")
        row = []
        for env_id, n, prec, tab, steps in WORKLOADS:
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "config_rate.py"), env_id, str(n), prec,
                                  tab, str(steps)], env=env, capture_output=True, text=True)
            try:
                j = json.loads(out.stdout.strip().splitlines()[-1])
                row.append("%.4g" % j["env_steps_per_s"])
            except Exception:
                row.append("FAILED " + out.stderr[-300:].replace("\n", " | "))
        names = ("v1@2^24", "v1@2^20", "dense64@2^22", "cont@2^24", "f64@2^23")
        print("%-32s " % os.path.basename(lib) + " | ".join("%s %s" % (a, b) for a, b in zip(names, row)), flush=True)
