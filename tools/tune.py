#!/usr/bin/env python
"""Time the build variants of libaquagym.so (make -C aquaticgymenv_b200/csrc variant NAME=.. FLAGS=..)
on ONE box, one subprocess per variant and workload: the bench headline (AquaEnv-v1, default
obstacles) and the 64-obstacle stress map.  Boxes differ by a few per cent, so variants are only
comparable within one call.  usage (GPU box): python tools/tune.py [envs] [steps]"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
envs = sys.argv[1] if len(sys.argv) > 1 else str(1 << 24)
steps = sys.argv[2] if len(sys.argv) > 2 else "200"
libs = sorted(glob.glob(os.path.join(ROOT, "aquaticgymenv_b200", "variants", "*.so")))
for lib in libs:
    env = dict(os.environ, AQUAGYM_LIB=lib)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-extras", "--e2e-steps", "0",
                          "--steps", steps, "--warmup", "20", "--envs-per-gpu", envs],
                         env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        print("%-34s %.4g env-steps/s  roofline %.3f  %.4f ms/step" % (
            os.path.basename(lib), j["value"], j["roofline"]["frac"], j["ms_per_step"]), flush=True)
    except Exception:
        print(os.path.basename(lib), "FAILED", out.stdout[-300:], out.stderr[-600:], flush=True)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "config_rate.py"), "AquaEnv-v0", str(1 << 22),
                          "float32", "dense64", "100"], env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        print("%-34s %.4g env-steps/s  dense64 map, 2^22 envs, %.1f us/step" % (
            "", j["env_steps_per_s"], j["us_per_step"]), flush=True)
    except Exception:
        print(os.path.basename(lib), "dense64 FAILED", out.stdout[-300:], out.stderr[-600:], flush=True)
