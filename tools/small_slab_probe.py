#!/usr/bin/env python
"""Small slabs (BASELINE configs 1-3 at their literal sizes): how a step's time splits into kernel
duration and launch gap.  Times back-to-back steps eagerly and as one CUDA graph of 100 steps; run
it under `ncu --metrics gpu__time_duration.sum` for the kernel durations.
usage (GPU box): python tools/small_slab_probe.py [envs ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aquaticgymenv_b200 import gym_aqua  # noqa: E402

geometry = os.environ.get("AQUA_PROBE_GEOMETRY", "auto")   # auto | wide | narrow
sizes = [int(a) for a in sys.argv[1:]] or [65536, 131072, 1 << 20]
for n in sizes:
    env = gym_aqua.make("AquaEnv-v1", num_envs=n, seed=0, geometry=geometry)
    env.reset()
    pool = [env.random_actions(k).clone() for k in range(4)]
    for k in range(20):
        env.step_flags(pool[k % 4])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(400):
        env.step_flags(pool[k % 4])
    b.record()
    torch.cuda.synchronize()
    eager_us = a.elapsed_time(b) * 1e3 / 400
    # the same launches through the bare C ABI with prebuilt arguments: what the Python wrapper adds
    import ctypes as C
    import time
    from aquaticgymenv_b200 import _native as N
    args = [(env._h, env._stream(), C.c_void_p(pool[k].data_ptr()), N.ACT_I32 if pool[k].dtype == torch.int32 else N.ACT_I64,
             C.c_void_p(env.reward.data_ptr()), C.c_void_p(env.flags.data_ptr()), None) for k in range(4)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(400):
        env.lib.aqua_step_ex(*args[k % 4])
    t_issue_raw = (time.perf_counter() - t0) * 1e6 / 400
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(400):
        env.step_flags(pool[k % 4])
    t_issue_wrapped = (time.perf_counter() - t0) * 1e6 / 400
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        env.step_flags(pool[0])
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        for k in range(100):
            env.step_flags(pool[k % 4])
    g.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    graph_us = a.elapsed_time(b) * 1e3 / 1000
    print(json.dumps({"envs": n, "geometry": geometry, "eager_us_per_step": eager_us, "host_us_per_call_raw_abi": t_issue_raw,
                      "host_us_per_call_python_wrapper": t_issue_wrapped, "graph_us_per_step": graph_us,
                      "graph_env_steps_per_s": n / graph_us * 1e6}), flush=True)
    env.close()
