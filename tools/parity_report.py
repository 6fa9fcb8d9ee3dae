#!/usr/bin/env python
"""Measured parity of the float32 (throughput) and float64 instantiations against the C oracle:
flag-flip rate, decision margins of the flips, and drift, over lock-step random rollouts with
teacher forcing after a flip (SURVEY.md section 7.3 H3 iv), for both kernel families.  Prints one
JSON object.
usage (GPU box): python tools/parity_report.py [envs] [steps]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aquaticgymenv_b200 import AquaVecEnv  # noqa: E402
from oracle import aqua_oracle as ao  # noqa: E402
from tests.helpers import SCALE_POS, SCALE_REWARD, err_angle, err_scaled  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
# optional third argument "soak": only the headline table (AquaEnv-v1) and the streaming kernel, both
# precisions -- for large slabs (e.g. 1048576 1000 soak: BASELINE's 1M envs, 1,000-step rollouts)
soak = len(sys.argv) > 3 and sys.argv[3] == "soak"
threads = len(os.sched_getaffinity(0))
out = {"envs": n, "steps": steps, "configs": []}
tables = [("AquaEnv-v1", False, "v1"), ("AquaContinuousEnv-v2", True, "v2"), ("AquaEnv dense64", False, "dense64")]
for name, cont, tab in (tables[:1] if soak else tables):
    table = np.load(os.path.join(ROOT, "tests", "golden", tab + "_table.npy"))
    # "wide": whole warps run step_stream_kernel (what large slabs run); "narrow": step_kernel, V = 1
    runs = [("float32", "wide"), ("float64", "wide"), ("float32", "narrow"), ("float64", "narrow")]
    for precision, geometry in (runs[:2] if soak else runs):
        env = AquaVecEnv(n, continuous=cont, obstacles=table, seed=11, precision=precision, geometry=geometry,
                         device="cuda:0")
        ref = ao.OracleVecEnv(n, continuous=cont, obstacles=table, seed=11)
        env.reset()
        ref.reset()
        flips, flip_margins, worst_pos, worst_ang, worst_rew = 0, [], 0.0, 0.0, 0.0
        for t in range(steps):
            act = env.random_actions(t)
            flags = env.step_flags(act).cpu().numpy()
            reward = env.reward.cpu().numpy().astype(np.float64)
            _, r_ref, f_ref, margin = ref.step(act.cpu().numpy(), autoreset=True, want_margin=True, nthreads=threads)
            differ = flags != f_ref
            ok = ~differ
            got, want = env.get_state().cpu().numpy(), ref.get_state()
            worst_pos = max(worst_pos, float(np.abs(got[:2, ok] - want[:2, ok]).max()))
            worst_ang = max(worst_ang, float(err_angle(got[2, ok], want[2, ok]).max() * 2 * np.pi))
            worst_rew = max(worst_rew, float(np.abs(reward[ok] - r_ref[ok]).max()))
            if differ.any():
                flips += int(differ.sum())
                flip_margins.extend(margin[differ].tolist())
                env.set_state(torch.as_tensor(want))
        out["configs"].append({
            "env": name, "precision": precision, "geometry": geometry,
            "kernel": "step_stream_kernel" if geometry == "wide" else "step_kernel (V = 1)",
            "env_steps": n * steps, "flag_flips": flips,
            "flips_per_env_step": flips / (n * steps),
            "max_margin_of_a_flip_px": max(flip_margins) if flip_margins else None,
            "max_abs_position_error_px": worst_pos, "max_abs_heading_error_rad": worst_ang,
            "max_abs_reward_error": worst_rew})
        env.close()
print(json.dumps(out, indent=1))
