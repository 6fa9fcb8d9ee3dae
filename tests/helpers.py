"""Shared parity-test helpers.

Tolerance policy (BASELINE.json north_star: "within
1e-5 relative"): rewards and poses cross zero, so "relative" is taken against the
scale of the quantity:  |got - ref| <= rel * max(|ref|, scale)  with scale = 100
for x, y and the goal, 2*pi for the heading (compared on the circle: -pi and +pi
are the same heading), 0.05 for the wave speed and 10 for the reward.
"""
import glob
import os

import numpy as np

SCALE_POS, SCALE_ANGLE, SCALE_WAVE, SCALE_REWARD = 100.0, 2.0 * np.pi, 0.05, 10.0


def load_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_names(golden_dir, prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(golden_dir, prefix + "*.npz")))


def err_scaled(got, ref, scale):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return np.abs(got - ref) / np.maximum(np.abs(ref), scale)


def err_angle(got, ref):
    d = np.asarray(got, np.float64) - np.asarray(ref, np.float64)
    d = (d + np.pi) % (2.0 * np.pi) - np.pi
    return np.abs(d) / SCALE_ANGLE


def table_to_reference_list(tab):
    out = []
    for cx, cy, typ, a, b in np.asarray(tab, np.float64).reshape(-1, 5):
        out.append((np.array([cx, cy]), "c", a) if typ == 0.0 else (np.array([cx, cy]), "r", (a, b)))
    return out


def reference_main():
    """The reference's own caller code (`main/impl`, `main/testing`) made importable next to the
    drop-in env: its root is baseline/_ref (shipped to the GPU box by __graft_entry__.build()) or
    /root/reference.  `gym` resolves to the stand-in of oracle/gym_standin with the DROP-IN's six
    ids registered, `gym_aqua` (imported by main/testing/__init__.py:2) to the drop-in package,
    and `tensorflow` (imported by main/impl/dqn.py:9 at module level, absent here) to an empty
    stub.  Returns (gym, root) or skips the calling test when no reference copy is around."""
    import sys
    import types

    import pytest
    root = None
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.path.join(here, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "main", "testing")):
            root = cand
            break
    if root is None:
        pytest.skip("no copy of the reference's main/ package (baseline/_ref is made by __graft_entry__.build())")
    for p in (os.path.join(root, "main"), os.path.join(here, "oracle", "gym_standin")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gym
    from aquaticgymenv_b200 import gym_aqua as dropin
    sys.modules["gym_aqua"] = dropin
    dropin._register_with_gym()
    if "impl.dqn" not in sys.modules:  # import it once under the stub, then take the stub away again
        stub = "tensorflow" not in sys.modules
        if stub:
            tf = types.ModuleType("tensorflow")
            tf.config = types.SimpleNamespace(list_physical_devices=lambda kind=None: [],
                                              experimental=types.SimpleNamespace(set_memory_growth=lambda *a: None))
            tf.keras = types.SimpleNamespace()
            sys.modules["tensorflow"] = tf
        try:
            import impl.dqn  # noqa: F401
            import impl.utils  # noqa: F401
        finally:
            if stub:
                del sys.modules["tensorflow"]
    return gym, root


def wide_actions(g, t0, t1):
    """Rows t0..t1-1 of the golden's action tape: stored for the discrete envs, regenerated from the
    ACTION stream (oracle draws, pinned in test_philox.py) for the continuous ones."""
    if not bool(g["continuous"]):
        return np.ascontiguousarray(g["actions"][t0:t1]).astype(np.int32)
    import ctypes as C
    from oracle import aqua_oracle as ao
    L = ao.lib()
    n = g["flags"].shape[1]
    out = np.zeros((t1 - t0, n, 2), np.float32)
    buf = (C.c_float * 2)()
    seed, base = int(g["seed"]), int(g["env_id_base"])
    for t in range(t0, t1):
        for i in range(n):
            L.aqo_draw_action_continuous(seed, base + i, t, buf)
            out[t - t0, i] = (buf[0], buf[1])
    return out
