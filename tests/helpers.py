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
