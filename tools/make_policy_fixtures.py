#!/usr/bin/env python
"""tests/golden/policy_{no_obs,with_obs}.npz <- the weights of the two policies the reference
ships (example_policies/*/models/*, TF SavedModels), read with aquaticgymenv_b200.savedmodel.
Build container only (needs /root/reference); the .npz fixtures are committed."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("_savedmodel", os.path.join(ROOT, "aquaticgymenv_b200", "savedmodel.py"))
sm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sm)

ref = os.environ.get("AQUA_REFERENCE", "/root/reference")
for name, model in (("no_obs", "example_no_obs/models/model-00030"), ("with_obs", "example_with_obs/models/model-00032")):
    (w0, b0), (w1, b1), (w2, b2) = sm.load_weights(os.path.join(ref, "example_policies", model))
    assert w0.shape == (5, 64) and w1.shape == (64, 64) and w2.shape == (64, 3), (w0.shape, w1.shape, w2.shape)
    out = os.path.join(ROOT, "tests", "golden", "policy_%s.npz" % name)
    np.savez_compressed(out, w0=w0, b0=b0, w1=w1, b1=b1, w2=w2, b2=b2, source=model)
    print(out, w0.shape, w1.shape, w2.shape)
