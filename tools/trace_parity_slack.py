"""Slack of the joint-inversion comparisons of tests/test_reference_traces.py (needs a GPU); development aid."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests"); sys.path.insert(0, ROOT + "/oracle")
import numpy as np
import test_reference_traces as T
for model, extra in (("exponential", 0.01), ("exp_pist_flow", 0.03)):
    for well in ("PLM1", "PLM6", "PLM7"):
        fx, t, names, tr = T._joint_run(model, well)
        if model == "exp_pist_flow":
            ie = names.index("eta1"); re_ = t["vars"]["eta1"]
            m = np.abs(tr[:, :, ie].mean(axis=0) - re_["mean"]) < 4.0 * re_["sd"] + 0.1
            print(model, well, "chains in the reference's mode: %d of %d" % (m.sum(), m.size))
            tr = tr[:, m, :]
        w = T._compare(fx, t, names, tr, skip=("nu_",), extra=extra)
        print(model, well, "worst (|F-p| - 4.5 sigma), allowed < %.2f:" % extra, {k: round(v, 4) for k, v in w.items()})
