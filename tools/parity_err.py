"""Worst relative error of the forward kernel vs the reference golden vectors (synthetic set); development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import GOLD, MODEL_CFGS, rel_err, synth_plan
from noblegas_rtd_mcmc_b200 import synthetic
z = np.load(os.path.join(GOLD, "forward_synth.npz"))
for name in ("cfg3", "cfg3i", "epm_dm", "dm", "emm0"):
    if name.startswith("cfg3"):
        m1, m2, pn = "exp_pist_flow", "dispersion", list(synthetic.PAR_NAMES_CFG3)
    else:
        m1, m2, pn = MODEL_CFGS[name]
    plan, _, _ = synth_plan(m1, m2, pn)
    out = plan.forward_host(z[name + "/theta"], pn)
    print(name, "max rel err %.2e" % max(rel_err(out[:, i], z[name + "/" + t]) for i, t in enumerate(synthetic.TRACERS_CFG3)))
