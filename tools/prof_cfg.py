"""Run a few forward+loglik launches of one model pair (target of ncu captures).  usage: prof_cfg.py {epm_dm|epm|dm} [B] [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import synthetic
from helpers import synth_plan
cfgs = {"epm_dm": ("exp_pist_flow", "dispersion", list(synthetic.PAR_NAMES_CFG3)),
        "epm": ("exp_pist_flow", False, ["tau1", "eta1", "J"]),
        "dm": ("dispersion", False, ["tau1", "D1", "J"])}
name = sys.argv[1] if len(sys.argv) > 1 else "epm_dm"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
m1, m2, pn = cfgs[name]
th7 = synthetic.theta_cfg3(B, 0)
cols = dict(zip(synthetic.PAR_NAMES_CFG3, th7.T)); cols["D1"] = cols["D2"]
plan, _, _ = synth_plan(m1, m2, pn)
theta = torch.from_numpy(np.ascontiguousarray(np.stack([cols[p] for p in pn], 1))).cuda()
logp = torch.empty(B, dtype=torch.float64, device="cuda")
for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
    plan.forward_loglik_dev(theta, pn, np.ones(7), np.ones(7) * 0.05, "normal", logp_t=logp)
torch.cuda.synchronize()
print("ok")
