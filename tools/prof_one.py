"""Run a few cfg-3 forward+loglik launches (target of ncu captures). Development aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import synthetic
from helpers import synth_plan
B = 65536
pn = list(synthetic.PAR_NAMES_CFG3)
plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
theta = torch.from_numpy(synthetic.theta_cfg3(B, 0)).cuda()
logp = torch.empty(B, dtype=torch.float64, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    plan.forward_loglik_dev(theta, pn, np.ones(7), np.ones(7) * 0.05, "normal", logp_t=logp)
torch.cuda.synchronize()
print("ok", float(torch.nansum(logp)))
