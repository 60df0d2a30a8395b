"""A few fused sampler launches on the cfg-3 workload (target of ncu captures). Development aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import synthetic
from noblegas_rtd_mcmc_b200.sampler import Sampler, prior
from helpers import synth_plan
B = 65536
pn = list(synthetic.PAR_NAMES_CFG3)
plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
truth = np.array([[180.0, 1500.0, 0.6, 0.4, 1.8, 0.4, synthetic.LOG10_J_MONTHLY]])
obs = plan.forward_host(truth, pn)[0]; sd = 0.05 * np.abs(obs)
pri = [prior("uniform", "tau1", 12, 12000), prior("beta", "nu_", 2.0, 0.1), prior("normal", "J", synthetic.LOG10_J_MONTHLY, 0.33),
       prior("uniform", "tau2", 600, 180000), prior("uniform", "f1", 0.01, 0.99), prior("uniform", "eta1", 1, 5),
       prior("uniform", "D2", 0.01, 2.0)]
q0 = [-3.0, 2.0, synthetic.LOG10_J_MONTHLY, -4.5, 0.3, -1.0, -1.2]
smp = Sampler(pri, obs, sd, B, plan=plan, lik="studentt", nu_range=(5.0, 30.0), f2_from_f1=True, tune_interval=100, hist_cap=64, seed=1, q0=q0, scaling=0.01)
for _ in range(3):
    smp.run(int(sys.argv[1]) if len(sys.argv) > 1 else 4, tune=True)
torch.cuda.synchronize()
print("ok", float(smp.get("accepted").mean()))
