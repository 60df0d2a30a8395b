"""Helper of test_sampler_gpu.py::test_age_sampler_table_variants_agree: a short cfg-3 sampler run in THIS process (the
library reads NGRTD_MCMC_TB11 once per process), final states printed as JSON."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), HERE, os.path.join(os.path.dirname(HERE), "oracle")):
    sys.path.insert(0, p)
import numpy as np

from helpers import synth_plan
from noblegas_rtd_mcmc_b200 import synthetic
from noblegas_rtd_mcmc_b200.sampler import Sampler, prior

pn = list(synthetic.PAR_NAMES_CFG3)
plan, series, tab = synth_plan("exp_pist_flow", "dispersion", pn)
truth = np.array([[180.0, 1500.0, 0.6, 0.4, 1.8, 0.4, synthetic.LOG10_J_MONTHLY]])
obs = plan.forward_host(truth, pn)[0]
sd = 0.05 * np.abs(obs)
pri = [prior("uniform", "tau1", 12, 12000), prior("beta", "nu_", 2.0, 0.1), prior("normal", "J", synthetic.LOG10_J_MONTHLY, 0.33),
       prior("uniform", "tau2", 600, 180000), prior("uniform", "f1", 0.01, 0.99), prior("uniform", "eta1", 1, 5),
       prior("uniform", "D2", 0.01, 2.0)]
q0 = [-3.0, 2.0, synthetic.LOG10_J_MONTHLY, -4.5, 0.3, -1.0, -1.2]
smp = Sampler(pri, obs, sd, 96, plan=plan, lik="studentt", nu_range=(5.0, 30.0), f2_from_f1=True, tune_interval=50,
              hist_cap=500, seed=77, q0=q0, scaling=0.01)
smp.run(120, tune=True)
print(json.dumps({"q": smp.get("q").cpu().numpy().tolist(), "logp": smp.get("logp").cpu().numpy().tolist(),
                  "accepted": smp.get("accepted").cpu().numpy().tolist()}))
