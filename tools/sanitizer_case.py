"""Small end-to-end case for compute-sanitizer (forward all classes, chunked path, CE, sampler both kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import json
import numpy as np
from helpers import MODEL_CFGS, real_plan, synth_plan, GOLD
from noblegas_rtd_mcmc_b200 import synthetic
from noblegas_rtd_mcmc_b200.noble_gas_mcmc import mcmc_model
from noblegas_rtd_mcmc_b200.sampler import Sampler, prior
rng = np.random.default_rng(0)
for name in ("epm_dm", "epm_pfm123", "dm_dm", "pfm"):
    m1, m2, pn = MODEL_CFGS[name]
    plan, _, _ = synth_plan(m1, m2, pn)
    z = np.load(os.path.join(GOLD, "forward_synth.npz"))
    out = plan.forward_host(z[name + "/theta"][:37], pn)
    lp = plan.forward_loglik_host(z[name + "/theta"][:37], pn, np.ones(7), np.ones(7), "studentt", nu=np.full(37, 7.0))
m1, m2, pn = MODEL_CFGS["epm123"]
plan, _ = real_plan(m1, m2, pn, ["CFC12", "SF6", "H3", "He4_ter"], L=2500)       # 3 chunks
z = np.load(os.path.join(GOLD, "forward_real.npz"))
plan.forward_host(z["epm123/theta"][:19], pn)
fx = json.load(open(os.path.join(GOLD, "ng_posterior.json")))["wells"]["PLM1"]
mdl = mcmc_model(fx["obs"], mcmc_model.well_elev["PLM1"])
s = Sampler(mdl.build_priors(), mdl.obs_mu, mdl.obs_sd, 37, plan=None, gases=mdl.gases, lik="studentt", nu_range=(1.0, 30.0), tune_interval=10, hist_cap=16, seed=1)
s.run(40, tune=True, record=True, keep_trace=True); s.stop_tuning(); s.run(10, record=True)
pn = list(synthetic.PAR_NAMES_CFG3)
plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
obs = plan.forward_host(np.array([[180.0, 1500.0, 0.6, 0.4, 1.8, 0.4, synthetic.LOG10_J_MONTHLY]]), pn)[0]
pri = [prior("uniform", "tau1", 12, 12000), prior("beta", "nu_", 2.0, 0.1), prior("normal", "J", synthetic.LOG10_J_MONTHLY, 0.33),
       prior("uniform", "tau2", 600, 180000), prior("uniform", "f1", 0.01, 0.99), prior("uniform", "eta1", 1, 5), prior("uniform", "D2", 0.01, 2.0)]
s = Sampler(pri, obs, 0.05 * np.abs(obs), 37, plan=plan, lik="studentt", nu_range=(5.0, 30.0), f2_from_f1=True, tune_interval=5, hist_cap=8, seed=1,
            q0=[-3.0, 2.0, synthetic.LOG10_J_MONTHLY, -4.5, 0.3, -1.0, -1.2], scaling=0.01)
s.run(12, tune=True, record=True, keep_trace=True); s.stop_tuning(); s.run(6, record=True)
m1, m2, pn = MODEL_CFGS["emm0"]
plan, _ = real_plan(m1, m2, pn, ["CFC12"], L=2500)
s = Sampler([prior("uniform", "tau1", 1.0, 1000.0)], [36.4], [3.0], 21, plan=plan, lik="normal", tune_interval=5, hist_cap=8, seed=2)
s.run(6, tune=True, record=True)
print("sanitizer case ok")
