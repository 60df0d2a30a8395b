"""Steady-state cycles per tile-group for the experiment builds (NGRTD_LIB selects the library)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import synthetic
from helpers import synth_plan
cfgs = {"epm_dm": ("exp_pist_flow", "dispersion", list(synthetic.PAR_NAMES_CFG3)),
        "epm": ("exp_pist_flow", False, ["tau1", "eta1", "J"]),
        "dm": ("dispersion", False, ["tau1", "D1", "J"])}
for B in (2368 * 16 * 8, 65536):
  th7 = synthetic.theta_cfg3(B, 0)
  cols = dict(zip(synthetic.PAR_NAMES_CFG3, th7.T)); cols["D1"] = cols["D2"]
  for name, (m1, m2, pn) in cfgs.items():
      plan, _, _ = synth_plan(m1, m2, pn)
      theta = torch.from_numpy(np.ascontiguousarray(np.stack([cols[p] for p in pn], 1))).cuda()
      logp = torch.empty(B, dtype=torch.float64, device="cuda")
      for _ in range(3): plan.forward_loglik_dev(theta, pn, np.ones(7), np.ones(7) * .05, "normal", logp_t=logp)
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(10): plan.forward_loglik_dev(theta, pn, np.ones(7), np.ones(7) * .05, "normal", logp_t=logp)
      e1.record(); torch.cuda.synchronize()
      ms = e0.elapsed_time(e1) / 10
      tg = B / 8 * 210 / 592          # tile-groups per SMSP
      print("%s %-7s B=%d %.4f ms  %.1f cycles per tile-group per SMSP  %.2f TF(8col)" % (os.environ.get("NGRTD_LIB", "default")[-9:-3], name, B, ms, ms * 1e-3 * 1.92e9 / tg, 2.0*840*8*(2 if m2 else 1)*B/ms/1e9), flush=True)

from noblegas_rtd_mcmc_b200.sampler import Sampler, prior
B = 65536
pn = list(synthetic.PAR_NAMES_CFG3)
plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
truth = np.array([[180.0, 1500.0, 0.6, 0.4, 1.8, 0.4, synthetic.LOG10_J_MONTHLY]])
obs = plan.forward_host(truth, pn)[0]; sd = 0.05 * np.abs(obs)
pri = [prior("uniform", "tau1", 12, 12000), prior("beta", "nu_", 2.0, 0.1), prior("normal", "J", synthetic.LOG10_J_MONTHLY, 0.33),
       prior("uniform", "tau2", 600, 180000), prior("uniform", "f1", 0.01, 0.99), prior("uniform", "eta1", 1, 5),
       prior("uniform", "D2", 0.01, 2.0)]
q0 = [-3.0, 2.0, synthetic.LOG10_J_MONTHLY, -4.5, 0.3, -1.0, -1.2]
for lik in ("normal", "studentt"):
    smp = Sampler(pri, obs, sd, B, plan=plan, lik=lik, nu_range=(5.0, 30.0), f2_from_f1=True, tune_interval=100, hist_cap=256, seed=1, q0=q0, scaling=0.01)
    smp.run(20, tune=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); smp.run(200, tune=True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    print("sampler %-8s %.4f ms/step  %.3e evals/s (6 counted)  acc %.3f" % (lik, ms, B * 6 / ms * 1e3, float(smp.get("accepted").mean()) / smp.info()["step"]), flush=True)
    smp.close()
