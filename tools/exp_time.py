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
