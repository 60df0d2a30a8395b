"""Time the cfg-3 forward kernel of one library build (NGRTD_LIB=<path>): 65,536-chain launch, steady state (303,104 chains),
dispersion-only and exponential-class-only loops; prints worst parity error against the golden vectors.  Development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import synthetic
from helpers import GOLD, rel_err, synth_plan
tag = os.path.basename(os.environ.get("NGRTD_LIB", "libngrtd.so"))
cfgs = {"epm_dm": ("exp_pist_flow", "dispersion", list(synthetic.PAR_NAMES_CFG3)),
        "epm": ("exp_pist_flow", False, ["tau1", "eta1", "J"]), "dm": ("dispersion", False, ["tau1", "D1", "J"])}
out = []
for name, (m1, m2, pn) in cfgs.items():
    plan, _, _ = synth_plan(m1, m2, pn)
    for B in (65536, 303104):
        th7 = synthetic.theta_cfg3(B, 0)
        cols = dict(zip(synthetic.PAR_NAMES_CFG3, th7.T)); cols["D1"] = cols["D2"]
        theta = torch.from_numpy(np.ascontiguousarray(np.stack([cols[p] for p in pn], 1))).cuda()
        logp = torch.empty(B, dtype=torch.float64, device="cuda")
        obs = np.ones(7); sd = np.ones(7) * 0.05
        for _ in range(5):
            plan.forward_loglik_dev(theta, pn, obs, sd, "normal", logp_t=logp)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                plan.forward_loglik_dev(theta, pn, obs, sd, "normal", logp_t=logp)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 20)
        out.append("%s B=%d %.4f ms" % (name, B, best))
z = np.load(os.path.join(GOLD, "forward_synth.npz"))
pn = list(synthetic.PAR_NAMES_CFG3)
plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
errs = []
for name in ("cfg3", "cfg3i"):
    o = plan.forward_host(z[name + "/theta"], pn)
    errs.append(max(rel_err(o[:, i], z[name + "/" + t]) for i, t in enumerate(synthetic.TRACERS_CFG3)))
print("%-22s %s | parity %.2e" % (tag, " | ".join(out), max(errs)), flush=True)
