"""Sweep (warps, NT, UA) variants of the cfg-3 forward kernel of a -DNGRTD_TUNE build (NGRTD_LIB=<path>): 65,536- and
303,104-chain launches, best of 5 x 40 back-to-back launches.  Development aid (r2 session 3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import synthetic
from helpers import synth_plan
pn = list(synthetic.PAR_NAMES_CFG3)
plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
variants = [(0, 0, 0), (8, 4, 1), (12, 2, 1), (12, 3, 1), (12, 4, 1), (8, 2, 1), (8, 2, 2), (16, 3, 1), (16, 2, 2), (0, 0, 0)]
obs = np.ones(7); sd = np.ones(7) * 0.05
for (w, nt, ua) in variants:
    os.environ["NGRTD_FWD_WARPS"] = str(w); os.environ["NGRTD_FWD_NT"] = str(nt); os.environ["NGRTD_FWD_UA"] = str(ua)
    out = []
    for B in (65536, 303104):
        NB = 8
        thetas = [torch.from_numpy(synthetic.theta_cfg3(B, i)).cuda() for i in range(NB)]
        logp = torch.empty(B, dtype=torch.float64, device="cuda")
        try:
            for i in range(10):
                plan.forward_loglik_dev(thetas[i % NB], pn, obs, sd, "normal", logp_t=logp)
            torch.cuda.synchronize()
        except Exception as e:
            out.append("B=%d failed: %s" % (B, str(e)[:60])); continue
        best = 1e9
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(40):
                plan.forward_loglik_dev(thetas[i % NB], pn, obs, sd, "normal", logp_t=logp)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 40)
        out.append("B=%d %.4f ms checksum %.10e" % (B, best, float(torch.nansum(logp))))
    print("warps=%-2d NT=%d UA=%d  %s" % (w, nt, ua, " | ".join(out)), flush=True)
