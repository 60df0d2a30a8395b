"""Sweep (warps, NT, UA) variants of the forward kernel (needs a -DNGRTD_TUNE build). Development aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import synthetic
from helpers import synth_plan

B = 65536 * 4
cfgs = {"epm_dm": ("exp_pist_flow", "dispersion", list(synthetic.PAR_NAMES_CFG3)),
        "epm": ("exp_pist_flow", False, ["tau1", "eta1", "J"]),
        "dm": ("dispersion", False, ["tau1", "D1", "J"])}
th7 = synthetic.theta_cfg3(B, 0)
cols = dict(zip(synthetic.PAR_NAMES_CFG3, th7.T))
cols["D1"] = cols["D2"]
variants = [(0, 0, 0), (16, 1, 1), (24, 1, 1), (20, 1, 1), (24, 2, 1)]
which = sys.argv[1:] or list(cfgs)
for name in which:
    m1, m2, pn = cfgs[name]
    plan, _, _ = synth_plan(m1, m2, pn)
    theta = torch.from_numpy(np.ascontiguousarray(np.stack([cols[p] for p in pn], 1))).cuda()
    obs = np.ones(7); sd = np.ones(7) * 0.05
    logp = torch.empty(B, dtype=torch.float64, device="cuda")
    ref = None
    for (w, nt, ua) in variants:
        os.environ["NGRTD_FWD_WARPS"] = str(w); os.environ["NGRTD_FWD_NT"] = str(nt); os.environ["NGRTD_FWD_UA"] = str(ua)
        for _ in range(3):
            plan.forward_loglik_dev(theta, pn, obs, sd, "normal", logp_t=logp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            plan.forward_loglik_dev(theta, pn, obs, sd, "normal", logp_t=logp)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        ncomp = 2 if m2 else 1
        flops = 2.0 * 840 * 8 * ncomp * B
        cs = float(torch.nansum(logp))
        if ref is None: ref = cs
        print("%-7s W=%2d NT=%d UA=%d  %.4f ms  %.2f TFLOP/s(8col)  checksum_rel=%.1e" % (name, w, nt, ua, ms, flops / ms / 1e9, abs(cs - ref) / abs(ref)), flush=True)

