"""Forward + likelihood kernel on long lag axes (chunk-streamed tables, BASELINE config 5): ms per launch and algorithmic
FP64 rate for L = 840 (resident), 4,096, 10,000 and 25,256 lags.  Development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import synthetic
from helpers import synth_plan
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cfgs = {"epm_dm": ("exp_pist_flow", "dispersion", list(synthetic.PAR_NAMES_CFG3)), "dm": ("dispersion", False, ["tau1", "D1", "J"]),
        "epm": ("exp_pist_flow", False, ["tau1", "eta1", "J"])}
th7 = synthetic.theta_cfg3(B, 0)
cols = dict(zip(synthetic.PAR_NAMES_CFG3, th7.T)); cols["D1"] = cols["D2"]
for L in (840, 1024, 2048, 4096, 10000, 25256):
    for name, (m1, m2, pn) in cfgs.items():
        plan, _, _ = synth_plan(m1, m2, pn, L=L)
        theta = torch.from_numpy(np.ascontiguousarray(np.stack([cols[p] for p in pn], 1))).cuda()
        logp = torch.empty(B, dtype=torch.float64, device="cuda")
        obs = np.ones(7); sd = np.ones(7) * 0.05
        for _ in range(2):
            plan.forward_loglik_dev(theta, pn, obs, sd, "normal", logp_t=logp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            plan.forward_loglik_dev(theta, pn, obs, sd, "normal", logp_t=logp)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        ncomp = 2 if m2 else 1
        print("L=%6d %-7s B=%d  %8.3f ms  %.2f TFLOP/s over 8 MMA columns (%.3f of 36.45)" % (
            L, name, B, ms, 2.0 * L * 8 * ncomp * B / ms / 1e9, 2.0 * L * 8 * ncomp * B / ms / 1e9 / 36.45), flush=True)
