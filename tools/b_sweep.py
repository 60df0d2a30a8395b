import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import synthetic
from helpers import synth_plan
pn = list(synthetic.PAR_NAMES_CFG3)
plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
for B in (592 * 16 * 1, 592 * 16 * 2, 592 * 16 * 4, 592 * 16 * 6, 65536, 592 * 16 * 7, 592 * 16 * 8, 592 * 16 * 14, 592 * 16 * 28):
    th = torch.from_numpy(synthetic.theta_cfg3(B, 0)).cuda()
    lp = torch.empty(B, dtype=torch.float64, device="cuda")
    for _ in range(3): plan.forward_loglik_dev(th, pn, np.ones(7), np.ones(7) * .05, "normal", logp_t=lp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): plan.forward_loglik_dev(th, pn, np.ones(7), np.ones(7) * .05, "normal", logp_t=lp)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("B=%7d (%.2f units/SMSP)  %.4f ms  %.3f ns/chain  %.2f TF(8col)" % (B, B / 16 / 592, ms, ms * 1e6 / B, 2.0 * 840 * 8 * 2 * B / ms / 1e9), flush=True)
