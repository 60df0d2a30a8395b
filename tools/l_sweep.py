"""Per-lag-group cost vs per-unit (prologue + epilogue) cost of k_forward: time at several L for fixed B (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import synthetic
from helpers import synth_plan
cfgs = {"epm_dm": ("exp_pist_flow", "dispersion", list(synthetic.PAR_NAMES_CFG3)),
        "epm": ("exp_pist_flow", False, ["tau1", "eta1", "J"]),
        "dm": ("dispersion", False, ["tau1", "D1", "J"])}
for B in (65536, 2368 * 16 * 8):
    th7 = synthetic.theta_cfg3(B, 0)
    cols = dict(zip(synthetic.PAR_NAMES_CFG3, th7.T)); cols["D1"] = cols["D2"]
    for name, (m1, m2, pn) in cfgs.items():
        res = []
        for L in (8, 212, 420, 840, 1024):
            plan, _, _ = synth_plan(m1, m2, pn, L=L)
            theta = torch.from_numpy(np.ascontiguousarray(np.stack([cols[p] for p in pn], 1))).cuda()
            logp = torch.empty(B, dtype=torch.float64, device="cuda")
            for _ in range(3): plan.forward_loglik_dev(theta, pn, np.ones(7), np.ones(7) * .05, "normal", logp_t=logp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): plan.forward_loglik_dev(theta, pn, np.ones(7), np.ones(7) * .05, "normal", logp_t=logp)
            e1.record(); torch.cuda.synchronize()
            res.append((L, e0.elapsed_time(e1) / 10))
        tiles = B / 8 / 592
        (l0, t0), (l1, t1) = res[1], res[3]
        slope = (t1 - t0) / ((l1 - l0) / 4)          # ms per lag-group (all tiles)
        icpt = t1 - slope * l1 / 4
        print("%-7s B=%d " % (name, B) + " ".join("L=%d: %.4f" % r for r in res) +
              "  | per group %.1f cyc, per tile overhead %.0f cyc (%.0f%% at L=840)" %
              (slope * 1e-3 * 1.92e9 / tiles, icpt * 1e-3 * 1.92e9 / tiles, 100 * icpt / t1), flush=True)
