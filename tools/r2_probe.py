"""r2 development probe: time the cfg-3 forward+loglik kernel of one library build at several batch sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import synthetic
from helpers import synth_plan
tag = os.path.basename(os.environ.get("NGRTD_LIB", "libngrtd.so")) + " tape_min=" + os.environ.get("NGRTD_TAPE_MIN", "-") + " pdl=" + os.environ.get("NGRTD_PDL", "-")
pn = list(synthetic.PAR_NAMES_CFG3)
plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
out = []
sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "65536,70000,131072,303104".split(","))]
for B in sizes:
    NB = 8
    thetas = [torch.from_numpy(synthetic.theta_cfg3(B, i)).cuda() for i in range(NB)]
    logp = torch.empty(B, dtype=torch.float64, device="cuda")
    obs = np.ones(7); sd = np.ones(7) * 0.05
    for i in range(10):
        plan.forward_loglik_dev(thetas[i % NB], pn, obs, sd, "normal", logp_t=logp)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(40):
            plan.forward_loglik_dev(thetas[i % NB], pn, obs, sd, "normal", logp_t=logp)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 40)
    out.append("B=%d %.4f ms (%.1f ns/unit-SMSP)" % (B, best, best * 1e6 / (B / 16 / 592)))
print("%-44s %s" % (tag, " | ".join(out)), flush=True)
