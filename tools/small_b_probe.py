import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import MODEL_CFGS, real_plan
from noblegas_rtd_mcmc_b200.sampler import Sampler, prior
m1, m2, pn = MODEL_CFGS["epm123"]
tracers = ["CFC12", "SF6", "H3", "He4_ter"]
plan, _ = real_plan(m1, m2, pn, tracers)
obs = np.array([36.38, 1.245, 4.869, 8.8e-9]); sd = 0.05 * obs
pri = [prior("uniform", "tau1", 1, 1000), prior("beta", "nu_", 2.0, 0.1), prior("normal", "J", -10.42, 0.33), prior("uniform", "eta1", 1, 5),
       prior("beta", "thalf_cfc", 2, 2, 5, 35), prior("halfnormal", "lamsf6", 0.5 / 3)]
for B in (3, 16, 256, 4096, 65536):
    for lik in ("studentt", "normal"):
        smp = Sampler(pri, obs, sd, B, plan=plan, lik=lik, nu_range=(5.0, 30.0), tune_interval=1000, hist_cap=4096, seed=1)
        smp.run(200, tune=True); torch.cuda.synchronize()
        t0 = time.perf_counter(); smp.run(2000, tune=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("B=%6d %-8s %.1f us/step  acc %.3f" % (B, lik, dt / 2000 * 1e6, float(smp.get("accepted").mean()) / 2200), flush=True)
        smp.close()
th = torch.from_numpy(np.array([[42.0, 1.7, -10.3, 20.0, 0.05]] * 3)).cuda()
o = plan.forward_dev(th, pn); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): plan.forward_dev(th, pn, o)
torch.cuda.synchronize(); print("forward_dev B=3: %.1f us/call" % ((time.perf_counter() - t0) / 200 * 1e6))
t0 = time.perf_counter()
for _ in range(200): plan.forward_host(th.cpu().numpy(), pn)
print("forward_host B=3: %.1f us/call" % ((time.perf_counter() - t0) / 200 * 1e6))
