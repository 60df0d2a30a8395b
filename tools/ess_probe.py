import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import synthetic, diagnostics as D
from noblegas_rtd_mcmc_b200.sampler import Sampler, prior
from helpers import synth_plan
B = int(sys.argv[1]); tune = int(sys.argv[2]); draws = int(sys.argv[3])
pn = list(synthetic.PAR_NAMES_CFG3)
plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
truth = np.array([[180.0, 1500.0, 0.6, 0.4, 1.8, 0.4, synthetic.LOG10_J_MONTHLY]])
obs = plan.forward_host(truth, pn)[0]
pri = [prior("uniform", "tau1", 12, 12000), prior("beta", "nu_", 2.0, 0.1), prior("normal", "J", synthetic.LOG10_J_MONTHLY, 0.33),
       prior("uniform", "tau2", 600, 180000), prior("uniform", "f1", 0.01, 0.99), prior("uniform", "eta1", 1, 5), prior("uniform", "D2", 0.01, 2.0)]
for ti in (100, 500):
    smp = Sampler(pri, obs, 0.05 * np.abs(obs), B, plan=plan, lik="studentt", nu_range=(5.0, 30.0), f2_from_f1=True, tune_interval=ti,
                  hist_cap=min(tune + draws, 4096), seed=1, scaling=0.01, q0=[-3.0, 2.0, synthetic.LOG10_J_MONTHLY, -4.5, 0.3, -1.0, -1.2])
    t0 = time.perf_counter()
    smp.run(tune, tune=True); smp.stop_tuning(); smp.run(draws, record=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    s = D.moments_summary(float(draws), smp.get("mean").cpu().numpy(), smp.get("m2").cpu().numpy())
    print("tune_interval", ti, "secs %.2f" % dt, "acc %.3f" % (float(smp.get("accepted").mean()) / (tune + draws)), "lamb %.3f" % float(smp.get("lamb").mean()))
    print("  mean", np.round(s["mean"], 4)); print("  sd  ", np.round(s["sd"], 4)); print("  rhat", np.round(s["r_hat"], 3)); print("  ess ", np.round(s["ess"]))
    smp.close()
