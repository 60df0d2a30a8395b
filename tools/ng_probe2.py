import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import distributed as ngdist
from noblegas_rtd_mcmc_b200.noble_gas_mcmc import mcmc_model
from noblegas_rtd_mcmc_b200.sampler import Sampler
fx = json.load(open(os.path.join(ROOT, "tests", "golden", "ng_posterior.json")))["wells"]["PLM1"]
mdl = mcmc_model(fx["obs"], mcmc_model.well_elev["PLM1"])
def T():
    torch.cuda.synchronize(); return time.perf_counter()
warm = Sampler(mdl.build_priors(), mdl.obs_mu, mdl.obs_sd, 64, plan=None, gases=mdl.gases, lik="studentt",
               nu_range=(1.0, 30.0), tune_interval=5000, hist_cap=8, seed=1)
warm.run(4, tune=True); warm.close()
for rep in range(2):
    t0 = T()
    ngs = Sampler(mdl.build_priors(), mdl.obs_mu, mdl.obs_sd, int(sys.argv[1]) if len(sys.argv) > 1 else 65536, plan=None, gases=mdl.gases, lik="studentt",
                  nu_range=(1.0, 30.0), tune_interval=5000, hist_cap=2048, seed=123423)
    t1 = T(); ngs.run(10000, tune=True); t2 = T(); ngs.stop_tuning(); t3 = T(); ngs.run(5000, tune=False, record=True); t4 = T()
    summ = ngdist.global_summary(5000, ngs.get("mean"), ngs.get("m2")); t5 = T()
    print("create %.3f  tune-run %.3f  stop_tuning %.3f  draw-run %.3f  summary %.3f" % (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4), flush=True)
    ngs.close()
