"""Throughput of the noble-gas CE sampler kernel (k_mcmc_ng, config 1) against the number of chains; development aid."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import distributed as ngdist
from noblegas_rtd_mcmc_b200.noble_gas_mcmc import mcmc_model
from noblegas_rtd_mcmc_b200.sampler import Sampler
fx = json.load(open(os.path.join(ROOT, "tests", "golden", "ng_posterior.json")))["wells"]["PLM1"]
mdl = mcmc_model(fx["obs"], mcmc_model.well_elev["PLM1"])
for NGC in [int(a) for a in sys.argv[1:]] or [32768, 65536, 131072, 262144]:
    ngs = Sampler(mdl.build_priors(), mdl.obs_mu, mdl.obs_sd, NGC, plan=None, gases=mdl.gases, lik="studentt",
                  nu_range=(1.0, 30.0), tune_interval=5000, hist_cap=2048, seed=123423)
    ngs.run(100, tune=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    ngs.run(10000, tune=True)
    ngs.stop_tuning()
    ngs.run(5000, tune=False, record=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    summ = ngdist.global_summary(5000, ngs.get("mean"), ngs.get("m2"))
    print("chains %7d: %.3f s for 15,000 steps = %.3e chain-steps/s; min ESS %.3e -> %.3e ESS/s; max r_hat %.4f" % (
        NGC, dt, NGC * 15000 / dt, np.min(summ["ess"]), np.min(summ["ess"]) / dt, np.max(summ["r_hat"])), flush=True)
    ngs.close()
