"""Print our config-1 posterior summaries next to the reference's ng_optPLM*.csv (development aid)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from noblegas_rtd_mcmc_b200 import diagnostics as D
from noblegas_rtd_mcmc_b200.noble_gas_mcmc import mcmc_model
fx = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/ng_posterior.json")))
for well in ("PLM1", "PLM7", "PLM6"):
    w = fx["wells"][well]
    mdl = mcmc_model(w["obs"], mcmc_model.well_elev[well])
    res = mdl.sample(chains=256, tune=10000, draws=int(sys.argv[1]) if len(sys.argv) > 1 else 6000, tune_interval=5000, hist_cap=66000, thin=4)
    post = res["posterior"]
    print(well, "accept", res["sample_stats"]["accept_rate"].mean())
    for var in ("T", "E", "Ae", "F", "m", "b", "nu", "Ae_beta", "F_beta"):
        r = w["summary"][var]; a = post[var]
        lo, hi = D.hdi(a)
        print("  %-8s mean %.6g / %.6g  sd %.4g / %.4g  median %.6g / %.6g  hdi [%.5g, %.5g] / [%.5g, %.5g]  ref mcse %.3g ess %.0f rhat(ours,64ch) %.4f" % (
            var, a.mean(), r["mean"], a.std(), r["sd"], np.median(a), r["median"], lo, hi, r["hdi_3%"], r["hdi_97%"], r["mcse_mean"], r["ess_bulk"], D.rhat(a[:64])))
