"""BASELINE config 1: the reference's noble-gas closed-equilibrium fit (ng_interp/noble_gas_mcmc.py) for one well with
its sampler settings (DEMetropolisZ(tune_interval=5000), tune 10,000 + 50,000 draws, 4 chains), printed next to the
posterior summary the reference ships (ng_interp/ng_opt<well>.csv).

    python examples/config1_noble_gas_fit.py [well] [chains]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from noblegas_rtd_mcmc_b200 import diagnostics
from noblegas_rtd_mcmc_b200.noble_gas_mcmc import mcmc_model


def main():
    well = sys.argv[1] if len(sys.argv) > 1 else "PLM1"
    chains = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "ng_posterior.json")))["wells"][well]
    mdl = mcmc_model(fx["obs"], mcmc_model.well_elev[well])
    t0 = time.perf_counter()
    res = mdl.sample(chains=chains, tune=10000, draws=50000, tune_interval=5000, random_seed=123423)
    dt = time.perf_counter() - t0
    summ = diagnostics.summary(res["posterior"])
    print("well %s  %d chains x 60,000 steps: %.2f s" % (well, chains, dt))
    print("%-8s %12s %12s %12s %12s %10s %8s" % ("", "mean", "ref mean", "sd", "ref sd", "ess_bulk", "r_hat"))
    for k in ("m", "Ae", "F", "E", "b", "T", "nu"):
        r, ref = summ[k], fx["summary"][k]
        print("%-8s %12.6g %12.6g %12.4g %12.4g %10.0f %8.4f" % (k, r["mean"], ref["mean"], r["sd"], ref["sd"], r["ess_bulk"], r["r_hat"]))


if __name__ == "__main__":
    main()
