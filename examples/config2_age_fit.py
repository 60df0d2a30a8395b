"""BASELINE config 2: the reference's age inversion (age_ens_runs_mcmc/run_age_mcmc.py) for one well on the real yearly
input series -- exp_pist_flow RTD, tracers CFC12 + SF6 + H3 + He4_ter, savenum [1,2,3] (variable J, CFC decay, SF6
contamination), DEMetropolisZ(tune_interval=1000), tune 10,000 + 10,000 draws, 3 chains, random_seed 123423.

The driver below mirrors run_age_mcmc.py:122-231 (okw / pkw / ckw construction) with our drop-in `conv_mcmc`.
Observations: ens_dict.pk is a missing blob of the reference, so the observation ensembles are rebuilt from what the
reference's own traces reveal: obs_mu = `observed_data/like` of conv_traces/<well>...123.netcdf (tests/golden/age_traces.json)
and obs_err = ens.std() + 0.05 obs_mu from tests/golden/age_obs_err.json (H3 a priori, the others estimated from the
single-tracer traces, oracle/fit_obs_err.py).  With them the run reproduces the reference's posterior (printed beside ours).

    python examples/config2_age_fit.py [well] [chains]      (keep chains <= ~4096: the full trace is downloaded and summarised on the host)

Reference wall time for the same run (pymc3 3.11.2, 3 CPU processes): 296-324 s (sampling_time attribute of
conv_traces/PLM*.CFC12.SF6.H3.He4_ter.exp_pist_flow.123.netcdf).
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pandas as pd

from helpers import load_c_in
from noblegas_rtd_mcmc_b200 import diagnostics
from noblegas_rtd_mcmc_b200 import noble_gas_utils as ng_utils
from noblegas_rtd_mcmc_b200.run_age_mcmc_utils import conv_mcmc



def main():
    ww = sys.argv[1] if len(sys.argv) > 1 else "PLM1"
    chains = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    mod_type1, mod_type2 = "exp_pist_flow", False
    tracers = ["CFC12", "SF6", "H3", "He4_ter"]
    savenum = [1, 2, 3]
    C = load_c_in()
    L = len(C["H3"])
    C_in_dict = {k: pd.DataFrame({k: v[::-1]}, index=np.arange(L - 1, -1, -1)) for k, v in C.items()}
    obs_perr = {"CFC12": 0.05, "SF6": 10.0 if ww == "PLM6" else 0.05, "H3": 0.05, "He4_ter": 0.05}      # run_age_mcmc.py:100-114
    import json
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "age_traces.json")))["traces"]["%s.%s.%s.123" % (ww, ".".join(tracers), mod_type1)]
    rel = json.load(open(os.path.join(ROOT, "tests", "golden", "age_obs_err.json")))["rel"]
    okw = {}
    for i, tt in enumerate(tracers):
        mu_t = ref["obs_mu"][i]
        std_t = max(rel[tt][ww] - obs_perr[tt], 0.0) * mu_t              # ens.std(): a two-point ensemble with this mean and spread
        okw[tt] = {"obs_df": np.array([mu_t - std_t, mu_t + std_t]), "obs_perr": obs_perr[tt]}
    pkw = {"tau1_low": 1.0, "tau1_high": 1000.0}
    par_names = ["tau1"]
    if mod_type1 == "exp_pist_flow":
        pkw["eta1_low"], pkw["eta1_high"] = 1.0, 5.0
        par_names += ["eta1"]
    if "He4_ter" in tracers and 1 in savenum:
        pkw["J_mu"] = np.log10(ng_utils.J_flux(Del=1., rho_r=2700, rho_w=1000, U=3.7, Th=10.2, phi=0.05))
        pkw["J_sd"] = 0.33
        par_names += ["J"]
    if "CFC12" in tracers and 2 in savenum:
        pkw["cfc_thalf_lo"], pkw["cfc_thalf_hi"] = 5.0, 35.0
        par_names += ["thalf_cfc"]
    if "SF6" in tracers and 3 in savenum:
        par_names += ["lamsf6"]
    pkw["par_names"] = par_names
    ckw = {"mod_type1": mod_type1, "mod_type2": mod_type2}
    for tt in tracers:
        if tt in ("CFC11", "CFC12", "CFC113", "SF6"):
            ckw[tt] = {"C_t": C_in_dict[tt]}
        elif tt == "He4_ter":
            ckw[tt] = {"C_t": C_in_dict[tt] * 0.0, "rad_accum": "4He"}
        elif tt == "H3":
            ckw[tt] = {"C_t": C_in_dict["H3"], "t_half": 12.34}
    mc_conv = conv_mcmc(ww, tracers, okw, ckw, pkw, "conv_traces", "".join(str(x) for x in savenum))
    t0 = time.perf_counter()
    idata = mc_conv.sample_mcmc(chains=chains, tune=10000, draws=10000, random_seed=123423, tune_interval=1000)
    dt = time.perf_counter() - t0
    summ = diagnostics.summary({k: v for k, v in idata["posterior"].items() if k in ("tau1", "eta1", "J", "thalf_cfc", "lamsf6", "nu")})
    print("well %s  %d chains x 20,000 steps x %d tracers: sampling_time %.2f s (reference, 3 chains: 296-324 s); "
          "whole sample_mcmc call incl. CUDA context, plan upload and trace download %.2f s" % (
              ww, chains, len(tracers), idata["sample_stats"]["sampling_time"], dt))
    print("%-10s %10s %10s %10s %10s %8s | %10s %10s %10s  (reference trace, pymc3)" % ("", "mean", "sd", "median", "ess_bulk", "r_hat", "mean", "sd", "median"))
    for k, r in summ.items():
        rv = ref["vars"].get(k)
        tail = " | %10.4g %10.4g %10.4g" % (rv["mean"], rv["sd"], rv["q"][9]) if rv else ""
        print("%-10s %10.4g %10.4g %10.4g %10.0f %8.3f%s" % (k, r["mean"], r["sd"], r["median"], r["ess_bulk"], r["r_hat"], tail))
    pp = mc_conv.posterior_predictive()
    mu, err = mc_conv.observations()
    for i, t in enumerate(tracers):
        print("posterior predictive %-8s median %.4g  obs %.4g +- %.2g" % (t, np.median(pp[t]), mu[i], err[i]))


if __name__ == "__main__":
    main()
