"""TEST INFRASTRUCTURE (build container only): summarise the posterior traces the reference ships for config 2
(age_ens_runs_mcmc/conv_traces/*.netcdf -- pymc3 3.11.2 DEMetropolisZ output of run_age_mcmc.py, 3 chains x 10,000 draws)
into tests/golden/age_traces.json.  The files are NetCDF-4/HDF5; the image has no h5py/netCDF4, so they are read with the
repo's own pure-Python reader (noblegas_rtd_mcmc_b200/netcdf4_reader.py).  /root/reference does not travel to the GPU box;
this fixture does.

    python oracle/gen_trace_fixtures.py [/root/reference]
"""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


reader = _load("netcdf4_reader", "noblegas_rtd_mcmc_b200/netcdf4_reader.py")       # no CUDA library needed for these two
diagnostics = _load("diagnostics", "noblegas_rtd_mcmc_b200/diagnostics.py")
QGRID = [0.5, 1, 2.5, 5, 10, 20, 25, 30, 40, 50, 60, 70, 75, 80, 90, 95, 97.5, 99, 99.5]


def summarise(path):
    tr = reader.read_trace(path)
    name = os.path.basename(path)[:-len(".netcdf")]
    well, rest = name.split(".", 1)
    parts = rest.split(".")
    savenum, model, tracers = parts[-1], parts[-2], parts[:-2]
    post = {k: np.asarray(v, dtype=np.float64) for k, v in tr["posterior"].items() if k not in ("chain", "draw")}
    summ = diagnostics.summary(post)
    out = {"well": well, "tracers": tracers, "model": model, "savenum": savenum,
           "obs_mu": [float(x) for x in np.asarray(tr["observed_data"]["like"]).ravel()],
           "sampling_time": float(np.asarray(tr["attrs"]["posterior"]["sampling_time"]).ravel()[0]),
           "tuning_steps": int(np.asarray(tr["attrs"]["posterior"]["tuning_steps"]).ravel()[0]),
           "chains": int(next(iter(post.values())).shape[0]), "draws": int(next(iter(post.values())).shape[1]),
           "accept_rate": float(np.asarray(tr["sample_stats"]["accepted"], dtype=np.float64).mean()),
           "lambda_final": [float(x) for x in np.asarray(tr["sample_stats"]["lambda"])[:, -1]],
           "scaling_final": [float(x) for x in np.asarray(tr["sample_stats"]["scaling"])[:, -1]],
           "vars": {}}
    for k, v in post.items():
        s = summ[k]
        out["vars"][k] = {"mean": float(v.mean()), "sd": float(v.std(ddof=1)), "q": [float(x) for x in np.percentile(v, QGRID)],
                          "ess_bulk": float(s["ess_bulk"]), "ess_tail": float(s["ess_tail"]), "r_hat": float(s["r_hat"]),
                          "mcse_mean": float(s["mcse_mean"]), "chain_means": [float(x) for x in v.mean(axis=1)]}
    return name, out


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    files = sorted(glob.glob(os.path.join(ref, "age_ens_runs_mcmc", "conv_traces", "*.netcdf")))
    res = {"source": "age_ens_runs_mcmc/conv_traces/*.netcdf of the untouched reference, read by noblegas_rtd_mcmc_b200/netcdf4_reader.py",
           "qgrid": QGRID, "traces": {}}
    for f in files:
        name, out = summarise(f)
        res["traces"][name] = out
        v = out["vars"]["tau1"]
        print("%-62s tau1 mean %9.3f sd %9.3f ess %7.0f rhat %.3f  t %.0f s" % (name, v["mean"], v["sd"], v["ess_bulk"], v["r_hat"], out["sampling_time"]))
    with open(os.path.join(ROOT, "tests", "golden", "age_traces.json"), "w") as fh:
        json.dump(res, fh, indent=0)
    print("wrote tests/golden/age_traces.json:", len(files), "traces")


if __name__ == "__main__":
    main()
