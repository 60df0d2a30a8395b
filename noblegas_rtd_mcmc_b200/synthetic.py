"""Deterministic synthetic workloads for the benchmark configs of BASELINE.json.

Config 3: L = 840 monthly lags, six input-series tracers
{H3, He3 (ingrowth of the H3 series), CFC11, CFC12, CFC113, SF6} plus He4_ter as the
implicit `k*J` column; EPM + dispersion two-component mixture; theta drawn from the
reference priors (age_ens_runs_mcmc/run_age_mcmc.py:145-183 of the reference) rescaled
from years to monthly steps.  Everything is a pure function of (L, seed): no data files.
"""
import numpy as np

SERIES_NAMES = ("H3", "CFC11", "CFC12", "CFC113", "SF6")
TRACERS_CFG3 = ("H3", "He3", "CFC11", "CFC12", "CFC113", "SF6", "He4_ter")
PAR_NAMES_CFG3 = ("tau1", "tau2", "f1", "f2", "eta1", "D2", "J")
T_HALF_H3_STEPS = 12.34 * 12.0          # tritium half-life in monthly steps
LOG10_J_MONTHLY = -10.424151075511698 - np.log10(12.0)


def input_series(L=840, seed=0):
    """Return dict name -> float64[L], newest-first (entry k = concentration k lags before sampling)."""
    rng = np.random.default_rng(seed)
    k = np.arange(L, dtype=np.float64)
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    s = L / 840.0
    base = {
        "H3": 6.0 + 2600.0 * np.exp(-(((k - 684.0 * s) / (30.0 * s)) ** 2)),
        "CFC11": 260.0 * sig((300.0 * s - k) / (60.0 * s)),
        "CFC12": 550.0 * sig((360.0 * s - k) / (60.0 * s)),
        "CFC113": 85.0 * sig((280.0 * s - k) / (50.0 * s)),
        "SF6": 10.0 * np.exp(-k / (180.0 * s)),
    }
    out = {}
    for name in SERIES_NAMES:
        noise = rng.lognormal(0.0, 0.05, size=L)
        out[name] = base[name] * noise + 1e-10
    return out


def theta_cfg3(B, seed=0):
    """theta[B, 7] in PAR_NAMES_CFG3 order: EPM(tau1, eta1) + DM(tau2, D2), f1, f2 = 1 - f1, log10 J."""
    rng = np.random.default_rng(seed + 1)
    tau1 = rng.uniform(12.0, 12000.0, B)
    tau2 = rng.uniform(600.0, 180000.0, B)
    f1 = rng.uniform(0.01, 0.99, B)
    eta1 = rng.uniform(1.0, 5.0, B)
    D2 = rng.uniform(0.01, 2.0, B)
    J = rng.normal(LOG10_J_MONTHLY, 0.33, B)
    return np.ascontiguousarray(np.stack([tau1, tau2, f1, 1.0 - f1, eta1, D2, J], axis=1))


def theta_cfg3_informative(B, seed=0):
    """Same layout as theta_cfg3 but with mean ages comparable to the record length, so every
    output is finite and well away from the all-weights-masked NaN region (used for tight parity)."""
    rng = np.random.default_rng(seed + 2)
    tau1 = rng.uniform(12.0, 600.0, B)
    tau2 = rng.uniform(100.0, 3000.0, B)
    f1 = rng.uniform(0.01, 0.99, B)
    eta1 = rng.uniform(1.0, 3.0, B)
    D2 = rng.uniform(0.01, 2.0, B)
    J = rng.normal(LOG10_J_MONTHLY, 0.33, B)
    return np.ascontiguousarray(np.stack([tau1, tau2, f1, 1.0 - f1, eta1, D2, J], axis=1))


def tracer_table_cfg3():
    """Tracer descriptors in the ckw layout of the reference driver (run_age_mcmc.py:200-224)."""
    lam = None
    return {
        "H3": dict(series="H3", t_half=T_HALF_H3_STEPS),
        "He3": dict(series="H3", t_half=T_HALF_H3_STEPS, rad_accum="3He"),
        "CFC11": dict(series="CFC11"),
        "CFC12": dict(series="CFC12"),
        "CFC113": dict(series="CFC113"),
        "SF6": dict(series="SF6"),
        "He4_ter": dict(series=None, rad_accum="4He"),
    }


def series_matrix_and_descs(par_names, tracers=None, L=840, seed=0):
    """(X [L, nseries] newest-first, tracer descriptor dicts) for the C ABI / the oracle, cfg-3 tracer table."""
    tracers = list(TRACERS_CFG3) if tracers is None else list(tracers)
    series = input_series(L, seed)
    names = list(SERIES_NAMES)
    X = np.ascontiguousarray(np.stack([series[n] for n in names], axis=1))
    tab = tracer_table_cfg3()
    descs = []
    for t in tracers:
        d = tab[t]
        descs.append(dict(series=names.index(d["series"]) if d["series"] is not None else -1,
                          rad_accum=d.get("rad_accum", False),
                          lam=float(np.log(2.0) / d["t_half"]) if "t_half" in d else 0.0,
                          use_thalf_cfc=(t == "CFC12" and "thalf_cfc" in par_names),
                          use_lamsf6=(t == "SF6")))
    return X, descs
