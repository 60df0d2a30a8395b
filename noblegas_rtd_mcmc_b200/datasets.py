"""Input series shipped with the package.

`load_c_in()` rebuilds the reference's yearly input series (`age_ens_runs_mcmc/C_in_dict.pk`: six tracers x 25,256 lags,
index = years before sampling, newest first) from a 7 KB head: only the newest <= 85 lags of every series vary, the
remaining ~25,000 rows are one constant background value per tracer (the back-extension of
`age_modeling_mcmc.prep.py:165-223`), so head + background reproduce the pickle bit for bit (checked by
tests/test_oracle_golden.py::test_c_in_head_matches_reference when the reference tree is present)."""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
SERIES_KEYS = ("CFC11", "CFC12", "CFC113", "SF6", "He4_ter", "H3")
REAL_SERIES_NAMES = ("CFC11", "CFC12", "CFC113", "SF6", "H3")

# tracer -> (series key, t_half [yr], rad_accum) on the yearly data (reference run_age_mcmc.py:200-224)
REAL_TRACERS = {
    "CFC11": ("CFC11", False, False),
    "CFC12": ("CFC12", False, False),
    "CFC113": ("CFC113", False, False),
    "SF6": ("SF6", False, False),
    "He4_ter": (None, False, "4He"),
    "He3": ("H3", 12.34, "3He"),
    "H3": ("H3", 12.34, False),
}


def load_c_in(L=None):
    """dict tracer -> float64[L] (newest first); L defaults to the reference's 25,256 lags."""
    z = np.load(os.path.join(_DATA, "c_in_head.npz"))
    Lfull = int(z["L"])
    L = Lfull if L is None else int(L)
    out = {}
    for k in SERIES_KEYS:
        v = np.full(L, float(z[k + "_bg"]))
        n = min(L, 128)
        v[:n] = z[k + "_head"][:n]
        out[k] = v
    return out


def real_series_matrix_and_descs(par_names, tracers, L=None):
    """(X [L, 5] newest-first, tracer descriptor dicts for _lib.Plan / the oracle, series dict) on the yearly data."""
    C = load_c_in(L)
    names = list(REAL_SERIES_NAMES)
    X = np.ascontiguousarray(np.stack([C[n] for n in names], axis=1))
    descs = []
    for t in tracers:
        s, th, ra = REAL_TRACERS[t]
        descs.append(dict(series=names.index(s) if s is not None else -1, rad_accum=ra,
                          lam=float(-1.0 * np.log(0.5) / th) if th else 0.0,
                          use_thalf_cfc=(t == "CFC12" and "thalf_cfc" in par_names),
                          use_lamsf6=(t == "SF6")))
    return X, descs, C
