import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import c_oracle
from noblegas_rtd_mcmc_b200 import _lib, synthetic
pn = list(synthetic.PAR_NAMES_CFG3)
X, descs = synthetic.series_matrix_and_descs(pn)
plan = _lib.Plan(X, descs, "exp_pist_flow", "dispersion")
for B in (5000, 4999, 2500, 7000):
    th = synthetic.theta_cfg3_informative(B, 40 + B)
    th[: B // 3] = synthetic.theta_cfg3(B // 3, 41 + B)
    a = plan.forward_host(th, pn)
    want = c_oracle.forward(X, descs, "exp_pist_flow", "dispersion", th, pn)
    na, nb = np.isnan(a), np.isnan(want)
    print("B", B, "nan mismatch", int((na != nb).sum()))
    m = ~(na | nb)
    rel = np.zeros_like(a); rel[m] = np.abs(a[m] - want[m]) / np.maximum(np.abs(want[m]), 1e-300)
    bad = np.argwhere(rel > 1e-10)
    print("  bad entries", len(bad), "chains", sorted(set(bad[:, 0].tolist()))[:20])
    for c in sorted(set(bad[:, 0].tolist()))[:4]:
        print("  chain", c, "unit", c // 16, "theta", th[c], "\n    got ", a[c], "\n    want", want[c])
