"""Time the dispersion plan of the bench's cfg2_dispersion block (real yearly series, L = 25,256, constant beyond lag 128:
lag loop over 128 lags + tail by quadrature) of one library build (NGRTD_LIB=<path>), for the bench's parameter ranges and for
narrow RTDs only.  Development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import _lib, datasets
B = 65536
tr2 = ["CFC12", "SF6", "H3", "He4_ter"]
pn = ["tau1", "D1", "J", "lamsf6"]
X, d, _ = datasets.real_series_matrix_and_descs(pn, tr2)
plan = _lib.Plan(X, d, "dispersion", False)
obs = np.array([300.0, 5.0, 8.0, 1e-8]); sd = 0.05 * obs
lp = torch.empty(B, dtype=torch.float64, device="cuda")
out = []
for name, dlo, dhi in (("bench D 0.01-2", 0.01, 2.0), ("narrow D 0.01-0.05", 0.01, 0.05), ("wide D 0.5-2", 0.5, 2.0)):
    rng = np.random.default_rng(11)
    th = np.stack([rng.uniform(1, 1000, B), rng.uniform(dlo, dhi, B), rng.normal(-10.42, 0.33, B), np.abs(rng.normal(0, 0.17, B))], axis=1)
    th_d = torch.from_numpy(th).cuda()
    for _ in range(3):
        plan.forward_loglik_dev(th_d, pn, obs, sd, "normal", logp_t=lp)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            plan.forward_loglik_dev(th_d, pn, obs, sd, "normal", logp_t=lp)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 5)
    out.append("%s: %.4f ms (checksum %.12e)" % (name, best, float(torch.nansum(lp))))
print("%-14s %s" % (os.path.basename(os.environ.get("NGRTD_LIB", "libngrtd.so")), " | ".join(out)), flush=True)
