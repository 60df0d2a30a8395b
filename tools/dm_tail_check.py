"""Parity and timing of a -DNGRTD_DM_TAIL build (NGRTD_LIB=<path>) on the real yearly series (L = 25,256, constant beyond lag 128):
golden vectors of the untouched reference for every configuration with a dispersion component, the oracle on parameters outside the
quadrature's validated domain (per-chain fallback), and launch times.  Development aid (r1: experimental build, see profiles/r1_notes.md)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
from helpers import GOLD, MODEL_CFGS as CONFIGS, REAL_TRACERS, rel_err, real_plan, load_c_in
import np_oracle as O
tag = os.path.basename(os.environ.get("NGRTD_LIB", "libngrtd.so"))
z = np.load(os.path.join(GOLD, "forward_real.npz"))
tracers = ["CFC12", "SF6", "H3", "He4_ter", "He3", "CFC11"]
for name in ("dm", "dm_dm", "epm_dm", "dm_emm"):
    m1, m2, pn = CONFIGS[name]
    if "thalf_cfc" in pn:
        print("%-14s %-8s skipped (per-chain decay constant: plan keeps the full loop)" % (tag, name)); continue
    plan, _ = real_plan(m1, m2, pn, tracers)
    th = z[name + "/theta"]
    out = plan.forward_host(th, pn)
    worst = 0.0
    for i, t in enumerate(tracers):
        want = z[name + "/" + t]
        assert np.array_equal(np.isnan(out[:, i]), np.isnan(want)), (name, t, "NaN pattern")
        worst = max(worst, rel_err(out[:, i], want))
    # timing: 16,384 chains drawn around the golden parameters
    rng = np.random.default_rng(1)
    big = th[rng.integers(0, len(th), 16384)] * rng.uniform(0.9, 1.1, (16384, th.shape[1]))
    for dn in ("D1", "D2"):
        if dn in pn:
            big[:, pn.index(dn)] = np.clip(big[:, pn.index(dn)], 0.01, 2.5)
    if "f1" in pn:
        big[:, pn.index("f1")] = np.clip(big[:, pn.index("f1")], 0.01, 0.99)
    bt = torch.from_numpy(np.ascontiguousarray(big)).cuda()
    o = torch.empty((16384, len(tracers)), dtype=torch.float64, device="cuda")
    for _ in range(2): plan.forward_dev(bt, pn, out_t=o)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): plan.forward_dev(bt, pn, out_t=o)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
    print("%-14s %-8s golden parity %.2e (%d thetas) | 16,384 chains %.3f ms per launch" % (tag, name, worst, len(th), ms), flush=True)
# fallback domain: compare with the oracle's full sums
pn = ["tau1", "D1"]
plan, _ = real_plan("dispersion", False, pn, tracers)
th = np.array([[150.0, 0.004], [131.0, 0.002], [9000.0, 0.003], [40000.0, 0.01], [60000.0, 0.02], [0.5, 1.0], [300.0, 3.0], [2000.0, 8.0]])
out = plan.forward_host(th, pn)
C = load_c_in()
worst = 0.0
for i, t in enumerate(tracers):
    key, thalf, ra = REAL_TRACERS[t]
    s = C[key] if key is not None else np.zeros(len(C["H3"]))
    want = O.forward_mod(th, pn, t, s, "dispersion", False, t_half=thalf, rad_accum=ra)
    ok = np.isfinite(want) & (want != 0)
    assert np.array_equal(np.isnan(out[:, i]), np.isnan(want)), (t, out[:, i], want)
    worst = max(worst, float(np.max(np.abs(out[ok, i] - want[ok]) / np.abs(want[ok]))) if ok.any() else 0.0)
print("%-14s outside the validated domain (8 parameter pairs) vs oracle: %.2e" % (tag, worst))
