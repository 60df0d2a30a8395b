"""GPU: edge cases -- empty and ragged batches, tiny and non-multiple-of-4 lag axes, one and eight tracers, NaN /
out-of-domain parameters, error paths of the C ABI -- against the numpy oracle (reference arithmetic)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle(X, descs, m1, m2, theta, pn):
    import np_oracle as O
    out = np.empty((theta.shape[0], len(descs)))
    for i, d in enumerate(descs):
        s = X[:, d["series"]] if d.get("series", -1) >= 0 else np.zeros(X.shape[0])
        lam = d.get("lam", 0.0)
        name = "CFC12" if d.get("use_thalf_cfc") else ("SF6" if d.get("use_lamsf6") else ("He4_ter" if d.get("rad_accum") == "4He" else "X"))
        out[:, i] = O.forward_mod(theta, pn, name, s, m1, m2, t_half=(np.log(2) / lam if lam else False),
                                  rad_accum=d.get("rad_accum", False))
    return out


@pytest.mark.parametrize("L", [1, 2, 3, 5, 64, 1023, 1024, 1025, 2049])
def test_lag_axis_lengths(L):
    """L not a multiple of 4 (zero padding), L = 1, and lengths around the 1024-lag chunk boundary."""
    from helpers import rel_err
    from noblegas_rtd_mcmc_b200 import _lib
    rng = np.random.default_rng(L)
    X = rng.uniform(0.5, 20.0, (L, 2))
    descs = [dict(series=0, lam=0.05), dict(series=1), dict(series=-1, rad_accum="4He"), dict(series=0, lam=0.05, rad_accum="3He")]
    B = 21
    for m1, m2, pn in (("exponential", False, ["tau1", "J"]), ("exp_pist_flow", "dispersion", ["tau1", "tau2", "f1", "f2", "eta1", "D2", "J"]),
                       ("piston", "exponential", ["tau1", "tau2", "f1", "f2", "J"]), ("dispersion", False, ["tau1", "D1", "J"])):
        f1 = rng.uniform(0.1, 0.9, B)
        cols = {"tau1": rng.uniform(0.3, max(2.0, 0.6 * L), B), "tau2": rng.uniform(1.0, max(3.0, 1.5 * L), B), "f1": f1, "f2": 1 - f1,
                "eta1": rng.uniform(1, 5, B), "D1": rng.uniform(0.02, 2, B), "D2": rng.uniform(0.02, 2, B), "J": rng.normal(-10.4, 0.3, B)}
        theta = np.stack([cols[p] for p in pn], axis=1)
        out = _lib.Plan(X, descs, m1, m2).forward_host(theta, pn)
        want = _oracle(X, descs, m1, m2, theta, pn)
        assert rel_err(out[:, :3], want[:, :3]) < 1e-10, (L, m1, m2)
        # 3He ingrowth at lag 0 is 1 - exp(-lambda * 1e-5): the reference's own subtraction cancels 7 digits there, so for
        # L <= 2 (where that lag dominates) one ulp of exp() is 2e-10 of the result -- conditioning, not a kernel error
        assert rel_err(out[:, 3], want[:, 3]) < (1e-8 if L <= 2 else 1e-10), (L, m1, m2)


def test_empty_batch_and_tracer_counts():
    from helpers import rel_err
    from noblegas_rtd_mcmc_b200 import _lib
    rng = np.random.default_rng(1)
    X = rng.uniform(0.5, 20.0, (300, 7))
    one = _lib.Plan(X, [dict(series=3)], "exponential")
    assert one.forward_host(np.empty((0, 1)), ["tau1"]).shape == (0, 1)
    eight = [dict(series=i) for i in range(7)] + [dict(series=-1, rad_accum="4He")]       # ones + 7 series would be 8 columns + index
    with pytest.raises(_lib.NgrtdError):
        _lib.Plan(X, eight, "exponential")                                                 # more than 7 folded columns
    seven = [dict(series=i) for i in range(6)] + [dict(series=-1, rad_accum="4He")] + [dict(series=2)]   # 8 tracers, 7 columns
    p8 = _lib.Plan(X, seven, "exp_pist_flow")
    theta = np.stack([rng.uniform(2, 200, 9), rng.uniform(1, 5, 9), rng.normal(-10.4, 0.3, 9)], axis=1)
    out = p8.forward_host(theta, ["tau1", "eta1", "J"])
    assert out.shape == (9, 8) and np.array_equal(out[:, 2], out[:, 7])
    assert rel_err(out, _oracle(X, seven, "exp_pist_flow", False, theta, ["tau1", "eta1", "J"])) < 1e-10
    with pytest.raises(_lib.NgrtdError):
        one.forward_host(np.ones((2, 1)), ["eta1"])                                        # tau1 is required
    with pytest.raises(ValueError):
        _lib.Plan(X, [dict(series=0, rad_accum="SF6")], "exponential")


def test_pathological_parameters_propagate_like_the_reference():
    """NaN, zero, negative and huge parameters: NaN/inf patterns equal the reference arithmetic (no exceptions)."""
    import np_oracle as O
    from noblegas_rtd_mcmc_b200 import _lib
    rng = np.random.default_rng(2)
    L = 500
    X = rng.uniform(0.5, 20.0, (L, 1))
    descs = [dict(series=0, lam=0.03)]
    # exponential-piston: all-masked (tau(1-1/eta) beyond the window) -> NaN; eta = 1 exact; tiny tau; NaN inputs
    theta = np.array([[1e6, 5.0], [700.0, 5.0], [40.0, 1.0], [1e-3, 2.0], [np.nan, 2.0], [30.0, np.nan], [499.0 / 0.8, 5.0], [623.76, 5.0]])
    out = _lib.Plan(X, descs, "exp_pist_flow").forward_host(theta, ["tau1", "eta1"])[:, 0]
    want = O.forward_mod(theta, ["tau1", "eta1"], "X", X[:, 0], "exp_pist_flow", False, t_half=np.log(2) / 0.03)
    assert np.array_equal(np.isnan(out), np.isnan(want)) and np.isnan(out).sum() >= 3
    ok = ~np.isnan(want)
    assert np.allclose(out[ok], want[ok], rtol=1e-10, atol=0)
    # dispersion: D <= 0, tau <= 0, NaN -> NaN; tau far beyond the window -> every weight underflows -> NaN (0/0)
    theta = np.array([[50.0, 0.0], [50.0, -0.3], [0.0, 0.3], [-5.0, 0.3], [np.nan, 0.3], [50.0, np.nan], [5e6, 0.01], [50.0, 0.3], [3000.0, 0.05]])
    out = _lib.Plan(X, descs, "dispersion").forward_host(theta, ["tau1", "D1"])[:, 0]
    want = O.forward_mod(theta, ["tau1", "D1"], "X", X[:, 0], "dispersion", False, t_half=np.log(2) / 0.03)
    assert np.array_equal(np.isnan(out), np.isnan(want)), (out, want)
    ok = ~np.isnan(want)
    assert np.allclose(out[ok], want[ok], rtol=1e-10, atol=0)
    # piston: tau beyond the window clamps to the last lag, ties pick the lower index, tau < 1 chooses between lags 0 and 1
    theta = np.array([[1e9], [10.5], [0.3], [0.6], [0.5000049], [-4.0], [498.5], [499.49], [np.nan]])
    out = _lib.Plan(X, descs, "piston").forward_host(theta, ["tau1"])[:, 0]
    want = O.forward_mod(theta, ["tau1"], "X", X[:, 0], "piston", False, t_half=np.log(2) / 0.03)
    assert np.allclose(out, want, rtol=1e-12, atol=0)


def test_ce_edges():
    import np_oracle as O
    from noblegas_rtd_mcmc_b200.noble_gas_utils import noble_gas_fun
    gases = ["He", "Ne", "Ar", "Kr", "Xe"]
    E = np.array([0.0, 3000.0, 3000.0, 3000.0, 50000.0, 3000.0])
    T = np.array([10.0, -0.5, 65.0, 99.0, 5.0, 100.0])
    Ae = np.array([0.0, 0.01, 0.01, 0.1, 0.01, 1e-4])
    F = np.array([0.5, 0.5, 10.0, 0.0, 0.5, 1.0])
    got = noble_gas_fun(gases, E, T, Ae, F, "lapse_rate").ce_exc(True)
    want = O.ce_exc(gases, E, T, Ae, F, True)
    for i, g in enumerate(gases):
        assert np.allclose(got[g], want[:, i], rtol=1e-12, atol=0, equal_nan=True), g
    one_atm = noble_gas_fun(["Ar"], 1234.0, 12.0, 0.002, 0.3, "1atm").equil_conc()["Ar"]
    assert abs(one_atm - O.equil_conc(["Ar"], 12.0, 0.000101325)[0]) < 1e-12 * one_atm


def test_constant_tail_closed_form_corners():
    """Series with a long constant tail (the reference's back-extension): analytic tail vs the oracle's full lag loop,
    including masks beyond the cut (k0 > Kc), masks beyond the window, nearly uniform weights (Taylor branch of the
    arithmetico-geometric sum), a non-unit lag-index slope and per-chain CFC decay."""
    import np_oracle as O
    from helpers import rel_err
    from noblegas_rtd_mcmc_b200 import _lib
    rng = np.random.default_rng(12)
    L, nv = 6000, 77
    X = np.empty((L, 3))
    X[:nv] = rng.uniform(1.0, 50.0, (nv, 3))
    X[nv:] = np.array([3.3273, 1e-10, 7.5])
    idx = 2.0 * np.arange(L) + 10.0                                   # regular lag index, slope 2, offset 10
    descs = [dict(series=0, lam=np.log(2) / 12.34), dict(series=0, lam=np.log(2) / 12.34, rad_accum="3He"), dict(series=1),
             dict(series=2, use_thalf_cfc=True), dict(series=-1, rad_accum="4He"), dict(series=1, use_lamsf6=True)]
    pn = ["tau1", "tau2", "f1", "f2", "eta1", "eta2", "J", "thalf_cfc", "lamsf6"]
    plan = _lib.Plan(X, descs, "exp_pist_flow", "exp_pist_flow", lag_index=idx)
    B = 600
    f1 = rng.uniform(0.05, 0.95, B)
    tau1 = np.exp(rng.uniform(0, np.log(2000), B))
    tau2 = np.exp(rng.uniform(np.log(50), np.log(15000), B))
    tau2[:8] = [1e6, 5e7, 1e9, 7000.0, 7499.9, 7500.1, 9000.0, 1e5]   # nearly uniform weights / masks around and beyond L
    eta2 = rng.uniform(1, 5, B)
    eta2[:8] = [1.0, 1.0, 1.0, 5.0, 5.0, 5.0, 5.0, 1.0001]
    theta = np.stack([tau1, tau2, f1, 1 - f1, rng.uniform(1, 5, B), eta2, rng.normal(-10.4, 0.3, B), rng.uniform(5, 35, B),
                      np.abs(rng.normal(0, 0.17, B))], axis=1)
    out = plan.forward_host(theta, pn)
    tp = O.lag_grid(L)
    want = np.empty_like(out)
    names = ["H3", "He3", "X", "CFC12", "He4_ter", "SF6"]
    for i, d in enumerate(descs):
        s = X[:, d["series"]] if d["series"] >= 0 else np.zeros(L)
        want[:, i] = O.forward_mod(theta, pn, names[i], s, "exp_pist_flow", "exp_pist_flow",
                                   t_half=(12.34 if d.get("lam") else False), rad_accum=d.get("rad_accum", False),
                                   index_newest_first=idx)
    assert rel_err(out, want) < 1e-10
    assert np.isnan(out[:, 0]).sum() >= 2 and np.isfinite(out[0]).all()
    # an irregular lag index disables the analytic tail for the 4He column but must still be exact
    idx2 = idx.copy()
    idx2[L - 3] += 0.5
    plan2 = _lib.Plan(X, descs, "exponential", False, lag_index=idx2)
    th2 = theta[:64][:, [0, 6, 7, 8]]
    got = plan2.forward_host(th2, ["tau1", "J", "thalf_cfc", "lamsf6"])
    w2 = O.forward_mod(th2, ["tau1", "J", "thalf_cfc", "lamsf6"], "He4_ter", np.zeros(L), "exponential", False, rad_accum="4He",
                       index_newest_first=idx2)
    assert rel_err(got[:, 4], w2) < 1e-10


def test_dispersion_extreme_parameters_on_the_real_series():
    """Very narrow / very broad dispersion RTDs and modes far beyond the lag window on the 25,256-lag series against the
    oracle's full sums.  In -DNGRTD_DM_TAIL builds (ngrtd_build_features() & 1) these parameters lie outside the quadrature's
    validated domain and exercise its lag-by-lag fallback; in the default build they take the ordinary lag loop."""
    import np_oracle as O
    from helpers import REAL_TRACERS, load_c_in, real_plan
    tracers = ["CFC12", "SF6", "H3", "He4_ter", "He3", "CFC11"]
    pn = ["tau1", "D1"]
    plan, C = real_plan("dispersion", False, pn, tracers)
    th = np.array([[150.0, 0.004], [131.0, 0.002], [9000.0, 0.003], [40000.0, 0.01], [60000.0, 0.02], [0.5, 1.0],
                   [300.0, 3.0], [2000.0, 8.0]])
    out = plan.forward_host(th, pn)
    for i, t in enumerate(tracers):
        key, thalf, ra = REAL_TRACERS[t]
        s = C[key] if key is not None else np.zeros(len(C["H3"]))
        want = O.forward_mod(th, pn, t, s, "dispersion", False, t_half=thalf, rad_accum=ra)
        assert np.array_equal(np.isnan(out[:, i]), np.isnan(want)), (t, out[:, i], want)
        ok = np.isfinite(want) & (want != 0)
        if ok.any():
            assert np.max(np.abs(out[ok, i] - want[ok]) / np.abs(want[ok])) < 1e-10, t
