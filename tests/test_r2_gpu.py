"""GPU: features added in round 2 -- the dispersion tail by quadrature as the DEFAULT build (incl. the per-chain decay
column), the tape schedule of k_forward (cut units, small batches, determinism), the f1-complement column alias, pooled
moments on the device (K6), objects on several devices in one thread, and the compiled C-ABI demo."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dm_tail_is_the_default_build():
    """ngrtd_build_features() bit 0: dispersion components integrate the constant tail of the reference's back-extended
    series (age_modeling_mcmc.prep.py:167-169) by quadrature instead of looping over ~25,000 constant lags."""
    from noblegas_rtd_mcmc_b200 import _lib
    assert _lib.lib.ngrtd_build_features() & 1


def test_dm_tail_real_series_goldens_and_loop_equivalence(monkeypatch):
    """The 4 golden configurations with a dispersion component on the real 25,256-lag series (outputs of the untouched
    reference) through the quadrature path, and the same plans with the analytic tails switched off (NGRTD_NO_TAIL: full
    lag loop) -- both within 1e-10 of the reference, and within 1e-11 of each other."""
    from helpers import GOLD, MODEL_CFGS, real_plan, rel_err
    z = np.load(os.path.join(GOLD, "forward_real.npz"))
    tracers = ["CFC12", "SF6", "H3", "He4_ter", "He3", "CFC11"]
    for name in ("dm", "dm_dm", "epm_dm", "dm_emm"):
        m1, m2, pn = MODEL_CFGS[name]
        th = z[name + "/theta"]
        monkeypatch.delenv("NGRTD_NO_TAIL", raising=False)
        plan, _ = real_plan(m1, m2, pn, tracers)
        quad = plan.forward_host(th, pn)
        monkeypatch.setenv("NGRTD_NO_TAIL", "1")
        plan_full, _ = real_plan(m1, m2, pn, tracers)
        full = plan_full.forward_host(th, pn)
        monkeypatch.delenv("NGRTD_NO_TAIL", raising=False)
        for i, t in enumerate(tracers):
            want = z[name + "/" + t]
            assert rel_err(quad[:, i], want) < 1e-10, (name, t)
            assert rel_err(full[:, i], want) < 1e-10, (name, t)
        assert rel_err(quad, full) < 1e-11, name


def test_dm_tail_with_per_chain_cfc_decay():
    """Dispersion + sampled thalf_cfc (run_age_mcmc_utils.py:107-109: CFC-12 decays with a per-chain constant) on the real
    series: the quadrature tail carries the per-chain-decay column; against the numpy oracle's full sums."""
    import np_oracle as O
    from helpers import REAL_TRACERS, real_plan
    tracers = ["CFC12", "SF6", "H3", "He4_ter"]
    pn = ["tau1", "D1", "J", "thalf_cfc", "lamsf6"]
    plan, C = real_plan("dispersion", False, pn, tracers)
    rng = np.random.default_rng(5)
    B = 48
    th = np.stack([rng.uniform(5, 900, B), rng.uniform(0.01, 2.0, B), rng.normal(-10.42, 0.33, B), rng.uniform(5, 35, B),
                   np.abs(rng.normal(0, 0.17, B))], axis=1)
    th[:4, 0] = [3000.0, 9000.0, 14000.0, 60.0]           # modes inside the constant tail
    out = plan.forward_host(th, pn)
    for i, t in enumerate(tracers):
        key, thalf, ra = REAL_TRACERS[t]
        s = C[key] if key is not None else np.zeros(len(C["H3"]))
        want = O.forward_mod(th, pn, t, s, "dispersion", False, t_half=thalf, rad_accum=ra)
        ok = np.isfinite(want) & (want != 0)
        assert np.array_equal(np.isnan(out[:, i]), np.isnan(want)), t
        assert np.max(np.abs(out[ok, i] - want[ok]) / np.abs(want[ok])) < 1e-10, t
    # the same through a two-component plan (exp_pist_flow closed-form tail + dispersion quadrature tail, per-chain decay)
    pn2 = ["tau1", "tau2", "f1", "f2", "eta1", "D2", "J", "thalf_cfc", "lamsf6"]
    plan2, _ = real_plan("exp_pist_flow", "dispersion", pn2, tracers)
    f1 = rng.uniform(0.05, 0.95, B)
    th2 = np.stack([rng.uniform(5, 900, B), rng.uniform(50, 12000, B), f1, 1 - f1, rng.uniform(1, 3, B), rng.uniform(0.01, 2.0, B),
                    rng.normal(-10.42, 0.33, B), rng.uniform(5, 35, B), np.abs(rng.normal(0, 0.17, B))], axis=1)
    out2 = plan2.forward_host(th2, pn2)
    for i, t in enumerate(tracers):
        key, thalf, ra = REAL_TRACERS[t]
        s = C[key] if key is not None else np.zeros(len(C["H3"]))
        want = O.forward_mod(th2, pn2, t, s, "exp_pist_flow", "dispersion", t_half=thalf, rad_accum=ra)
        ok = np.isfinite(want) & (want != 0)
        assert np.array_equal(np.isnan(out2[:, i]), np.isnan(want)), t
        assert np.max(np.abs(out2[ok, i] - want[ok]) / np.abs(want[ok])) < 1e-10, t


def test_dm_tail_sampler_kernel_uses_it_and_matches_forward():
    """k_mcmc_age is compiled with the same tail code: logp of the sampler's starting point == forward + likelihood."""
    import torch
    from helpers import real_plan
    from noblegas_rtd_mcmc_b200.sampler import Sampler, prior
    tracers = ["CFC12", "SF6", "H3", "He4_ter"]
    pn = ["tau1", "D1"]
    plan, _ = real_plan("dispersion", False, pn, tracers)
    truth = np.array([[120.0, 0.3]])
    obs = plan.forward_host(truth, pn)[0]
    sd = 0.05 * np.abs(obs)
    pri = [prior("uniform", "tau1", 1, 1000), prior("uniform", "D1", 0.01, 2.0)]
    smp = Sampler(pri, obs, sd, 32, plan=plan, lik="normal", tune_interval=100, hist_cap=64, seed=3)
    lp0 = smp.get("logp").cpu().numpy()
    q0 = smp.get("q").cpu().numpy()
    sig = 1.0 / (1.0 + np.exp(-q0))
    nat = np.stack([1 + 999 * sig[:, 0], 0.01 + 1.99 * sig[:, 1]], axis=1)
    lik = plan.forward_loglik_host(nat, pn, obs, sd, "normal")
    jac = (np.log(sig) + np.log1p(-sig)).sum(axis=1)              # uniform prior + interval transform: log sig(x) + log sig(-x)
    assert np.allclose(lp0, lik + jac, rtol=1e-12, atol=1e-9)
    smp.run(50, tune=True)
    torch.cuda.synchronize()
    assert np.isfinite(smp.get("logp").cpu().numpy()).all()
    smp.close()


def test_tape_schedule_small_and_ragged_batches():
    """k_forward cuts a unit's lag range between the warps of a CTA (tape schedule): batches smaller than a unit, smaller
    than the grid, and ragged ones must give the oracle's results; a repeated call is bitwise identical (static cuts)."""
    import c_oracle
    from helpers import rel_err
    from noblegas_rtd_mcmc_b200 import _lib, synthetic
    pn = list(synthetic.PAR_NAMES_CFG3)
    X, descs = synthetic.series_matrix_and_descs(pn)
    plan = _lib.Plan(X, descs, "exp_pist_flow", "dispersion")
    for B in (1, 3, 16, 17, 100, 148 * 16 + 5, 5000):
        th = synthetic.theta_cfg3_informative(B, 40 + B)
        th[: B // 3] = synthetic.theta_cfg3(B // 3, 41 + B) if B >= 3 else th[: B // 3]
        a = plan.forward_host(th, pn)
        b = plan.forward_host(th, pn)
        assert np.array_equal(a, b, equal_nan=True), B
        want = c_oracle.forward(X, descs, "exp_pist_flow", "dispersion", th, pn)
        assert rel_err(a, want) < 1e-10, B
    # single-component and piston plans take the same schedule (no lag loop for piston: a unit is one tape position)
    for m1, m2, names in (("dispersion", False, ["tau1", "D1", "J"]), ("piston", False, ["tau1", "J"]),
                          ("piston", "exp_pist_flow", ["tau1", "tau2", "f1", "f2", "eta2", "J"])):
        X1, d1 = synthetic.series_matrix_and_descs(names)
        p1 = _lib.Plan(X1, d1, m1, m2)
        rng = np.random.default_rng(9)
        B = 777
        cols = {"tau1": rng.uniform(2, 700, B), "tau2": rng.uniform(20, 900, B), "D1": rng.uniform(0.01, 2, B),
                "J": rng.normal(synthetic.LOG10_J_MONTHLY, 0.3, B), "f1": rng.uniform(0.1, 0.9, B), "eta2": rng.uniform(1, 3, B)}
        cols["f2"] = 1 - cols["f1"]
        th = np.ascontiguousarray(np.stack([cols[n] for n in names], axis=1))
        want = c_oracle.forward(X1, d1, m1, m2, th, names)
        assert rel_err(p1.forward_host(th, names), want) < 1e-10, (m1, m2)


def test_f1_complement_column_is_bitwise_the_seven_column_call():
    """NGRTD_P_F1_COMPLEMENT ("f1_f2c"): theta without the f2 column, f2 = 1 - f1 formed on the device
    (run_age_mcmc_utils.py:304 makes f2 a Deterministic(1 - f1))."""
    from noblegas_rtd_mcmc_b200 import _lib, synthetic
    pn = list(synthetic.PAR_NAMES_CFG3)
    X, descs = synthetic.series_matrix_and_descs(pn)
    plan = _lib.Plan(X, descs, "exp_pist_flow", "dispersion")
    th = synthetic.theta_cfg3_informative(4099, 77)
    full = plan.forward_host(th, pn)
    pn6 = ["tau1", "tau2", "f1_f2c", "eta1", "D2", "J"]
    six = plan.forward_host(np.ascontiguousarray(np.delete(th, 3, axis=1)), pn6)
    assert np.array_equal(full, six, equal_nan=True)
    with pytest.raises(_lib.NgrtdError):
        plan.forward_host(th, ["tau1", "tau2", "f1_f2c", "f2", "eta1", "D2", "J"])


def test_pooled_moments_match_per_chain_moments():
    """K6 on the device: ngrtd_sampler_pooled_moments + the 3*nd+1 number summary == the summary from per-chain moments."""
    import json
    from noblegas_rtd_mcmc_b200 import distributed as ngdist
    from noblegas_rtd_mcmc_b200.noble_gas_mcmc import mcmc_model
    from noblegas_rtd_mcmc_b200.sampler import Sampler
    fx = json.load(open(os.path.join(ROOT, "noblegas_rtd_mcmc_b200", "data", "ng_obs_plm.json")))["wells"]["PLM1"]
    mdl = mcmc_model(fx["obs"], mcmc_model.well_elev["PLM1"])
    smp = Sampler(mdl.build_priors(), mdl.obs_mu, mdl.obs_sd, 3001, plan=None, gases=mdl.gases, lik="studentt",
                  nu_range=(1.0, 30.0), tune_interval=500, hist_cap=512, seed=11)
    smp.run(1500, tune=True)
    smp.stop_tuning()
    smp.run(400, tune=False, record=True)
    a = ngdist.pooled_summary(smp, 400)
    b = ngdist.global_summary(400, smp.get("mean"), smp.get("m2"))
    assert a["chains"] == b["chains"] == 3001
    for k in ("mean", "sd", "r_hat", "mcse_mean"):
        assert np.allclose(a[k], b[k], rtol=1e-9, atol=0), k
    assert np.allclose(a["ess"], b["ess"], rtol=1e-6), (a["ess"], b["ess"])      # B/n from sum mean^2: cancellation ~1e-9
    v1 = smp.pooled_moments().cpu().numpy()
    v2 = smp.pooled_moments().cpu().numpy()
    assert np.array_equal(v1, v2)                                                # deterministic two-stage reduction
    smp.close()


def test_objects_on_two_devices_in_one_thread():
    """the dynamic-shared-memory attribute is per device and the *_dev entry points must run on the object's
    device whatever the caller's current device is (and leave it unchanged)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from noblegas_rtd_mcmc_b200 import _lib, synthetic
    pn = list(synthetic.PAR_NAMES_CFG3)
    X, descs = synthetic.series_matrix_and_descs(pn)
    th = synthetic.theta_cfg3_informative(1000, 5)
    outs = []
    torch.cuda.set_device(0)
    plans = [_lib.Plan(X, descs, "exp_pist_flow", "dispersion", device=d) for d in (0, 1)]
    assert torch.cuda.current_device() == 0
    for d, plan in enumerate(plans):
        t = torch.from_numpy(th).to("cuda:%d" % d)
        with torch.cuda.device(d):
            s = torch.cuda.current_stream()
        torch.cuda.set_device(0)                            # the caller's current device stays 0 for both plans
        o = plan.forward_dev(t, pn, stream=s)
        torch.cuda.synchronize(d)
        assert torch.cuda.current_device() == 0
        outs.append(o.cpu().numpy())
    assert np.array_equal(outs[0], outs[1], equal_nan=True)


def test_c_abi_demo_runs(tmp_path):
    """examples/c_abi_demo.c (the plain-C caller of INTEGRATION.md) is compiled against include/ngrtd.h + libngrtd.so and RUN."""
    src = os.path.join(ROOT, "examples", "c_abi_demo.c")
    exe = str(tmp_path / "c_abi_demo")
    libdir = os.path.join(ROOT, "noblegas_rtd_mcmc_b200")
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-I" + os.path.join(ROOT, "include"), src, "-o", exe, "-L" + libdir, "-lngrtd",
                           "-Wl,-rpath," + libdir, "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "worst relative error" in out.stdout, out.stdout


def test_propagate_obs_ensembles_matches_the_reference_lines():
    """(f2) prep.propagate_obs_ensembles == age_modeling_mcmc.prep.py:242-489 run verbatim on the same synthetic CE posteriors
    (tests/golden/ens_dict_small.npz): same random stream, same members, batched device calls instead of per-member loops."""
    from noblegas_rtd_mcmc_b200 import prep
    z = np.load(os.path.join(ROOT, "tests", "golden", "ens_dict_small.npz"))
    wells = [str(w) for w in z["wells"]]
    draws = {w: z["draws/" + w] for w in wells}
    map_dict, ens_dict, extras = prep.propagate_obs_ensembles(draws)
    assert abs(extras["Rterr"] - float(z["Rterr"])) <= 1e-14 * extras["Rterr"]
    n = 0
    for k in z.files:
        parts = k.split("/")
        if parts[0] == "ens":
            got = np.asarray(ens_dict[parts[1]][parts[2]]).ravel()
        elif parts[0] == "map":
            got = np.atleast_1d(np.asarray(map_dict[parts[1]].loc[parts[2]])).ravel()
        elif parts[0] == "marg":
            key = min(extras["he3_ens_marg"], key=lambda r: abs(r - float(parts[1])))
            got = np.asarray(extras["he3_ens_marg"][key][parts[2]]).ravel()
        else:
            continue
        want = z[k]
        assert got.shape == want.shape, k
        assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)) < 1e-10, k
        n += 1
    assert n >= 60
    # the structure the age inversion reads (run_age_mcmc_utils.py:353-356)
    mu, sd = prep.observation_moments(ens_dict, "PLM1", ["CFC12", "SF6", "H3", "He4_ter"], {"CFC12": 0.05, "SF6": 0.05, "H3": 0.05, "He4_ter": 0.05})
    assert mu.shape == (4,) and np.all(sd > 0)
    assert list(ens_dict["CFC"]["PLM1"].columns) == ["CFC11", "CFC12", "CFC113"] and ens_dict["CFC12"]["PLM6"].shape == (64, 1)
    # full size: 50,000 members per well in one call per (well, quantity)
    rng = np.random.default_rng(3)
    big = {w: draws[w][rng.integers(0, 64, 60000)] * rng.uniform(0.98, 1.02, (60000, 6)) for w in wells}
    _, ens_big, _ = prep.propagate_obs_ensembles(big, marginal=False, as_frames=False)
    assert ens_big["SF6"]["PLM7"].shape == (50000,) and np.isfinite(ens_big["He3"]["PLM1"]).all()


def test_salinity_goldens_on_device():
    """S != 0: Setchenow terms of the noble-gas solubilities (utils/noble_gas_utils.py:145-159) and the salinity terms of the
    CFC / SF6 solubilities (utils/cfc_utils.py:62-83,196-209) against the untouched reference."""
    from helpers import rel_err
    from noblegas_rtd_mcmc_b200.cfc_utils import cfc_ce_corr, sf6_ce_corr
    from noblegas_rtd_mcmc_b200.noble_gas_utils import noble_gas_fun
    z = np.load(os.path.join(ROOT, "tests", "golden", "ce_salinity.npz"))
    gases = ["He", "Ne", "Ar", "Kr", "Xe"]
    for S in (5.0, 35.0):
        m = z["S"] == S
        E, T, Ae, F = z["E"][m], z["T"][m], z["Ae"][m], z["F"][m]
        ng = noble_gas_fun(gases=gases, E=E, T=T, Ae=Ae, F=F, P="lapse_rate", S=S)
        for j, g in enumerate(gases):
            assert rel_err(ng.solubility(g), z["K"][m][:, j]) < 1e-12, (S, g)
        ce, dry, wet = ng.ce_exc(True), ng.equil_conc_dry(), ng.equil_conc()
        for j, g in enumerate(gases):
            assert rel_err(ce[g], z["ce_true"][m][:, j]) < 1e-12 and rel_err(dry[g], z["eq_dry"][m][:, j]) < 1e-12
            assert rel_err(wet[g], z["eq_wet"][m][:, j]) < 1e-12
        Tc = np.minimum(T, 30.0)
        c = cfc_ce_corr(cfc_num=[11, 12, 113], E=E, T=Tc, Ae=Ae, F=F, S=S)
        assert rel_err(c.solubility_cfc(), z["cfc_K"][m]) < 1e-12 and rel_err(c.equil_air_conc_cfc(z["Cm"][m]), z["cfc_air"][m]) < 1e-12
        s6 = sf6_ce_corr(E=E, T=Tc, Ae=Ae, F=F, S=S)
        assert rel_err(s6.solubility_sf6(), z["sf6_K"][m]) < 1e-12 and rel_err(s6.equil_air_conc_sf6(z["Cs"][m]), z["sf6_air"][m]) < 1e-12


def test_g_tp_holds_the_decayed_weights():
    """convolve() leaves in .g_tp what the reference stores there (utils/convolution_integral_utils.py:300-316): the
    normalised weights times exp(-lam tp) (or 1 - exp(-lam tp) for 3He); materialised lazily; C_i == dot(flip(C), g_tp)."""
    import pandas as pd
    from helpers import load_c_in
    from noblegas_rtd_mcmc_b200.convolution_integral_utils import tracer_conv_integral
    C = load_c_in(600)
    idx = np.arange(599, -1, -1, dtype=float)
    df = pd.DataFrame({"H3_tu": C["H3"][::-1]}, index=idx)
    for ra in (False, "3He"):
        m = tracer_conv_integral(df, 0)
        m.update_pars(tau=37.5, mod_type="exp_pist_flow", eta=1.5, t_half=12.34, rad_accum=ra)
        c = m.convolve()
        g = m.g_tp
        tp = np.arange(600, dtype=float); tp[0] += 1e-5
        lam = np.log(2) / 12.34
        m2 = tracer_conv_integral(df, 0)
        m2.update_pars(tau=37.5, mod_type="exp_pist_flow", eta=1.5)
        g0 = m2.gen_g_tp()
        want = g0 * (1 - np.exp(-lam * tp)) if ra == "3He" else g0 * np.exp(-lam * tp)
        assert np.allclose(g, want, rtol=1e-12, atol=0)
        assert abs(np.dot(C["H3"], g) - c) <= 1e-11 * abs(c)
        # external weights: stored after the decay as well
        m3 = tracer_conv_integral(df, 0)
        m3.update_pars(tau=37.5, mod_type="exp_pist_flow", eta=1.5, t_half=12.34, rad_accum=ra)
        c3 = m3.convolve(g_tau=g0)
        assert np.allclose(m3.g_tp, want, rtol=1e-12) and abs(c3 - c) <= 1e-10 * abs(c)
    # the plan cache is keyed on the CONTENT of the series: an in-place edit of C_t is seen by the next convolve()
    m = tracer_conv_integral(df.copy(), 0)
    m.update_pars(tau=20.0, mod_type="exponential")
    a = m.convolve()
    m.C_t.iloc[:, 0] *= 2.0
    assert abs(m.convolve() - 2.0 * a) <= 1e-12 * abs(a)


_NG_STATE_SCRIPT = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from noblegas_rtd_mcmc_b200.noble_gas_mcmc import mcmc_model
from noblegas_rtd_mcmc_b200.sampler import Sampler
fx = json.load(open(os.path.join(sys.argv[1], "noblegas_rtd_mcmc_b200", "data", "ng_obs_plm.json")))["wells"]["PLM1"]
mdl = mcmc_model(fx["obs"], mcmc_model.well_elev["PLM1"])
out = {}
for name, kw in (("own", dict(hist_cap=96)), ("pool", dict(hist_cap=64))):
    smp = Sampler(mdl.build_priors(), mdl.obs_mu, mdl.obs_sd, 777, plan=None, gases=mdl.gases, lik="studentt",
                  nu_range=(1.0, 30.0), tune_interval=50, seed=5, **kw)
    if name == "pool":
        smp.set_population(111)
    for _ in range(6):
        smp.run(37, tune=True)              # ring wraps (cap < steps), launches of odd length
    smp.stop_tuning()
    tr = smp.run(40, tune=False, record=True, keep_trace=True)
    out[name + "_q"] = smp.get("q").cpu().numpy(); out[name + "_trace"] = tr.cpu().numpy()
    out[name + "_mean"] = smp.get("mean").cpu().numpy(); out[name + "_lamb"] = smp.get("lamb").cpu().numpy()
    smp.close()
np.savez(sys.argv[2], **out)
"""


def test_ng_register_resident_kernel_is_bitwise_the_generic_one(tmp_path):
    """k_mcmc_ng_r<6> (state in registers, next step's history rows fetched by cp.async one step ahead) against the
    run-time-nd kernel k_mcmc_ng (NGRTD_NG_GENERIC=1): identical positions, trace, moments and tuned lambda, with the
    chain's own history (incl. the hazard of selecting the row just appended, and ring wrap-around) and with a shared
    population archive."""
    import sys
    res = {}
    for tag, env in (("r", {}), ("generic", {"NGRTD_NG_GENERIC": "1"})):
        path = str(tmp_path / (tag + ".npz"))
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, "-c", _NG_STATE_SCRIPT, ROOT, path], check=True, env=e, timeout=600)
        res[tag] = np.load(path)
    for k in res["r"].files:
        assert np.array_equal(res["r"][k], res["generic"][k]), k
    assert np.isfinite(res["r"]["own_trace"]).all() and res["r"]["own_trace"].std() > 0


def test_library_host_buffers_pinned_and_write_combined():
    """ngrtd_host_alloc: page-locked (optionally write-combined) buffers for the *_host calls; the blocking call reads them in
    place, the submit / wait form copies from them -- same numbers as pageable numpy arrays."""
    import gc
    from noblegas_rtd_mcmc_b200 import _lib, synthetic
    pn = list(synthetic.PAR_NAMES_CFG3)
    X, descs = synthetic.series_matrix_and_descs(pn)
    plan = _lib.Plan(X, descs, "exp_pist_flow", "dispersion")
    th = synthetic.theta_cfg3_informative(5000, 3)
    obs = np.ones(len(descs)); sd = 0.05 * np.ones(len(descs))
    want = plan.forward_loglik_host(th, pn, obs, sd, "normal")
    for wc in (False, True):
        buf = _lib.host_array(th.shape, write_combined=wc)
        assert buf.shape == th.shape and buf.dtype == np.float64
        buf[...] = th
        out = _lib.host_array(len(th))
        got = plan.forward_loglik_host(buf, pn, obs, sd, "normal", logp_out=out)
        assert np.allclose(got, want, rtol=1e-12, atol=0, equal_nan=True) and np.isfinite(want).any()
        out2 = _lib.host_array(len(th))
        plan.forward_loglik_host_submit(buf, pn, obs, sd, "normal", logp_out=out2, slot=1)
        plan.host_wait(1)
        assert np.allclose(out2, want, rtol=1e-12, atol=0, equal_nan=True)
        del buf, out, out2, got
        gc.collect()                                     # blocks are freed with their last view
    with pytest.raises(_lib.NgrtdError):
        _lib.check(_lib.lib.ngrtd_host_alloc(None, 8, 0))
