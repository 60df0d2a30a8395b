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
