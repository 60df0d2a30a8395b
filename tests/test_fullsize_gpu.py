"""GPU: parity at BASELINE.json's full sizes through size-independent properties (batch-split invariance, mixture and
input linearity, J scaling, eta=1 identity) plus oracle spot checks on random subsets, for cfg 3 (65,536 chains, L=840),
the real yearly series (L=25,256, chunk-streamed path) and a cfg-5 style long lag axis (L=10,000)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_subset(X, descs, m1, m2, theta, pn, idx):
    import c_oracle
    return c_oracle.forward(X, descs, m1, m2, theta[idx], pn)


def test_cfg3_full_batch_properties():
    from helpers import rel_err, synth_descs
    from noblegas_rtd_mcmc_b200 import _lib, synthetic
    pn = list(synthetic.PAR_NAMES_CFG3)
    X, descs = synth_descs(pn)
    plan = _lib.Plan(X, descs, "exp_pist_flow", "dispersion")
    B = 65536
    theta = synthetic.theta_cfg3_informative(B, 9)
    theta[: B // 2] = synthetic.theta_cfg3(B // 2, 9)             # half from the wide prior (NaN-producing region)
    full = plan.forward_host(theta, pn)
    # (i) batch-split invariance (ragged splits, not multiples of the 16-chain unit).  r2: the tape schedule of k_forward
    # cuts a unit's lag range between warps at positions that depend on the batch size, so a chain's sums are associated
    # differently in a different batch: identical NaN pattern, values equal to rounding (a repeated call is bitwise equal)
    cuts = [0, 7, 4099, 33333, B]
    parts = np.concatenate([plan.forward_host(theta[a:b], pn) for a, b in zip(cuts[:-1], cuts[1:])], axis=0)
    assert rel_err(full, parts) < 1e-13
    assert np.array_equal(full, plan.forward_host(theta, pn), equal_nan=True)
    # (ii) mixture linearity: f1*EPM + f2*DM with single-component plans
    p1 = _lib.Plan(X, descs, "exp_pist_flow", False)
    p2 = _lib.Plan(X, descs, "dispersion", False)
    c1 = p1.forward_host(theta[:, [0, 4, 6]], ["tau1", "eta1", "J"])
    c2 = p2.forward_host(theta[:, [1, 5, 6]], ["tau1", "D1", "J"])
    mix = theta[:, [2]] * c1 + theta[:, [3]] * c2
    assert rel_err(full, mix) < 1e-12
    # (iii) input linearity: doubling every series doubles every series-driven tracer exactly
    plan2 = _lib.Plan(2.0 * X, descs, "exp_pist_flow", "dispersion")
    dbl = plan2.forward_host(theta[:8192], pn)
    assert np.array_equal(dbl[:, :6], 2.0 * plan.forward_host(theta[:8192], pn)[:, :6], equal_nan=True)   # same batch: bitwise
    # (iv) He4_ter is proportional to J = 10**theta_J
    th2 = theta[:8192].copy()
    th2[:, 6] += np.log10(4.0)
    he = plan.forward_host(th2, pn)[:, 6]
    assert rel_err(he, 4.0 * full[:8192, 6]) < 2e-13
    # (v) oracle spot check on a random subset (C port of the reference arithmetic)
    idx = np.random.default_rng(0).choice(B, 768, replace=False)
    want = _oracle_subset(X, descs, "exp_pist_flow", "dispersion", theta, pn, idx)
    assert rel_err(full[idx], want) < 1e-10
    assert np.isnan(full).any() and np.isfinite(full).any()


def test_real_series_full_length_properties():
    """L = 25,256 (reference C_in_dict), 4,096 chains of the shipped `.123` EPM+PFM configuration."""
    from helpers import MODEL_CFGS, load_c_in, real_plan, rel_err
    m1, m2, pn = MODEL_CFGS["epm_pfm123"]
    tracers = ["CFC12", "SF6", "H3", "He4_ter", "He3", "CFC11", "CFC113"]
    plan, C = real_plan(m1, m2, pn, tracers)
    rng = np.random.default_rng(4)
    B = 4096
    f1 = rng.uniform(0.01, 0.99, B)
    theta = np.stack([rng.uniform(1, 1000, B), rng.uniform(50, 15000, B), f1, 1 - f1, rng.uniform(1, 5, B),
                      rng.normal(-10.42, 0.33, B), rng.uniform(5, 35, B), np.abs(rng.normal(0, 0.17, B))], axis=1)
    full = plan.forward_host(theta, pn)
    halves = np.concatenate([plan.forward_host(theta[:1111], pn), plan.forward_host(theta[1111:], pn)], axis=0)
    assert rel_err(full, halves) < 1e-13           # equal to rounding (batch-size dependent association, see the cfg-3 test)
    # eta1 = 1 makes exp_pist_flow identical to exponential (verified on the reference)
    th1 = theta.copy()
    th1[:, 4] = 1.0
    a = plan.forward_host(th1, pn)
    plan_e, _ = real_plan("exponential", "piston", [p for p in pn if p != "eta1"], tracers)
    b = plan_e.forward_host(np.delete(th1, 4, axis=1), [p for p in pn if p != "eta1"])
    assert np.array_equal(a, b, equal_nan=True)
    # oracle spot check
    import c_oracle
    names = ["CFC11", "CFC12", "CFC113", "SF6", "H3"]
    X = np.stack([C[n] for n in names], axis=1)
    from test_oracle_golden import _real_descs
    _, descs = _real_descs(tracers, pn)
    idx = rng.choice(B, 96, replace=False)
    want = c_oracle.forward(X, descs, m1, m2, theta[idx], pn)
    assert rel_err(full[idx], want) < 1e-10


def test_cfg5_long_lag_axis_mixture():
    """cfg 5: L = 10,000 lags, EPM + dispersion mixture with 4He accumulation (chunk-streamed, 10 chunks)."""
    import c_oracle
    from helpers import rel_err
    from noblegas_rtd_mcmc_b200 import _lib
    L = 10000
    rng = np.random.default_rng(8)
    k = np.arange(L)
    X = np.stack([5 + 2000 * np.exp(-((k - 60.0) / 8.0) ** 2), 500 / (1 + np.exp((k - 50.0) / 8.0)) + 1e-10,
                  9 * np.exp(-k / 25.0) + 1e-10], axis=1) * rng.lognormal(0, 0.05, (L, 3))
    descs = [dict(series=0, lam=np.log(2) / 12.34), dict(series=0, lam=np.log(2) / 12.34, rad_accum="3He"),
             dict(series=1), dict(series=2, use_lamsf6=True), dict(series=-1, rad_accum="4He")]
    pn = ["tau1", "tau2", "f1", "f2", "eta1", "D2", "J", "lamsf6"]
    plan = _lib.Plan(X, descs, "exp_pist_flow", "dispersion")
    B = 2048
    f1 = rng.uniform(0.01, 0.99, B)
    theta = np.stack([rng.uniform(1, 1000, B), rng.uniform(50, 15000, B), f1, 1 - f1, rng.uniform(1, 5, B),
                      rng.uniform(0.01, 2, B), rng.normal(-10.42, 0.33, B), np.abs(rng.normal(0, 0.17, B))], axis=1)
    out = plan.forward_host(theta, pn)
    want = c_oracle.forward(X, descs, "exp_pist_flow", "dispersion", theta, pn)
    assert rel_err(out, want) < 1e-10
    assert np.isfinite(out).mean() > 0.9


def test_posterior_predictive_batch_matches_loop(tmp_path, monkeypatch):
    """the posterior-predictive sweep of run_age_mcmc.py:243-319 as one batched launch."""
    import pandas as pd
    from helpers import load_c_in
    from noblegas_rtd_mcmc_b200.convolution_integral_utils import tracer_conv_integral
    from noblegas_rtd_mcmc_b200.run_age_mcmc_utils import conv_mcmc
    C = load_c_in(3000)
    L = 3000
    df = lambda v, n: pd.DataFrame({n: v[::-1]}, index=np.arange(L - 1, -1, -1))
    ckw = {"mod_type1": "dispersion", "mod_type2": False, "CFC12": dict(C_t=df(C["CFC12"], "CFC12")),
           "H3": dict(C_t=df(C["H3"], "H3_tu"), t_half=12.34)}
    okw = {t: dict(obs_df=np.array([v, v * 1.02, v * 0.98]), obs_perr=0.05) for t, v in (("CFC12", 36.4), ("H3", 4.87))}
    pkw = dict(tau1_low=1.0, tau1_high=1000.0, D1_low=0.01, D1_high=2.0, par_names=["tau1", "D1"])
    monkeypatch.chdir(tmp_path)
    mc = conv_mcmc("PLM1", ["CFC12", "H3"], okw, ckw, pkw, "conv_traces", "0")
    idata = mc.sample_mcmc(chains=8, tune=1500, draws=500, tune_interval=250)
    post = idata["posterior"]
    # the finished trace is written where the reference writes it (run_age_mcmc_utils.py:242-256,425), as NetCDF-4
    from noblegas_rtd_mcmc_b200 import diagnostics
    assert mc.trace_name == "./conv_traces/PLM1.CFC12.H3.dispersion.0.netcdf"
    back = diagnostics.load_trace(mc.trace_name)
    assert np.array_equal(back["posterior"]["tau1"], post["tau1"]) and np.array_equal(back["posterior"]["nu"], post["nu"])
    assert abs(float(np.ravel(back["attrs"]["sampling_time"])[0]) - mc.sampling_time) < 1e-12
    assert post["tau1"].shape == (8, 500) and np.all((post["tau1"] > 1) & (post["tau1"] < 1000))
    assert np.all((post["nu"] >= 5) & (post["nu"] <= 30))
    pp = mc.posterior_predictive(chain=0)
    assert pp["CFC12"].shape == (500,)
    m = tracer_conv_integral(ckw["H3"]["C_t"], 0)
    for s in (0, 17, 499):                                              # the reference's per-draw loop, 3 draws
        m.update_pars(tau=post["tau1"][0, s], mod_type="dispersion", t_half=12.34, D=post["D1"][0, s], bbar=False, Phi_im=False)
        assert abs(m.convolve() - pp["H3"][s]) <= 1e-12 * abs(pp["H3"][s])
    # ... and against the oracle (numpy restatement of the reference), every draw of chain 0
    import np_oracle as O
    th0 = np.stack([post["tau1"][0], post["D1"][0]], axis=1)
    for t, key, th in (("CFC12", "CFC12", False), ("H3", "H3", 12.34)):
        want = O.forward_mod(th0, ["tau1", "D1"], t, C[key], "dispersion", False, t_half=th)
        assert np.max(np.abs(pp[t] - want) / np.abs(want)) < 1e-10, t
    mu, err = mc.observations()
    assert abs(np.median(pp["CFC12"]) - mu[0]) < 4 * err[0]          # the fit explains the observation


def test_pinned_host_zero_copy_and_staging_modes_are_bitwise_identical(monkeypatch):
    from helpers import rel_err
    """The *_host entry points read pinned (mapped) host buffers straight from the kernel (TMA bulk copy per 16-chain unit,
    next unit prefetched) and fall back to staged copies for pageable memory; the device entry points can stage theta the
    same way (NGRTD_STAGE).  Every route must give the same bits, at ragged batch sizes (partial last unit, odd row count
    x odd ndim => lane-load staging of the last unit) and with the per-chain nu of the Student-T likelihood."""
    import torch
    from helpers import synth_plan
    from noblegas_rtd_mcmc_b200 import synthetic
    pn = list(synthetic.PAR_NAMES_CFG3)
    plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
    obs = np.array([8.0, 40.0, 150.0, 300.0, 50.0, 5.0, 1e-8])
    sd = 0.05 * obs
    for B in (1, 15, 16, 37, 4099, 33333, 65536):
        theta = synthetic.theta_cfg3_informative(B, 21)
        theta[: B // 2] = synthetic.theta_cfg3(B // 2, 22)
        nu = np.random.default_rng(B).uniform(5.0, 30.0, B)
        monkeypatch.setenv("NGRTD_STAGE", "0")
        th_d = torch.from_numpy(theta).cuda()
        nu_d = torch.from_numpy(nu).cuda()
        model_d = torch.empty((B, 7), dtype=torch.float64, device="cuda")
        ref_t = plan.forward_loglik_dev(th_d, pn, obs, sd, "studentt", nu_t=nu_d, model_t=model_d).cpu().numpy()
        ref_n = plan.forward_loglik_dev(th_d, pn, obs, sd, "normal").cpu().numpy()
        ref_m = model_d.cpu().numpy()
        for stage in ("1", "2"):
            monkeypatch.setenv("NGRTD_STAGE", stage)
            model_d.zero_()
            got = plan.forward_loglik_dev(th_d, pn, obs, sd, "studentt", nu_t=nu_d, model_t=model_d).cpu().numpy()
            assert np.array_equal(got, ref_t, equal_nan=True), (B, stage)
            assert np.array_equal(model_d.cpu().numpy(), ref_m, equal_nan=True), (B, stage)
        monkeypatch.setenv("NGRTD_STAGE", "0")
        th_p = torch.from_numpy(theta).pin_memory()
        nu_p = torch.from_numpy(nu).pin_memory()
        lp_p = torch.empty(B, dtype=torch.float64).pin_memory()
        # one launch over the whole batch (zero-copy) is bitwise equal to the device call; the staged-copy pipeline splits
        # batches >= 32,768 chains into two launches, whose tape cuts differ -> equal to rounding there
        def same(a, b, split):
            if not split:
                return np.array_equal(a, b, equal_nan=True)
            return rel_err(a, b) < 1e-12                     # logp: sum of z^2 terms, z up to ~20
        for mode in ("mapped", "copy"):
            monkeypatch.setenv("NGRTD_HOST_MODE", mode)
            split = mode == "copy" and B >= 32768
            lp_p.fill_(7.0)
            plan.forward_loglik_host(th_p.numpy(), pn, obs, sd, "studentt", nu=nu_p.numpy(), logp_out=lp_p.numpy())
            assert same(lp_p.numpy(), ref_t, split), (B, mode)
            lp, model = plan.forward_loglik_host(th_p.numpy(), pn, obs, sd, "normal", want_model=True)   # pageable outputs
            assert same(lp, ref_n, split) and same(model, ref_m, split), (B, mode)
            assert same(plan.forward_host(th_p.numpy(), pn), ref_m, split), (B, mode)
        # pageable theta always takes the staged-copy pipeline
        assert same(plan.forward_loglik_host(theta, pn, obs, sd, "normal"), ref_n, B >= 32768)
        # a pinned slice whose rows start at an address that is not a multiple of 16 bytes (lane-load staging)
        if B > 16:
            flat = torch.empty(B * 7 + 1, dtype=torch.float64).pin_memory()
            flat[1:] = torch.from_numpy(theta).reshape(-1)
            odd = flat.numpy()[1:].reshape(B, 7)
            monkeypatch.setenv("NGRTD_HOST_MODE", "mapped")
            assert np.array_equal(plan.forward_loglik_host(odd, pn, obs, sd, "normal"), ref_n, equal_nan=True)


def test_host_submit_wait_pipeline_matches_sync_calls():
    """ngrtd_forward_loglik_host_submit / ngrtd_host_wait: several independent batches in flight (copy-in, kernel and
    copy-out of neighbouring batches overlap) give exactly the results of the synchronous call; a busy slot is refused."""
    import torch
    from helpers import synth_plan
    from noblegas_rtd_mcmc_b200 import _lib, synthetic
    pn = list(synthetic.PAR_NAMES_CFG3)
    plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
    obs = np.array([8.0, 40.0, 150.0, 300.0, 50.0, 5.0, 1e-8])
    sd = 0.05 * obs
    sizes = [65536, 4099, 33333, 16, 65536, 1, 20000, 65536]
    thetas = [torch.from_numpy(synthetic.theta_cfg3_informative(B, 100 + i)).pin_memory() for i, B in enumerate(sizes)]
    nus = [torch.from_numpy(np.random.default_rng(i).uniform(5.0, 30.0, B)).pin_memory() for i, B in enumerate(sizes)]
    want = [plan.forward_loglik_host(t.numpy(), pn, obs, sd, "studentt", nu=n.numpy(), want_model=True) for t, n in zip(thetas, nus)]
    outs = [torch.empty(B, dtype=torch.float64).pin_memory() for B in sizes]
    mods = [torch.empty((B, 7), dtype=torch.float64).pin_memory() for B in sizes]
    depth = plan.HOST_SLOTS
    for i in range(len(sizes) + depth):
        if i >= depth:
            k = i - depth
            plan.host_wait(k % depth)
            assert np.array_equal(outs[k].numpy(), want[k][0], equal_nan=True), k
            assert np.array_equal(mods[k].numpy(), want[k][1], equal_nan=True), k
        if i < len(sizes):
            plan.forward_loglik_host_submit(thetas[i].numpy(), pn, obs, sd, "studentt", nu=nus[i].numpy(),
                                            logp_out=outs[i].numpy(), model_out=mods[i].numpy(), slot=i % depth)
    plan.forward_loglik_host_submit(thetas[3].numpy(), pn, obs, sd, "normal", logp_out=outs[3].numpy(), slot=1)
    with pytest.raises(_lib.NgrtdError):
        plan.forward_loglik_host_submit(thetas[3].numpy(), pn, obs, sd, "normal", logp_out=outs[3].numpy(), slot=1)
    plan.host_wait(1)
    plan.host_wait(1)                                   # waiting on an idle slot is a no-op
    assert np.array_equal(outs[3].numpy(), plan.forward_loglik_host(thetas[3].numpy(), pn, obs, sd, "normal"), equal_nan=True)
    with pytest.raises(_lib.NgrtdError):
        plan.host_wait(99)
