"""CPU: the sampler restatement (oracle/np_sampler.py) -- Philox known answers, transforms against scipy.stats,
tuning rule, diagnostics against analytic cases, and the multi-rank statistics gather over gloo (world_size 2)."""
import os
import sys

import numpy as np
import pytest

import np_sampler as S


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    assert S.philox4x32_10([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert S.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert S.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = [S.u01(*S.chain_rng(7, c, s, 0)[:2]) for c in range(50) for s in range(50)]
    assert 0.0 < min(u) and max(u) < 1.0 and abs(np.mean(u) - 0.5) < 0.02


def test_transforms_are_densities_in_transformed_space():
    """log prior + log|Jacobian| of each transform equals scipy's logpdf(v) + log|dv/dx| (numerical derivative)."""
    from scipy import stats
    cases = [(dict(kind="uniform", p0=1.0, p1=1000.0), lambda v: stats.uniform.logpdf(v, 1.0, 999.0)),
             (dict(kind="beta", p0=2.0, p1=4.0, lo=2776.9, hi=3300.0),
              lambda v: stats.beta.logpdf((v - 2776.9) / (3300 - 2776.9), 2, 4)),     # pymc3 keeps the density of the [0,1] variable
             (dict(kind="normal", p0=-146.0, p1=17.0), lambda v: stats.norm.logpdf(v, -146, 17)),
             (dict(kind="halfnormal", p0=0.5 / 3), lambda v: stats.halfnorm.logpdf(v, scale=0.5 / 3))]
    for pr, logpdf in cases:
        for x in (-3.0, -0.4, 0.0, 0.7, 2.5):
            xx = x if pr["kind"] != "normal" else -146.0 + 10 * x
            v, lp = S.transform_dim(pr, xx)
            h = 1e-6
            dv = (S.transform_dim(pr, xx + h)[0] - S.transform_dim(pr, xx - h)[0]) / (2 * h)
            if pr["kind"] == "beta":
                dv = dv / (pr["hi"] - pr["lo"])
            assert abs(lp - (logpdf(v) + np.log(abs(dv)))) < 1e-6, (pr["kind"], x)


def test_test_point_and_tune_rule():
    pri = [dict(kind="uniform", p0=1, p1=1000), dict(kind="beta", p0=2.0, p1=0.1, lo=0, hi=1),
           dict(kind="normal", p0=-10.4, p1=0.33), dict(kind="halfnormal", p0=0.5 / 3)]
    q = S.test_point(pri)
    assert q[0] == 0.0 and abs(q[1] - np.log((2 / 2.1) / (0.1 / 2.1))) < 1e-12 and q[2] == -10.4
    assert abs(np.exp(q[3]) - (0.5 / 3) * np.sqrt(2 / np.pi)) < 1e-15
    assert [S.tune_factor(r) for r in (0.0005, 0.01, 0.1, 0.3, 0.6, 0.8, 0.99)] == [0.1, 0.5, 0.9, 1.0, 1.1, 2.0, 10.0]


def test_oracle_sampler_recovers_gaussian_posterior():
    """DE-MC-Z restatement on a conjugate problem: Normal prior x Normal likelihood."""
    pri = [dict(kind="normal", target="x", p0=0.0, p1=2.0)]
    obs, sd = 1.5, 0.5
    post_var = 1 / (1 / 4.0 + 1 / 0.25)
    post_mu = post_var * (obs / 0.25)
    Q, LP, AC = S.run_chain(pri, lambda v: S.normal_logp(np.array([obs]), np.array([v["x"]]), np.array([sd])),
                            6000, seed=5, chain=0, tune_steps=2000, tune_interval=500)
    x = Q[2000:, 0]
    assert abs(x.mean() - post_mu) < 0.08 and abs(x.std() - np.sqrt(post_var)) < 0.08
    assert 0.1 < AC[2000:].mean() < 0.9


def test_diagnostics_on_known_processes():
    from noblegas_rtd_mcmc_b200 import diagnostics as D
    rng = np.random.default_rng(0)
    iid = rng.normal(size=(4, 5000))
    assert abs(D.rhat(iid) - 1.0) < 0.01
    assert 0.8 * 20000 < D.ess_bulk(iid) < 1.2 * 20000
    phi = 0.9                                                   # AR(1): ESS = N (1-phi)/(1+phi)
    ar = np.zeros((4, 20000))
    e = rng.normal(size=ar.shape)
    for t in range(1, ar.shape[1]):
        ar[:, t] = phi * ar[:, t - 1] + e[:, t]
    want = 80000 * (1 - phi) / (1 + phi)
    assert 0.75 * want < D.ess_mean(ar) < 1.3 * want
    shifted = iid + np.arange(4)[:, None]
    assert D.rhat(shifted) > 1.3
    lo, hi = D.hdi(rng.normal(size=200000))
    assert abs(lo + 1.88) < 0.05 and abs(hi - 1.88) < 0.05
    # moment-based estimators agree with the trace-based ones when chains are many
    many = rng.normal(3.0, 2.0, size=(2000, 50))
    ms = D.moments_summary(50.0, many.mean(axis=1, keepdims=True), ((many - many.mean(axis=1, keepdims=True)) ** 2).sum(axis=1, keepdims=True))
    assert abs(ms["mean"][0] - 3.0) < 0.02 and abs(ms["sd"][0] - 2.0) < 0.03 and abs(ms["r_hat"][0] - 1.0) < 0.02
    assert 0.8 * 100000 < ms["ess"][0] < 1.25 * 100000


def test_shard_partition():
    from noblegas_rtd_mcmc_b200.distributed import shard
    for total, world in ((65536, 8), (10, 3), (7, 8), (1048576, 8)):
        parts = [shard(total, r, world) for r in range(world)]
        assert sum(c for _, c in parts) == total
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from noblegas_rtd_mcmc_b200.distributed import gather_chain_stats, global_summary, shard
    rng = np.random.default_rng(123)
    total = 11
    mean_all = rng.normal(size=(total, 3))
    m2_all = rng.uniform(1, 2, size=(total, 3))
    off, cnt = shard(total, rank, world)
    m, v = gather_chain_stats(torch.from_numpy(mean_all[off:off + cnt]), torch.from_numpy(m2_all[off:off + cnt]))
    ok = np.array_equal(m.numpy(), mean_all) and np.array_equal(v.numpy(), m2_all)
    s = global_summary(20, torch.from_numpy(mean_all[off:off + cnt]), torch.from_numpy(m2_all[off:off + cnt]))
    q.put((rank, ok, float(s["r_hat"][0]), int(s["chains"])))
    dist.destroy_process_group()


def test_chain_stats_gather_gloo_world2():
    """N > 1 path on CPU: ragged shards (6 + 5 chains) all-gathered over gloo, identical summary on every rank."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == res[1][2] and res[0][3] == 11


def _gloo_pooled_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from noblegas_rtd_mcmc_b200.distributed import allreduce_pooled, shard, summary_from_pooled
    rng = np.random.default_rng(77)
    total, nd = 13, 3
    mean_all = rng.normal(size=(total, nd))
    m2_all = rng.uniform(1, 2, size=(total, nd))
    off, cnt = shard(total, rank, world)
    m, v = mean_all[off:off + cnt], m2_all[off:off + cnt]
    # what ngrtd_sampler_pooled_moments leaves on the device of a rank: [sum mean, sum mean^2, sum M2, chains]
    vec = torch.from_numpy(np.concatenate([m.sum(0), (m * m).sum(0), v.sum(0), [float(cnt)]]))
    s = summary_from_pooled(20, allreduce_pooled(vec).numpy())
    q.put((rank, s["mean"].tolist(), s["r_hat"].tolist(), s["ess"].tolist(), s["chains"]))
    dist.destroy_process_group()


def test_pooled_summary_gloo_world2():
    """K6 across ranks on CPU: one all-reduce of 3*nd + 1 pooled moments gives, on every rank, the summary that the
    all-gathered per-chain moments give (diagnostics.moments_summary)."""
    import socket
    import torch.multiprocessing as mp
    from noblegas_rtd_mcmc_b200 import diagnostics as dg
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_pooled_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res[0][1:] == res[1][1:] and res[0][4] == 13
    rng = np.random.default_rng(77)
    mean_all = rng.normal(size=(13, 3))
    m2_all = rng.uniform(1, 2, size=(13, 3))
    ref = dg.moments_summary(20.0, mean_all, m2_all)
    assert np.allclose(res[0][1], ref["mean"], rtol=1e-13)
    assert np.allclose(res[0][2], ref["r_hat"], rtol=1e-12)


def test_trace_npz_roundtrip(tmp_path):
    from noblegas_rtd_mcmc_b200 import diagnostics as dg
    rng = np.random.default_rng(3)
    post = {"tau1": rng.normal(size=(4, 50)), "f1": rng.uniform(size=(4, 50))}
    dg.save_trace(tmp_path / "t.npz", post, sample_stats={"accepted": rng.integers(0, 2, (4, 50))}, attrs={"sampling_time": 1.25})
    tr = dg.load_trace(tmp_path / "t.npz")
    assert set(tr["posterior"]) == {"tau1", "f1"} and tr["attrs"]["sampling_time"] == 1.25
    for k in post:
        assert np.array_equal(tr["posterior"][k], post[k])
    assert np.allclose(dg.summary(tr["posterior"])["tau1"]["mean"], post["tau1"].mean())
    with pytest.raises(ValueError):
        dg.save_trace(tmp_path / "bad.npz", {"x": np.zeros(5)})


def test_summary_columns_match_reference_tables(tmp_path):
    """diagnostics.summary carries the columns of the reference's ng_optPLM*.csv (az.summary + median), in order; for
    iid draws ESS ~ N and MCSE follow their textbook values."""
    from noblegas_rtd_mcmc_b200 import diagnostics as dg
    assert dg.SUMMARY_COLUMNS == ("mean", "sd", "hdi_3%", "hdi_97%", "mcse_mean", "mcse_sd", "ess_bulk", "ess_tail", "r_hat", "median")
    rng = np.random.default_rng(11)
    x = rng.normal(3.0, 2.0, size=(4, 2500))
    row = dg.summary({"x": x})["x"]
    assert tuple(row) == dg.SUMMARY_COLUMNS
    n = x.size
    assert 0.7 * n < row["ess_bulk"] < 1.3 * n and 0.6 * n < row["ess_tail"] < 1.4 * n
    assert abs(row["mcse_mean"] / (2.0 / np.sqrt(n)) - 1) < 0.2
    assert abs(row["mcse_sd"] / (2.0 / np.sqrt(2 * n)) - 1) < 0.35
    assert abs(row["r_hat"] - 1) < 0.01
    # an AR(1) chain has far fewer effective draws
    y = np.zeros((4, 2500))
    e = rng.normal(size=y.shape)
    for t in range(1, y.shape[1]):
        y[:, t] = 0.9 * y[:, t - 1] + e[:, t]
    r2 = dg.summary({"y": y})["y"]
    assert r2["ess_bulk"] < 0.15 * n and r2["ess_tail"] < 0.5 * n
    dg.summary_csv(tmp_path / "s.csv", {"x": x})
    head = open(tmp_path / "s.csv").readline().strip()
    assert head == ",mean,sd,hdi_3%,hdi_97%,mcse_mean,mcse_sd,ess_bulk,ess_tail,r_hat,median"


def test_bind_to_gpu_numa_is_conservative(tmp_path):
    """distributed.bind_to_gpu_numa acts only on a real multi-node topology: strict subset of the allowed CPUs and a
    NUMA node >= 0; everything else leaves the affinity untouched."""
    import os
    from noblegas_rtd_mcmc_b200 import distributed as D
    before = os.sched_getaffinity(0)
    try:
        assert D._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
        assert D.bind_to_gpu_numa("0000:ff:1f.0", sysfs_root=str(tmp_path)).startswith("unchanged")      # no sysfs entry
        dev = tmp_path / "0000:c0:00.0"
        dev.mkdir()
        allcpus = ",".join(str(c) for c in sorted(before))
        (dev / "numa_node").write_text("-1\n"); (dev / "local_cpulist").write_text(allcpus + "\n")
        assert D.bind_to_gpu_numa("0000:C0:00.0", sysfs_root=str(tmp_path)).startswith("unchanged")      # VM: node -1
        (dev / "numa_node").write_text("0\n")
        assert D.bind_to_gpu_numa("0000:c0:00.0", sysfs_root=str(tmp_path)).startswith("unchanged")      # single node
        assert os.sched_getaffinity(0) == before
        if len(before) >= 2:
            keep = sorted(before)[: len(before) // 2]
            (dev / "numa_node").write_text("1\n"); (dev / "local_cpulist").write_text(",".join(map(str, keep)) + ",9999\n")
            msg = D.bind_to_gpu_numa("0000:c0:00.0", sysfs_root=str(tmp_path))
            assert msg.startswith("bound to numa node 1") and os.sched_getaffinity(0) == set(keep)
    finally:
        os.sched_setaffinity(0, before)
