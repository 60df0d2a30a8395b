"""Development probe: exact (importance-sampling) posterior of the BMM inversions vs the device sampler with per-chain and
shared DE-MC-Z archives."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import bmm_common as C
from noblegas_rtd_mcmc_b200.sampler import Sampler
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for model1 in ("exp_pist_flow", "exponential"):
    plan, pn, pri, obs, sd, J_mu = C.setup(model1)
    t0 = time.time()
    th, w, ess = C.exact_posterior(plan, pn, obs, sd, J_mu, n_batches=nb)
    print(model1, "IS: kept", len(w), "ESS %.0f" % ess, "time %.1f s" % (time.time() - t0), flush=True)
    names = pn + ["nu_"]
    ex = {n: (np.sum(w * th[:, i]), np.sqrt(np.sum(w * (th[:, i] - np.sum(w * th[:, i])) ** 2))) for i, n in enumerate(names)}
    for pool in (0, 2048):
        smp = Sampler(pri, obs, sd, 2048, plan=plan, lik="studentt", nu_range=(5.0, 30.0), f2_from_f1=True, tune_interval=1000,
                      hist_cap=4096, seed=123423)
        smp.set_population(pool)
        t0 = time.time()
        for _ in range(20):
            smp.run(500, tune=True)
        smp.stop_tuning()
        parts = [smp.run(500, tune=False, record=True, thin=10, keep_trace=True) for _ in range(20)]
        tr = torch.cat(parts, 0).cpu().numpy()
        dt = time.time() - t0
        from noblegas_rtd_mcmc_b200 import distributed as D
        print("  pool", pool, "time %.1f s" % dt, "r_hat", np.round(D.pooled_summary(smp, 10000)["r_hat"], 3), "acc", float(smp.get("accepted").mean()) / 20000)
        snames = [q["target"] for q in pri]
        for n in names:
            if n == "f2": continue
            a = tr[:, :, snames.index(n)].ravel()
            i = names.index(n)
            qs = np.percentile(a, [5, 25, 50, 75, 95])
            F = C.weighted_cdf(th[:, i], w, qs)
            cm = tr[:, :, snames.index(n)].mean(axis=0)
            print("    %-10s exact %10.4g +- %9.4g | sampler %10.4g +- %9.4g | exact CDF at sampler quantiles %s | chain-mean spread %.3g"
                  % (n, ex[n][0], ex[n][1], a.mean(), a.std(), np.round(F, 3), cm.std()))
        smp.close()
