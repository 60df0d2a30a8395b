"""BASELINE config 4: joint fit over all wells x recharge-ensemble members, ~1M chains sharded over the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 \
        examples/config4_joint_fit.py [total_chains=131072 x GPUs] [chains_per_group=256] [tune=10000] [draws=10000]

The reference runs ONE inversion per (well, tracer set) on the ensemble mean (run_age_mcmc.py:122-231, 296-324 s each).
Config 4 is the scale-out of that design: every ensemble member of every well gets its own population of chains, i.e.
group g = (well g % 3, member g // 3) is fitted to its own observation row; all groups advance in the same fused sampler
launches (k_mcmc_age, per-group observation rows: ngrtd_sampler_set_obs_groups).  Chains are keyed by global id, so the
result does not depend on the sharding; the only collective is the all-gather of per-chain moments at the end.

Model: exp_pist_flow, tracers CFC12 + SF6 + H3 + He4_ter on the reference's 25,256-lag yearly series, parameters
tau1, eta1, J, thalf_cfc, lamsf6 (+ nu), priors of run_age_mcmc.py:145-196.

r2: the whole chain of the reference runs on the device.  ens_dict.pk is a missing blob of the reference, so the members
are PRODUCED: (1) the closed-equilibrium noble-gas fit of every well (config 1, ng_interp/noble_gas_mcmc.py: 4 chains x
(10,000 + 50,000) steps); (2) `prep.propagate_obs_ensembles` turns those posteriors and the field observations into the
50,000-member observation ensembles of age_modeling_mcmc.prep.py:242-489 (ens_dict); (3) member m of well w supplies the
observation row of group (w, m), with the 5 % analytical error (SF6 of PLM6: 1000 %, run_age_mcmc.py:100-114).
Every group is one POPULATION of DE-MC-Z chains sharing an archive (ngrtd_sampler_set_population), so its chains find each
other's modes; the report is the R-hat of every group over its chains.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

from helpers import real_plan
from noblegas_rtd_mcmc_b200 import distributed as D
from noblegas_rtd_mcmc_b200 import prep
from noblegas_rtd_mcmc_b200.noble_gas_mcmc import mcmc_model
from noblegas_rtd_mcmc_b200 import noble_gas_utils as ng_utils
from noblegas_rtd_mcmc_b200.sampler import Sampler, prior

WELLS = ["PLM1", "PLM6", "PLM7"]
TRACERS = ["CFC12", "SF6", "H3", "He4_ter"]


def main():
    # default: 131,072 chains per GPU = BASELINE's 1,048,576 on the 8 GPUs of a box
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 131072 * int(os.environ.get("WORLD_SIZE", 1))
    cpg = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    # the reference runs 10,000 + 10,000 steps (run_age_mcmc_utils.py:416); the bimodal (tau1, eta1) posteriors of single
    # ensemble members need ~10x that for EVERY chain of a population to visit both modes (R-hat over chains < 1.05)
    tune = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
    draws = int(sys.argv[4]) if len(sys.argv) > 4 else 200000
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ngroups = total // cpg
    total = ngroups * cpg
    # ---- observation rows of all groups: config 1 -> prep -> ens_dict (identical on every rank: seeded) ----
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "age_traces.json")))["traces"]
    ngobs = json.load(open(os.path.join(ROOT, "noblegas_rtd_mcmc_b200", "data", "ng_obs_plm.json")))["wells"]
    t_pre = time.perf_counter()
    draws_ce = {}
    for w in WELLS:
        post = mcmc_model(ngobs[w]["obs"], mcmc_model.well_elev[w]).sample(chains=4, tune=10000, draws=50000)["posterior"]
        draws_ce[w] = np.stack([post[p].ravel() for p in prep.PARS], axis=1)             # prep.py:130
    _, ens_dict, _ = prep.propagate_obs_ensembles(draws_ce, marginal=False, as_frames=False)
    t_pre = time.perf_counter() - t_pre
    perr = {w: {"CFC12": 0.05, "SF6": 10.0 if w == "PLM6" else 0.05, "H3": 0.05, "He4_ter": 0.05} for w in WELLS}
    obs = np.empty((ngroups, len(TRACERS)))
    sd = np.empty_like(obs)
    for g in range(ngroups):
        w, m = WELLS[g % 3], g // 3
        obs[g] = np.abs([ens_dict[t][w][m % len(ens_dict[t][w])] for t in TRACERS])
        sd[g] = np.array([perr[w][t] for t in TRACERS]) * obs[g]
    # ---- model, priors (run_age_mcmc.py:145-196), sampler shard ----
    pn = ["tau1", "eta1", "J", "thalf_cfc", "lamsf6"]
    plan, _ = real_plan("exp_pist_flow", False, pn, TRACERS)
    J_mu = np.log10(ng_utils.J_flux(Del=1., rho_r=2700, rho_w=1000, U=3.7, Th=10.2, phi=0.05))
    pri = [prior("uniform", "tau1", 1.0, 1000.0), prior("beta", "nu_", 2.0, 0.1), prior("normal", "J", J_mu, 0.33),
           prior("uniform", "eta1", 1.0, 5.0), prior("beta", "thalf_cfc", 2.0, 2.0, lo=5.0, hi=35.0),
           prior("halfnormal", "lamsf6", 0.5 / 3)]
    off, cnt = D.shard(total, rank, world)
    smp = Sampler(pri, obs[0], sd[0], cnt, plan=plan, lik="studentt", nu_range=(5.0, 30.0), tune_interval=1000,
                  hist_cap=min(tune + draws, 2048), seed=123423, chain_offset=off, device=local)
    smp.set_obs_groups(obs, sd, cpg)
    smp.set_population(cpg)                    # one shared DE-MC-Z archive per (well, member) group; shards hold whole groups
    assert off % cpg == 0 and cnt % cpg == 0, "shards must hold whole populations"
    smp.run(8, tune=True)                      # first launch (module load) outside the timed region
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    LAUNCH = 500                               # the archive a launch reads is what the population had written before it
    for _ in range(tune // LAUNCH):
        smp.run(LAUNCH, tune=True)
    smp.stop_tuning()
    for _ in range(draws // LAUNCH):
        smp.run(LAUNCH, tune=False, record=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
    mean_all, m2_all = D.gather_chain_stats(smp.get("mean"), smp.get("m2"))        # the one collective: [total, ndim] x 2
    acc = smp.get("accepted").double().mean()
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        acc = acc / world
    if rank == 0:
        m = mean_all.cpu().numpy().reshape(ngroups, cpg, -1)
        v = (m2_all.cpu().numpy() / (draws - 1.0)).reshape(ngroups, cpg, -1)
        gmean = m.mean(axis=1)                                                   # posterior mean per group
        W, Bn = v.mean(axis=1), m.var(axis=1, ddof=1)
        rhat = np.sqrt(((draws - 1.0) / draws * W + Bn) / W)                     # per group, over its cpg chains
        names = [p["target"] for p in pri]
        steps = tune + draws + 8
        print("config 4: %d chains = %d groups (3 wells x %d members) x %d chains, %d GPUs, %d steps per chain, 4 tracers, L = 25,256"
              % (total, ngroups, (ngroups + 2) // 3, cpg, world, tune + draws))
        print("sampling %.2f s  ->  %.3e tracer-likelihood evals/s (chains x steps x tracers), acceptance %.3f"
              % (dt, total * (tune + draws) * len(TRACERS) / dt, float(acc) / steps))
        print("noble-gas fits of %d wells + ensemble propagation (50,000 members per well): %.1f s" % (len(WELLS), t_pre))
        print("R-hat of every group over its %d chains -- median: " % cpg + ", ".join("%s %.3f" % (n, np.median(rhat[:, i])) for i, n in enumerate(names)))
        print("                                           99th pct: " + ", ".join("%s %.3f" % (n, np.percentile(rhat[:, i], 99)) for i, n in enumerate(names)))
        print("groups with every R-hat < 1.05: %.1f %%" % (100.0 * np.mean(np.all(rhat < 1.05, axis=1))))
        from noblegas_rtd_mcmc_b200 import diagnostics
        ma, m2a = mean_all.cpu().numpy().reshape(ngroups, cpg, -1), m2_all.cpu().numpy().reshape(ngroups, cpg, -1)
        nr = np.array([diagnostics.nested_rhat(draws, ma[g], m2a[g], 8) for g in range(min(ngroups, 2048))])
        print("nested R-hat (8 superchains of %d chains per group; Margossian et al. 2022) -- max over %d groups: " % (cpg // 8, len(nr))
              + ", ".join("%s %.4f" % (n, nr[:, i].max()) for i, n in enumerate(names)))
        gvar = W + Bn                                                            # posterior variance per group (within + between chains)
        for k, w in enumerate(WELLS):
            sel = np.arange(ngroups) % 3 == k
            ref = fx["%s.CFC12.SF6.H3.He4_ter.exp_pist_flow.123" % w]["vars"]
            lowmode = float((gmean[sel, names.index("eta1")] < 1.6).mean())
            print("%s (%d members; %.0f %% of them with posterior-mean eta1 < 1.6): posterior pooled over the members | the reference's "
                  "single inversion of the ensemble mean with obs_err = ens.std() + 5 %%" % (w, int(sel.sum()), 100 * lowmode))
            for i, n in enumerate(names):
                if n == "nu_":
                    continue
                pm = gmean[sel, i].mean()
                psd = np.sqrt(gvar[sel, i].mean() + gmean[sel, i].var())
                print("    %-10s %10.4g +- %-10.3g | %10.4g +- %.3g" % (n, pm, psd, ref[n]["mean"], ref[n]["sd"]))
    smp.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
