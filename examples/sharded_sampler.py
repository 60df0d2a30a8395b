"""Chains sharded over the GPUs of one box (one process per GPU), NCCL all-gather of per-chain statistics.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        examples/sharded_sampler.py [total_chains] [steps]

Noble-gas closed-equilibrium fit of well PLM1 (config 1 of the reference) with `total_chains` chains; prints the pooled
posterior summary (R-hat / ESS from Welford moments of ALL ranks) and checks shard invariance against rank 0's own
re-computation of a block owned by another rank.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from noblegas_rtd_mcmc_b200 import distributed as D
from noblegas_rtd_mcmc_b200.noble_gas_mcmc import mcmc_model
from noblegas_rtd_mcmc_b200.sampler import Sampler


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "ng_posterior.json")))["wells"]["PLM1"]
    mdl = mcmc_model(fx["obs"], mcmc_model.well_elev["PLM1"])
    kw = dict(plan=None, gases=mdl.gases, lik="studentt", nu_range=(1.0, 30.0), tune_interval=1000, hist_cap=2 * steps, seed=7,
              device=local)
    off, cnt = D.shard(total, rank, world)
    smp = Sampler(mdl.build_priors(), mdl.obs_mu, mdl.obs_sd, cnt, chain_offset=off, **kw)
    smp.run(steps, tune=True)
    smp.stop_tuning()
    smp.run(steps, tune=False, record=True)
    summ = D.global_summary(steps, smp.get("mean"), smp.get("m2"))
    ok = True
    if world > 1 and rank == 0:      # recompute the first 16 chains of rank 1's block locally: must be bit-identical
        o1, c1 = D.shard(total, 1, world)
        ref = Sampler(mdl.build_priors(), mdl.obs_mu, mdl.obs_sd, 16, chain_offset=o1, **kw)
        ref.run(steps, tune=True); ref.stop_tuning(); ref.run(steps, tune=False, record=True)
        mine = ref.get("mean")
    if world > 1:
        allm, _ = D.gather_chain_stats(smp.get("mean"), smp.get("m2"))
        if rank == 0:
            ok = bool(torch.equal(allm[o1:o1 + 16], mine))
    if rank == 0:
        print(json.dumps({"world": world, "chains": int(summ["chains"]), "draws_per_chain": steps, "names": smp.names,
                          "mean": summ["mean"].tolist(), "sd": summ["sd"].tolist(), "r_hat": summ["r_hat"].tolist(),
                          "ess": summ["ess"].tolist(), "shard_invariant": ok}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
