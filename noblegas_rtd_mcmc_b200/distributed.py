"""Multi-GPU plumbing: chains shard over ranks (one process per GPU, torch.distributed over NCCL/NVLink); the only
collective is the reduction of chain statistics for R-hat / ESS: ONE all-reduce of 3*ndim + 1 pooled
moments that every rank reduced on its own device (`pooled_summary`), or -- when per-chain moments are wanted -- an
all-gather of them (`global_summary`).  On CPU test runs the same code goes through the gloo backend."""
import numpy as np


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(pci_bus_id, sysfs_root="/sys/bus/pci/devices"):
    """Pin the calling process to the CPUs next to its GPU (one process per GPU on a multi-socket box), so that pinned host
    buffers allocated afterwards are first-touched on the GPU's own NUMA node and the host-buffer entry points do not pull
    their parameter batches across the socket interconnect.  Deliberately conservative: acts only when sysfs reports a NUMA
    node >= 0 for the device and its local CPU list is a non-empty strict subset of the CPUs this process may run on; in
    every other case (single-node hosts, VMs that hide the topology, unreadable sysfs) it changes nothing.
    Returns a short description of what was done (for logs / the bench line)."""
    import os
    try:
        dev = os.path.join(sysfs_root, pci_bus_id.lower())          # "dddd:bb:dd.f", e.g. 0000:c0:00.0
        node = int(open(os.path.join(dev, "numa_node")).read().strip())
        local = _parse_cpulist(open(os.path.join(dev, "local_cpulist")).read())
    except (OSError, ValueError):
        return "unchanged (no sysfs topology for %s)" % pci_bus_id
    allowed = os.sched_getaffinity(0)
    target = local & allowed
    if node < 0 or not target or target == allowed:
        return "unchanged (numa_node %d, %d local of %d allowed cpus)" % (node, len(target), len(allowed))
    os.sched_setaffinity(0, target)
    return "bound to numa node %d (%d of %d cpus)" % (node, len(target), len(allowed))


def shard(total, rank, world):
    """Contiguous block of global chain ids for `rank`: (offset, count).  Philox counters use global ids, so a chain's
    trajectory does not depend on the sharding."""
    base, rem = divmod(int(total), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def gather_chain_stats(mean_t, m2_t):
    """All-gather per-chain Welford moments [chains_local, ndim] from every rank -> ([chains_total, ndim], same).
    Shards may differ in size by one chain; tensors are padded to the maximum and trimmed after the gather."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return mean_t, m2_t
    world = dist.get_world_size()
    n_local = torch.tensor([mean_t.shape[0]], dtype=torch.int64, device=mean_t.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts)
    packed = torch.zeros((nmax, 2 * mean_t.shape[1]), dtype=mean_t.dtype, device=mean_t.device)
    packed[:mean_t.shape[0], :mean_t.shape[1]] = mean_t
    packed[:mean_t.shape[0], mean_t.shape[1]:] = m2_t
    bufs = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(bufs, packed)
    allp = torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
    nd = mean_t.shape[1]
    return allp[:, :nd].contiguous(), allp[:, nd:].contiguous()


def global_summary(n_draws, mean_t, m2_t):
    """R-hat / ESS / pooled moments over the chains of ALL ranks (identical on every rank)."""
    from . import diagnostics
    mean_all, m2_all = gather_chain_stats(mean_t, m2_t)
    return diagnostics.moments_summary(float(n_draws), mean_all.cpu().numpy(), m2_all.cpu().numpy())


def allreduce_pooled(vec_t):
    """Sum the pooled-moment vectors [3*nd + 1] of all ranks (in place; no-op without a process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec_t, op=dist.ReduceOp.SUM)
    return vec_t


def summary_from_pooled(n_draws, vec):
    """Pooled mean / sd, classic R-hat and the many-chain ESS estimate from [sum mean, sum mean^2, sum M2, chains]
    (the same quantities diagnostics.moments_summary derives from per-chain moments)."""
    vec = np.asarray(vec, dtype=np.float64)
    nd = (len(vec) - 1) // 3
    M = float(vec[3 * nd])
    n = float(n_draws)
    s1, s2, sm2 = vec[:nd], vec[nd:2 * nd], vec[2 * nd:3 * nd]
    mean = s1 / M
    Bn = np.maximum(s2 - s1 * s1 / M, 0.0) / (M - 1.0)         # B / n: variance of the chain means
    W = sm2 / (M * (n - 1.0))
    var_plus = (n - 1.0) / n * W + Bn
    with np.errstate(all="ignore"):
        return {"mean": mean, "sd": np.sqrt(var_plus), "r_hat": np.sqrt(var_plus / W), "ess": M * var_plus / Bn,
                "mcse_mean": np.sqrt(Bn / M), "chains": int(M), "draws_per_chain": n}


def pooled_summary(sampler, n_draws, stream=None):
    """R-hat / ESS / pooled moments over the chains of ALL ranks without moving per-chain data: every rank reduces its
    chains on the device (Sampler.pooled_moments), one all-reduce adds the 3*ndim + 1 numbers (identical on every rank)."""
    vec = allreduce_pooled(sampler.pooled_moments(stream))
    return summary_from_pooled(n_draws, vec.cpu().numpy())
