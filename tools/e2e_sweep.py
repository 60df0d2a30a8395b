"""Host-buffer (e2e) call timing over NGRTD_HOST_PARTS settings (development aid; bench.py is the contract).

    python tools/e2e_sweep.py [B]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from noblegas_rtd_mcmc_b200 import synthetic
from helpers import synth_plan

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
pn = list(synthetic.PAR_NAMES_CFG3)
plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
obs = np.array([8.0, 40.0, 150.0, 300.0, 50.0, 5.0, 1e-8])
sd = 0.05 * obs
thetas = [torch.from_numpy(synthetic.theta_cfg3(B, seed=77 + i)).pin_memory() for i in range(8)]
logp = torch.empty(B, dtype=torch.float64).pin_memory()
hl = logp.numpy()
th_d = thetas[0].cuda()
ref = plan.forward_loglik_dev(th_d, pn, obs, sd, "normal").cpu().numpy()
for rnd in range(1):
    for parts in ("1", "2", "4", "mapped"):
        if parts == "mapped":
            os.environ["NGRTD_HOST_MODE"] = "mapped"
        else:
            os.environ["NGRTD_HOST_MODE"] = "copy"
            os.environ["NGRTD_HOST_PARTS"] = parts
        for i in range(5):
            plan.forward_loglik_host(thetas[i % 8].numpy(), pn, obs, sd, "normal", logp_out=hl)
        ts = []
        for rep in range(5):
            t0 = time.perf_counter()
            n = 50
            for i in range(n):
                plan.forward_loglik_host(thetas[i % 8].numpy(), pn, obs, sd, "normal", logp_out=hl)
            ts.append((time.perf_counter() - t0) / n)
        plan.forward_loglik_host(thetas[0].numpy(), pn, obs, sd, "normal", logp_out=hl)
        same = np.array_equal(ref, hl, equal_nan=True)
        print("round %d parts=%s B=%d  min %.4f  median %.4f ms/call  %.3e evals/s (6 counted)  identical_to_device_path=%s" % (
            rnd, parts, B, min(ts) * 1e3, sorted(ts)[2] * 1e3, B * 6 / sorted(ts)[2], same), flush=True)
# mapped-mode anatomy: fixed call overhead (16 chains), lane-load staging, kernel time with host-resident theta / logp
os.environ["NGRTD_HOST_MODE"] = "mapped"
for nb in (16, 4096, 16384, 32768):
    ts = []
    for rep in range(5):
        t0 = time.perf_counter()
        for i in range(50):
            plan.forward_loglik_host(thetas[i % 8].numpy()[:nb], pn, obs, sd, "normal", logp_out=hl[:nb])
        ts.append((time.perf_counter() - t0) / 50)
    print("mapped B=%d: median %.4f ms/call" % (nb, sorted(ts)[2] * 1e3))
for stg in ("2", "0", "1"):
    os.environ["NGRTD_STAGE_MAPPED"] = stg
    ts = []
    for rep in range(5):
        t0 = time.perf_counter()
        for i in range(20):
            plan.forward_loglik_host(thetas[i % 8].numpy(), pn, obs, sd, "normal", logp_out=hl)
        ts.append((time.perf_counter() - t0) / 20)
    print("mapped, staging mode %s: median %.4f ms/call" % (stg, sorted(ts)[2] * 1e3))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lp_d = torch.empty(B, dtype=torch.float64, device="cuda")
os.environ["NGRTD_STAGE"] = "1"
import ctypes
from noblegas_rtd_mcmc_b200 import _lib
_lib.dptr = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())     # pinned host tensors are UVA-addressable
for name, th_src, lp_dst in (("theta host, logp device", thetas, lp_d), ("theta host, logp host", thetas, logp),
                             ("theta device, logp host", [t.cuda() for t in thetas], logp)):
    for i in range(3):
        plan.forward_loglik_dev(th_src[i], pn, obs, sd, "normal", logp_t=lp_dst)
    torch.cuda.synchronize()
    e0.record()
    for i in range(20):
        plan.forward_loglik_dev(th_src[i % 8], pn, obs, sd, "normal", logp_t=lp_dst)
    e1.record()
    torch.cuda.synchronize()
    print("kernel (stage 1) %s: %.4f ms" % (name, e0.elapsed_time(e1) / 20))
os.environ["NGRTD_STAGE"] = "0"
# copy-only and kernel-only references
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
d = torch.empty_like(th_d)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50):
    d.copy_(thetas[i % 8], non_blocking=True)
torch.cuda.synchronize()
print("H2D only: %.4f ms per %d bytes" % ((time.perf_counter() - t0) / 50 * 1e3, thetas[0].numel() * 8))
lp = torch.empty(B, dtype=torch.float64, device="cuda")
t0 = time.perf_counter()
for i in range(50):
    logp.copy_(lp, non_blocking=True)
torch.cuda.synchronize()
print("D2H only: %.4f ms per %d bytes" % ((time.perf_counter() - t0) / 50 * 1e3, B * 8))
for stage, n in [(st, n) for n in (65536,) for st in ("0", "1")]:
    os.environ["NGRTD_STAGE"] = stage
    t = th_d[:n].contiguous()
    for _ in range(3):
        plan.forward_loglik_dev(t, pn, obs, sd, "normal", logp_t=lp[:n])
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        plan.forward_loglik_dev(t, pn, obs, sd, "normal", logp_t=lp[:n])
    e1.record()
    torch.cuda.synchronize()
    os.environ["NGRTD_STAGE"] = "0"
    same = np.array_equal(plan.forward_loglik_dev(t, pn, obs, sd, "normal").cpu().numpy(), lp[:n].cpu().numpy(), equal_nan=True)
    print("kernel only stage=%s B=%d: %.4f ms  identical_to_stage0=%s" % (stage, n, e0.elapsed_time(e1) / 20, same))
