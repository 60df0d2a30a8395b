"""Where the time of the submit / wait host path goes (development aid): per-step wall time, CPU time inside submit and
wait, over repeated runs; NUMA placement of the process and of the GPU."""
import os, sys, time, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import synthetic
from helpers import synth_plan
B = 65536
pn = list(synthetic.PAR_NAMES_CFG3)
plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
obs = np.array([8.0, 40.0, 150.0, 300.0, 50.0, 5.0, 1e-8]); sd = 0.05 * obs
thetas = [torch.from_numpy(synthetic.theta_cfg3(B, seed=77 + i)).pin_memory() for i in range(8)]
tn = [t.numpy() for t in thetas]
DEPTH = int(os.environ.get("DEPTH", "3"))
hls = [torch.empty(B, dtype=torch.float64).pin_memory().numpy() for _ in range(DEPTH)]
print("cpus allowed", len(os.sched_getaffinity(0)), "of", os.cpu_count(), flush=True)
try:
    busid = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), "pci_bus_id") else None
    for p in glob.glob("/sys/bus/pci/devices/*/numa_node"):
        d = os.path.dirname(p)
        if open(os.path.join(d, "vendor")).read().strip() == "0x10de" and open(os.path.join(d, "class")).read().startswith("0x0302"):
            print("gpu", os.path.basename(d), "numa", open(p).read().strip(), "local cpus", open(os.path.join(d, "local_cpulist")).read().strip())
    print("this process last ran on cpu", open("/proc/self/stat").read().split()[38])
except Exception as e:
    print("numa probe failed", e)
def run(n):
    ts = tw = 0.0
    for i in range(n + DEPTH):
        if i >= DEPTH:
            t0 = time.perf_counter(); plan.host_wait((i - DEPTH) % DEPTH); tw += time.perf_counter() - t0
        if i < n:
            t0 = time.perf_counter()
            plan.forward_loglik_host_submit(tn[i % 8], pn, obs, sd, "normal", logp_out=hls[i % DEPTH], slot=i % DEPTH)
            ts += time.perf_counter() - t0
    return ts, tw
run(8)
for rep in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ts, tw = run(200); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("rep %d: %.4f ms/step (%.3e evals/s); in submit %.1f us/step, in wait %.1f us/step" % (rep, dt / 200 * 1e3, B * 6 * 200 / dt, ts / 200 * 1e6, tw / 200 * 1e6), flush=True)
# raw copy rates of the pinned buffers used above
d = torch.empty_like(thetas[0], device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("H2D", lambda i: d.copy_(thetas[i % 8], non_blocking=True)),):
    for i in range(5): fn(i)
    torch.cuda.synchronize(); e0.record()
    for i in range(50): fn(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print("%s %.4f ms per %d bytes = %.1f GB/s" % (name, ms, thetas[0].numel() * 8, thetas[0].numel() * 8 / ms / 1e6))
