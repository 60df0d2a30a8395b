"""Analytic-tail path vs the full lag loop (NGRTD_NO_TAIL=1) on the real yearly series: agreement and speed-up."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, time
from helpers import MODEL_CFGS, real_plan
if len(sys.argv) > 1 and sys.argv[1] == "child":
    rng = np.random.default_rng(5)
    out = {}
    for name in ("emm123", "epm123", "epm_pfm123", "emm_pfm123", "pfm_epm", "emm0"):
        m1, m2, pn = MODEL_CFGS[name]
        tracers = ["CFC12", "SF6", "H3", "He4_ter", "He3", "CFC11", "CFC113"]
        plan, _ = real_plan(m1, m2, pn, tracers)
        B = 16384
        f1 = rng.uniform(0.01, 0.99, B)
        cols = {"tau1": np.exp(rng.uniform(0, np.log(1000), B)), "tau2": np.exp(rng.uniform(np.log(50), np.log(15000), B)), "f1": f1, "f2": 1 - f1,
                "eta1": rng.uniform(1, 5, B), "eta2": rng.uniform(1, 5, B), "J": rng.normal(-10.4, 0.33, B), "thalf_cfc": rng.uniform(5, 35, B),
                "lamsf6": np.abs(rng.normal(0, 0.17, B))}
        theta = np.ascontiguousarray(np.stack([cols[p] for p in pn], axis=1))
        th = torch.from_numpy(theta).cuda()
        o = plan.forward_dev(th, pn); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3): plan.forward_dev(th, pn, o)
        torch.cuda.synchronize()
        out[name] = dict(ms=(time.perf_counter() - t0) / 3 * 1e3)
        np.save("/tmp/tail_%s_%s.npy" % (name, os.environ.get("NGRTD_NO_TAIL", "0")), o.cpu().numpy())
    print(json.dumps(out))
else:
    res = {}
    for flag in ("0", "1"):
        env = dict(os.environ)
        if flag == "1": env["NGRTD_NO_TAIL"] = "1"
        else: env.pop("NGRTD_NO_TAIL", None)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        res[flag] = json.loads(r.stdout.strip().splitlines()[-1])
    for name in res["0"]:
        a = np.load("/tmp/tail_%s_0.npy" % name); b = np.load("/tmp/tail_%s_1.npy" % name)
        assert np.array_equal(np.isnan(a), np.isnan(b)), name
        m = ~np.isnan(b)
        err = np.max(np.abs(a[m] - b[m]) / np.abs(b[m]))
        print("%-12s tail %.3f ms  full loop %.3f ms  speed-up %.0fx  max rel diff %.2e  (16,384 chains x 7 tracers, L=25,256)" % (
            name, res["0"][name]["ms"], res["1"][name]["ms"], res["1"][name]["ms"] / res["0"][name]["ms"], err))
