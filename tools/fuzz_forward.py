"""Randomised parity sweep of the forward model against the numpy oracle (cases: tests/fuzz_cases.py).
Development aid (GPU); usage: fuzz_forward.py [cases=200] [seed=1]."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

from fuzz_cases import one_case


def main():
    ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    t0 = time.time()
    fails, worst_all = 0, 0.0
    for c in range(ncase):
        bad_nan, worst, tag, nkeep, e = one_case(rng, c)
        worst_all = max(worst_all, e if worst <= 1.0 else worst_all)
        if bad_nan or worst > 1.0:
            fails += 1
            print("FAIL case %d: %s | NaN mismatches %d, worst err/tol %.3g (err %.3g), compared %d chains" % (c, tag, bad_nan, worst, e, nkeep), flush=True)
    print("fuzz: %d cases, %d failures, worst relative error among passing cases %.3g, %.1f s" % (ncase, fails, worst_all, time.time() - t0))
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
