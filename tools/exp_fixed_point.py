"""Derive and verify the fixed-point constants of exp_scaled_fx (ngrtd_common.cuh).

p(f) = exp(f ln2 / N) = 1 + m(f),  m = f (a1 + f (a2 + f a3)),  f in [0, 1) given as a 32-bit fraction F = f 2^32.
The device evaluates m 2^52 with 32-bit integer multiplies:
    s3 = mulhi(F, A3); t2 = A2 + s3; s1 = mulhi(F, t2); T1 = A1 + s1 * SH (64 bit); M = F*T1_hi + mulhi(F, T1_lo)
with A3, A2 at scale 2^S2 (S2 = 34 + 2 b), A1 at scale 2^52, SH = 2^(52 - S2).  This script fits a1..a3 (interpolation of
m(f)/f at Chebyshev nodes), prints the integer constants and the maximum relative error of 1 + M 2^-52 against exp().
"""
import sys
import numpy as np
import mpmath as mp

mp.mp.prec = 200


def fit(b):
    kap = mp.log(2) / (1 << b)
    nodes = [(mp.mpf(1) + mp.cos(mp.pi * (2 * i + 1) / 6)) / 2 for i in range(3)]
    A = mp.matrix([[1, x, x * x] for x in nodes])
    y = mp.matrix([mp.expm1(kap * x) / x for x in nodes])
    a = mp.lu_solve(A, y)
    return kap, [a[0], a[1], a[2]]


def consts(b):
    kap, (a1, a2, a3) = fit(b)
    S2 = 34 + 2 * b
    A3 = int(mp.nint(a3 * mp.mpf(2) ** S2))
    A2 = int(mp.nint(a2 * mp.mpf(2) ** S2))
    A1 = int(mp.nint(a1 * mp.mpf(2) ** 52))
    SH = 1 << (52 - S2)
    assert A3 < 2**32 and A2 + A3 < 2**32, (A2, A3)
    return kap, A1, A2, A3, SH


def emulate(F, A1, A2, A3, SH):
    F = F.astype(object)
    s3 = (F * A3) >> 32
    t2 = A2 + s3
    s1 = (F * t2) >> 32
    T1 = A1 + s1 * SH
    hi, lo = T1 >> 32, T1 & 0xFFFFFFFF
    return F * hi + ((F * lo) >> 32)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for b in (7, 8, 9):
        kap, A1, A2, A3, SH = consts(b)
        F = np.concatenate([rng.integers(0, 2**32, 20000, dtype=np.uint64), np.array([0, 1, 2**31, 2**32 - 1], dtype=np.uint64)])
        M = emulate(F, A1, A2, A3, SH)
        worst = 0
        for Fi, Mi in zip(F.tolist(), M.tolist()):
            assert Mi < 2**52
            ex = mp.exp(kap * mp.mpf(Fi) / 2**32)
            worst = max(worst, abs((1 + mp.mpf(Mi) / 2**52) / ex - 1))
        print(f"b={b}: A1=0x{A1:X}ull (hi 0x{A1 >> 32:X}, lo 0x{A1 & 0xFFFFFFFF:X}) A2={A2}u A3={A3}u SH={SH}  max rel err {float(worst):.3e}")
