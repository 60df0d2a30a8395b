"""Coefficients of exp_scaled_bits (ngrtd_common.cuh): q(g) ~= exp((g - 1) ln2 / N) on g in [1, 2), cubic through the
Chebyshev nodes of the interval (near-minimax).  g = 1 + f is assembled from the 32 fraction bits of the exponent, so the
polynomial needs no integer->double conversion.  Prints the constants and the maximum relative error."""
import mpmath as mp
mp.mp.prec = 200
for b in (7, 8, 9):
    kap = mp.log(2) / (1 << b)
    nodes = [mp.mpf(3) / 2 + mp.cos(mp.pi * (2 * i + 1) / 8) / 2 for i in range(4)]
    A = mp.matrix([[1, x, x * x, x ** 3] for x in nodes])
    y = mp.matrix([mp.exp(kap * (x - 1)) for x in nodes])
    c = [float(v) for v in mp.lu_solve(A, y)]
    worst = 0
    for i in range(4001):
        g = 1 + mp.mpf(i) / 4001
        q = c[0] + g * (c[1] + g * (c[2] + g * c[3]))
        worst = max(worst, abs(q / mp.exp(kap * (g - 1)) - 1))
    print("b=%d: %s  max rel err %.2e" % (b, ", ".join("%.17g" % v for v in c), float(worst)))
