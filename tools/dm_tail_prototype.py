"""float64 prototype of the r2 dispersion-tail quadrature, checked against the REFERENCE's golden vectors (CPU only).

Companion of tools/dm_tail_quadrature_study.py.  For the single-component dispersion configuration of tests/golden/forward_real.npz
("dm": 24 (tau, D) pairs x 6 tracers, produced by the untouched reference on the real 25,256-lag series) the convolution is evaluated
as   head [0, Kc) summed directly  +  tail [Kc, L) by Gauss-Legendre panels with analytic midpoint Euler-Maclaurin end corrections,
in plain float64 with the arithmetic a kernel would use (one exp for the weight, one per decay constant, t^-1.5 = 1/(t sqrt t)).
Column types as in PlanView::ct: ones, bg*exp(-lam t), bg*(1-exp(-lam t)), (i0 + s t)*exp(-lam t).  Prints the worst relative error
against the golden vectors and the number of nodes; the result is kept in profiles/r1_dm_tail_quadrature_study.txt.
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from helpers import GOLD, REAL_TRACERS, load_c_in
import np_oracle as O

ORDER, PPS = 16, 2.0
XG, WG = np.polynomial.legendre.leggauss(ORDER)


def base(t, a, c):
    """w(t) = t^-1.5 exp(-(a/t + c t)) and its logarithmic derivative polynomial pieces."""
    return np.exp(-(a / t + c * t)) / (t * np.sqrt(t))


def derivs(t, a, c, lam):
    """g = w(t) e^{-lam t}: returns g, g', g''' (analytic: g' = g p, g''' = g (p^3 + 3 p p' + p''))."""
    g = base(t, a, c) * np.exp(-lam * t)
    p = -1.5 / t + a / t ** 2 - c - lam
    p1 = 1.5 / t ** 2 - 2 * a / t ** 3
    p2 = -3.0 / t ** 3 + 6 * a / t ** 4
    return g, g * p, g * (p * p + p1), g * (p ** 3 + 3 * p * p1 + p2)


def end_corr(t, a, c, cols):
    """midpoint Euler-Maclaurin term  g'(t)/24 - 7 g'''(t)/5760  for every column at one end point"""
    out = []
    for ty, bg, lam, i0, s in cols:
        if ty == 0:
            g, g1, g2, g3 = derivs(t, a, c, 0.0); out.append(g1 / 24 - 7 * g3 / 5760)
        elif ty == 1:
            g, g1, g2, g3 = derivs(t, a, c, lam); out.append(bg * (g1 / 24 - 7 * g3 / 5760))
        elif ty == 2:
            g, g1, g2, g3 = derivs(t, a, c, 0.0); h, h1, h2, h3 = derivs(t, a, c, lam)
            out.append(bg * ((g1 - h1) / 24 - 7 * (g3 - h3) / 5760))
        else:       # (i0 + s t) h(t):  ' = s h + q h',  ''' = 3 s h'' + q h'''
            h, h1, h2, h3 = derivs(t, a, c, lam); q = i0 + s * t
            out.append((s * h + q * h1) / 24 - 7 * (3 * s * h2 + q * h3) / 5760)
    return np.array(out)


def col_values(t, cols):
    out = []
    for ty, bg, lam, i0, s in cols:
        if ty == 0: out.append(np.ones_like(t))
        elif ty == 1: out.append(bg * np.exp(-lam * t))
        elif ty == 2: out.append(-bg * np.expm1(-lam * t))
        else: out.append((i0 + s * t) * np.exp(-lam * t))
    return np.stack(out, -1)


def tail_quad(tau, D, Kc, L, cols):
    a, c = tau / (4 * D), 1 / (4 * D * tau)
    lo, hi = Kc - 0.5, L - 0.5
    sig = max(tau * np.sqrt(2 * D), 1.0)
    wlo = max(lo, tau - 12 * sig)
    whi = min(hi, max(tau + 60 * sig, tau + 200 * D * tau, lo + 1))
    if whi <= wlo:
        return np.zeros(len(cols)), 0
    edges, w = [wlo], sig / PPS
    while edges[-1] < whi:
        x = edges[-1]
        step = w if abs(x - tau) < 6 * sig else max(w, 0.25 * abs(x - tau))
        edges.append(min(whi, x + min(step, 0.35 * x)))
    edges = np.array(edges)
    mid, half = 0.5 * (edges[1:] + edges[:-1])[:, None], 0.5 * (edges[1:] - edges[:-1])[:, None]
    t = (mid + half * XG[None, :]).ravel()
    wt = (half * WG[None, :]).ravel()
    val = ((base(t, a, c) * wt)[:, None] * col_values(t, cols)).sum(0)
    if wlo == lo: val += end_corr(lo, a, c, cols)
    if whi == hi: val -= end_corr(hi, a, c, cols)
    return val, t.size


def head_direct(tau, D, Kc, X):
    a, c = tau / (4 * D), 1 / (4 * D * tau)
    tp = np.arange(Kc, dtype=np.float64); tp[0] += 1e-5
    return (base(tp, a, c)[:, None] * X[:Kc]).sum(0)


def run(verbose=True):
    global cols, Xh, L, Kc
    z = np.load(os.path.join(GOLD, "forward_real.npz"))
    series = load_c_in()
    L, Kc = len(series["H3"]), 128
    lamH = np.log(2) / 12.34
    tracers = ["CFC12", "SF6", "H3", "He4_ter", "He3", "CFC11"]
    # folded columns: ones + one per tracer
    cols = [(0, 0, 0, 0, 0)]
    Xh = [np.ones(Kc)]
    tp = np.arange(Kc, dtype=np.float64); tp[0] += 1e-5
    for tr in tracers:
        key, th, ra = REAL_TRACERS[tr]
        if ra == "4He":
            cols.append((3, 0.0, 0.0, 0.0, 1.0)); Xh.append(np.arange(Kc, dtype=np.float64))          # index * J (:323)
        elif ra == "3He":
            cols.append((2, float(series[key][-1]), lamH, 0, 0)); Xh.append(series[key][:Kc] * (1 - np.exp(-lamH * tp)))
        elif th:
            cols.append((1, float(series[key][-1]), lamH, 0, 0)); Xh.append(series[key][:Kc] * np.exp(-lamH * tp))
        else:
            cols.append((1, float(series[key][-1]), 0.0, 0, 0)); Xh.append(series[key][:Kc])
    Xh = np.stack(Xh, -1)
    for k in ("CFC12", "SF6", "H3", "CFC11"):
        assert np.all(series[k][Kc:] == series[k][-1]), "series not constant beyond Kc"
    theta = z["dm/theta"]
    worst, nodes, skipped = 0.0, [], 0
    J = 10 ** O.DEFAULT_LOG10_J
    for i, (tau, D) in enumerate(theta):
        if not (0.01 <= D <= 2.5 and tau >= 1.0):
            skipped += 1; continue
        tq, n = tail_quad(tau, D, Kc, L, cols)
        tot = head_direct(tau, D, Kc, Xh) + tq
        out = tot[1:] / tot[0]
        for jx, tr in enumerate(tracers):
            want = z["dm/" + tr][i]
            got = out[jx] * (J if tr == "He4_ter" else 1.0)
            if np.isfinite(want) and want != 0:
                e = abs(got - want) / abs(want)
                worst = max(worst, e)
                if e > 1e-10: print("  tau %.4g D %.4g %s: got %.15g want %.15g rel %.2e" % (tau, D, tr, got, want, e))
        nodes.append(n)
    if verbose:
        print("golden 'dm' (reference output, real series L = %d, Kc = %d): %d parameter pairs (%d outside the validated domain skipped), "
              "worst relative error %.2e, nodes median %d max %d (order %d, %.0f panels per sigma) vs %d direct tail terms" % (
                  L, Kc, len(nodes), skipped, worst, np.median(nodes), max(nodes), ORDER, PPS, L - Kc))
    return worst, int(np.median(nodes)), int(max(nodes)), skipped


if __name__ == "__main__":
    run()
