"""Feasibility study (numpy, CPU): can the constant tail of a DISPERSION component be integrated instead of summed?

The real input series are constant beyond lag Kc = 128 ("constant-tail closed form"), which the exponential-class
components already exploit analytically.  Dispersion components still loop over all L = 25,256 lags.  Beyond Kc every term is a
smooth function of the lag,
    g(t) = t^-1.5 exp(-(a/t + c t)) * col(t),   a = tau/(4D), c = 1/(4 D tau),   col(t) in {1, e^-lam t, 1 - e^-lam t, t e^-lam t},
so  sum_{k=Kc}^{L-1} g(k) = int_{Kc-1/2}^{L-1/2} g dt + midpoint Euler-Maclaurin end corrections (exponentially small interior error
for peaks wider than ~2 lags).  This script measures, over the reference's prior ranges, how many Gauss-Legendre nodes reach 1e-12 of
the FULL normalised output (head [0, Kc) summed directly as today), using windowed panels of width ~sigma = tau sqrt(2D) around the
peak.  Result: profiles/r1_dm_tail_quadrature_study.txt.  Nothing here is product code.
"""
import numpy as np

L, Kc = 25256, 128
LAM = np.log(2) / 12.34          # 3H decay per yearly lag
ld = np.longdouble


def g_terms(t, a, c, lam, ty=ld):
    t = t.astype(ty)
    e = np.exp(-(ty(a) / t + ty(c) * t)) * t ** ty(-1.5)
    d = np.exp(-ty(lam) * t)
    return np.stack([e, e * d, e * (1 - d), e * t * d], -1)     # ones, decay, ingrowth, lag-index*decay columns


def direct(a, c, lam):
    k = np.arange(Kc, L)
    return g_terms(k, a, c, lam).sum(0)


def head(a, c, lam):
    k = np.arange(1, Kc)
    return g_terms(k, a, c, lam).sum(0)


def dg(t, a, c, lam, h=1e-3):     # first derivative for the h^2/24 midpoint correction (central difference in long double)
    t = np.array([t])
    return ((g_terms(t + h, a, c, lam) - g_terms(t - h, a, c, lam)) / (2 * h))[0]


def d3g(t, a, c, lam, h=1e-2):
    t = np.array([t])
    f = lambda x: g_terms(x, a, c, lam)
    return ((f(t + 2 * h) - 2 * f(t + h) + 2 * f(t - h) - f(t - 2 * h)) / (2 * h ** 3))[0]


def quad(a, c, lam, tau, D, npanel_per_sigma=1.0, order=8):
    lo, hi = Kc - 0.5, L - 0.5
    sig = max(tau * np.sqrt(2 * D), 1.0)
    # window where the weight can matter: [tau - 12 sig_left, tau + 40 sig] clipped (right tail of the inverse Gaussian is fat)
    wlo = max(lo, tau - 12 * sig)
    whi = min(hi, max(tau + 60 * sig, tau + 200 * D * tau, lo + 1))   # right tail ~ exp(-c t), 1/c = 4 D tau
    if whi <= wlo:
        return np.zeros(4, dtype=ld), 0
    # geometric panels to the right of the peak (the tail decays like exp(-c t)), uniform ~sigma panels near it
    edges = [wlo]
    w = sig / npanel_per_sigma
    while edges[-1] < whi:
        x = edges[-1]
        step = w if abs(x - tau) < 6 * sig else max(w, 0.25 * abs(x - tau))
        step = min(step, 0.35 * x)            # t^-1.5 exp(-a/t) varies on the scale of t itself: geometric panels from the left
        edges.append(min(whi, x + step))
    edges = np.array(edges)
    xg, wg = np.polynomial.legendre.leggauss(order)
    mid = 0.5 * (edges[1:] + edges[:-1])[:, None]
    half = 0.5 * (edges[1:] - edges[:-1])[:, None]
    t = (mid + half * xg[None, :]).ravel()
    wt = (half * wg[None, :]).ravel()
    val = (g_terms(t, a, c, lam, np.float64) * wt[:, None]).sum(0).astype(ld)     # float64, as a kernel would
    # midpoint Euler-Maclaurin: sum g(k) = int_{m-1/2}^{n+1/2} g - (1/24)[g'] + (7/5760)[g'''] - ...
    if wlo == lo:
        val += dg(lo, a, c, lam) / 24 - 7 * d3g(lo, a, c, lam) / 5760
    if whi == hi:
        val -= dg(hi, a, c, lam) / 24 - 7 * d3g(hi, a, c, lam) / 5760
    return val, t.size


import sys
DLO, DHI = (float(sys.argv[1]), float(sys.argv[2])) if len(sys.argv) > 2 else (0.01, 2.0)
print("tau ~ logU(1, 15000) lags, D ~ logU(%g, %g), L = %d, Kc = %d, 3H decay; quadrature in float64, reference sums in long double" % (DLO, DHI, L, Kc))
rng = np.random.default_rng(0)
rows = []
for trial in range(400):
    tau = float(np.exp(rng.uniform(np.log(1.0), np.log(15000.0))))
    D = float(np.exp(rng.uniform(np.log(DLO), np.log(DHI))))
    a, c = tau / (4 * D), 1 / (4 * D * tau)
    full_h = head(a, c, LAM)
    ex = direct(a, c, LAM)
    tot = full_h + ex
    if not np.isfinite(tot[0]) or tot[0] < 1e-300:
        continue
    best = None
    for pps, order in ((1.0, 8), (1.0, 12), (2.0, 12), (2.0, 16), (4.0, 16)):
        q, n = quad(a, c, LAM, tau, D, pps, order)
        # error of the NORMALISED outputs col/ones, relative
        out_ex = tot[1:] / tot[0]
        tq = full_h + q
        out_q = tq[1:] / tq[0]
        err = float(np.max(np.abs(out_q - out_ex) / np.abs(out_ex)))
        errn = float(abs(tq[0] - tot[0]) / tot[0])
        best = (pps, order, n, max(err, errn))
        if best[3] < 1e-12:
            break
    rows.append((tau, D, float(ex[0] / tot[0]), *best))
rows = np.array(rows)
print("cases %d; tail share of the normalisation sum: median %.3f" % (len(rows), np.median(rows[:, 2])))
ok = rows[:, 6] < 1e-12
print("reached 1e-12 of the normalised outputs: %d of %d;  nodes used: median %d, 90%% %d, max %d  (direct tail: %d terms)" % (
    ok.sum(), len(rows), np.median(rows[ok, 5]), np.percentile(rows[ok, 5], 90), rows[ok, 5].max(), L - Kc))
for pps, order in ((1.0, 8), (1.0, 12), (2.0, 12), (2.0, 16), (4.0, 16)):
    sel = ok & (rows[:, 3] == pps) & (rows[:, 4] == order)
    print("  first setting that reached it: %.0f panels per sigma, order %2d: %d cases" % (pps, order, sel.sum()))
bad = rows[~ok]
for r in bad[np.argsort(-bad[:, 6])][:12]:
    print("  not reached: tau %.1f D %.3f tail share %.2e nodes %d err %.1e" % (r[0], r[1], r[2], r[5], r[6]))
