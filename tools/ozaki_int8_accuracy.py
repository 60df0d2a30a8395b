"""Numerical feasibility of taking the lag reduction off the FP64 pipe ("what would move the ceiling"):
out[chain, col] = sum_k W[chain, k] X[k, col] with W and X split into 8-bit integer slices (Ozaki scheme), every slice-pair
product accumulated exactly in int32 (what tcgen05 kind::i8 does in TMEM) and recombined in FP64.  CPU / numpy only.

Prints, for the cfg-3 workload (EPM + dispersion weights of 2,048 informative chains, L = 840, the 8 folded columns), the worst
relative error of the NORMALISED outputs (column c / column 0) against the plain FP64 reduction, for several slice counts and truncations of the slice pairs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from noblegas_rtd_mcmc_b200 import synthetic

L, B = 840, 2048
ser = synthetic.input_series(L, 0)
tp = np.arange(L, dtype=np.float64); tp[0] = 1e-5
lam = np.log(2.0) / 148.08
cols = [np.ones(L), ser["H3"] * np.exp(-lam * tp), ser["H3"] * (1 - np.exp(-lam * tp)), ser["CFC11"], ser["CFC12"], ser["CFC113"],
        ser["SF6"], np.arange(L, dtype=np.float64)]
X = np.stack(cols, axis=1)                                   # [L, 8]
th = synthetic.theta_cfg3_informative(B, 5)
p = dict(zip(synthetic.PAR_NAMES_CFG3, th.T))
# unnormalised weights (conv utils :186-196)
eta, tau1 = p["eta1"][:, None], p["tau1"][:, None]
Wg = np.where(tp[None, :] >= tau1 * (1 - 1 / eta), (eta / tau1) * np.exp(-eta * tp[None, :] / tau1 + eta - 1), 0.0)
tau2, D = p["tau2"][:, None], p["D2"][:, None]
x = tp[None, :] / tau2
Wd = (1 / tau2) / np.sqrt(4 * np.pi * D * x) * (1 / x) * np.exp(-((1 - x) ** 2) / (4 * D * x))


def slices_unsigned(A, axis, S):
    """A >= 0:  A ~ scale * sum_i d_i 2^(-8 i), d_i in [0, 255] (floor digits: 8 new bits per slice), scale a power of two."""
    amax = np.max(A, axis=axis, keepdims=True)
    scale = 2.0 ** (np.floor(np.log2(np.where(amax > 0, amax, 1.0))) + 1.0)      # A / scale in [0, 1)
    r = A / scale
    out = []
    for i in range(1, S + 1):
        d = np.floor(r * 2.0 ** (8 * i))
        assert d.min() >= 0 and d.max() <= 255
        r = r - d * 2.0 ** (-8 * i)
        out.append(d.astype(np.int64))
    return scale, out, [2.0 ** (-8 * i) for i in range(1, S + 1)]


def slices_signed(A, axis, S):
    """A ~ scale * sum_i d_i 2^(-7 i), d_i in [-64, 64] (round-to-nearest digits: 7 new bits per slice)."""
    amax = np.max(np.abs(A), axis=axis, keepdims=True)
    scale = 2.0 ** (np.floor(np.log2(np.where(amax > 0, amax, 1.0))) + 2.0)      # |A / scale| < 1/2
    r = A / scale
    out = []
    for i in range(1, S + 1):
        d = np.rint(r * 2.0 ** (7 * i))
        assert np.abs(d).max() <= 64
        r = r - d * 2.0 ** (-7 * i)
        out.append(d.astype(np.int64))
    return scale, out, [2.0 ** (-7 * i) for i in range(1, S + 1)]


def ozaki(W, X, SW, SX, budget_bits):
    """all slice pairs whose weight 2^-(8 i + 7 j) is above 2^-budget_bits"""
    sw, dw, ww = slices_unsigned(W, 1, SW)
    sx, dx, wx = slices_signed(X, 0, SX)
    acc = np.zeros((W.shape[0], X.shape[1]))
    npairs = 0
    for i in range(SW):
        for j in range(SX):
            if ww[i] * wx[j] < 2.0 ** (-budget_bits):
                continue
            part = dw[i] @ dx[j]                             # exact integers, int32 range checked
            assert np.abs(part).max() < 2 ** 31
            acc += part.astype(np.float64) * (ww[i] * wx[j])
            npairs += 1
    return acc * sw * sx, npairs


for name, W in (("exponential-piston", Wg), ("dispersion", Wd)):
    ref = W @ X
    ok = ref[:, 0] > 0
    refn = ref[ok, 1:] / ref[ok, :1]
    for SW, SX, bits in ((4, 5, 36), (5, 6, 44), (6, 7, 52), (6, 7, 60), (7, 8, 60)):
        got, npairs = ozaki(W, X, SW, SX, bits)
        gn = got[ok, 1:] / got[ok, :1]
        err = np.max(np.abs(gn - refn) / np.maximum(np.abs(refn), 1e-300))
        print("%-18s %d weight slices (u8) x %d input slices (s8), pairs above 2^-%d: %2d int8 products, worst relative error %.2e" % (
            name, SW, SX, bits, npairs, err))
