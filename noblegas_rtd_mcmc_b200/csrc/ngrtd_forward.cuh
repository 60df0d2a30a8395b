// ngrtd_forward.cuh -- fused RTD-weight generation (K1) + lag-axis reduction (K2) + mixing/likelihood epilogue.
//
// Restates (as a different algorithm, same arithmetic result to ~1e-13):
//   utils/convolution_integral_utils.py:168-196,270   gen_g_tp   (weights, normalisation)
//   utils/convolution_integral_utils.py:300-340       convolve   (decay/ingrowth, 4He input, dot)
//   age_ens_runs_mcmc/run_age_mcmc_utils.py:81-163    ForwardMod.perform (two components, f1/f2, SF6/CFC rules)
//
// Mapping to the machine:
//   out[chain, col] = sum_k W[chain, k] * X[k, col] is a skinny GEMM (N = 8 folded columns shared by all
//   chains, K = lags).  tcgen05 has no FP64 kind, so the reduction runs on the FP64 pipe as
//   mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4): one warp instruction = 8 chains x 4 lags x 8 columns.  Lane
//   (r = lane>>2, j = lane&3) GENERATES the weight of chain r at lag 4g+j in registers (A fragment), loads
//   X[4g+j][r] from shared memory (B fragment; layout: xf_index() in ngrtd_common.cuh), and owns output columns
//   2j, 2j+1 of chain r (C fragment).  Weights never touch memory.
//     * exponential / exp_pist_flow: geometric recurrence v <- v*r^4, re-anchored by a true exp() at every
//       chunk start; the mask tp >= tau(1-1/eta) is an integer compare against k0.
//     * dispersion: w = tp^-1.5 * exp(-(1-x)^2/(4Dx)); the tp^-1.5 factor is folded into a second copy of X
//       (Xd), the exponent is two FMAs on the per-lag table {1/tp, tp}, exp() is the table-driven, conversion-free
//       exp_scaled_bits().
//     * piston: a gather of one row of X in the epilogue.
//   Normalisation sum = column 0 (ones) of the same MMA.
#pragma once
#include "ngrtd_common.cuh"

#ifndef NGRTD_LOOP_UNROLL
#define NGRTD_LOOP_UNROLL 2      // lag groups per iteration of the steady-state loop (2: the compiler pairs the weight chains)
#endif

namespace ngrtd {

constexpr int LOOP_UNROLL = NGRTD_LOOP_UNROLL;

// All kernels of the library use this one dynamic shared-memory array; sub-arrays are addressed as OFFSETS (in
// doubles) from it, so that every access is provably in the shared address space (pointers kept in a struct were
// demoted to generic LD.E loads inside the sampler kernel: profiles/r1_ncu_mcmc_summary.txt).
extern __shared__ __align__(128) double ngrtd_smem[];

struct SmemView {
    int Xf;
    int Xd;
    int itp;
    int xraw;
    int xrawd;
    int tbl;
    int scratch;   // [warps][NT][8 chains][8 cols]
    int bar;       // mbarrier of the bulk-copy staging (2 doubles reserved)
    int part;      // tape schedule: [warps][NT * NPART][32 lanes] accumulators of partial lag ranges
    int flag;      // tape schedule: [warps] int, 1 = the warp's partial is published
    int park;      // PK: [warps][NT][8 chains][PARK_DOUBLES] epilogue-only values parked across the lag loop
    int ctab;      // CT_DOUBLES column descriptors of the analytic tail (dm_tail)
};

struct LikPar {
    int kind;                  // -1 none, 0 normal, 1 student-t
    double obs[MAX_TRACER];
    double isd[MAX_TRACER];    // 1 / sd
    double lc[MAX_TRACER];     // normal: -0.5 log(2 pi sd^2); student-t: -log(sd)   (per-tracer constants, host)
    const double* nu;          // [B] (student-t)
};

// ---------------------------------------------------------------- per-component lane state
template <int CLS>
struct Comp {};

template <>
struct Comp<CLS_NONE> {
    __device__ __forceinline__ void init(double, double, double, double, int) {}
};

template <>
struct Comp<CLS_P> {
    int ix;
    __device__ __forceinline__ void init(double tau, double, double, double dtp, int L) {
        // argmin_k |tp_k - tau| with tp_0 = 1e-5 + dtp, tp_k = k + dtp; first index wins ties (numpy argmin)
        double s = tau - dtp;
        if (s >= 1.0) {
            double n = ceil(s - 0.5);
            n = fmin(n, (double)(L - 1));
            ix = (int)n;
        } else {
            double d0 = fabs((1e-5 + dtp) - tau);
            double d1 = fabs((1.0 + dtp) - tau);
            ix = (L > 1 && !(d0 <= d1) && d1 == d1) ? 1 : 0;
        }
    }
};

template <>
struct Comp<CLS_G> {
    double v, r4, er, tpk0;
    int k0;
    __device__ __forceinline__ void init(double tau, double eta, double, double dtp, int) {
        // mask threshold with the reference's operation order and roundings (conv utils :189)
        double thr = __dmul_rn(tau, __dsub_rn(1.0, __ddiv_rn(1.0, eta)));
        double tp0 = 1e-5 + dtp;
        if (thr <= tp0) {
            k0 = 0;
        } else {
            double c = ceil(thr - dtp);
            k0 = (c >= 1073741824.0) ? 1073741824 : max(1, (int)c);
        }
        tpk0 = (k0 == 0) ? tp0 : (double)k0 + dtp;
        er = eta / tau;
        r4 = exp(-4.0 * er);
        v = 0.0;
    }
    // same, with the quotients 1/eta (IEEE-rounded: it enters the reference's mask threshold), eta/tau and
    // exp(-4 eta/tau) supplied by the lane-split prologue of WarpTiles::begin
    __device__ __forceinline__ void init_q(double tau, double inv_eta, double er_, double r4_, double dtp) {
        double thr = __dmul_rn(tau, __dsub_rn(1.0, inv_eta));
        double tp0 = 1e-5 + dtp;
        if (thr <= tp0) {
            k0 = 0;
        } else {
            double c = ceil(thr - dtp);
            k0 = (c >= 1073741824.0) ? 1073741824 : max(1, (int)c);
        }
        tpk0 = (k0 == 0) ? tp0 : (double)k0 + dtp;
        er = er_;
        r4 = r4_;
        v = 0.0;
    }
    // direct evaluation at lag k (first group of a chunk); primes the recurrence for lag k+4
    __device__ __forceinline__ double first(int k, double dtp) {
        double vk = exp(-er * (((double)k + dtp) - tpk0));
        v = vk * r4;
        return (k >= k0) ? ((k == 0) ? 1.0 : vk) : 0.0;
    }
    __device__ __forceinline__ double next(int k) {
        double w = (k >= k0) ? v : 0.0;
        v *= r4;
        return w;
    }
};

// Dispersion component.  TB selects the exp table (ExpCfg<TB>).
// The exponent e(tp) = -(tp - tau)^2 / (4 D tau tp) is evaluated in table units (x exp_k<TB>()) as ap/tp + bp*tp + cp.
template <int TB>
struct CompD {
    double ap, bp, cp;
    // exp_scaled_bits() does not propagate NaN -> WarpTiles::end() poisons the sums of dead chains instead
    __device__ __forceinline__ bool dead() const { return ap != ap; }
    // i4D = 1/(4D) and bq = -exp_k/(4 D tau) are supplied by the lane-split prologue of WarpTiles::begin
    __device__ __forceinline__ void init_q(double tau, double D, double i4D, double bq, const PlanView& pv) {
        constexpr double K = exp_k<TB>();
        bool ok = (tau > 0.0) && (D > 0.0) && (tau < 1.0e300) && (D < 1.0e300);
        ap = -K * tau * i4D;
        bp = bq;
        cp = K * 2.0 * i4D;
        // Dead-chain rule: below 2^-1022 every reference weight is (sub)denormal or exactly zero and g/g.sum() is NaN or
        // precision-less -> the chain is declared dead (NaN output) when the LARGEST exponent over the lag grid is below the
        // clamp.  e(tp) is unimodal with its maximum (0) at tp = tau, so the discrete maximum sits at one of the two grid
        // lags around tau (the last lag when tau lies beyond the grid).  Fast path: for 1 <= tau <= tp[L-1] a grid lag lies
        // within 1/2 of tau, where e >= -(1/4D) * 0.25 / (tau (tau - 1/2)); only chains that fail this bound (or lie
        // outside the grid) evaluate the two candidate lags (pv.itp holds the {1/tp, tp} pairs the lag loop reads).
        if (ok && !(tau >= 1.0 + pv.dtp && tau <= pv.tpl &&
                    K * i4D * 0.25 < -(double)exp_clamp<TB>() * (tau * (tau - 0.5)))) {
            const double s = fmin(fmax(floor(tau - pv.dtp), 0.0), (double)(pv.L - 1));
            const int ka = (int)s, kb = min(ka + 1, pv.L - 1);
            const double2 ia = *reinterpret_cast<const double2*>(pv.itp + 2 * ka);
            const double2 ib = *reinterpret_cast<const double2*>(pv.itp + 2 * kb);
            const double emax = fmax(fma(ap, ia.x, fma(bp, ia.y, cp)), fma(ap, ib.x, fma(bp, ib.y, cp)));
            ok = emax >= (double)exp_clamp<TB>();
            // Such chains are evaluated RELATIVE to their largest weight (the normalisation g / g.sum() cancels the factor):
            // exp_scaled_bits clamps at ~2^-1022 instead of flushing to zero, and when the largest weight itself is within
            // a few hundred e-folds of that floor (mode far beyond the lag window: e^-680 at the last lag, the floor
            // hundreds of times across the window) the clamped weights would no longer be negligible.  After the shift the
            // floor sits 708 e-folds below the largest weight.
            if (ok) cp -= emax;
        }
        cp += FX_MAGIC;     // exp_scaled_bits takes ep + FX_MAGIC: folded into the constant term
        if (!ok) ap = __longlong_as_double(0x7ff8000000000000LL);
    }
    // natural-log scale of the loop's weights relative to the unshifted RTD weight: w_loop = w * exp(-shift()) (the analytic
    // tail is summed in absolute terms and rescaled with it); i4D as in init_q.  0 on the fast path.
    __device__ __forceinline__ double shift(double i4D) const {
        constexpr double K = exp_k<TB>();
        return (K * 2.0 * i4D - (cp - FX_MAGIC)) * (1.0 / K);           // cp - FX_MAGIC is exact
    }
    // it = {1/tp, tp} of the lag (shared-memory table): two independent FMAs, no carried state
    __device__ __forceinline__ double weight(double2 it, unsigned int tbl_lane) const {
        return exp_scaled_bits<TB>(fma(ap, it.x, fma(bp, it.y, cp)), ngrtd_smem, tbl_lane);
    }
};
template <int CLS, int TB>
struct CompSel { using type = Comp<CLS>; };
template <int TB>
struct CompSel<CLS_D, TB> { using type = CompD<TB>; };

// order-16 Gauss-Legendre nodes / weights on [-1, 1] (positive half) for dm_tail
__constant__ double DM_GX[8] = {0.095012509837637441, 0.28160355077925892, 0.45801677765722737, 0.61787624440264377,
                                0.755404408355003, 0.86563120238783176, 0.9445750230732326, 0.98940093499164994};
__constant__ double DM_GW[8] = {0.18945061045506864, 0.18260341504492364, 0.16915651939500265, 0.14959598881657671,
                                0.12462897125553407, 0.095158511682492605, 0.062253523938647456, 0.027152459411754176};

// ---------------------------------------------------------------- analytic tail of geometric weights
// G0(x, n) = sum_{i<n} e^{-i x},  G1(x, n) = sum_{i<n} i e^{-i x}   (x = eta/tau + lambda >= 0)
__device__ __forceinline__ double geo_sum0(double x, double n) {
    if (x == 0.0) return n;
    return expm1(-n * x) / expm1(-x);
}
__device__ __forceinline__ double geo_sum1(double x, double n) {
    double nx = n * x;
    if (nx < 1e-4) {            // nearly uniform weights: Taylor in x (the closed form cancels like 2/(n x))
        double m = n - 1.0;
        double S1 = 0.5 * n * m, S2 = m * n * (2.0 * n - 1.0) / 6.0, S3 = S1 * S1;
        return S1 - x * S2 + 0.5 * x * x * S3;
    }
    double om = -expm1(-x);     // 1 - q
    double q = exp(-x), qn = exp(-nx);
    return (q * geo_sum0(x, n) - n * qn) / om;
}
// sum_{k=a}^{L-1} w_k x_k for one folded column; w_k = Wa * exp(-er (k - a)), a = first tail lag with non-zero weight
__device__ __forceinline__ double col_tail(const ColTail& ct, double er, double Wa, double a, double n, double dtp) {
    if (ct.type < 0) return 0.0;
    double g0 = geo_sum0(er, n);
    if (ct.type == 0) return Wa * g0;
    double Da = exp(-ct.lam * (a + dtp));
    double g0l = (ct.lam == 0.0) ? g0 : geo_sum0(er + ct.lam, n);
    if (ct.type == 1) return Wa * ct.bg * Da * g0l;
    if (ct.type == 2) return Wa * ct.bg * (g0 - Da * g0l);
    return Wa * Da * ((ct.i0 + ct.s * a) * g0l + ct.s * geo_sum1(er + ct.lam, n));
}

// ---------------------------------------------------------------- dispersion tail by quadrature (out of line)
// Beyond Kc every folded column is an analytic function of the lag (PlanView::ct), so
//   sum_{k=Kc}^{L-1} w(tp_k) col(tp_k),   w(t) = t^-1.5 exp(-(t - tau)^2 / (4 D tau t))   (the loop's weight incl. Xd's t^-1.5)
// equals the integral over [tp_Kc - 1/2, tp_{L-1} + 1/2] minus the midpoint Euler-Maclaurin end terms g1/24 - 7 g3/5760
// (g1, g3: first and third derivative; the interior error is exponentially small for peaks wider than ~2 lags).
// Order-16 Gauss-Legendre panels, 2 per sigma = tau sqrt(2D) near the mode, widening geometrically away from it and never
// wider than 0.35 t; the 4 lanes of a chain take 4 of the 16 nodes of every panel and all 8 columns, partial sums are
// combined by shuffles.  tools/dm_tail_prototype.py is the float64 prototype (8e-15 against the reference's golden vectors).
// Outside the validated domain (D outside [0.01, 2.5], mode more than 12 sigma beyond the last lag) the same analytic
// terms are summed lag by lag.
// r2: a __noinline__ function shared by every kernel (r1 inlined it per tile and component: 3x the code of the headline
// kernel and extra register pressure around the lag loop).  It reads the column descriptors from a 48-double table that
// FwdCta::setup copies into shared memory (a reference to the kernel-parameter struct would force a local-memory copy).
constexpr int CT_DOUBLES = 48;      // [8 columns][type, bg, lam, i0, s] + {Kc, L, dtp, dyn_bg}
struct CtView {
    const double* t;
    __device__ __forceinline__ int type(int c) const { return (int)t[c * 5]; }
    __device__ __forceinline__ double bg(int c) const { return t[c * 5 + 1]; }
    __device__ __forceinline__ double lam(int c) const { return t[c * 5 + 2]; }
    __device__ __forceinline__ double i0(int c) const { return t[c * 5 + 3]; }
    __device__ __forceinline__ double s(int c) const { return t[c * 5 + 4]; }
    __device__ __forceinline__ double Kc() const { return t[40]; }
    __device__ __forceinline__ double L() const { return t[41]; }
    __device__ __forceinline__ double dtp() const { return t[42]; }
    __device__ __forceinline__ double dyn_bg() const { return t[43]; }
};
__device__ __forceinline__ void ct_fill(double* tab, const PlanView& pv, int tid) {
    if (tid < NCOL) {
        const ColTail c = pv.ct[tid];
        tab[tid * 5] = (double)c.type; tab[tid * 5 + 1] = c.bg; tab[tid * 5 + 2] = c.lam; tab[tid * 5 + 3] = c.i0; tab[tid * 5 + 4] = c.s;
    }
    if (tid == NCOL) { tab[40] = (double)pv.Kc; tab[41] = (double)pv.L; tab[42] = pv.dtp; tab[43] = pv.dyn_bg; }
}
// lam_dyn / accd: per-chain decay constant (thalf_cfc) and the sum of its column dyn_bg * exp(-lam_dyn t) (DYN plans)
// sh: natural-log scale of the lag loop's weights (CompD::shift), so that loop and tail sums share one scale
template <bool DYN>
__device__ __forceinline__ void dm_node(const CtView& cv, double t, double wt, double tau, double c4, double sh,
                                        double (&acc)[NCOL], double lam_dyn, double& accd) {
    const double dt = t - tau;
    const double rs = rsqrt(t);                   // t^-1.5 without a square root and a division (tail launch -6 %)
    const double wgt = wt * exp(-dt * dt * c4 / t - sh) * (rs * rs * rs);
    acc[0] += wgt;
    if constexpr (DYN) accd = fma(wgt, cv.dyn_bg() * exp(-lam_dyn * t), accd);
    double last = 0.0, d = 1.0;
#pragma unroll
    for (int c = 1; c < NCOL; c++) {
        const int ty = cv.type(c);
        if (ty < 0) continue;
        const double lam = cv.lam(c);
        if (lam != last) { d = exp(-lam * t); last = lam; }
        double v;
        if (ty == 1) v = cv.bg(c) * d;
        else if (ty == 2) v = cv.bg(c) * (1.0 - d);
        else v = (cv.i0(c) + cv.s(c) * (t - cv.dtp())) * d;
        acc[c] = fma(wgt, v, acc[c]);
    }
}
// sgn * (g1/24 - 7 g3/5760) of every column at the end point t
template <bool DYN>
__device__ __forceinline__ void dm_end(const CtView& cv, double t, double sgn, double tau, double a, double c4, double sh,
                                       double (&acc)[NCOL], double lam_dyn, double& accd) {
    const double dt = t - tau, it = 1.0 / t;
    const double w = exp(-dt * dt * c4 * it - sh) * it / sqrt(t);
    const double p1 = (1.5 - 2.0 * a * it) * it * it, p2 = (-3.0 + 6.0 * a * it) * it * it * it;
    const double pb = (-1.5 + a * it) * it - c4;                  // log-derivative without the decay constant
    auto terms = [&](double lam, double& g, double& g1, double& g2, double& g3) {
        const double p = pb - lam;
        g = w * exp(-lam * t);
        g1 = g * p; g2 = g * (p * p + p1); g3 = g * (p * p * p + 3.0 * p * p1 + p2);
    };
    double g, g1, g2, g3;
    terms(0.0, g, g1, g2, g3);
    const double k1 = sgn / 24.0, k3 = -sgn * 7.0 / 5760.0;
    acc[0] += k1 * g1 + k3 * g3;
    if constexpr (DYN) {
        double h, h1, h2, h3;
        terms(lam_dyn, h, h1, h2, h3);
        accd += cv.dyn_bg() * (k1 * h1 + k3 * h3);
    }
#pragma unroll 1
    for (int c = 1; c < NCOL; c++) {
        const int ty = cv.type(c);
        if (ty < 0) continue;
        double h, h1, h2, h3;
        terms(cv.lam(c), h, h1, h2, h3);
        if (ty == 1) acc[c] += cv.bg(c) * (k1 * h1 + k3 * h3);
        else if (ty == 2) acc[c] += cv.bg(c) * (k1 * (g1 - h1) + k3 * (g3 - h3));
        else {
            const double q = cv.i0(c) + cv.s(c) * (t - cv.dtp());
            acc[c] += k1 * (cv.s(c) * h + q * h1) + k3 * (3.0 * cv.s(c) * h2 + q * h3);
        }
    }
}
// Tail sums of one dispersion component of one chain, computed by the chain's 4 lanes (j = lane & 3; every lane of the warp
// calls).  res[0], res[1]: the tail of the lane's two output columns (2j, 2j+1); res[2]: the per-chain-decay column
// (lane 0 of the chain carries the total, the others 0).  ct_off: offset of the CT_DOUBLES table in ngrtd_smem.
template <bool DYN>
__device__ __noinline__ void dm_tail(int ct_off, double tau, double D, int dead_j, double sh, double lam_dyn, double* res) {
    const CtView cv{ngrtd_smem + ct_off};
    const int j = dead_j & 3;
    const bool dead = (dead_j >> 2) != 0;
    double acc[NCOL];
    double accd = 0.0;
#pragma unroll
    for (int c = 0; c < NCOL; c++) acc[c] = 0.0;
    if (!dead) {
        const double Kc = cv.Kc(), Ld = cv.L(), dtp = cv.dtp();
        const double lo = Kc - 0.5 + dtp, hi = Ld - 0.5 + dtp;
        const double c4 = 1.0 / (4.0 * D * tau), a = tau / (4.0 * D);
        const double sig = fmax(tau * sqrt(2.0 * D), 1.0);
        bool quad = D >= 0.01 && D <= 2.5 && tau - 12.0 * sig <= hi;
        double wlo = fmax(lo, tau - 12.0 * sig);
        double whi = fmin(hi, fmax(fmax(tau + 60.0 * sig, tau + 200.0 * D * tau), lo + 1.0));
        // |d log w / dt| at t: the rate at which the integrand changes per lag
        auto rate = [&](double t) { const double it = 1.0 / t; return fabs((-1.5 + a * it) * it - c4); };
        // The two-term Euler-Maclaurin end correction is only as good as the integrand is smooth AT the end point (next term
        // 31 g5 / 967680 ~ 3.2e-5 rate^5 g): when an end carries weight and the weights change fast there -- mode far beyond
        // the last lag, weights growing by a factor > 1.05 per lag towards it -- the end point is moved inwards, 4 lags at a
        // time, until that term is below 1e-14 of the window's largest weight, and the lags cut off are summed one by one.
        // (found by tools/fuzz_forward.py: 2e-7 .. 8e-7 in the lag-index column for tau = 11 .. 26 window lengths, D = 0.01)
        int mlo = 0, mhi = 0;
        const double tmx = fmin(fmax(tau, 1e-5 + dtp), hi - 0.5);
        const double dm = tmx - tau, lwm = -dm * dm * c4 / tmx;    // log of the largest weight (without its t^-1.5)
        if (quad) {
            const double ltm = 1.5 * log(tmx);
            auto em_bad = [&](double t) {
                const double dt = t - tau;
                const double lw = -dt * dt * c4 / t - lwm + ltm - 1.5 * log(t);             // log(weight / largest weight)
                return -10.35 + 5.0 * log(rate(t)) + lw > -32.2;                           // log 3.2e-5, log 1e-14
            };
            constexpr int EM_CAP = 4096;
            if (wlo == lo) while (mlo < EM_CAP && em_bad(lo + mlo)) mlo += 4;
            if (whi == hi) while (mhi < EM_CAP && em_bad(hi - mhi)) mhi += 4;
            if (mlo >= EM_CAP || mhi >= EM_CAP || lo + mlo + 1.0 >= hi - mhi) quad = false;
        }
        if (quad) {
            if (mlo | mhi) {
                const int ka = (int)Kc, kb = (int)Ld;
#pragma unroll 1
                for (int k = ka + j; k < ka + mlo; k += 4) dm_node<DYN>(cv, (double)k + dtp, 1.0, tau, c4, sh, acc, lam_dyn, accd);
#pragma unroll 1
                for (int k = kb - mhi + j; k < kb; k += 4) dm_node<DYN>(cv, (double)k + dtp, 1.0, tau, c4, sh, acc, lam_dyn, accd);
            }
            const double elo = lo + mlo, ehi = hi - mhi;          // end points of the integral (half a lag outside the kept lags)
            if (wlo == lo) wlo = elo;
            if (whi == hi) whi = ehi;
            if (whi > wlo) {
                const double w = 0.5 * sig;
                double x = wlo;
                const double m15 = (tmx > lo) ? 1.5 * log(tmx / lo) : 0.0;
                const double skip0 = 56.0 + m15 - lwm;             // lwm <= 0
                double lam_max = DYN ? lam_dyn : 0.0;
#pragma unroll
                for (int c = 1; c < NCOL; c++) if (cv.type(c) >= 0) lam_max = fmax(lam_max, cv.lam(c));
                while (x < whi) {
                    const double ad = fabs(x - tau);
                    double step = (ad < 6.0 * sig) ? w : fmax(w, 0.25 * ad);
                    step = fmin(step, 0.35 * x);
                    double x1 = fmin(whi, x + step);
                    // A panel that cannot matter to any column is skipped (far from the mode of a narrow RTD the rate bound below
                    // would otherwise cut tens of panels out of nothing).  Bound of its integrand against the integrand at the
                    // window's largest weight, t = tmx, in logs: the exponential factor at the panel's point nearest the mode,
                    // + the growth of t^-1.5 over the window (m15), + the growth of the fastest-decaying column towards small t
                    // (lam_max (tmx - x)); ingrowth / lag-index columns grow by less than e^6 towards large t.  Below e^-56 the
                    // panels skipped over a window of e^10 lags add < e^-46 = 1e-20 of any column sum.
                    const double tn = (x1 <= tau) ? x1 : ((x >= tau) ? x : tau);
                    const double dn = tn - tau;
                    if (dn * dn * c4 > (skip0 + lam_max * fmax(tmx - x, 0.0)) * tn) { x = x1; continue; }    // no division
                    // a 16-node panel integrates exp(c x) on [-1, 1] to 1e-14 up to c ~ 4: keep rate * half-width below that
                    // (left of the mode the rate falls with t, right of it it is bounded by c4 + 1.5 / t)
                    step = fmin(step, 8.0 / fmax(x < tau ? rate(x) : c4 + 1.5 / x, 1e-300));
                    x1 = fmin(whi, x + step);
                    {   // the 4 lanes of the chain take 4 of the panel's 16 nodes each (no divergence inside a chain)
                        const double mid = 0.5 * (x1 + x), half = 0.5 * (x1 - x);
#pragma unroll 1
                        for (int i = 0; i < 4; i++) {
                            const int q = 4 * j + i;
                            const double gx = (q & 1) ? -DM_GX[q >> 1] : DM_GX[q >> 1];
                            dm_node<DYN>(cv, fma(half, gx, mid), half * DM_GW[q >> 1], tau, c4, sh, acc, lam_dyn, accd);
                        }
                    }
                    x = x1;
                }
                if (j == 0 && wlo == elo) dm_end<DYN>(cv, elo, 1.0, tau, a, c4, sh, acc, lam_dyn, accd);
                if (j == 1 && whi == ehi) dm_end<DYN>(cv, ehi, -1.0, tau, a, c4, sh, acc, lam_dyn, accd);
            }
        } else {
            const int k0 = (int)Kc, k1 = (int)Ld;
#pragma unroll 1
            for (int k = k0 + j; k < k1; k += 4) dm_node<DYN>(cv, (double)k + dtp, 1.0, tau, c4, sh, acc, lam_dyn, accd);
        }
    }
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int c = 0; c < NCOL; c++) {
        acc[c] += __shfl_xor_sync(full, acc[c], 1);
        acc[c] += __shfl_xor_sync(full, acc[c], 2);
    }
    res[0] = (j == 0) ? acc[0] : (j == 1) ? acc[2] : (j == 2) ? acc[4] : acc[6];
    res[1] = (j == 0) ? acc[1] : (j == 1) ? acc[3] : (j == 2) ? acc[5] : acc[7];
    res[2] = 0.0;
    if constexpr (DYN) {          // the chain's 4 lanes are summed again in end(): hand the total to lane 0 only
        accd += __shfl_xor_sync(full, accd, 1);
        accd += __shfl_xor_sync(full, accd, 2);
        if (j == 0) res[2] = accd;
    }
}

// ---------------------------------------------------------------- one warp = NT tiles of 8 chains
// TM: 1 = the analytic constant tail is decided at run time (PlanView::Kc), 0 = compiled out (k_forward instantiates both
//     and the launcher picks: without the tail code the lag loop keeps fewer values alive and schedules tighter).
// PK: the values only the epilogue needs (f1, f2, lamsf6, J) are parked in shared memory across the lag loop instead of
//     occupying 16 registers (k_forward; the sampler kernel has no shared memory to spare and keeps them in registers).
constexpr int PARK_DOUBLES = 4;
template <int C1, int C2, bool DYN, int NT, int UA, int TB, int TM = 1, bool PK = false>
struct WarpTiles {
    static constexpr bool LOOP1 = (C1 == CLS_G || C1 == CLS_D);
    static constexpr bool LOOP2 = (C2 == CLS_G || C2 == CLS_D);
    static constexpr bool ANY_LOOP = LOOP1 || LOOP2;
    static constexpr bool ANY_D = (C1 == CLS_D || C2 == CLS_D);
    static constexpr bool ANY_G = (C1 == CLS_G || C2 == CLS_G);
    static constexpr int TBITS = TB, NTILES = NT;
    static constexpr bool DYNAMIC = DYN, PARK = PK;
    static constexpr int TAILMODE = TM;

    static constexpr double EXP_K = exp_k<TB>();
    // doubles per lane and tile that a partial lag range hands to the owner of its unit (tape schedule of k_forward)
    static constexpr int NPART = (LOOP1 ? 2 : 0) + (LOOP2 ? 2 : 0) + ((DYN && LOOP1) ? 1 : 0) + ((DYN && LOOP2) ? 1 : 0);

    typename CompSel<C1, TB>::type c1[NT];
    typename CompSel<C2, TB>::type c2[NT];
    double a1[NT][UA][2], a2[NT][UA][2];   // UA independent DMMA accumulator chains per tile and component
    double dv[NT], d4[NT], lam[NT], ad1[NT], ad2[NT];
    double Jl[NT];                          // J = 10**log10J (run_age_mcmc_utils.py:101)

    // Per-chain set-up.  The 4 lanes (r, 0..3) of a chain would otherwise repeat the same divisions and exponentials
    // (~25-40 FP64 instructions each on the pipe the lag loops of the other warps need), so the work is split: every
    // lane performs ONE division and ONE exp() with its own operands and the results are exchanged by shuffles.
    //   division slots: lanes 0,1 component 1, lanes 2,3 component 2
    //       exponential class: 1/eta (IEEE-rounded, it enters the mask threshold) and eta/tau
    //       dispersion:        0.25/D = 1/(4D) and -(N/ln2)/4 / (D tau)
    //   exp slots: lane 0 J = exp(ln10 * log10 J) (argument product carried in two pieces), lane 1 / 2 the 4-lag decay
    //       factor exp(-4 eta/tau) of component 1 / 2, lane 3 exp(-4 lambda) of a per-chain decay constant.
    __device__ __forceinline__ void begin(const ChainPar (&p)[NT], const PlanView& pv, int lane, bool need_J,
                                          double* park_warp = nullptr) {
        const int j = lane & 3, base = lane & ~3;
        const unsigned full = 0xffffffffu;
#pragma unroll
        for (int t = 0; t < NT; t++) {
            double q[4] = {0.0, 0.0, 0.0, 0.0};
            if constexpr (LOOP1 || LOOP2) {
                double num = 1.0, den = 1.0;
                if constexpr (C1 == CLS_G) { if (j == 0) den = p[t].eta1; if (j == 1) { num = p[t].eta1; den = p[t].tau1; } }
                if constexpr (C1 == CLS_D) { if (j == 0) { num = 0.25; den = p[t].D1; } if (j == 1) { num = -0.25 * EXP_K; den = p[t].D1 * p[t].tau1; } }
                if constexpr (C2 == CLS_G) { if (j == 2) den = p[t].eta2; if (j == 3) { num = p[t].eta2; den = p[t].tau2; } }
                if constexpr (C2 == CLS_D) { if (j == 2) { num = 0.25; den = p[t].D2; } if (j == 3) { num = -0.25 * EXP_K; den = p[t].D2 * p[t].tau2; } }
                const double qq = __ddiv_rn(num, den);
                if constexpr (LOOP1) { q[0] = __shfl_sync(full, qq, base); q[1] = __shfl_sync(full, qq, base + 1); }
                if constexpr (LOOP2) { q[2] = __shfl_sync(full, qq, base + 2); q[3] = __shfl_sync(full, qq, base + 3); }
            }
            double e[4] = {0.0, 0.0, 0.0, 0.0};
            {
                constexpr double LN10_HI = 2.302585092994045901, LN10_LO = -2.1707562233822494e-16;
                double ea = 0.0, el = 0.0;
                if (j == 0 && need_J) {
                    ea = p[t].log10J * LN10_HI;
                    el = fma(p[t].log10J, LN10_HI, -ea) + p[t].log10J * LN10_LO;
                }
                if constexpr (C1 == CLS_G) { if (j == 1) ea = -4.0 * q[1]; }
                if constexpr (C2 == CLS_G) { if (j == 2) ea = -4.0 * q[3]; }
                if (DYN) { if (j == 3) ea = -4.0 * p[t].lam_cfc; }
                double ee = exp(ea);
                ee = fma(ee, el, ee);
                Jl[t] = need_J ? __shfl_sync(full, ee, base) : 0.0;
                if constexpr (PK) {
                    if (j == 0) {
                        double* pk = park_warp + (t * 8 + (lane >> 2)) * PARK_DOUBLES;
                        pk[0] = p[t].f1; pk[1] = p[t].f2; pk[2] = p[t].lamsf6; pk[3] = Jl[t];
                    }
                }
                if constexpr (C1 == CLS_G) e[1] = __shfl_sync(full, ee, base + 1);
                if constexpr (C2 == CLS_G) e[2] = __shfl_sync(full, ee, base + 2);
                if (DYN) e[3] = __shfl_sync(full, ee, base + 3);
            }
            if constexpr (C1 == CLS_G) c1[t].init_q(p[t].tau1, q[0], q[1], e[1], pv.dtp);
            else if constexpr (C1 == CLS_D) c1[t].init_q(p[t].tau1, p[t].D1, q[0], q[1], pv);
            else c1[t].init(p[t].tau1, p[t].eta1, p[t].D1, pv.dtp, pv.L);
            if constexpr (C2 == CLS_G) c2[t].init_q(p[t].tau2, q[2], q[3], e[2], pv.dtp);
            else if constexpr (C2 == CLS_D) c2[t].init_q(p[t].tau2, p[t].D2, q[2], q[3], pv);
            else c2[t].init(p[t].tau2, p[t].eta2, p[t].D2, pv.dtp, pv.L);
#pragma unroll
            for (int u = 0; u < UA; u++) a1[t][u][0] = a1[t][u][1] = a2[t][u][0] = a2[t][u][1] = 0.0;
            if (DYN) {
                lam[t] = p[t].lam_cfc;
                d4[t] = e[3];
                dv[t] = 0.0;
                ad1[t] = ad2[t] = 0.0;
            }
        }
    }

    // accumulate lags [kc, kc + 4*ngroups); shared memory holds lag kc at local lag index kl (a multiple of 4: 0 for a
    // streamed chunk, kc when the whole table is resident)
    __device__ __forceinline__ void chunk(const SmemView& s, const PlanView& pv, int kc, int ngroups, int lane, int kl = 0) {
        if (!ANY_LOOP) return;
        const int j = lane & 3, r = lane >> 2;
        int k = kc + j;
        const double* pf = ngrtd_smem + s.Xf + xf_index(kl + j, r);     // chunks start at multiples of 4 lags
        const double* pd = ngrtd_smem + s.Xd + xf_index(kl + j, r);
        const double2* pi = reinterpret_cast<const double2*>(ngrtd_smem + s.itp) + kl + j;
        const double* px = ngrtd_smem + s.xraw + kl + j;
        const double* pxd = ngrtd_smem + s.xrawd + kl + j;
        const unsigned int tbl = (unsigned int)(lane & (ExpCfg<TB>::REP - 1)) << 3;   // byte offset of this lane's copy of the exp table
        const double dtp = pv.dtp;
        {   // first group of the chunk: direct evaluation (handles tp_0 = 1e-5 and re-anchors the recurrences)
            double bf = pf[0];
            double bd = ANY_D ? pd[0] : 0.0;
            double2 it = ANY_D ? pi[0] : make_double2(0.0, 0.0);
            double xr = DYN ? px[0] : 0.0;
            double xrd = (DYN && ANY_D) ? pxd[0] : 0.0;
#pragma unroll
            for (int t = 0; t < NT; t++) {
                double w1 = 0.0, w2 = 0.0;
                if constexpr (C1 == CLS_G) { w1 = c1[t].first(k, dtp); dmma884(a1[t][0][0], a1[t][0][1], w1, bf); }
                if constexpr (C1 == CLS_D) { w1 = c1[t].weight(it, tbl); dmma884(a1[t][0][0], a1[t][0][1], w1, bd); }
                if constexpr (C2 == CLS_G) { w2 = c2[t].first(k, dtp); dmma884(a2[t][0][0], a2[t][0][1], w2, bf); }
                if constexpr (C2 == CLS_D) { w2 = c2[t].weight(it, tbl); dmma884(a2[t][0][0], a2[t][0][1], w2, bd); }
                if constexpr (DYN) {
                    double dvk = exp(-lam[t] * ((double)k + dtp));
                    double du = (k == 0) ? exp(-lam[t] * (1e-5 + dtp)) : dvk;
                    dv[t] = dvk * d4[t];
                    if constexpr (LOOP1) ad1[t] = fma(w1 * du, (C1 == CLS_D) ? xrd : xr, ad1[t]);
                    if constexpr (LOOP2) ad2[t] = fma(w2 * du, (C2 == CLS_D) ? xrd : xr, ad2[t]);
                }
            }
        }
        // steady state: UA groups per iteration, each group feeding its own accumulator set so that
        // consecutive DMMAs of one tile are independent (hides the DMMA dependent-issue latency)
        int g = 1;
        // {1/tp, tp} of the lane's lag.  Beyond the first group of a chunk tp = k + dtp is an exactly representable
        // integer, so with -DNGRTD_TP_DADD it is carried in a register (+= 4.0, exact) and only 1/tp is loaded: one LDS.64
        // (2 wavefronts) instead of one LDS.128 (4 wavefronts) per group, for one DADD per group shared by the NT tiles.
        // Experiment (profiles/r1_notes.md, "session 3"): forward launch -1.6 %, sampler step +4.5 % -> not the default.
#ifdef NGRTD_TP_DADD
        double tpk = (double)k + dtp;
#define NGRTD_LOAD_IT(it_) do { tpk += 4.0; (it_).x = pi[0].x; (it_).y = tpk; } while (0)
#else
#define NGRTD_LOAD_IT(it_) do { (it_) = pi[0]; } while (0)
#endif
#ifdef NGRTD_BURST
        // experiment (profiles/r1_notes.md, "burst"): weights of TWO lag groups first, then their DMMAs back to back, so that
        // the shared pipe alternates less often between DFMA and DMMA work
        if constexpr (!DYN) {
            for (; g + 2 <= ngroups; g += 2) {
                double bf[2], bd[2] = {0.0, 0.0};
                double w1[NT][2], w2[NT][2];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    k += 4;
                    pf += 4 * NCOL;
                    bf[u] = pf[0];
                    double2 it = make_double2(0.0, 0.0);
                    if constexpr (ANY_D) { pd += 4 * NCOL; pi += 4; bd[u] = pd[0]; it = pi[0]; }
#pragma unroll
                    for (int t = 0; t < NT; t++) {
                        w1[t][u] = w2[t][u] = 0.0;
                        if constexpr (C1 == CLS_G) w1[t][u] = c1[t].next(k);
                        if constexpr (C1 == CLS_D) w1[t][u] = c1[t].weight(it, tbl);
                        if constexpr (C2 == CLS_G) w2[t][u] = c2[t].next(k);
                        if constexpr (C2 == CLS_D) w2[t][u] = c2[t].weight(it, tbl);
                    }
                }
#pragma unroll
                for (int t = 0; t < NT; t++)
                    asm volatile("" : "+d"(w1[t][0]), "+d"(w1[t][1]), "+d"(w2[t][0]), "+d"(w2[t][1]));
#pragma unroll
                for (int u = 0; u < 2; u++) {
#pragma unroll
                    for (int t = 0; t < NT; t++) {
                        if constexpr (LOOP1) dmma884(a1[t][0][0], a1[t][0][1], w1[t][u], (C1 == CLS_D) ? bd[u] : bf[u]);
                        if constexpr (LOOP2) dmma884(a2[t][0][0], a2[t][0][1], w2[t][u], (C2 == CLS_D) ? bd[u] : bf[u]);
                    }
                }
            }
        }
#endif
#pragma unroll LOOP_UNROLL
        for (; g + UA <= ngroups; g += UA) {
#pragma unroll
            for (int u = 0; u < UA; u++) {
                k += 4;
                pf += 4 * NCOL;
                double bf = pf[0];
                double bd = 0.0, xr = 0.0, xrd = 0.0;
                double2 it = make_double2(0.0, 0.0);
                if constexpr (ANY_D) { pd += 4 * NCOL; pi += 4; bd = pd[0]; NGRTD_LOAD_IT(it); }
                if constexpr (DYN) { px += 4; xr = px[0]; if constexpr (ANY_D) { pxd += 4; xrd = pxd[0]; } }
#pragma unroll
                for (int t = 0; t < NT; t++) {
                    double w1 = 0.0, w2 = 0.0;
                    if constexpr (C1 == CLS_G) { w1 = c1[t].next(k); dmma884(a1[t][u][0], a1[t][u][1], w1, bf); }
                    if constexpr (C1 == CLS_D) { w1 = c1[t].weight(it, tbl); dmma884(a1[t][u][0], a1[t][u][1], w1, bd); }
                    if constexpr (C2 == CLS_G) { w2 = c2[t].next(k); dmma884(a2[t][u][0], a2[t][u][1], w2, bf); }
                    if constexpr (C2 == CLS_D) { w2 = c2[t].weight(it, tbl); dmma884(a2[t][u][0], a2[t][u][1], w2, bd); }
                    if constexpr (DYN) {
                        double du = dv[t];
                        dv[t] *= d4[t];
                        if constexpr (LOOP1) ad1[t] = fma(w1 * du, (C1 == CLS_D) ? xrd : xr, ad1[t]);
                        if constexpr (LOOP2) ad2[t] = fma(w2 * du, (C2 == CLS_D) ? xrd : xr, ad2[t]);
                    }
                }
            }
        }
        if constexpr (UA > 1)
        for (; g < ngroups; g++) {   // remainder groups
            k += 4;
            pf += 4 * NCOL;
            double bf = pf[0];
            double bd = 0.0, xr = 0.0, xrd = 0.0;
                double2 it = make_double2(0.0, 0.0);
            if constexpr (ANY_D) { pd += 4 * NCOL; pi += 4; bd = pd[0]; NGRTD_LOAD_IT(it); }
            if constexpr (DYN) { px += 4; xr = px[0]; if constexpr (ANY_D) { pxd += 4; xrd = pxd[0]; } }
#pragma unroll
            for (int t = 0; t < NT; t++) {
                double w1 = 0.0, w2 = 0.0;
                if constexpr (C1 == CLS_G) { w1 = c1[t].next(k); dmma884(a1[t][0][0], a1[t][0][1], w1, bf); }
                if constexpr (C1 == CLS_D) { w1 = c1[t].weight(it, tbl); dmma884(a1[t][0][0], a1[t][0][1], w1, bd); }
                if constexpr (C2 == CLS_G) { w2 = c2[t].next(k); dmma884(a2[t][0][0], a2[t][0][1], w2, bf); }
                if constexpr (C2 == CLS_D) { w2 = c2[t].weight(it, tbl); dmma884(a2[t][0][0], a2[t][0][1], w2, bd); }
                if constexpr (DYN) {
                    double du = dv[t];
                    dv[t] *= d4[t];
                    if constexpr (LOOP1) ad1[t] = fma(w1 * du, (C1 == CLS_D) ? xrd : xr, ad1[t]);
                    if constexpr (LOOP2) ad2[t] = fma(w2 * du, (C2 == CLS_D) ? xrd : xr, ad2[t]);
                }
            }
        }
    }

#undef NGRTD_LOAD_IT

    // ---- tape schedule of k_forward: a unit's lag range may be cut between warps.  The warp that holds the range starting
    // at lag 0 owns the unit (it runs end()); every other range stores its accumulators (NPART doubles per lane and tile,
    // lane-contiguous: conflict-free) for the owner to add.  All ranges use the same weight scale (geometric weights are
    // relative to tp_k0, dispersion weights are absolute), so partial sums simply add.
    __device__ __forceinline__ void store_partial(double* slot, int lane) const {
        int i = 0;
#pragma unroll
        for (int t = 0; t < NT; t++) {
            if constexpr (LOOP1) {
                double x0 = a1[t][0][0], x1 = a1[t][0][1];
#pragma unroll
                for (int u = 1; u < UA; u++) { x0 += a1[t][u][0]; x1 += a1[t][u][1]; }
                slot[(i++) * 32 + lane] = x0; slot[(i++) * 32 + lane] = x1;
                if constexpr (DYN) slot[(i++) * 32 + lane] = ad1[t];
            }
            if constexpr (LOOP2) {
                double x0 = a2[t][0][0], x1 = a2[t][0][1];
#pragma unroll
                for (int u = 1; u < UA; u++) { x0 += a2[t][u][0]; x1 += a2[t][u][1]; }
                slot[(i++) * 32 + lane] = x0; slot[(i++) * 32 + lane] = x1;
                if constexpr (DYN) slot[(i++) * 32 + lane] = ad2[t];
            }
        }
    }
    __device__ __forceinline__ void add_partial(const double* slot, int lane) {
        int i = 0;
#pragma unroll
        for (int t = 0; t < NT; t++) {
            if constexpr (LOOP1) {
                a1[t][0][0] += slot[(i++) * 32 + lane]; a1[t][0][1] += slot[(i++) * 32 + lane];
                if constexpr (DYN) ad1[t] += slot[(i++) * 32 + lane];
            }
            if constexpr (LOOP2) {
                a2[t][0][0] += slot[(i++) * 32 + lane]; a2[t][0][1] += slot[(i++) * 32 + lane];
                if constexpr (DYN) ad2[t] += slot[(i++) * 32 + lane];
            }
        }
    }

    // normalise, mix the two components, apply tracer rules; lane (r, j) returns the outputs of tracers
    // j, j+4 of chain r in val[0..1] (NaN-propagating exactly like f1*cout1 + f2*cout2 of the reference)
    __device__ __forceinline__ void end(const ChainPar (&p)[NT], const PlanView& pv, double* scratch_warp, int lane,
                                        double (&val)[NT][2], const double* park_warp = nullptr, int ct_off = 0) {
        const int j = lane & 3, r = lane >> 2;
        const unsigned full = 0xffffffffu;
        if constexpr (PK) __syncwarp();
#pragma unroll
        for (int t = 0; t < NT; t++) {
            double pf1, pf2, plam, pJ;
            if constexpr (PK) {
                const double* pk = park_warp + (t * 8 + r) * PARK_DOUBLES;
                pf1 = pk[0]; pf2 = pk[1]; plam = pk[2]; pJ = pk[3];
            } else {
                pf1 = p[t].f1; pf2 = p[t].f2; plam = p[t].lamsf6; pJ = Jl[t];
            }
#pragma unroll
            for (int u = 1; u < UA; u++) {
                a1[t][0][0] += a1[t][u][0]; a1[t][0][1] += a1[t][u][1];
                a2[t][0][0] += a2[t][u][0]; a2[t][0][1] += a2[t][u][1];
            }
            if (TM != 0 && tail_active(pv, ANY_G, ANY_D)) {         // analytic tail [Kc, L): closed form (G), quadrature (D, NGRTD_DM_TAIL)
                const double dtp = pv.dtp;
                if constexpr (C1 == CLS_G) {
                    double a = (double)max(c1[t].k0, pv.Kc), n = (double)pv.L - a;
                    if (n > 0.0) {
                        double Wa = exp(-c1[t].er * ((a + dtp) - c1[t].tpk0));
                        a1[t][0][0] += col_tail(pv.ct[2 * j], c1[t].er, Wa, a, n, dtp);
                        a1[t][0][1] += col_tail(pv.ct[2 * j + 1], c1[t].er, Wa, a, n, dtp);
                        if (DYN && j == 0)
                            ad1[t] += Wa * pv.dyn_bg * exp(-p[t].lam_cfc * (a + dtp)) * geo_sum0(c1[t].er + p[t].lam_cfc, n);
                    }
                }
                if constexpr (C2 == CLS_G) {
                    double a = (double)max(c2[t].k0, pv.Kc), n = (double)pv.L - a;
                    if (n > 0.0) {
                        double Wa = exp(-c2[t].er * ((a + dtp) - c2[t].tpk0));
                        a2[t][0][0] += col_tail(pv.ct[2 * j], c2[t].er, Wa, a, n, dtp);
                        a2[t][0][1] += col_tail(pv.ct[2 * j + 1], c2[t].er, Wa, a, n, dtp);
                        if (DYN && j == 0)
                            ad2[t] += Wa * pv.dyn_bg * exp(-p[t].lam_cfc * (a + dtp)) * geo_sum0(c2[t].er + p[t].lam_cfc, n);
                    }
                }
                if constexpr (DM_TAIL && C1 == CLS_D) {
                    double res[3];
                    dm_tail<DYN>(ct_off, p[t].tau1, p[t].D1, j | (c1[t].dead() ? 4 : 0), c1[t].shift(__ddiv_rn(0.25, p[t].D1)),
                                 p[t].lam_cfc, res);
                    a1[t][0][0] += res[0]; a1[t][0][1] += res[1];
                    if (DYN) ad1[t] += res[2];
                }
                if constexpr (DM_TAIL && C2 == CLS_D) {
                    double res[3];
                    dm_tail<DYN>(ct_off, p[t].tau2, p[t].D2, j | (c2[t].dead() ? 4 : 0), c2[t].shift(__ddiv_rn(0.25, p[t].D2)),
                                 p[t].lam_cfc, res);
                    a2[t][0][0] += res[0]; a2[t][0][1] += res[1];
                    if (DYN) ad2[t] += res[2];
                }
            }
            double m[2], md = 0.0;
            double x1[2], x2[2] = {0.0, 0.0}, xd1 = 0.0, xd2 = 0.0;
            // Normalisation: the sums S1, S2 (column 0) are inverted ONCE per chain -- lane 0 takes 1/S1, lane 1 1/S2 --
            // and broadcast, instead of 4-6 divisions in every lane.  a * (1/S) differs from a / S by <= 1.5 ulp; when
            // 1/S would leave the normal range (S subnormal: every weight underflowed) the plain division is kept so
            // that such chains behave exactly as before (decided per chain: results never depend on warp neighbours).
            double S1 = 1.0, S2 = 1.0, r1 = 1.0, r2 = 1.0;
            bool use_div = false, any_div = false;
            if constexpr (LOOP1) {
                S1 = __shfl_sync(full, a1[t][0][0], lane & ~3);
                if constexpr (C1 == CLS_D) { if (c1[t].dead()) S1 = c1[t].ap; }
            }
            if constexpr (LOOP2) {
                S2 = __shfl_sync(full, a2[t][0][0], lane & ~3);
                if constexpr (C2 == CLS_D) { if (c2[t].dead()) S2 = c2[t].ap; }
            }
            if constexpr (LOOP1 || LOOP2) {
                const double den = (LOOP1 && (!LOOP2 || (j & 1) == 0)) ? S1 : S2;
                const double rr = __ddiv_rn(1.0, den);
                // (S == 0 exactly -- every weight masked or flushed -- needs no division either: all weights are >= 0, so the
                //  column sums are 0 as well and 0 * (1/0) = 0 * inf = NaN is the reference's 0/0 = NaN)
                use_div = (LOOP1 && S1 != 0.0 && fabs(S1) < 1e-290) || (LOOP2 && S2 != 0.0 && fabs(S2) < 1e-290);   // per chain, not per warp
                any_div = __any_sync(full, use_div);
                if constexpr (LOOP1) r1 = __shfl_sync(full, rr, lane & ~3);
                if constexpr (LOOP2) r2 = __shfl_sync(full, rr, (lane & ~3) + (LOOP1 ? 1 : 0));
            }
            // component 1
            if constexpr (C1 == CLS_P) {
                int ix = c1[t].ix;
                x1[0] = pv.Xf[xf_index(ix, 2 * j)];
                x1[1] = pv.Xf[xf_index(ix, 2 * j + 1)];
                if (DYN) {
                    double tp = ((ix == 0) ? 1e-5 : (double)ix) + pv.dtp;
                    xd1 = pv.xraw[ix] * exp(-p[t].lam_cfc * tp);
                }
            } else {
                double s = 0.0;
                if (DYN) {
                    s = ad1[t];
                    s += __shfl_xor_sync(full, s, 1);
                    s += __shfl_xor_sync(full, s, 2);
                }
                x1[0] = a1[t][0][0] * r1;
                x1[1] = a1[t][0][1] * r1;
                if (DYN) xd1 = s * r1;
                if (any_div) {                  // warp-uniform: the IEEE divisions are not even issued in the common case
                    if (use_div) {
                        x1[0] = a1[t][0][0] / S1;
                        x1[1] = a1[t][0][1] / S1;
                        if (DYN) xd1 = s / S1;
                    }
                }
            }
            if constexpr (C2 == CLS_P) {
                int ix = c2[t].ix;
                x2[0] = pv.Xf[xf_index(ix, 2 * j)];
                x2[1] = pv.Xf[xf_index(ix, 2 * j + 1)];
                if (DYN) {
                    double tp = ((ix == 0) ? 1e-5 : (double)ix) + pv.dtp;
                    xd2 = pv.xraw[ix] * exp(-p[t].lam_cfc * tp);
                }
            } else if constexpr (C2 != CLS_NONE) {
                double s = 0.0;
                if (DYN) {
                    s = ad2[t];
                    s += __shfl_xor_sync(full, s, 1);
                    s += __shfl_xor_sync(full, s, 2);
                }
                x2[0] = a2[t][0][0] * r2;
                x2[1] = a2[t][0][1] * r2;
                if (DYN) xd2 = s * r2;
                if (any_div) {
                    if (use_div) {
                        x2[0] = a2[t][0][0] / S2;
                        x2[1] = a2[t][0][1] / S2;
                        if (DYN) xd2 = s / S2;
                    }
                }
            }
            // cout = f1*cout1 + f2*cout2 (run_age_mcmc_utils.py:154); cout2 = 0.0 without a second component
            m[0] = pf1 * x1[0] + pf2 * x2[0];
            m[1] = pf1 * x1[1] + pf2 * x2[1];
            if (DYN) md = pf1 * xd1 + pf2 * xd2;
            double* sc = scratch_warp + (t * 8 + r) * NCOL;
            sc[2 * j] = m[0];
            sc[2 * j + 1] = m[1];
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 2; q++) {
                int tr = j + 4 * q;
                double v = 0.0;
                if (tr < pv.ntracer) {
                    TracerDev td = pv.tr[tr];
                    if (td.dyn) {
                        v = md;
                    } else {
                        double va = td.col_a >= 0 ? sc[td.col_a] : 0.0;
                        v = td.col_b >= 0 ? va + pJ * sc[td.col_b] : va;
                    }
                    if (td.sf6) v *= (1.0 + plam);
                }
                val[t][q] = v;
            }
            __syncwarp();
        }
    }
};

// ---------------------------------------------------------------- likelihood terms (pymc3 3.11.2 formulae)
// Normal:    -0.5 log(2 pi sd^2) - z^2/2,  z = (obs - mu)/sd           (the log is a per-tracer host constant)
// Student-T, lam = sd^-2:  lgamma((nu+1)/2) - lgamma(nu/2) + 0.5 log(lam/(nu pi)) - (nu+1)/2 log1p(lam (x-mu)^2/nu)
//            = [lgamma((nu+1)/2) - lgamma(nu/2) - 0.5 log(nu pi)] - log(sd) - (nu+1)/2 log1p(z^2/nu)
__device__ __forceinline__ double lik_term_normal(double obs, double mu, double isd, double lc) {
    double z = (obs - mu) * isd;
    return lc - 0.5 * z * z;
}
__device__ __forceinline__ double lik_studentt_const(double nu) {
    return lgamma(0.5 * (nu + 1.0)) - lgamma(0.5 * nu) - 0.5 * log(nu * 3.14159265358979323846);
}
__device__ __forceinline__ double lik_term_studentt(double obs, double mu, double isd, double lc, double nu, double cst) {
    double z = (obs - mu) * isd;
    return cst + lc - 0.5 * (nu + 1.0) * log1p(z * z / nu);
}

// ---------------------------------------------------------------- TMA bulk copies (cp.async.bulk, SASS UBLKCP)
// One elected thread arms an mbarrier with the byte count and issues 1-D bulk copies global -> shared; every thread
// then waits on the barrier phase.  Sizes and addresses are multiples of 16 bytes by construction (Lpad % 4 == 0).
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned int bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

// ---------------------------------------------------------------- per-CTA scaffolding shared by the kernels
// Shared-memory carve-up, chunk streaming of the lag tables and the unit schedule.  k_forward evaluates one
// parameter vector per chain; k_mcmc_age (ngrtd_mcmc.cuh) runs whole Metropolis steps around eval().
// Shared-memory footprint (doubles) of the forward tables + per-warp scratch (+ the hand-off slots and flags of the tape
// schedule); the launchers size the dynamic shared memory with the same function FwdCta::setup lays it out with.
template <class WT>
__host__ __device__ inline int fwd_smem_doubles(int nwarps, int lc_cap, bool tape) {
    int p = ExpCfg<WT::TBITS>::DOUBLES + 2 + nwarps * WT::NTILES * 8 * NCOL + lc_cap * NCOL;
    if (WT::ANY_D) p += lc_cap * NCOL + 2 * lc_cap;
    if (WT::DYNAMIC) p += lc_cap + (WT::ANY_D ? lc_cap : 0);
    if (tape) p += nwarps * WT::NTILES * WT::NPART * 32 + ((nwarps + 2) >> 1);
    if (WT::PARK) p += nwarps * WT::NTILES * 8 * PARK_DOUBLES;
    if (WT::TAILMODE != 0 && WT::ANY_D && DM_TAIL) p += CT_DOUBLES;
    return p;
}

template <int C1, int C2, bool DYN, int NT, int UA, int TB, int TM = 1, bool PK = false>
struct FwdCta {
    using WT = WarpTiles<C1, C2, DYN, NT, UA, TB, TM, PK>;
    SmemView s;
    const PlanView& pv;
    int lc_cap, nchunks, nwarps, nthreads, tid, lane, warp;
    int Lloop;       // lags covered by the lag loop: Lpad, or Kc when the analytic tail applies
    int scratch_off, park_off;
    unsigned int phase;
    bool need_J, pending;   // pending: resident tables issued (TMA in flight), not yet waited for
    bool tape;              // tape schedule active (k_forward, resident tables): hand-off slots are laid out

    __device__ __forceinline__ FwdCta(const PlanView& pv_) : pv(pv_) {}

    // returns the offset (in doubles) of the first shared-memory double not used by the forward tables.
    // want_tape: reserve the hand-off slots of the tape schedule (k_forward with resident tables)
    __device__ __forceinline__ int setup(int lc_cap_, bool want_tape = false) {
        lc_cap = lc_cap_;
        nthreads = blockDim.x;
        nwarps = nthreads >> 5;
        tid = threadIdx.x;
        lane = tid & 31;
        // broadcast from lane 0: tells the compiler the warp index is warp-uniform, so everything derived from it (the tape
        // cuts, loop trip counts) stays on the uniform datapath and mma.sync needs no WARPSYNC in front of it
        warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
        int p = 0;
        s.tbl = p; p += ExpCfg<TB>::DOUBLES;   // first: 128-byte aligned, so copy c of every entry sits in banks 2c, 2c+1
        s.bar = p; p += 2;
        s.scratch = p; p += nwarps * NT * 8 * NCOL;
        s.Xf = p; p += lc_cap * NCOL;
        s.Xd = p; if (WT::ANY_D) p += lc_cap * NCOL;
        s.itp = p; if (WT::ANY_D) p += 2 * lc_cap;
        s.xraw = p; if (DYN) p += lc_cap;
        s.xrawd = p; if (DYN && WT::ANY_D) p += lc_cap;
        scratch_off = s.scratch + warp * NT * 8 * NCOL;
        Lloop = (TM != 0 && tail_active(pv, WT::ANY_G, WT::ANY_D)) ? pv.Kc : pv.Lpad;
        nchunks = WT::ANY_LOOP ? (Lloop + lc_cap - 1) / lc_cap : 1;
        tape = want_tape && nchunks == 1;
        s.part = p; if (tape) p += nwarps * NT * WT::NPART * 32;
        s.flag = p; if (tape) p += (nwarps + 2) >> 1;
        s.park = p; if (PK) p += nwarps * NT * 8 * PARK_DOUBLES;
        park_off = s.park + warp * NT * 8 * PARK_DOUBLES;
        s.ctab = p;
        if (TM != 0 && WT::ANY_D && DM_TAIL) { p += CT_DOUBLES; ct_fill(ngrtd_smem + s.ctab, pv, tid); }
        need_J = false;
        for (int t = 0; t < pv.ntracer; t++) need_J |= (pv.tr[t].col_b >= 0);
        phase = 0;
        if (tid == 0) mbar_init(reinterpret_cast<unsigned long long*>(ngrtd_smem + s.bar), 1);
        if (WT::ANY_D) {
            const double* tg = (TB == 11) ? pv.tbl11 : pv.tbl7;
            for (int i = tid; i < ExpCfg<TB>::DOUBLES; i += nthreads) ngrtd_smem[s.tbl + i] = tg[i >> ExpCfg<TB>::REP_BITS];
        }
        if (tape && tid < nwarps) reinterpret_cast<volatile int*>(ngrtd_smem + s.flag)[tid] = 0;
        __syncthreads();
        pending = false;
        if (nchunks == 1 && WT::ANY_LOOP) {           // resident tables: one TMA load per launch, waited for lazily so
            issue_chunk(0, Lloop);                    // that it overlaps the first unit's parameter loads and prologue
            pending = true;
        }
        return p;
    }

    // stage lags [kc, kc+len) of the plan tables into shared memory with TMA bulk copies (one elected thread issues
    // them, all threads wait on the mbarrier phase).  Callers guarantee no thread still reads the previous chunk.
    __device__ __forceinline__ void issue_chunk(int kc, int len) {
        unsigned long long* bar = reinterpret_cast<unsigned long long*>(ngrtd_smem + s.bar);
        if (tid == 0) {
            const unsigned int bx = (unsigned int)len * NCOL * 8u, bl = (unsigned int)len * 8u;
            unsigned int total = bx;
            if (WT::ANY_D) total += bx + 2u * bl;
            if (DYN) total += bl + (WT::ANY_D ? bl : 0u);
            mbar_expect_tx(bar, total);
            bulk_g2s(ngrtd_smem + s.Xf, pv.Xf + (size_t)kc * NCOL, bx, bar);
            if (WT::ANY_D) {
                bulk_g2s(ngrtd_smem + s.Xd, pv.Xd + (size_t)kc * NCOL, bx, bar);
                bulk_g2s(ngrtd_smem + s.itp, pv.itp + 2 * (size_t)kc, 2u * bl, bar);
            }
            if (DYN) {
                bulk_g2s(ngrtd_smem + s.xraw, pv.xraw + kc, bl, bar);
                if (WT::ANY_D) bulk_g2s(ngrtd_smem + s.xrawd, pv.xrawd + kc, bl, bar);
            }
        }
    }
    __device__ __forceinline__ void wait_chunk() {
        mbar_wait(reinterpret_cast<unsigned long long*>(ngrtd_smem + s.bar), phase);
        phase ^= 1u;
    }
    __device__ __forceinline__ void load_chunk(int kc, int len) {
        if (!WT::ANY_LOOP) return;
        issue_chunk(kc, len);
        wait_chunk();
    }

    // forward model of NT tiles: lane (r, j) receives tracers j, j+4 of chain r in val[t][0..1].
    // In lock-step mode every warp of the CTA must call this the same number of times (inactive warps only
    // take part in the chunk loads and barriers).
    __device__ __forceinline__ void eval(const ChainPar (&par)[NT], bool active, bool lockstep, double (&val)[NT][2]) {
        WT w;
        w.begin(par, pv, lane, need_J, ngrtd_smem + park_off);
        if (!lockstep) {
            if (pending) { wait_chunk(); pending = false; }
            w.chunk(s, pv, 0, Lloop / 4, lane);
        } else {
            for (int c = 0; c < nchunks; c++) {
                int kc = c * lc_cap;
                int len = min(lc_cap, Lloop - kc);
                __syncthreads();
                load_chunk(kc, len);
                __syncthreads();
                if (active) w.chunk(s, pv, kc, len / 4, lane);
            }
        }
        if (active) w.end(par, pv, ngrtd_smem + scratch_off, lane, val, ngrtd_smem + park_off, s.ctab);
    }
    // lock-step rounds only (k_forward with streamed tables; its resident-table path is the tape schedule)
    __device__ __forceinline__ void eval_lockstep(const ChainPar (&par)[NT], bool active, double (&val)[NT][2]) {
        WT w;
        w.begin(par, pv, lane, need_J, ngrtd_smem + park_off);
        for (int c = 0; c < nchunks; c++) {
            int kc = c * lc_cap;
            int len = min(lc_cap, Lloop - kc);
            __syncthreads();
            load_chunk(kc, len);
            __syncthreads();
            if (active) w.chunk(s, pv, kc, len / 4, lane);
        }
        if (active) w.end(par, pv, ngrtd_smem + scratch_off, lane, val, ngrtd_smem + park_off, s.ctab);
    }

    // Unit schedule.  Resident tables: static and balanced -- units are dealt round-robin to the 4*gridDim.x SM
    // sub-partitions (warp & 3 selects the sub-partition) and, within one, round-robin to its warps.  All units
    // cost the same, so every sub-partition carries floor or ceil of nunits/(4*grid) units and its FP64 pipe stays
    // shared by >= 3 warps until the end (a dynamic counter let the last partial round pile onto random
    // sub-partitions: profiles/r1_notes.md).  Long lag axis: CTA-wide rounds in lock step (inactive warps still
    // iterate so that they take part in the chunk barriers).
    // Used as:  for (cta.sched_begin(n); cta.sched_valid(); cta.sched_next()) { u = cta.unit; active = cta.active; ... }
    // (a plain loop, not a callback: a lambda capturing the kernel-parameter structs by reference was not inlined in
    // the full build and forced a 1.5 KB local-memory copy of the parameters per thread).
    long long unit, sched_i, sched_n;
    bool active, lockstep;
    __device__ __forceinline__ void sched_begin(long long nunits) {
        sched_n = nunits;
        lockstep = nchunks != 1;
        sched_i = lockstep ? (long long)blockIdx.x : (long long)(warp / (nwarps >= 4 ? 4 : nwarps));
        sched_update();
    }
    __device__ __forceinline__ long long unit_of(long long si) const {
        if (!lockstep) {
            const int spc = nwarps >= 4 ? 4 : nwarps;
            return ((long long)blockIdx.x * spc + (warp % spc)) + si * ((long long)gridDim.x * spc);
        }
        return si * nwarps + warp;
    }
    __device__ __forceinline__ long long sched_step() const {
        return lockstep ? (long long)gridDim.x : (long long)(nwarps / (nwarps >= 4 ? 4 : nwarps));
    }
    __device__ __forceinline__ void sched_update() {
        unit = unit_of(sched_i);
        active = lockstep ? unit < sched_n : true;
    }
    // the unit this warp will work on after the current one (>= sched_n: none) -- used to prefetch its parameters
    __device__ __forceinline__ long long sched_peek() const { return unit_of(sched_i + sched_step()); }
    __device__ __forceinline__ bool sched_valid() const {
        return lockstep ? sched_i < (sched_n + nwarps - 1) / nwarps : unit < sched_n;
    }
    __device__ __forceinline__ void sched_next() {
        sched_i += sched_step();
        sched_update();
    }
};

// sum of the likelihood terms of the tracers a lane owns (j, j+4), reduced over the 4 lanes of a chain.
// ob/is/lc: observation, 1/sd and per-tracer constant of the lane's two tracers (registers, no indexed structs).
// The nu-dependent Student-T constant is evaluated once per chain: lane 0 takes lgamma((nu+1)/2), lane 1 lgamma(nu/2),
// lane 2 the log, and the pieces are combined with the same shuffles that reduce the tracer terms.
__device__ __forceinline__ double lik_reduce(int kind, int ntracer, int j, const double (&v)[2], const double (&ob)[2],
                                             const double (&is)[2], const double (&lc)[2], double nu) {
    double acc = 0.0;
    if (kind == 1) {
        if (j == 0) acc = (double)ntracer * lgamma(0.5 * (nu + 1.0));
        else if (j == 1) acc = -(double)ntracer * lgamma(0.5 * nu);
        else if (j == 2) acc = -0.5 * (double)ntracer * log(nu * 3.14159265358979323846);
    }
#pragma unroll
    for (int q = 0; q < 2; q++) {
        if (j + 4 * q < ntracer)
            acc += (kind == 1) ? lik_term_studentt(ob[q], v[q], is[q], lc[q], nu, 0.0) : lik_term_normal(ob[q], v[q], is[q], lc[q]);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    return acc;
}

// ---------------------------------------------------------------- the forward (+ likelihood) kernel
// Parameter staging (stage != 0).  The parameter rows of a unit (NT*8 chains x ndim doubles, contiguous in theta) are
// brought into a per-warp shared-memory slot by ONE TMA bulk copy issued by lane 0 against the warp's own mbarrier, and
// the copy for the warp's NEXT unit is issued as soon as the current rows are in registers, so it has a whole unit
// (~50 us) to land.  This makes theta traffic one large request per unit instead of 7-11 scattered 8-byte loads per lane,
// which is what lets the *_host entry points hand the kernel a pointer to PINNED HOST memory (zero-copy over PCIe: the
// first round of units streams in while the SMs start, the second round is prefetched under the first) instead of a
// staged cudaMemcpyAsync.  Units whose byte count or source address is not a multiple of 16 (ragged last unit, odd
// row offsets) are staged with plain lane loads.
//
// Schedule with resident tables (r2): the TAPE.  A CTA takes a contiguous block of units; their lag loops are laid end to
// end on a tape of (units x lag groups) and the tape is cut into one equal segment per warp.  A unit whose lag range is cut
// is finished by the warp that holds its first lags (the owner); the warps holding the other ranges publish their DMMA
// accumulators through shared memory (WarpTiles::store_partial) -- they do so at the START of their segment, the owner
// collects at the END of its own, so the flags are practically never waited for and no cycle of waits exists.
// What it buys over the r1 schedule (units dealt round-robin to warps, every warp whole units):
//   * every warp carries the same number of lag groups, so the four warps of a sub-partition share the FP64/DMMA pipe
//     until the last group (r1: 1.73 units per warp at 65,536 chains -> the second round ran 3 of 4 warps);
//   * the cuts fall at different lags in different warps, so their prologues (parameter loads, divisions, exp) and
//     epilogues (normalisation, mixing, likelihood, stores) no longer coincide: while one warp is outside its lag loop the
//     others keep the pipe busy (r1: all four in lock step, ~7 us of idle pipe per round, profiles/r1_notes.md);
//   * a batch smaller than the grid still uses every warp of the SMs it touches.
// Results are deterministic for a given (B, grid); they differ from a different cut of the same chain by rounding only
// (different association of the same sums), which the parity tests bound at 1e-10 against the reference.
// Tables streamed in chunks (long lag axes): CTA-wide rounds in lock step as in r1.
//
// Programmatic dependent launch: the kernel triggers its dependents at once and waits for its own predecessor
// (griddepcontrol.wait) only after the table loads, so with the launch attribute set by launch_forward_t the set-up of
// launch i+1 runs under the tail of launch i.  Without the attribute both instructions are no-ops.
constexpr int TAPE_MIN_GROUPS = 8;     // a cut closer than this to a unit boundary is moved onto the boundary

template <int C1, int C2, bool DYN, int NT, int UA, int MAXW, bool TAIL>
__global__ void __launch_bounds__(MAXW * 32, 1)
k_forward(PlanView pv, SlotMap sm, const double* __restrict__ theta, long long B, double* __restrict__ out,
          double* __restrict__ logp, LikPar lik, int lc_cap, int stage, int tape_min) {
    using CTA = FwdCta<C1, C2, DYN, NT, UA, FWD_TB, TAIL ? 1 : 0, true>;
    using WT = typename CTA::WT;
    asm volatile("griddepcontrol.launch_dependents;");
    CTA cta(pv);
    int p_end = cta.setup(lc_cap, true);
    p_end = (p_end + 1) & ~1;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int j = cta.lane & 3, r = cta.lane >> 2;
    const long long nunits = (B + NT * 8 - 1) / (NT * 8);
    const int th_unit = NT * 8 * sm.ndim;       // doubles per staged unit
    double* const stg = ngrtd_smem + p_end + cta.warp * th_unit;
    unsigned long long* const tbar = reinterpret_cast<unsigned long long*>(ngrtd_smem + p_end + cta.nwarps * th_unit) + cta.warp;
    unsigned int tphase = 0;
    bool staged_tma = false;
    // stage the parameter rows of unit un into this warp's slot (all lanes call; previous readers are past a __syncwarp)
#define NGRTD_STAGE_UNIT(un)                                                                            \
    do {                                                                                                \
        const long long c0_ = (un) * (NT * 8);                                                          \
        const long long n_ = min((long long)(NT * 8), B - c0_) * sm.ndim;                               \
        const double* src_ = theta + c0_ * sm.ndim;                                                     \
        staged_tma = stage == 1 && ((n_ & 1) == 0) && ((reinterpret_cast<unsigned long long>(src_) & 15ull) == 0); \
        if (staged_tma) {                                                                               \
            if (cta.lane == 0) {                                                                        \
                mbar_expect_tx(tbar, (unsigned int)n_ * 8u);                                            \
                bulk_g2s(stg, src_, (unsigned int)n_ * 8u, tbar);                                       \
            }                                                                                           \
        } else {                                                                                        \
            for (int i_ = cta.lane; i_ < (int)n_; i_ += 32) stg[i_] = src_[i_];                         \
        }                                                                                               \
    } while (0)
    // outputs + likelihood of one finished unit (lane (r, j): tracers j, j+4 of chain r of every tile)
#define NGRTD_EMIT_UNIT()                                                                               \
    do {                                                                                                \
        /* the lane's two tracers (j, j+4): read from the kernel parameters here, not held across the lag loop */ \
        double ob[2], is[2], lc[2];                                                                     \
        _Pragma("unroll") for (int q = 0; q < 2; q++) {                                                 \
            int tr = min(j + 4 * q, MAX_TRACER - 1);                                                    \
            ob[q] = lik.obs[tr]; is[q] = lik.isd[tr]; lc[q] = lik.lc[tr];                               \
        }                                                                                               \
        _Pragma("unroll") for (int t = 0; t < NT; t++) {                                                \
            bool ok = chain[t] < B;                                                                     \
            if (out != nullptr && ok) {                                                                 \
                if (j < pv.ntracer) out[chain[t] * pv.ntracer + j] = val[t][0];                         \
                if (j + 4 < pv.ntracer) out[chain[t] * pv.ntracer + j + 4] = val[t][1];                 \
            }                                                                                           \
            if (logp != nullptr) {                                                                      \
                double nu = (lik.kind == 1) ? lik.nu[ok ? chain[t] : B - 1] : 0.0;                      \
                double acc = lik_reduce(lik.kind, pv.ntracer, j, val[t], ob, is, lc, nu);               \
                if (j == 0 && ok) logp[chain[t]] = acc;                                                 \
            }                                                                                           \
        }                                                                                               \
    } while (0)

    if (cta.tape) {
        // ---- tape schedule: this CTA's units [ub0, ub1), ng lag groups each, cut into nwarps segments
        // (tape positions are 32-bit: the launcher keeps units-per-CTA x lag groups below 2^31)
        const int ng = WT::ANY_LOOP ? cta.Lloop / 4 : 1;
        const long long ub0 = nunits * blockIdx.x / gridDim.x, ub1 = nunits * (blockIdx.x + 1) / gridDim.x;
        const int G = (int)(ub1 - ub0) * ng;
        const int nw = cta.nwarps;
        auto cut = [&](int w) -> int {
            int c = (int)((long long)G * w / nw);
            const int rem = c % ng;
            if (rem < tape_min) c -= rem;
            else if (ng - rem < tape_min) c += ng - rem;
            return c;
        };
        int pos = cut(cta.warp);
        const int s1 = cut(cta.warp + 1);
        volatile int* const flags = reinterpret_cast<volatile int*>(ngrtd_smem + cta.s.flag);
        const int part_sz = NT * WT::NPART * 32;
        if (stage) {
            if (cta.lane == 0) mbar_init(tbar, 1);
            __syncwarp();
            if (pos < s1) NGRTD_STAGE_UNIT(ub0 + pos / ng);
        }
        while (pos < s1) {
            const int ul = pos / ng;
            const int g0 = pos - ul * ng;
            const int g1 = min(ng, g0 + (s1 - pos));
            const long long u = ub0 + ul;
            ChainPar par[NT];
            if (stage) {
                if (staged_tma) { mbar_wait(tbar, tphase); tphase ^= 1u; }
                __syncwarp();
            }
#pragma unroll
            for (int t = 0; t < NT; t++) {
                const long long ch = (u * NT + t) * 8 + r;
                long long cl = ch < B ? ch : B - 1;
                if (stage) par[t] = load_chain_par(stg, sm, cl - u * (NT * 8), pv, cta.need_J);
                else par[t] = load_chain_par(theta, sm, cl, pv, cta.need_J);
            }
            pos += g1 - g0;
            if (stage) {
                __syncwarp();                                   // every lane holds its rows: the slot is free again
                if (pos < s1) NGRTD_STAGE_UNIT(ub0 + pos / ng);
            }
            WT w;
            w.begin(par, pv, cta.lane, cta.need_J, ngrtd_smem + cta.park_off);
            if (cta.pending) { cta.wait_chunk(); cta.pending = false; }
            w.chunk(cta.s, pv, 4 * g0, g1 - g0, cta.lane, 4 * g0);
            if (g0 != 0) {
                // a later range of a unit owned by a lower warp: publish the accumulators
                w.store_partial(ngrtd_smem + cta.s.part + cta.warp * part_sz, cta.lane);
                __threadfence_block();
                __syncwarp();
                if (cta.lane == 0) flags[cta.warp] = 1;
            } else {
                if (g1 < ng) {
                    // owner of a cut unit: the remaining ranges belong to the next warps (each publishes at most one partial)
                    const int uend = (ul + 1) * ng;
                    int c2 = s1;                                // = cut(warp + 1): start of the next warp's segment
                    for (int w2 = cta.warp + 1; w2 < nw && c2 < uend; w2++) {
                        const int c3 = cut(w2 + 1);
                        if (c3 > c2) {                          // non-empty segment: its first range continues this unit
                            while (flags[w2] == 0) {}
                            __threadfence_block();
                            w.add_partial(ngrtd_smem + cta.s.part + w2 * part_sz, cta.lane);
                        }
                        c2 = c3;
                    }
                }
                double val[NT][2];
                w.end(par, pv, ngrtd_smem + cta.scratch_off, cta.lane, val, ngrtd_smem + cta.park_off, cta.s.ctab);
                long long chain[NT];
#pragma unroll
                for (int t = 0; t < NT; t++) chain[t] = ((ub0 + ul) * NT + t) * 8 + r;
                NGRTD_EMIT_UNIT();
            }
        }
    } else {
        // ---- tables streamed in chunks: CTA-wide rounds in lock step
        cta.sched_begin(nunits);
        if (stage) {
            if (cta.lane == 0) mbar_init(tbar, 1);
            __syncwarp();
            if (cta.sched_valid() && cta.active) NGRTD_STAGE_UNIT(cta.unit);
        }
        for (; cta.sched_valid(); cta.sched_next()) {
            const long long u = cta.unit;
            const bool active = cta.active;
            ChainPar par[NT];
            long long chain[NT];
            if (stage && active) {
                if (staged_tma) { mbar_wait(tbar, tphase); tphase ^= 1u; }
                __syncwarp();
            }
#pragma unroll
            for (int t = 0; t < NT; t++) {
                chain[t] = (u * NT + t) * 8 + r;
                long long cl = chain[t] < B ? chain[t] : B - 1;
                if (stage && active) par[t] = load_chain_par(stg, sm, cl - u * (NT * 8), pv, cta.need_J);
                else par[t] = load_chain_par(theta, sm, active ? cl : 0, pv, cta.need_J);
            }
            if (stage && active) {
                __syncwarp();                                   // every lane holds its rows: the slot is free again
                const long long un = cta.sched_peek();
                if (un < nunits) NGRTD_STAGE_UNIT(un);
            }
            double val[NT][2];
            cta.eval_lockstep(par, active, val);
            if (!active) continue;
            NGRTD_EMIT_UNIT();
        }
    }
#undef NGRTD_STAGE_UNIT
#undef NGRTD_EMIT_UNIT
    if (cta.pending) cta.wait_chunk();      // warps without work must not exit under an in-flight bulk copy
}

}  // namespace ngrtd
