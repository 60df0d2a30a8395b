// ngrtd_common.cuh -- shared constants, device-side plan view and small device helpers.
// sm_100a only.  See DESIGN.md for the data layout and the roofline of each kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace ngrtd {

constexpr int NCOL = 8;                 // folded input columns per lag: col 0 = ones (normalisation), 1..7 tracers
constexpr int NSLOT = 11;               // ForwardMod.p_dict slots
constexpr int MAX_TRACER = 8;
constexpr int LC_MAX = 1024;            // lags resident in shared memory per chunk

// Layout of the folded tables Xf / Xd (global and shared memory alike), addressed through xf_index(lag, column).
// Default: plain [lag][8 columns] rows.  The B fragment of DMMA.8x8x4 makes lane (r = lane >> 2, j = lane & 3) read
// X[4g + j][r]; a 64-bit shared load is served per half-warp (lanes 0-15: r = 0..3), which with plain rows covers
// 4 x 32 bytes at a 64-byte stride -- lags j and j + 2 fall on the same banks and every B-fragment load costs 4 wavefronts
// instead of 2 (ncu source page).  -DNGRTD_XF_SWIZZLE stores each group of 4 lags as [column half][lag][column & 3] so
// that a half-warp reads one contiguous 128-byte line (2 wavefronts per load).  Measured on the B200 (profiles/r1_notes.md,
// "session 3"): bit-identical results, all GPU tests green, forward launch 0.1153 -> 0.1170 ms, sampler -2 % -- the lag
// loop is held by the FP64/DMMA pipe, not by the LSU data pipe, so the plain layout stays the default.
__host__ __device__ __forceinline__ int xf_index(int k, int c) {
#ifdef NGRTD_XF_SWIZZLE
    return (k >> 2) * (4 * NCOL) + (c >> 2) * 16 + (k & 3) * 4 + (c & 3);
#else
    return k * NCOL + c;
#endif
}

// table-driven exp(): exp(e) = 2^(n/N) * q(g), n = floor(e*N/ln2), g = 1 + frac(e*N/ln2)  (exp_scaled_bits below).
// The N-entry table of 2^(j/N) lives in shared memory as 8-byte entries, replicated TBL_REP times: lane l gathers from
// copy l & (TBL_REP-1) at double index j*TBL_REP + copy.  A 64-bit shared load is served per half-warp; with 16 copies
// every lane of a half-warp owns its own pair of banks and the gather is conflict-free (2 wavefronts per warp).
// Measured (profiles/r1_notes.md): one copy as two 32-bit word arrays, N = 256: ~6.8 wavefronts per gather pair, LSU data
// pipe 64 % busy; 16 copies: 46 % -- but the lag loop is held by the shared FP64/DMMA pipe (math-pipe-throttle stalls
// at 67 % pipe activity), so throughput did not move, and 16 KB more shared memory pushed the sampler kernel into
// chunked streaming.  Default: 4 copies (4 KB), N = 128.
#ifndef NGRTD_TBL_BITS
#define NGRTD_TBL_BITS 7
#endif
#ifndef NGRTD_TBL_REP_BITS
#define NGRTD_TBL_REP_BITS 2
#endif
constexpr int TBL_BITS = NGRTD_TBL_BITS;
constexpr int TBL_N = 1 << TBL_BITS;
constexpr int TBL_REP_BITS = NGRTD_TBL_REP_BITS;
constexpr int TBL_REP = 1 << TBL_REP_BITS;
constexpr int TBL_DOUBLES = TBL_N * TBL_REP;      // shared-memory footprint in doubles
constexpr double LN2 = 0.693147180559945309417232121458;
constexpr double EXP_K = TBL_N / LN2;
// exp_scaled_bits: q(g) ~= exp((g - 1) ln2 / N) on g in [1, 2)  (tools/exp_poly_g.py: constants and error bounds)
#if NGRTD_TBL_BITS == 7
constexpr double EXQ_C0 = 0.99459942332194162, EXQ_C1 = 0.0053859675366433966, EXQ_C2 = 1.4582602945559778e-05,
                 EXQ_C3 = 2.6538188920206277e-08;      // max rel err 2.8e-13
#elif NGRTD_TBL_BITS == 8
constexpr double EXQ_C0 = 0.99729605607536986, EXQ_C1 = 0.0027002849873872833, EXQ_C2 = 3.6556244405132574e-06,
                 EXQ_C3 = 3.3127848075725495e-09;      // 1.8e-14
#elif NGRTD_TBL_BITS == 9
constexpr double EXQ_C0 = 0.99864711289033903, EXQ_C1 = 0.0013519715460713645, EXQ_C2 = 9.1514977059835842e-07,
                 EXQ_C3 = 4.1381786370901806e-10;      // 1.2e-15
#else
constexpr double EXQ_C0 = 0, EXQ_C1 = 0, EXQ_C2 = 0, EXQ_C3 = 0;
#endif
constexpr double FX_MAGIC = 1572864.0;               // 1.5 * 2^20: ulp 2^-32, integer part biased by 2^19
constexpr int EXP_NMIN = -1022 * TBL_N;             // below 2^-1022: clamp (see DESIGN.md "underflow")

enum Cls : int { CLS_NONE = 0, CLS_P = 1, CLS_G = 2, CLS_D = 3 };

struct TracerDev {
    int col_a;   // folded series column (or -1)
    int col_b;   // folded lag-index column for '4He' (or -1)
    int dyn;     // 1: per-chain lambda (thalf_cfc) -> uses the dyn accumulators
    int sf6;     // 1: *= 1 + lamsf6
};

struct ColTail {
    int type;        // 0 ones, 1 decay, 2 ingrowth, 3 lag-index * decay, -1 unused column
    double bg;       // constant series value beyond Kc
    double lam;      // decay constant folded into the column
    double i0, s;    // type 3: lag_index[k] = i0 + s*k for k >= Kc
};

// Device view of a plan (passed by value to kernels).
struct PlanView {
    int L;          // true number of lags
    int Lpad;       // padded to a multiple of 4 (pad rows are zero)
    double dtp;     // integer-valued shift of the lag grid
    const double* Xf;     // [Lpad, 8] folded columns
    const double* Xd;     // [Lpad, 8] Xf * tp^-1.5 (dispersion component)
    const double* itp;    // [Lpad][2] {1/tp, tp} (pad: 0, 0)
    const double* xraw;   // [Lpad] raw series of the per-chain-lambda tracer
    const double* xrawd;  // [Lpad] xraw * tp^-1.5
    const double* tbl;    // [TBL_N] 2^(j/TBL_N) with j << (20 - TBL_BITS) subtracted from the high word (see exp_scaled_bits)
    int ntracer;
    TracerDev tr[MAX_TRACER];
    int eta1_is_one, eta2_is_one;   // 'exponential' == exp_pist_flow with eta = 1 (bit-identical in the reference)
    double default_log10J;          // run_age_mcmc_utils.py:90-91
    // Constant-tail closed form (exponential-class components only).  The reference back-extends every input series by
    // ~25,000 constant rows (age_modeling_mcmc.prep.py:165-223): beyond lag Kc every folded column is bg*exp(-lam tp),
    // bg*(1-exp(-lam tp)) or (i0 + s k)*exp(-lam tp), whose products with geometric weights sum in closed form.  The lag
    // loop then covers only [0, Kc) and the tail [Kc, L) is added analytically in the epilogue (exact to rounding).
    int Kc;                         // first lag of the analytic tail (multiple of 4); Kc >= Lpad: no tail
    ColTail ct[NCOL];
    double dyn_bg;                  // constant value of the per-chain-lambda series beyond Kc
    double tpl, itpl;               // last lag of the grid, tp[L-1], and its reciprocal (dispersion dead-chain rule)
};

// Which plans cut the lag loop at Kc and add the tail [Kc, L) analytically in the epilogue: exponential-class components always
// (closed form); dispersion components only in -DNGRTD_DM_TAIL builds (Gauss-Legendre quadrature of the smooth tail,
// WarpTiles::dm_tail in ngrtd_forward.cuh; experimental in r1, see profiles/r1_notes.md).  Plan creation leaves Kc = Lpad when
// the plan does not qualify.
#ifdef NGRTD_DM_TAIL
constexpr bool DM_TAIL = true;
#else
constexpr bool DM_TAIL = false;
#endif
__host__ __device__ __forceinline__ bool tail_active(const PlanView& pv, bool any_g, bool any_d) {
    return (any_g || any_d) && (!any_d || DM_TAIL) && pv.Kc < pv.L;
}

struct SlotMap {
    int ndim;
    signed char col_of_slot[NSLOT];   // -1: not in par_names -> p_dict default
};

struct ChainPar {
    double tau1, tau2, f1, f2, eta1, eta2, D1, D2, log10J, lam_cfc, lamsf6;   // J = 10**log10J is formed in WarpTiles::begin
};

__device__ __forceinline__ ChainPar load_chain_par(const double* __restrict__ theta, const SlotMap& sm,
                                                   long long chain, const PlanView& pv, bool need_J) {
    const double* row = theta + chain * sm.ndim;
    auto get = [&](int slot, double dflt) -> double {
        int c = sm.col_of_slot[slot];
        return c >= 0 ? row[c] : dflt;
    };
    ChainPar p;
    p.tau1 = get(0, 0.0);
    p.tau2 = get(1, 0.0);     // p_dict default 0.0
    p.f1 = get(2, 1.0);       // default 1.0
    p.f2 = get(3, 0.0);       // default 0.0
    p.eta1 = pv.eta1_is_one ? 1.0 : get(4, 0.0);
    p.eta2 = pv.eta2_is_one ? 1.0 : get(5, 0.0);
    p.D1 = get(6, 0.0);
    p.D2 = get(7, 0.0);
    p.log10J = need_J ? get(8, pv.default_log10J) : 0.0;         // J = 10**p_dict['J'] (:101), see WarpTiles::begin
    {
        int c = sm.col_of_slot[9];
        p.lam_cfc = c >= 0 ? (LN2 / row[c]) : 0.0;               // thalf_2_lambda (:146-152)
    }
    p.lamsf6 = get(10, 0.0);   // False -> x(1+0)
    return p;
}

// exp(e) for e <= ~0, given t = ep + FX_MAGIC with ep = e*N/ln2.  The addition is folded into the caller's FMA chain
// (cp + FX_MAGIC is a per-chain constant), so it costs nothing.  With 2^20 <= t < 2^21 the mantissa of t IS ep in fixed
// point: high word bits 0..19 = floor(ep) + 2^19, low word = the 32-bit fraction F of ep.
//   * table slot / exponent insertion come from the high word on the integer pipes (the 2^19 bias and the exponent
//     field of t fold into one immediate; the table stores 2^(j/N) with j << (20 - log2 N) subtracted from its high word);
//   * g = 1 + F 2^-32 is assembled from the bits of F (two shifts, one OR) and p = q(g) = exp((g-1) ln2/N) is a cubic
//     in g: no F2I / I2F conversions and no DADD for the reduced argument.
// Shared FP64/DMMA-pipe cost: 3 DFMA + 1 DMUL.  History (profiles/r1_notes.md): the first version rounded with F2I/I2F
// (XU pipe) and a centred polynomial; a fully integer polynomial with IMAD.HI / IMAD.WIDE measured SLOWER (wide integer
// multiplies issue at ~4.5 cycles and contend with the FP64 pipe -- tools/microbench/imad_peak.cu).
// Error: ep is rounded to 2^-32 table units twice (<= 2^-32 ln2/N relative, unbiased) + the polynomial.
// Below EXP_NMIN the high word is clamped (result ~2^-1022) instead of flushing to zero: one VIMNMX instead of a compare
// and two selects.  NaN does NOT propagate through the integer path: chains whose largest weight would be below 2^-1022,
// or with NaN parameters, are declared dead in Comp<CLS_D>::init and poisoned in the epilogue, which reproduces the
// reference's 0/0 = NaN when every weight underflows.
__device__ __forceinline__ double exp_scaled_bits(double t, const double* __restrict__ tbl) {
    constexpr int S = 20 - TBL_BITS;
    constexpr int HI_MAGIC = 0x41380000;                                  // high word of FX_MAGIC
    constexpr int HI_MIN = HI_MAGIC + EXP_NMIN;                           // high word of EXP_NMIN + FX_MAGIC
    constexpr unsigned int FOLD = (unsigned int)(((unsigned long long)HI_MAGIC << S) & 0xffffffffull);
    const int ht = max(__double2hiint(t), HI_MIN);
    const unsigned int F = (unsigned int)__double2loint(t);
    const double g = __hiloint2double((int)(0x3FF00000u | (F >> 12)), (int)(F << 20));
#ifdef NGRTD_ESTRIN
    const double p = fma(g * g, fma(g, EXQ_C3, EXQ_C2), fma(g, EXQ_C1, EXQ_C0));      // depth 2 instead of 3, one more FP64 op
#else
    const double p = fma(g, fma(g, fma(g, EXQ_C3, EXQ_C2), EXQ_C1), EXQ_C0);
#endif
    // tbl points at this lane's copy: entry j is at tbl[j * TBL_REP]
    const int off = (ht << (3 + TBL_REP_BITS)) & ((TBL_N - 1) << (3 + TBL_REP_BITS));
    const double T = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(tbl) + off);
    const unsigned int hi = (unsigned int)__double2hiint(T) + ((unsigned int)ht << S) - FOLD;
    return __hiloint2double((int)hi, __double2loint(T)) * p;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

}  // namespace ngrtd
