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

// table-driven exp(): exp(e) = 2^(n/N) * p(r),  n = rint(e*N/ln2),  r = e*N/ln2 - n in [-1/2, 1/2].
// The N-entry table of 2^(j/N) is stored as two 32-bit arrays (high / low words).  With N = 32 any 32-lane gather from
// a 128-byte array is bank-conflict free (one 4-byte slot per bank, equal slots broadcast): exactly two shared-memory
// wavefronts per lookup.  (A 2048-entry double table measured 8 wavefronts per lookup and made the kernel LSU-bound.)
// p(r) interpolates exp at Chebyshev nodes (near-minimax); max relative error: N=32/deg 4: 7.8e-14, N=64/deg 3: 4.5e-12,
// N=128/deg 3: 2.8e-13.  Measured on cfg 3 (profiles/r1_notes.md): N=32/deg 4 72.5 cycles per tile-group, N=64/deg 3
// 72.1, N=128/deg 3 70.5 (the extra bank conflicts of the 512-byte arrays cost less than the fourth DFMA): default N=128.
#ifndef NGRTD_TBL_BITS
#define NGRTD_TBL_BITS 8
#endif
constexpr int TBL_BITS = NGRTD_TBL_BITS;
constexpr int TBL_N = 1 << TBL_BITS;
constexpr int TBL_DOUBLES = TBL_N;      // shared-memory footprint in doubles (hi[N] + lo[N] as uint32)
constexpr double LN2 = 0.693147180559945309417232121458;
constexpr double EXP_K = TBL_N / LN2;
#if NGRTD_TBL_BITS == 5
constexpr int EXP_DEG = 4;
constexpr double EXP_C0 = 1.0, EXP_C1 = 0.02166084939172217, EXP_C2 = 0.00023459619819944503,
                 EXP_C3 = 1.693863390316062e-06, EXP_C4 = 9.172607532633896e-09;
#elif NGRTD_TBL_BITS == 6
constexpr int EXP_DEG = 3;
constexpr double EXP_C0 = 0.9999999999955212, EXP_C1 = 0.010830424696239445, EXP_C2 = 5.864919287197594e-05,
                 EXP_C3 = 2.1173168199978455e-07, EXP_C4 = 0.0;
#elif NGRTD_TBL_BITS == 7
constexpr int EXP_DEG = 3;
constexpr double EXP_C0 = 0.9999999999997201, EXP_C1 = 0.005415212348124269, EXP_C2 = 1.4662271345222707e-05,
                 EXP_C3 = 2.64664311467397e-08, EXP_C4 = 0.0;
#elif NGRTD_TBL_BITS == 8
constexpr int EXP_DEG = 3;
constexpr double EXP_C0 = 0.9999999999999825, EXP_C1 = 0.0027076061740622769, EXP_C2 = 3.6655661567589337e-06,
                 EXP_C3 = 3.3083029837113949e-09, EXP_C4 = 0.0;
#elif NGRTD_TBL_BITS == 9
constexpr int EXP_DEG = 3;
constexpr double EXP_C0 = 0.99999999999999891, EXP_C1 = 0.0013538030870311429, EXP_C2 = 9.1639143421807685e-07,
                 EXP_C3 = 4.1353784454173436e-10, EXP_C4 = 0.0;
#else
#error "NGRTD_TBL_BITS must be 5..9"
#endif
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
    const double* tbl;    // [TBL_N] doubles = hi[TBL_N], lo[TBL_N] words of 2^(j/TBL_N)
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
};

struct SlotMap {
    int ndim;
    signed char col_of_slot[NSLOT];   // -1: not in par_names -> p_dict default
};

struct ChainPar {
    double tau1, tau2, f1, f2, eta1, eta2, D1, D2, Jlin, lam_cfc, lamsf6;
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
    p.Jlin = 0.0;
    if (need_J) p.Jlin = exp10(get(8, pv.default_log10J));       // J = 10**p_dict['J'] (:101)
    {
        int c = sm.col_of_slot[9];
        p.lam_cfc = c >= 0 ? (LN2 / row[c]) : 0.0;               // thalf_2_lambda (:146-152)
    }
    p.lamsf6 = get(10, 0.0);   // False -> x(1+0)
    return p;
}

// exp(e) for e <= ~0 given ep = e * N/ln2.  FP64-pipe cost: 1 DADD + deg DFMA + 1 DMUL (+ 2 conversions).
//   * rounding: F2I.F64 / I2F.F64 on the conversion unit.  The magic-number alternative (ep + 1.5*2^52, two more
//     DADDs, no conversions) measured 8 % slower in the full kernel (72.4 vs 78.2 cycles per tile-group,
//     profiles/r1_notes.md) although the conversions also occupy the FP64 pipe for ~3.5 cycles each.
//     F2I saturates for hugely negative exponents and maps NaN to 0 (r = NaN then poisons the result).
//   * table: hi'[N] (uint32) followed by lo[N] (uint32), hi'[j] = hi(2^(j/N)) - (j << (20 - log2 N)), so the exponent
//     insertion is ONE integer multiply-add: hi'[j] + n*2^(20 - log2 N) = hi[j] + ((n >> log2 N) << 20).
//   * n is clamped at EXP_NMIN (result ~2^-1022) instead of flushing to zero: one IMNMX instead of a compare and two
//     selects.  Chains whose largest weight would be below 2^-1022 are declared dead (NaN) in Comp<CLS_D>::init, which
//     reproduces the reference's 0/0 = NaN when every weight underflows.
__device__ __forceinline__ double exp_scaled(double ep, const double* __restrict__ tbl) {
    const unsigned int* th = reinterpret_cast<const unsigned int*>(tbl);
    int n = __double2int_rn(ep);
    double r = ep - __int2double_rn(n);
    double p = (EXP_DEG == 4) ? fma(r, fma(r, fma(r, fma(r, EXP_C4, EXP_C3), EXP_C2), EXP_C1), EXP_C0)
                              : fma(r, fma(r, fma(r, EXP_C3, EXP_C2), EXP_C1), EXP_C0);
    int nc = max(n, EXP_NMIN);
    int off = (nc << 2) & ((TBL_N - 1) << 2);    // byte offset of the table slot
    const char* tb = reinterpret_cast<const char*>(th);
    int hi = (int)*reinterpret_cast<const unsigned int*>(tb + off) + nc * (1 << (20 - TBL_BITS));
    int lo = (int)*reinterpret_cast<const unsigned int*>(tb + off + TBL_N * 4);
    return __hiloint2double(hi, lo) * p;
}

// Conversion-free variant used by the dispersion lag loop.  t = ep + FX_MAGIC; the addition is folded into the caller's
// FMA chain (cp + FX_MAGIC is a per-chain constant), so it costs nothing.  With 2^20 <= t < 2^21 the mantissa of t IS ep
// in fixed point: high word bits 0..19 = floor(ep) + 2^19, low word = the 32-bit fraction F of ep.
//   * table slot / exponent insertion come from the high word on the integer pipes (the 2^19 bias and the exponent
//     field of t fold into one immediate);
//   * g = 1 + F 2^-32 is assembled from the bits of F (two shifts, one OR) and p = q(g) = exp((g-1) ln2/N) is a cubic
//     in g: no F2I / I2F (each costs ~3.5 cycles of the shared FP64/DMMA pipe on top of the XU slot) and no
//     DADD for the reduced argument.
// Shared-pipe cost: 3 DFMA + 1 DMUL = 8 cycles per warp, against 1 DADD + 3 DFMA + 1 DMUL + 2 conversions = 17 for
// exp_scaled.  (A fully integer polynomial with IMAD.HI / IMAD.WIDE measured SLOWER: wide integer multiplies issue at
// ~4.5 cycles and contend with the FP64 pipe -- tools/microbench/imad_peak.cu, profiles/r1_notes.md.)
// Error: ep is rounded to 2^-32 table units twice (<= 2^-32 ln2/N relative, unbiased) + the polynomial.
// Below EXP_NMIN the high word is clamped (result ~2^-1022, as exp_scaled); NaN does NOT propagate: callers flag dead
// chains (Comp<CLS_D>::dead).
__device__ __forceinline__ double exp_scaled_bits(double t, const double* __restrict__ tbl) {
    constexpr int S = 20 - TBL_BITS;
    constexpr int HI_MAGIC = 0x41380000;                                  // high word of FX_MAGIC
    constexpr int HI_MIN = HI_MAGIC + EXP_NMIN;                           // high word of EXP_NMIN + FX_MAGIC
    constexpr unsigned int FOLD = (unsigned int)(((unsigned long long)HI_MAGIC << S) & 0xffffffffull);
    const int ht = max(__double2hiint(t), HI_MIN);
    const unsigned int F = (unsigned int)__double2loint(t);
    const double g = __hiloint2double((int)(0x3FF00000u | (F >> 12)), (int)(F << 20));
    const double p = fma(g, fma(g, fma(g, EXQ_C3, EXQ_C2), EXQ_C1), EXQ_C0);
    const char* tb = reinterpret_cast<const char*>(tbl);
    const int off = (ht << 2) & ((TBL_N - 1) << 2);
    const unsigned int hi = *reinterpret_cast<const unsigned int*>(tb + off) + ((unsigned int)ht << S) - FOLD;
    const unsigned int lo = *reinterpret_cast<const unsigned int*>(tb + off + TBL_N * 4);
    return __hiloint2double((int)hi, (int)lo) * p;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

}  // namespace ngrtd
