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
// The N-entry table of 2^(j/N) lives in shared memory as 8-byte entries, replicated REP times: lane l gathers from
// copy l & (REP-1) at double index j*REP + copy.  Two configurations (ExpCfg<TB>, TB = log2 N), chosen per kernel:
//   TB = 7  (k_mcmc_age: the sampler kernel has ~3 KB of shared memory to spare): 128 entries x 4 copies = 4 KB, cubic q,
//           3 DFMA + 1 DMUL on the shared FP64/DMMA pipe per weight, max rel err of q 2.8e-13;
//   TB = 11 (k_forward, r2): 2,048 entries x 1 copy = 16 KB, QUADRATIC q -- one DFMA less per dispersion weight on the pipe
//           that bounds the lag loop -- max rel err of q 2.0e-13.  The exponent is carried in units of 4 table steps
//           (SUB = 2) so that its integer part still fits the 20 mantissa bits of the high word over the whole double
//           range (1022 * 2048 / 4 < 2^19); the two missing index bits are the top bits of the low word.
// tools/exp_poly_g.py prints the constants and error bounds.
// r1 measurements on the replication (profiles/r1_notes.md): the gather is never the bound (LSU data pipe 46-64 % busy, the
// loop is held by the shared FP64/DMMA pipe), so the larger table is not replicated (r2: 2 copies measured slower).
constexpr double LN2 = 0.693147180559945309417232121458;
#ifndef NGRTD_TB11_REP_BITS
#define NGRTD_TB11_REP_BITS 0
#endif
template <int TB>
struct ExpCfg;
template <>
struct ExpCfg<7> {
    static constexpr int BITS = 7, N = 128, SUB = 0, REP_BITS = 2, REP = 4, DOUBLES = N * REP, DEG = 3;
    static constexpr double C0 = 0.99459942332194162, C1 = 0.0053859675366433966, C2 = 1.4582602945559778e-05,
                            C3 = 2.6538188920206277e-08;      // max rel err 2.8e-13
};
template <>
struct ExpCfg<11> {
    static constexpr int BITS = 11, N = 2048, SUB = 2, REP_BITS = NGRTD_TB11_REP_BITS, REP = 1 << REP_BITS, DOUBLES = N * REP, DEG = 2;
    static constexpr double C0 = 0.99966160651623492, C1 = 0.00033833619981139687, C2 = 5.7284155667395806e-08,
                            C3 = 0.0;                         // max rel err 2.0e-13
    // -DNGRTD_EXP_R32 (experiment, not the default): q(g) = fma(g, C1, r) with r = C0 + C2 g^2 assembled from bits instead of
    // a second DFMA.  r stays in one binade
    // ([0.5, 1), ulp 2^-53) and varies by 3 C2 = 1.7e-7 = 1.55e9 ulps < 2^32: r = R0 + N 2^-44, R0 = C0 + C2,
    // N = round(C2 2^44 (g^2 - 1)) < 2^22 from one FMUL + one FFMA (against the 1.5 2^23 rounding constant) on the FP32 pipe,
    // added to the LOW word of R0 by one integer shift-add (no carry: lo(R0) + 1.55e9 < 2^32).  The constants and the error
    // (max rel err of q 2.5e-13, difference to the two-DFMA form zero-mean, <= 5e-14) are printed by tools/exp_r32_prototype.py.
    // Measured on the B200: parity 5.87e-14, but the cfg-3 launch goes from 0.1014 to 0.1051 ms -- one DFMA less, 19 more
    // instructions per loop trip (93 -> 112): the loop is as sensitive to issued instructions as to FP64-pipe operations.
    static constexpr unsigned int R_HI = 0x3feffd3au, R_LO_ADD = 0xf522cf52u;
    static constexpr float R_KN = 1007753.5f, R_K0 = 11575158.0f;
};
// scale of the exponent handed to exp_scaled_bits: t = e * exp_k + FX_MAGIC (units of 2^SUB table steps)
template <int TB>
__host__ __device__ constexpr double exp_k() { return (ExpCfg<TB>::N >> ExpCfg<TB>::SUB) / LN2; }
// below this (in the same units) the result is clamped at ~2^-1022 instead of flushing to zero
template <int TB>
__host__ __device__ constexpr int exp_clamp() { return -1022 * (ExpCfg<TB>::N >> ExpCfg<TB>::SUB); }
// what ngrtd_plan_create subtracts from the high word of every table entry (see exp_scaled_bits)
template <int TB>
__host__ __device__ constexpr int exp_ins_shift() { return 20 - (TB - ExpCfg<TB>::SUB); }
template <int TB>
__host__ __device__ constexpr unsigned int exp_tbl_fold() {
    return (unsigned int)(((unsigned long long)0x41380000u << exp_ins_shift<TB>()) & 0xffffffffull);
}
#ifndef NGRTD_FWD_TB
#define NGRTD_FWD_TB 11
#endif
constexpr int FWD_TB = NGRTD_FWD_TB;      // k_forward
constexpr int MCMC_TB = 7;      // k_mcmc_age when shared memory is short (many sampler dimensions, streamed lag tables)
constexpr int MCMC_TB_BIG = 11; // k_mcmc_age when the 16 KB table fits next to the resident lag tables (launch_mcmc_age_t)
constexpr double FX_MAGIC = 1572864.0;               // 1.5 * 2^20: ulp 2^-32, integer part biased by 2^19

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
    const double* tbl7;   // [128]  2^(j/N) with j << (20 - log2 N) subtracted from the high word (see exp_scaled_bits)
    const double* tbl11;  // [2048] same, N = 2048
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

// Which plans cut the lag loop at Kc and add the tail [Kc, L) analytically in the epilogue: exponential-class components by
// closed form, dispersion components by Gauss-Legendre quadrature of the smooth tail (WarpTiles::dm_tail in
// ngrtd_forward.cuh; opt-in experiment in r1, the default since r2 -- -DNGRTD_NO_DM_TAIL builds keep the full lag loop for
// dispersion plans, for A/B measurements).  Plan creation leaves Kc = Lpad when the plan does not qualify.
#ifdef NGRTD_NO_DM_TAIL
constexpr bool DM_TAIL = false;
#else
constexpr bool DM_TAIL = true;
#endif
__host__ __device__ __forceinline__ bool tail_active(const PlanView& pv, bool any_g, bool any_d) {
    return (any_g || any_d) && (!any_d || DM_TAIL) && pv.Kc < pv.L;
}

struct SlotMap {
    int ndim;
    signed char col_of_slot[NSLOT];   // -1: not in par_names -> p_dict default
    signed char f2_complement;        // 1: f2 = 1 - f1 (NGRTD_P_F1_COMPLEMENT column)
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
    p.f2 = sm.f2_complement ? 1.0 - p.f1 : get(3, 0.0);       // default 0.0
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

// exp(e) for e <= ~0, given t = ep + FX_MAGIC with ep = e * exp_k<TB>() (= e N / (2^SUB ln2)).  The addition is folded into
// the caller's FMA chain (cp + FX_MAGIC is a per-chain constant), so it costs nothing.  With 2^20 <= t < 2^21 the mantissa
// of t IS ep in fixed point: high word bits 0..19 = floor(ep) + 2^19, low word = the 32-bit fraction of ep, whose top SUB
// bits are the low bits of the table index and whose remaining bits F are the reduced argument.
//   * table slot / exponent insertion come from the high word on the integer pipes: the table stores 2^(j/N) with
//     (j >> SUB) << (20 - (TB - SUB)) and the image of FX_MAGIC's high word (exp_tbl_fold<TB>(): the 2^19 bias and the
//     exponent field of t) subtracted from its high word, so the insertion is ONE integer multiply-add;
//   * g = 1 + F 2^-(32-SUB) is assembled from the bits of F (shifts, one OR) and p = q(g) = exp((g-1) ln2/N) is a polynomial
//     in g (cubic for N = 128, quadratic for N = 2048): no F2I / I2F conversions and no DADD for the reduced argument.
// Shared FP64/DMMA-pipe cost: DEG DFMA + 1 DMUL.  History (profiles/r1_notes.md): the first version rounded with F2I/I2F
// (XU pipe) and a centred polynomial; a fully integer polynomial with IMAD.HI / IMAD.WIDE measured SLOWER (wide integer
// multiplies issue at ~4.5 cycles and contend with the FP64 pipe -- tools/microbench/imad_peak.cu).
// Error: ep is rounded to 2^-32 twice (<= 2^-32 2^SUB ln2/N relative, unbiased) + the polynomial.
// Below exp_clamp<TB>() the high word is clamped (result ~2^-1022) instead of flushing to zero: one VIMNMX instead of a
// compare and two selects.  NaN does NOT propagate through the integer path: chains whose largest weight would be below
// 2^-1022, or with NaN parameters, are declared dead in CompD::init_q and poisoned in the epilogue, which reproduces the
// reference's 0/0 = NaN when every weight underflows.
template <int TB>
__device__ __forceinline__ double exp_scaled_bits(double t, const double* __restrict__ tbl0, unsigned int tbl_lane) {
    using E = ExpCfg<TB>;
    constexpr int S = exp_ins_shift<TB>();
    constexpr int HI_MAGIC = 0x41380000;                                  // high word of FX_MAGIC
    constexpr int HI_MIN = HI_MAGIC + exp_clamp<TB>();                    // high word of exp_clamp + FX_MAGIC
    const int ht = max(__double2hiint(t), HI_MIN);
    const unsigned int lo = (unsigned int)__double2loint(t);
    // g = 1 + F 2^-(32-SUB): mantissa = the low (32 - SUB) bits of lo, left-aligned
#ifdef NGRTD_EXPV1
    // right shifts are funnel shifts (SHF) in SASS; (x >> s) + c is one LEA.HI
    const unsigned int Fm = lo & (0xffffffffu >> E::SUB);
    const double g = __hiloint2double((int)(0x3FF00000u + (Fm >> (12 - E::SUB))), (int)(lo << (20 + E::SUB)));
#else
    const double g = __hiloint2double((int)(0x3FF00000u | ((lo << E::SUB) >> 12)), (int)(lo << (20 + E::SUB)));
#endif
    double p;
    if constexpr (E::DEG == 3) p = fma(g, fma(g, fma(g, E::C3, E::C2), E::C1), E::C0);
#ifdef NGRTD_EXP_R32
    else {
        // experiment (r2 session 3, measured SLOWER: 0.1014 -> 0.1051 ms, see ExpCfg<11>): r = C0 + C2 g^2 from bits
        const float gf = __uint_as_float(0x3F800000u | ((lo << E::SUB) >> 9));
        const float mf = __fmaf_rn(__fmul_rn(gf, gf), E::R_KN, E::R_K0);
        p = fma(g, E::C1, __hiloint2double((int)E::R_HI, (int)((__float_as_uint(mf) << 9) + E::R_LO_ADD)));
    }
#else
    else p = fma(g, fma(g, E::C2, E::C1), E::C0);
#endif
    // entry j = (low bits of floor(ep)) : (top SUB bits of lo) of this lane's copy.  tbl_lane = byte offset of the lane's
    // copy from the start of the kernel's dynamic shared memory (the table sits at its very beginning): the masked index
    // and the copy select merge into one LOP3, and the base is the immediate of the load.
    constexpr int OSH = 3 + E::REP_BITS + E::SUB;
#ifdef NGRTD_EXPV1
    const unsigned int sh = (E::SUB == 0) ? ((unsigned int)ht << OSH) : (((unsigned int)ht << OSH) + (lo >> (32 - OSH)));
#else
    const unsigned int sh = (E::SUB == 0) ? ((unsigned int)ht << OSH) : __funnelshift_l(lo, (unsigned int)ht, OSH);
#endif
    const unsigned int off = (sh & ((unsigned int)(E::N - 1) << (3 + E::REP_BITS))) | tbl_lane;
    const double T = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(tbl0) + off);
    const unsigned int hi = (unsigned int)__double2hiint(T) + ((unsigned int)ht << S);   // the table's high words carry -fold
    return __hiloint2double((int)hi, __double2loint(T)) * p;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

}  // namespace ngrtd
