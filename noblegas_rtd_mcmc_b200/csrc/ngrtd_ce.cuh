// ngrtd_ce.cuh -- closed-equilibrium noble-gas model (K3), one chain per thread.
// Restates utils/noble_gas_utils.py:103-253 of the reference (lapse_rate, solubility, vapor_pressure,
// equil_conc, equil_conc_dry, ce_exc) with the same operation order so results agree to a few ulp.
#pragma once
#include "ngrtd_common.cuh"

namespace ngrtd {

// gas ids: 0 He, 1 Ne, 2 Ar, 3 Kr, 4 Xe
__constant__ double c_atm_std[5] = {5.24e-6, 1.818e-5, 9.34e-3, 1.14e-6, 8.7e-8};   // noble_gas_utils.py:38-49
__constant__ double c_sol[5][4] = {{-0.00953, 0.107722, 0.001969, -0.043825},        // :138-142
                                   {-7.259, 6.95, -1.3826, 0.0538},
                                   {-9.52, 8.83, -1.8959, 0.0698},
                                   {-6.292, 5.612, -0.8881, -0.0458},
                                   {-3.902, 2.439, 0.3863, -0.221}};
__constant__ double c_setch[5][3] = {{-10.081, 15.1068, 4.8127},                     // :145-149
                                     {-11.9556, 18.4062, 5.5464},
                                     {-10.6951, 16.7513, 4.9551},
                                     {-9.9787, 15.7619, 4.6181},
                                     {-14.5524, 22.5255, 6.7513}};

__device__ __forceinline__ double ce_lapse_rate(double E) {   // :112
    return pow(1.0 - .0065 * E / 288.15, 5.2561) * 0.000101325;
}

__device__ __forceinline__ double ce_poly(const double* A, double T_k) {   // A0 + A1/t + A2/t^2 + A3/t^3, t = .001 T_k
    double t = .001 * T_k;
    return A[0] + (A[1] / t) + (A[2] / (t * t)) + (A[3] / (t * t * t));
}

__device__ __forceinline__ double ce_solubility(int gas, double T, double S) {   // :117-180
    double T_k = T + 273.15;
    double gamma = 1.0;
    if (T < 65.0) {
        double setch = c_setch[gas][0] + (c_setch[gas][1] / (.01 * T_k)) + (c_setch[gas][2] * log(.01 * T_k));
        gamma = exp(S * setch);
    }
    double K_h;
    if (gas == 0) {
        double F = exp(ce_poly(c_sol[0], T_k));
        double Frac_He_gas = 5.24e-6 / 9.31e-3;
        double X_Ar_water = 1.0 / exp(ce_poly(c_sol[2], T_k)) * 9.31e-3;
        double X_he_water = F * Frac_He_gas * X_Ar_water;
        K_h = 5.24e-6 / X_he_water;
    } else {
        K_h = exp(ce_poly(c_sol[gas], T_k));
    }
    return gamma * K_h;
}

__device__ __forceinline__ double ce_vapor_pressure(double T) {   // :184-199
    double A, Bc, C;
    if (T <= 99.0) { A = 8.07131; Bc = 1730.63; C = 233.426; }
    else { A = 8.14019; Bc = 1810.94; C = 244.485; }
    double P = exp10(A - (Bc / (C + T)));
    P = P / 760. * 101325;
    return P / 1.0e9;
}

// what: 0 ce_exc(True), 1 ce_exc(False), 2 equil_conc_dry, 3 equil_conc (wet), 4 solubility,
//       5 total pressure P as used (lapse_rate(), :103-113, when no pressure is given), 6 vapor_pressure() (:184-199)
__device__ __forceinline__ double ce_eval(int what, int gas, double E, double T, double Ae, double F, double P,
                                          double S) {
    if (what == 5) return P;
    if (what == 6) return ce_vapor_pressure(T);
    double K = ce_solubility(gas, T, S);
    if (what == 4) return K;
    double z = c_atm_std[gas];
    if (what == 3) return (z * P / K) * (22414. / 18.);                        // :209-211
    double pv = ce_vapor_pressure(T);
    double p_i = z * (P - pv);                                                  // :225
    double C_eq = (T < 0.0) ? -9999.0 : (p_i / K) * (22414. / 18.);            // :226-230
    if (what == 2) return C_eq;
    double C_ex = ((1 - F) * Ae * z) / (1 + ((F * Ae * z) / C_eq));            // :248
    return what == 0 ? C_ex + C_eq : C_ex;
}

struct GasList { int n; int id[5]; };

// ---- sampler path (k_mcmc_ng): ce_exc(True) of several gases at one parameter vector, S = 0.  Same formulae as ce_eval,
// with the work that does not depend on the gas done once per step (1/t, vapour pressure) and the divisions folded:
//   ln K = A0 + it (A1 + it (A2 + it A3)),  C_eq = z (P - P_v) exp(-ln K) 22414/18,
//   C_ex = (1-F) Ae z / (1 + F Ae z / C_eq) = (1-F) Ae z C_eq / (C_eq + F Ae z)          (also for the -9999 sentinel)
// One exp and one division per gas instead of one exp, one log, one more exp (the Setchenow factor, == 1 at S = 0) and
// eight divisions; results agree with ce_eval to a few ulp (tests/test_sampler_gpu.py compares trajectories with the
// numpy restatement of the reference formulae).  Helium goes through ce_eval (its solubility is derived from argon's).
struct CeStep {
    double it, pda, num_c;     // 1/t (t = T_K / 1000), P - P_v, 22414/18
    bool neg, bad;             // T < 0 -> -9999 sentinel (:227-228); T_K <= 0 -> NaN (log of a negative number in the reference)
};
__device__ __forceinline__ CeStep ce_step(double T, double P) {
    CeStep c;
    const double T_k = T + 273.15;
    c.it = 1000.0 / T_k;
    c.pda = P - ce_vapor_pressure(T);
    c.num_c = 22414. / 18.;
    c.neg = T < 0.0;
    c.bad = !(T_k > 0.0);
    return c;
}
__device__ __forceinline__ double ce_exc_step(int gas, const CeStep& c, double E, double T, double Ae, double F, double P) {
    if (gas == 0) return ce_eval(0, 0, E, T, Ae, F, P, 0.0);
    if (c.bad) return __longlong_as_double(0x7ff8000000000000LL);
    const double lnK = c_sol[gas][0] + c.it * (c_sol[gas][1] + c.it * (c_sol[gas][2] + c.it * c_sol[gas][3]));
    const double z = c_atm_std[gas];
    const double C_eq = c.neg ? -9999.0 : (z * c.pda) * exp(-lnK) * c.num_c;
    const double fa = F * Ae * z;
    return ((1.0 - F) * Ae * z) * C_eq / (C_eq + fa) + C_eq;
}
// lapse_rate (:112) with pow(x, a) as exp(a log x): |log x| < 0.1 here, so the product a log x carries < 1e-16 absolute error
__device__ __forceinline__ double ce_lapse_rate_step(double E) {
    return exp(5.2561 * log(1.0 - .0065 * E / 288.15)) * 0.000101325;
}

#ifndef NGRTD_NO_AUX_KERNELS   // compiled into part 0 only (ngrtd_api.cu, build partitioning)
__global__ void k_ce(int what, GasList gl, const double* __restrict__ E, const double* __restrict__ T,
                     const double* __restrict__ Ae, const double* __restrict__ F, const double* __restrict__ P,
                     double S, long long B, double* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= B) return;
    double e = E ? E[i] : 0.0, t = T[i];
    double ae = Ae ? Ae[i] : 0.0, f = F ? F[i] : 0.0;
    double p = P ? P[i] : ce_lapse_rate(e);
    for (int g = 0; g < gl.n; g++) out[i * gl.n + g] = ce_eval(what, gl.id[g], e, t, ae, f, p, S);
}
#endif

// ce_exc_wrapper (ng_interp/noble_gas_mcmc.py:205-213): theta = [log10 Ae, log10 F, E, T]
#ifndef NGRTD_NO_AUX_KERNELS   // compiled into part 0 only (ngrtd_api.cu, build partitioning)
__global__ void k_ce_wrapper(GasList gl, const double* __restrict__ theta, long long B, double* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= B) return;
    double ae = exp10(theta[i * 4 + 0]), f = exp10(theta[i * 4 + 1]);
    double e = theta[i * 4 + 2], t = theta[i * 4 + 3];
    double p = ce_lapse_rate(e);
    for (int g = 0; g < gl.n; g++) out[i * gl.n + g] = ce_eval(0, gl.id[g], e, t, ae, f, p, 0.0);
}
#endif

// ---------------------------------------------------------------- CFC / SF6 solubility and excess-air corrections
// Restates utils/cfc_utils.py (cfc_ce_corr :25-152, sf6_ce_corr :160-306): the batch callers are the 50,000-draw
// observation-ensemble loops of age_modeling_mcmc.prep.py:242-303.  species: 11, 12, 113 (CFCs), 6 (SF6).
__device__ __forceinline__ double cfc_solubility(int sp, double T, double S) {      // :62-83, :196-209  [mol atm^-1 kg^-1]
    double a1, a2, a3, b1, b2, b3;
    switch (sp) {
        case 11: a1 = -136.2685; a2 = 206.1150; a3 = 57.2805; b1 = -.148598; b2 = 0.095114; b3 = -0.0163396; break;
        case 12: a1 = -124.4395; a2 = 185.4299; a3 = 51.6383; b1 = -0.149779; b2 = 0.094668; b3 = -0.0160043; break;
        case 113: a1 = -136.129; a2 = 206.475; a3 = 55.8957; b1 = -0.02754; b2 = 0.006033; b3 = 0.0; break;
        default: a1 = -98.7264000; a2 = 142.803; a3 = 38.8746; b1 = 0.0268696; b2 = -0.0334407; b3 = 0.0070843; break;
    }
    double T_k = T + 273.15;
    double th = T_k / 100;
    return exp(a1 + a2 * (100 / T_k) + a3 * log(T_k / 100) + S * (b1 + b2 * th + b3 * (th * th)));
}
__device__ __forceinline__ double cfc_vapor_pressure_atm(double T) {                 // :35-52
    return ce_vapor_pressure(T) / 0.000101325;
}
__device__ __forceinline__ double cfc_lapse_rate_atm(double E) { return pow(1.0 - .0065 * E / 288.15, 5.2561); }   // :54-60

struct SpeciesList { int n; int id[4]; };

// what: 0 equil_air_conc (measured aqueous -> atmospheric mixing ratio, :85-105 / :228-247),
//       1 equil_aq_conc  (mixing ratio -> aqueous, :107-124 / :262-278), 2 ce_exc_conc (:126-143 / :280-296), 3 solubility
// Ae is the constructor argument in ccSTP/g (the classes multiply by 1000 themselves, :29 / :164).
#ifndef NGRTD_NO_AUX_KERNELS   // compiled into part 0 only (ngrtd_api.cu, build partitioning)
__global__ void k_cfc(int what, SpeciesList sl, const double* __restrict__ E, const double* __restrict__ T,
                      const double* __restrict__ Ae_, const double* __restrict__ F_, const double* __restrict__ X, double S,
                      long long B, double* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= B) return;
    const double t = T[i], e = E ? E[i] : 0.0;
    const double Ae = Ae_ ? Ae_[i] * 1000. : 0.0, F = F_ ? F_[i] : 0.0;
    const double P = cfc_lapse_rate_atm(e), Pv = cfc_vapor_pressure_atm(t);
    const double P_da = P - Pv;
    for (int k = 0; k < sl.n; k++) {
        const int sp = sl.id[k];
        const bool sf6 = sp == 6;
        const double Ki = cfc_solubility(sp, t, S);
        const double x = X ? X[i * sl.n + k] : 0.0;
        double v;
        if (what == 3) {
            v = Ki;
        } else if (what == 0) {
            const double molar_volume = 22414.1;
            v = (x + ((x * F * (Ae / molar_volume)) / (Ki * P_da))) / (Ki * P_da + (Ae / molar_volume));
            if (sf6) v = v * 0.001;
        } else if (what == 1) {
            v = Ki * x * P_da;
            if (sf6) v = v / 0.001;
        } else {
            double C_eq = Ki * x * (P - Pv);
            double A = Ae / 22414 / 1e-12;
            v = ((1 - F) * A * (x * 1e-12)) / (1 + F * A * ((x * 1e-12 / C_eq)));
            if (sf6) v = 1000 * v;
        }
        out[i * sl.n + k] = v;
    }
}
#endif

}  // namespace ngrtd
