/*
 * ngrtd.h -- C ABI of libngrtd.so: the B200-native (sm_100a) batched likelihood hot path of
 * uz226/NobleGas_RTD_MCMC (lumped-parameter convolution integral + closed-equilibrium noble-gas
 * model + log-likelihood + device-resident Metropolis step).
 *
 * The reference is pure Python and has no FFI; each entry point below names the reference
 * interface it replaces (paths relative to the reference root).  The reference-side binding a
 * maintainer would add is a ctypes stub -- see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, a negative NGRTD_E* code otherwise, never throws;
 *     ngrtd_last_error() returns a thread-local message for the last failure.
 *   - `*_dev` entry points take DEVICE pointers plus a CUDA stream handle (cudaStream_t cast to
 *     void*; NULL = default stream); they enqueue work and return without synchronising.
 *   - `*_host` entry points take HOST pointers and synchronise before returning.  Pinned (page-locked) parameter and
 *     logp buffers are read / written by the kernel directly (zero-copy over PCIe); pageable buffers go through a
 *     staged copy-in / kernel / copy-out pipeline.  Same results either way.
 *   - all floating point data is IEEE double; matrices are row-major; batches are [B, n].
 *   - numerical pathologies propagate as NaN / the -9999 sentinel exactly like the reference
 *     (no error is raised for bad parameter values).
 */
#ifndef NGRTD_H_
#define NGRTD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGRTD_VERSION 100

/* error codes */
#define NGRTD_OK 0
#define NGRTD_EINVAL (-1)   /* bad argument (unknown model, too many columns, null pointer ...) */
#define NGRTD_ECUDA (-2)    /* CUDA runtime error; see ngrtd_last_error() */
#define NGRTD_ENOMEM (-3)

/* RTD model types: `mod_type` strings of utils/convolution_integral_utils.py:178-196 */
#define NGRTD_MOD_NONE 0            /* mod_type2 = False */
#define NGRTD_MOD_PISTON 1          /* 'piston'        :178-181 */
#define NGRTD_MOD_EXPONENTIAL 2     /* 'exponential'   :183-184 */
#define NGRTD_MOD_EXP_PIST_FLOW 3   /* 'exp_pist_flow' :186-190 */
#define NGRTD_MOD_DISPERSION 4      /* 'dispersion'    :192-196 */

/* parameter slots = keys of ForwardMod.p_dict, age_ens_runs_mcmc/run_age_mcmc_utils.py:73-79 */
#define NGRTD_P_TAU1 0
#define NGRTD_P_TAU2 1
#define NGRTD_P_F1 2
#define NGRTD_P_F2 3
#define NGRTD_P_ETA1 4
#define NGRTD_P_ETA2 5
#define NGRTD_P_D1 6
#define NGRTD_P_D2 7
#define NGRTD_P_J 8            /* log10 of the 4He production rate (run_age_mcmc_utils.py:101) */
#define NGRTD_P_THALF_CFC 9    /* CFC-12 first-order decay half life (:107-109) */
#define NGRTD_P_LAMSF6 10      /* SF6 contamination factor (:160-161) */
#define NGRTD_NSLOT 11
/* column alias: this column is f1 AND f2 = 1 - f1 is formed on the device -- the reference's model makes f2 a deterministic
 * 1 - f1 (run_age_mcmc_utils.py:304) and still hands both to the Op; a batched caller that owns theta can leave the
 * redundant column out (one eighth fewer bytes over the host link).  Not allowed together with NGRTD_P_F1 / NGRTD_P_F2. */
#define NGRTD_P_F1_COMPLEMENT 11

/* rad_accum modes of tracer_conv_integral.update_pars (utils/convolution_integral_utils.py:123,313-333) */
#define NGRTD_ACC_NONE 0
#define NGRTD_ACC_3HE 1   /* g *= 1 - exp(-lambda tp) */
#define NGRTD_ACC_4HE 2   /* C_t += index * J        */

/* One modelled tracer = one ForwardMod instance of the reference (run_age_mcmc_utils.py:52-79,
 * conv_kwgs built at run_age_mcmc.py:200-224). */
typedef struct ngrtd_tracer {
    int32_t series;          /* column of the input-series matrix, or -1 for an all-zero series */
    int32_t rad_accum;       /* NGRTD_ACC_* */
    double lambda;           /* decay constant from thalf_2_lambda(t_half); 0 = no decay */
    int32_t use_thalf_cfc;   /* 1: lambda = ln2 / theta[thalf_cfc] per chain (the CFC12 rule, :107-109) */
    int32_t use_lamsf6;      /* 1: output *= 1 + theta[lamsf6]                (the SF6 rule, :160-161) */
} ngrtd_tracer;

/* likelihood kinds (pymc3 call sites run_age_mcmc_utils.py:389-396, ng_interp/noble_gas_mcmc.py:264-266) */
#define NGRTD_LIK_NORMAL 0
#define NGRTD_LIK_STUDENTT 1

typedef struct ngrtd_plan ngrtd_plan;   /* opaque: input series + lag tables resident in HBM */

int ngrtd_version(void);
/* compile-time options of this build (bit mask): experimental kernel variants kept behind macros, see README.md */
#define NGRTD_FEATURE_DM_TAIL 1     /* -DNGRTD_DM_TAIL: constant tail of dispersion components by quadrature */
#define NGRTD_FEATURE_XF_SWIZZLE 2  /* -DNGRTD_XF_SWIZZLE: half-warp-contiguous layout of the folded tables */
#define NGRTD_FEATURE_TP_DADD 4     /* -DNGRTD_TP_DADD: lag value carried in a register */
int ngrtd_build_features(void);

/* Page-locked host buffers for the *_host entry points, for callers that have no CUDA runtime binding of their own (the
 * reference is numpy: `np.frombuffer((ctypes.c_double * n).from_address(p))` views the block).  write_combined != 0 asks for
 * write-combined memory (cudaHostAllocWriteCombined): not snooped on the way to the device -- meant for theta buffers the CPU
 * only WRITES; reading it back on the CPU is slow.  Replaces nothing in the reference (its arrays are pageable numpy). */
int ngrtd_host_alloc(void** out, size_t bytes, int32_t write_combined);
int ngrtd_host_free(void* p);

/* Measured FP64 peak of a device in TFLOP/s (kind 0: DFMA, 1: DMMA m8n8k4 -- the two share one pipe on sm_100): the roofline
 * denominator bench.py reports against, measured in the same process (the pool's MEASURED_PEAKS.json has no FP64 entry).
 * device < 0: current device.  Measurement aid, replaces nothing in the reference. */
int ngrtd_fp64_peak_probe(int32_t device, int32_t kind, double* tflops_out);
const char* ngrtd_last_error(void);

/* ---- plan: replaces tracer_conv_integral.__init__(C_t, t_samp) (conv utils :105-108) and the
 *      per-call C_t.copy() of ForwardMod.perform (run_age_mcmc_utils.py:94) for ALL tracers at once.
 *  series      host [L, nseries], row k = concentrations k lags before sampling (i.e. after the
 *              reference's np.flip, conv utils :336)
 *  lag_index   host [L] DataFrame index values after the flip (used by '4He': C_t + index*J, :323);
 *              NULL = 0,1,2,...
 *  dtp         floor(t_samp - C_t.index[-1]) (:172); 0 at every call site of the reference
 *  device      CUDA device ordinal (-1 = current device)                                           */
int ngrtd_plan_create(ngrtd_plan** plan, int32_t L, int32_t nseries, const double* series,
                      const double* lag_index, double dtp, int32_t ntracer, const ngrtd_tracer* tracers,
                      int32_t mod_type1, int32_t mod_type2, int32_t device);
int ngrtd_plan_destroy(ngrtd_plan* plan);
int ngrtd_plan_ntracer(const ngrtd_plan* plan);

/* ---- batched ForwardMod.perform (run_age_mcmc_utils.py:81-163) for every tracer of the plan.
 *  theta        [B, ndim]; column i is parameter slot slot_of_col[i] (= par_names order)
 *  slot_of_col  HOST int32[ndim], values NGRTD_P_*
 *  out          [B, ntracer] modelled concentrations                                               */
int ngrtd_forward_dev(ngrtd_plan* plan, const double* theta_d, int64_t B, int32_t ndim,
                      const int32_t* slot_of_col, double* out_d, void* stream);
int ngrtd_forward_host(ngrtd_plan* plan, const double* theta_h, int64_t B, int32_t ndim,
                       const int32_t* slot_of_col, double* out_h);

/* ---- fused forward model + log-likelihood (the MCMC likelihood evaluation).
 *  obs_mu, obs_sd  HOST double[ntracer]  (run_age_mcmc_utils.py:353-356)
 *  nu_d            [B] Student-T degrees of freedom (ignored for NGRTD_LIK_NORMAL; may be NULL)
 *  logp            [B]; model_out optional [B, ntracer] (NULL to skip)                              */
int ngrtd_forward_loglik_dev(ngrtd_plan* plan, const double* theta_d, int64_t B, int32_t ndim,
                             const int32_t* slot_of_col, int32_t lik_kind, const double* obs_mu,
                             const double* obs_sd, const double* nu_d, double* logp_d, double* model_out_d,
                             void* stream);
int ngrtd_forward_loglik_host(ngrtd_plan* plan, const double* theta_h, int64_t B, int32_t ndim,
                              const int32_t* slot_of_col, int32_t lik_kind, const double* obs_mu,
                              const double* obs_sd, const double* nu_h, double* logp_h, double* model_out_h);
/* ---- the same call split into submit + wait, for callers that evaluate INDEPENDENT batches back to back (the
 *  reference's Monte-Carlo sweeps and posterior-predictive loops: aux_scripts/age_modeling_mcmc.rtd_explore.py:277-359,
 *  run_age_mcmc.py:243-319).  Up to NGRTD_HOST_SLOTS batches are in flight: the host->device copy of batch i+1 and the
 *  device->host copy of batch i-1 run under the kernel of batch i.  theta_h / nu_h must stay valid and logp_h /
 *  model_out_h must not be read until ngrtd_host_wait(plan, slot) returns; pinned host memory is needed for the overlap
 *  (pageable buffers work, the copies then block the submitting thread).  A busy slot is an error (wait first).       */
#define NGRTD_HOST_SLOTS 4
int ngrtd_forward_loglik_host_submit(ngrtd_plan* plan, const double* theta_h, int64_t B, int32_t ndim,
                                     const int32_t* slot_of_col, int32_t lik_kind, const double* obs_mu,
                                     const double* obs_sd, const double* nu_h, double* logp_h, double* model_out_h,
                                     int32_t slot);
int ngrtd_host_wait(ngrtd_plan* plan, int32_t slot);

/* ---- tracer_conv_integral.gen_g_tp() (conv utils :155-281): normalised RTD weights g[B, L],
 *      one model, parameters tau/eta/D [B] (eta, D may be NULL when unused).                        */
int ngrtd_rtd_weights_dev(int32_t mod_type, int32_t L, double dtp, const double* tau_d, const double* eta_d,
                          const double* D_d, int64_t B, double* g_d, void* stream);
/* ---- 'frac_inf_diff' RTD: frac_rtd_numba_disp (conv utils :36-63, numba in the reference) + gen_g_tp post-processing
 *      (:238-270).  tau, D, bbar, Phi_im: [B] (B <= 65,535); g: [B, L] normalised weights; fm_mu: [B] mean travel time
 *      (the reference's `FM_mu` attribute) or NULL.  Feed g to ngrtd_convolve_g_dev for the concentration.              */
int ngrtd_rtd_weights_fdm_dev(int32_t L, double dtp, const double* tau_d, const double* D_d, const double* bbar_d,
                              const double* phi_d, int64_t B, double* g_d, double* fm_mu_d, void* stream);
/* ---- same with an externally supplied advective RTD (frac_rtd_numba, conv utils :66-97; update_pars(f_tadv_ext=...) :142,
 *      :250-252): f_tadv_ext [L] on the lag grid, shared by the B (bbar, Phi_im) sets, linearly interpolated (np.interp). */
int ngrtd_rtd_weights_fdm_ext_dev(int32_t L, double dtp, const double* f_tadv_ext_d, const double* bbar_d,
                                  const double* phi_d, int64_t B, double* g_d, double* fm_mu_d, void* stream);
/* ---- tracer_conv_integral.convolve(g_tau=...) tail (:305-340): decay/ingrowth + input assembly + dot
 *      for externally supplied weights g[B, L]; series [L] newest-first, lag_index [L] or NULL.      */
int ngrtd_convolve_g_dev(int32_t L, double dtp, const double* g_d, int64_t B, const double* series_d,
                         const double* lag_index_d, const double* lambda_d /*[B]*/, int32_t rad_accum,
                         const double* J_d /*[B] linear J or NULL*/, double* out_d, void* stream);

/* ---- closed-equilibrium noble-gas model: noble_gas_fun(...).ce_exc / equil_conc[_dry]
 *      (utils/noble_gas_utils.py:103-253).  gases: HOST int32[ngas] with 0=He 1=Ne 2=Ar 3=Kr 4=Xe.
 *  what: 0 = ce_exc(add_eq_conc=True), 1 = ce_exc(False), 2 = equil_conc_dry(), 3 = equil_conc(),
 *        4 = solubility K, 5 = total pressure as used (lapse_rate(), :103-113, when P is NULL),
 *        6 = vapor_pressure() (:184-199) -- 5 and 6 do not depend on the gas (one column per listed gas all the same)
 *  E, T, Ae, F, P: [B]; P may be NULL = 'lapse_rate' (:97-98, :112);  S = salinity            */
int ngrtd_ce_dev(int32_t what, int32_t ngas, const int32_t* gases, const double* E_d, const double* T_d,
                 const double* Ae_d, const double* F_d, const double* P_d, double S, int64_t B,
                 double* out_d /*[B, ngas]*/, void* stream);
int ngrtd_ce_host(int32_t what, int32_t ngas, const int32_t* gases, const double* E_h, const double* T_h,
                  const double* Ae_h, const double* F_h, const double* P_h, double S, int64_t B, double* out_h);
/* ce_exc_wrapper(theta) of ng_interp/noble_gas_mcmc.py:205-213: theta[B,4] = log10 Ae, log10 F, E, T */
int ngrtd_ce_wrapper_dev(int32_t ngas, const int32_t* gases, const double* theta_d, int64_t B, double* out_d,
                         void* stream);

/* ---- CFC-11/12/113 and SF6 solubility + closed-system excess-air corrections: cfc_ce_corr / sf6_ce_corr of
 *      utils/cfc_utils.py:25-152,160-306 (batch callers: age_modeling_mcmc.prep.py:242-303).
 *  species: HOST int32[nspecies] with values 11, 12, 113 (CFCs) or 6 (SF6)
 *  what: 0 = equil_air_conc_*(C_meas), 1 = equil_aq_conc_*(z_i), 2 = ce_exc_conc_*(z_i), 3 = solubility_*()
 *  E, T, Ae (ccSTP/g, as passed to the constructors), F: [B];  X: [B, nspecies] C_meas or z_i;  S = salinity      */
int ngrtd_cfc_dev(int32_t what, int32_t nspecies, const int32_t* species, const double* E_d, const double* T_d,
                  const double* Ae_d, const double* F_d, const double* X_d, double S, int64_t B, double* out_d,
                  void* stream);
int ngrtd_cfc_host(int32_t what, int32_t nspecies, const int32_t* species, const double* E_h, const double* T_h,
                   const double* Ae_h, const double* F_h, const double* X_h, double S, int64_t B, double* out_h);

/* ---- stand-alone log-likelihood over model outputs mu[B, T] (pymc3 Normal / StudentT logp). */
int ngrtd_loglik_dev(int32_t lik_kind, int32_t T, const double* mu_d, const double* obs_mu, const double* obs_sd,
                     const double* nu_d, int64_t B, double* logp_d, void* stream);

/* ---- device-resident batched sampler: replaces pymc3's DEMetropolisZ / metrop_select loop that the reference
 *      configures at age_ens_runs_mcmc/run_age_mcmc_utils.py:412-417 and ng_interp/noble_gas_mcmc.py:408-415, and the
 *      model assembly (priors, transforms, likelihood) of run_age_mcmc_utils.py:275-397 / noble_gas_mcmc.py:216-267.
 *      pymc3 3.11.2 is a third-party dependency absent from the reference tree; semantics: SURVEY.md App. B.      */
#define NGRTD_PRIOR_UNIFORM 0      /* mc.Uniform(p0, p1)          -> interval transform          */
#define NGRTD_PRIOR_BETA 1         /* mc.Beta(p0, p1) mapped affinely onto [lo, hi] -> log-odds  */
#define NGRTD_PRIOR_NORMAL 2       /* mc.Normal(p0, p1)           -> untransformed               */
#define NGRTD_PRIOR_HALFNORMAL 3   /* mc.HalfNormal(p0)           -> log transform               */
#define NGRTD_VAL_NU 11            /* value register of the Student-T nu_ variable (after NGRTD_P_* 0..10) */
/* noble-gas model value registers (ng_interp/noble_gas_mcmc.py:224-250) */
#define NGRTD_NG_LOG10AE 0
#define NGRTD_NG_LOG10F 1
#define NGRTD_NG_E 2
#define NGRTD_NG_M 3
#define NGRTD_NG_B 4
#define NGRTD_MAX_DIM 10

typedef struct ngrtd_prior {
    int32_t kind;      /* NGRTD_PRIOR_* */
    int32_t target;    /* value register: NGRTD_P_* / NGRTD_VAL_NU (age model) or NGRTD_NG_* / NGRTD_VAL_NU (noble-gas model) */
    double p0, p1;
    double lo, hi;     /* beta only */
} ngrtd_prior;

typedef struct ngrtd_sampler_cfg {
    int32_t ndim;
    ngrtd_prior prior[NGRTD_MAX_DIM];
    int32_t lik_kind;              /* NGRTD_LIK_* */
    int32_t nu_sampled;            /* 1: nu = nu_lo + (nu_hi - nu_lo) * nu_  (run_age_mcmc_utils.py:291-292) */
    double nu_lo, nu_hi, nu_fixed;
    int32_t nobs;                  /* tracers of the plan, or gases of the noble-gas model */
    double obs_mu[8], obs_sd[8];
    int32_t f2_from_f1;            /* f2 = 1 - f1 (run_age_mcmc_utils.py:304) */
    int32_t proposal_dist;         /* 0 Uniform(-1,1) (pymc3 DEMetropolisZ default), 1 Normal(0,1) */
    int32_t de_mcz;                /* 1 DE-MC-Z, 0 random-walk Metropolis */
    int32_t tune_target;           /* 0 lambda (pymc3 default for DEMetropolisZ), 1 scaling */
    int32_t tune_interval;
    double scaling;                /* 0.001 in pymc3 */
    double lamb;                   /* <= 0: 2.38 / sqrt(2 ndim) */
    double tune_drop_fraction;     /* 0.9 in pymc3 */
    int32_t hist_cap;              /* history ring capacity per chain; >= tune+draws reproduces pymc3 exactly */
    uint64_t seed;
    int64_t chain_offset;          /* global id of local chain 0 (multi-GPU sharding; Philox counters use global ids) */
    int32_t ngas;                  /* noble-gas model only */
    int32_t gases[5];
} ngrtd_sampler_cfg;

typedef struct ngrtd_sampler ngrtd_sampler;

/* plan == NULL selects the noble-gas closed-equilibrium model.  q0: HOST [ndim] start in transformed space or NULL
 * for pymc3's model.test_point.  All chains start at the same point, as in the reference. */
int ngrtd_sampler_create(ngrtd_sampler** s, const ngrtd_sampler_cfg* cfg, ngrtd_plan* plan, int64_t nchains,
                         const double* q0, int32_t device);
int ngrtd_sampler_destroy(ngrtd_sampler* s);
/* BASELINE config 4 (joint fit over wells x recharge-ensemble members): every group of `chains_per_group` consecutive
 * GLOBAL chain ids gets its own observation row (the reference loops wells serially, run_age_mcmc.py:122, and builds
 * obs_mu/obs_err per well from ens_dict, run_age_mcmc_utils.py:353-356).  obs_mu, obs_sd: HOST [ngroups, nobs]. */
int ngrtd_sampler_set_obs_groups(ngrtd_sampler* s, const double* obs_mu, const double* obs_sd, int64_t ngroups,
                                 int64_t chains_per_group);
/* DE-MC-Z with a SHARED archive (the original scheme of ter Braak & Vrugt 2008; pymc3's DEMetropolisZ keeps one archive per
 * chain because its chains run in separate processes, run_age_mcmc_utils.py:412-417): the two history entries of a proposal
 * are drawn from the archives of the `chains_per_population` consecutive GLOBAL chains of the chain's population, using the
 * entries that were complete when the current launch started.  Lets a population share the modes its members found
 * (BASELINE config 4: one population per (well, ensemble member) group).  0 restores per-chain archives.  Populations
 * must not straddle shards; call between launches. */
int ngrtd_sampler_set_population(ngrtd_sampler* s, int64_t chains_per_population);
/* advance every chain by nsteps Metropolis steps in ONE kernel launch.  tune: tuning phase; record: update the
 * per-chain Welford statistics and, if trace_d != NULL, write natural-space draws trace_d[ceil(nsteps/thin), B, ndim]. */
int ngrtd_sampler_run(ngrtd_sampler* s, int64_t nsteps, int32_t tune, int32_t record, int32_t thin, double* trace_d,
                      void* stream);
int ngrtd_sampler_stop_tuning(ngrtd_sampler* s);   /* DEMetropolisZ.stop_tuning: drop the oldest fraction of the history */
/* what: 0 q [B,ndim] (transformed), 1 logp [B], 2 lamb [B], 3 scaling [B], 4 accepted count [B], 5 mean [B,ndim],
 *       6 M2 [B,ndim] (Welford, natural values), 7 history ring [hist_cap,B,ndim], 8 accepted since the last tuning point [B]
 *       -- copied to out_d (device) on `stream`.  Together with ngrtd_sampler_info these are the complete sampler state
 *       (checkpoint; the reference only saves finished traces, run_age_mcmc_utils.py:425).                                 */
int ngrtd_sampler_get(ngrtd_sampler* s, int32_t what, double* out_d, void* stream);
/* restore: same selectors; setting q (0) re-evaluates logp unless logp (1) is set afterwards */
int ngrtd_sampler_set(ngrtd_sampler* s, int32_t what, const double* in_d, void* stream);
/* K6 on the device ("ncclAllReduce of pooled moments"): sums over the sampler's chains of the per-chain
 * Welford statistics -- out_d[0..nd) = sum_c mean, [nd..2nd) = sum_c mean^2, [2nd..3nd) = sum_c M2, out_d[3nd] = chains --
 * 3*ndim + 1 doubles on the device (deterministic two-stage reduction, same bits for the same chains).  Ranks add these
 * vectors with ONE all-reduce and derive the pooled mean / R-hat / ESS that ArviZ's az.summary reports in the reference
 * (run_age_mcmc.py:234, noble_gas_mcmc.py:450) -- see noblegas_rtd_mcmc_b200/distributed.py:pooled_summary. */
int ngrtd_sampler_pooled_moments(ngrtd_sampler* s, double* out_d, void* stream);
int ngrtd_sampler_set_counters(ngrtd_sampler* s, int64_t step, int64_t ndraws, int64_t hist_start);
int ngrtd_sampler_info(const ngrtd_sampler* s, int64_t* step, int64_t* ndraws, int64_t* hist_start);
int ngrtd_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);   /* known-answer hook */

#ifdef __cplusplus
}
#endif
#endif /* NGRTD_H_ */
