/* Plain-C caller of libngrtd.so (include/ngrtd.h): what a non-Python host of the reference's hot path binds.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/c_abi_demo.c -o /tmp/c_abi_demo -Lnoblegas_rtd_mcmc_b200 -lngrtd \
 *       -Wl,-rpath,$PWD/noblegas_rtd_mcmc_b200 -lm
 *   /tmp/c_abi_demo            (needs a B200; there is no CPU fallback)
 *
 * Evaluates tracer_conv_integral(...).convolve() of the reference (utils/convolution_integral_utils.py:155-340) for a batch of
 * exponential-model residence times on a synthetic input series, plus the fused Normal log-likelihood, through the host-buffer
 * entry points, and checks the result against the direct sum  C = sum_k g_k x_k,  g_k = exp(-tp_k / tau) / sum_j exp(-tp_j / tau). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ngrtd.h"

#define L 200
#define B 1000

int main(void) {
    static double series[L], theta[B], out[B], logp[B];
    for (int k = 0; k < L; k++) series[k] = 100.0 / (1.0 + exp((k - 40.0) / 6.0)) + 1.0;   /* newest lag first (np.flip, :336) */
    for (int b = 0; b < B; b++) theta[b] = 2.0 + 0.25 * b;                                   /* tau1 */
    ngrtd_tracer tr = {0, NGRTD_ACC_NONE, 0.0, 0, 0};
    ngrtd_plan* plan = NULL;
    int rc = ngrtd_plan_create(&plan, L, 1, series, NULL, 0.0, 1, &tr, NGRTD_MOD_EXPONENTIAL, NGRTD_MOD_NONE, -1);
    if (rc) { fprintf(stderr, "plan: %s\n", ngrtd_last_error()); return 1; }
    const int32_t slots[1] = {NGRTD_P_TAU1};
    rc = ngrtd_forward_host(plan, theta, B, 1, slots, out);
    if (rc) { fprintf(stderr, "forward: %s\n", ngrtd_last_error()); return 1; }
    const double obs[1] = {30.0}, sd[1] = {1.5};
    rc = ngrtd_forward_loglik_host(plan, theta, B, 1, slots, NGRTD_LIK_NORMAL, obs, sd, NULL, logp, NULL);
    if (rc) { fprintf(stderr, "loglik: %s\n", ngrtd_last_error()); return 1; }
    double worst = 0.0, worst_lp = 0.0;
    for (int b = 0; b < B; b++) {
        double s = 0.0, c = 0.0;
        for (int k = 0; k < L; k++) {
            double tp = k == 0 ? 1e-5 : (double)k;                                           /* gen_g_tp lag grid, :168-173 */
            double g = exp(-tp / theta[b]);
            s += g;
            c += g * series[k];
        }
        c /= s;
        double e = fabs(out[b] - c) / fabs(c);
        if (e > worst) worst = e;
        double z = (obs[0] - c) / sd[0], lp = -0.5 * log(2.0 * 3.14159265358979323846 * sd[0] * sd[0]) - 0.5 * z * z;
        e = fabs(logp[b] - lp) / fmax(fabs(lp), 1.0);
        if (e > worst_lp) worst_lp = e;
    }
    printf("libngrtd %d: %d chains x %d lags, worst relative error vs the direct sum: model %.2e, logp %.2e\n",
           ngrtd_version(), B, L, worst, worst_lp);
    ngrtd_plan_destroy(plan);
    return (worst < 1e-10 && worst_lp < 1e-9) ? 0 : 2;
}
