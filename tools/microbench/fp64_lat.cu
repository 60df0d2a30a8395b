// FP64 latency / operand-pattern microbenchmark for B200: dependent-chain latency of DFMA, DMUL and DMMA.8x8x4, and
// DFMA issue cost with register vs constant operands.  One CTA per SM, `warps` warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// MODE 0: NCH independent chains, acc = fma(acc, ca, cb)      (constant operands)
// MODE 1: NCH independent chains, acc = fma(acc, x[i], y[i])  (three register operands)
// MODE 2: NCH independent DMMA accumulate chains
// MODE 3: NCH chains of [DMUL -> DMMA]: v *= r; dmma(c, v, b)  (the geometric-weight loop)
// MODE 4: NCH chains of DMUL only (2 register operands)
template <int MODE, int NCH>
__global__ void k(double* out, double ca, double cb) {
    double acc[NCH], x[NCH], y[NCH], c0[NCH], c1[NCH];
    for (int i = 0; i < NCH; i++) { acc[i] = 1.0 + threadIdx.x * 1e-3 + i; x[i] = 0.999 + i * 1e-4 + threadIdx.x * 1e-7; y[i] = 1e-7 * (i + 1) + threadIdx.x * 1e-12; c0[i] = c1[i] = 0; }
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            if (MODE == 0) acc[i] = fma(acc[i], ca, cb);
            if (MODE == 1) acc[i] = fma(acc[i], x[i], y[i]);
            if (MODE == 2) dmma(c0[i], c1[i], x[i], y[i]);
            if (MODE == 3) { acc[i] *= x[i]; dmma(c0[i], c1[i], acc[i], y[i]); }
            if (MODE == 4) acc[i] *= x[i];
        }
    }
    double s = 0;
    for (int i = 0; i < NCH; i++) s += acc[i] + c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize(); float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}
template <int MODE, int NCH> void run(const char* name, double* out) {
    for (int wps = 1; wps <= 4; wps *= 2) {
        float ms = timeit([&] { k<MODE, NCH><<<148, 128 * wps>>>(out, 0.999, 1e-7); });
        double cyc = ms * 1e-3 * 1.92e9 / ITERS;     // cycles per loop iteration (all warps of an SMSP run concurrently)
        printf("%-34s NCH=%d warps/SMSP=%d  %.1f cycles per iteration per warp  = %.2f per chain-step per SMSP\n", name, NCH, wps, cyc, cyc / (NCH * wps));
    }
}
int main() {
    double* out; cudaMalloc(&out, 8 * 148 * 1024);
    run<0, 1>("DFMA const operands", out);
    run<1, 1>("DFMA 3 register operands", out);
    run<4, 1>("DMUL 2 register operands", out);
    run<2, 1>("DMMA accumulate chain", out);
    run<3, 1>("DMUL -> DMMA chain", out);
    run<0, 8>("DFMA const operands", out);
    run<1, 8>("DFMA 3 register operands", out);
    run<4, 8>("DMUL 2 register operands", out);
    run<2, 4>("DMMA accumulate chain", out);
    run<3, 2>("DMUL -> DMMA chain", out);
    run<3, 4>("DMUL -> DMMA chain", out);
    return 0;
}
