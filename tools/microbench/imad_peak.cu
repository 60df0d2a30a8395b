// Integer-multiply microbenchmark: warp-instruction issue cost of IMAD / IMAD.HI.U32 / IMAD.WIDE.U32 on B200, alone
// and next to a DFMA stream (feeds the exp_scaled_fx experiment: can the polynomial move off the FP64 pipe?).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
template <int NF, int NI, int MODE>
__global__ void k(double* out, double a, double b, unsigned int c) {
    double acc[NF > 0 ? NF : 1];
    unsigned int x[NI > 0 ? NI : 1];
    unsigned long long y[NI > 0 ? NI : 1];
    for (int i = 0; i < NF; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < NI; i++) { x[i] = threadIdx.x * 2654435761u + i; y[i] = x[i]; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NF; i++) acc[i] = fma(acc[i], a, b);
#pragma unroll
        for (int i = 0; i < NI; i++) {
            if (MODE == 0) x[i] = x[i] * c + 12345u;                                   // IMAD
            if (MODE == 1) x[i] = __umulhi(x[i], c) + 0x9e3779b9u;                     // IMAD.HI.U32
            if (MODE == 2) y[i] = (unsigned long long)(unsigned int)y[i] * c + y[i];   // IMAD.WIDE.U32
            if (MODE == 3) x[i] = (x[i] ^ c) + (x[i] >> 3);                            // alu pipe (LOP3/SHF/IADD3)
        }
    }
    double s = 0;
    for (int i = 0; i < NF; i++) s += acc[i];
    for (int i = 0; i < NI; i++) s += (double)x[i] + (double)y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize(); float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}
int main() {
    double* out; cudaMalloc(&out, 8 * 148 * 8 * 256);
    int blocks = 148 * 4, threads = 256; double lanes = (double)blocks * threads;
    double cyc = 1.92e9;
    auto rep = [&](const char* name, float ms, int n) {
        double iters_per_smsp = lanes / 32 * ITERS / (148 * 4);
        double c = ms * 1e-3 * cyc / iters_per_smsp;
        printf("%-28s %.3f ms  %.1f cycles per iteration per SMSP  (%.2f per instr of the varied kind)\n", name, ms, c, c / n);
    };
    unsigned int c = 0x9e3779b1u;
    rep("imad x8", timeit([&] { k<0, 8, 0><<<blocks, threads>>>(out, 0.999, 1e-7, c); }), 8);
    rep("imad.hi x8", timeit([&] { k<0, 8, 1><<<blocks, threads>>>(out, 0.999, 1e-7, c); }), 8);
    rep("imad.wide x8", timeit([&] { k<0, 8, 2><<<blocks, threads>>>(out, 0.999, 1e-7, c); }), 8);
    rep("alu x8 (3 ops each)", timeit([&] { k<0, 8, 3><<<blocks, threads>>>(out, 0.999, 1e-7, c); }), 8);
    rep("dfma8", timeit([&] { k<8, 0, 0><<<blocks, threads>>>(out, 0.999, 1e-7, c); }), 8);
    rep("dfma8 + imad x8", timeit([&] { k<8, 8, 0><<<blocks, threads>>>(out, 0.999, 1e-7, c); }), 8);
    rep("dfma8 + imad.hi x4", timeit([&] { k<8, 4, 1><<<blocks, threads>>>(out, 0.999, 1e-7, c); }), 4);
    rep("dfma8 + imad.hi x8", timeit([&] { k<8, 8, 1><<<blocks, threads>>>(out, 0.999, 1e-7, c); }), 8);
    rep("dfma8 + imad.wide x4", timeit([&] { k<8, 4, 2><<<blocks, threads>>>(out, 0.999, 1e-7, c); }), 4);
    rep("dfma8 + imad.wide x8", timeit([&] { k<8, 8, 2><<<blocks, threads>>>(out, 0.999, 1e-7, c); }), 8);
    rep("dfma8 + alu x8", timeit([&] { k<8, 8, 3><<<blocks, threads>>>(out, 0.999, 1e-7, c); }), 8);
    return 0;
}
