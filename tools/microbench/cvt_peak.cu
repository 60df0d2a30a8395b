// Conversion-unit microbenchmark: is F2I.F64 / I2F.F64 off the FP64 pipe on B200? (feeds exp_scaled design)
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
template <int NF, int NC, int MODE>
__global__ void k(double* out, double a, double b) {
    double acc[NF > 0 ? NF : 1]; double x[NC > 0 ? NC : 1];
    for (int i = 0; i < NF; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < NC; i++) x[i] = threadIdx.x * 1.37 + i * 11.1;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NF; i++) acc[i] = fma(acc[i], a, b);
#pragma unroll
        for (int i = 0; i < NC; i++) {
            if (MODE == 0) { int n = __double2int_rn(x[i]); x[i] = x[i] * 0.5 + 3.0 - __int2double_rn(n) * 0.25; }       // F2I + I2F (+2 FP64)
            if (MODE == 1) { int n = __double2int_rn(x[i]); x[i] = __hiloint2double(__double2hiint(x[i]) ^ (n & 1), __double2loint(x[i]) + n); }  // F2I only
            if (MODE == 2) { double t = x[i] + 6755399441055744.0; int n = __double2loint(t); x[i] = x[i] * 0.5 + 3.0 - (t - 6755399441055744.0) * 0.25 + (n & 1); } // magic
        }
    }
    double s = 0;
    for (int i = 0; i < NF; i++) s += acc[i];
    for (int i = 0; i < NC; i++) s += x[i];
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
    double cyc = 1.92e9;   // SM clock seen under FP64 load
    auto rep = [&](const char* name, float ms, int nf, int nc) {
        double warp_instr_groups = lanes / 32 * ITERS;   // per loop iteration per warp
        double cycles_per_iter_per_smsp = ms * 1e-3 * cyc / (warp_instr_groups / (148 * 4));
        printf("%-34s %.3f ms  -> %.1f cycles per loop iteration per SMSP-warp (NF=%d DFMA, NC=%d conv groups)\n", name, ms, cycles_per_iter_per_smsp, nf, nc);
    };
    rep("dfma8 only", timeit([&] { k<8, 0, 0><<<blocks, threads>>>(out, 0.999, 1e-7); }), 8, 0);
    rep("f2i+i2f x4 (+2 fp64 each)", timeit([&] { k<0, 4, 0><<<blocks, threads>>>(out, 0.999, 1e-7); }), 0, 4);
    rep("f2i only x4", timeit([&] { k<0, 4, 1><<<blocks, threads>>>(out, 0.999, 1e-7); }), 0, 4);
    rep("magic x4 (4 fp64 each)", timeit([&] { k<0, 4, 2><<<blocks, threads>>>(out, 0.999, 1e-7); }), 0, 4);
    rep("dfma8 + f2i+i2f x1", timeit([&] { k<8, 1, 0><<<blocks, threads>>>(out, 0.999, 1e-7); }), 8, 1);
    rep("dfma8 + f2i+i2f x2", timeit([&] { k<8, 2, 0><<<blocks, threads>>>(out, 0.999, 1e-7); }), 8, 2);
    rep("dfma8 + f2i x2", timeit([&] { k<8, 2, 1><<<blocks, threads>>>(out, 0.999, 1e-7); }), 8, 2);
    rep("dfma8 + magic x2", timeit([&] { k<8, 2, 2><<<blocks, threads>>>(out, 0.999, 1e-7); }), 8, 2);
    rep("dfma16 + f2i+i2f x1", timeit([&] { k<16, 1, 0><<<blocks, threads>>>(out, 0.999, 1e-7); }), 16, 1);
    return 0;
}
