// Micro-benchmark: fp64 FMA / ADD / MUL issue rate per SM as a function of resident warps and ILP.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int ILP>
__global__ void k(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = a + i + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) x[i] = fma(x[i], a, b);
            else if (OP == 1) x[i] = x[i] + b;
            else x[i] = x[i] * a;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP, int ILP> void run(const char* name, int threads) {
    double* out; cudaMalloc(&out, 148 * 1024 * 8);
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP, ILP><<<148, threads>>>(out, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<OP, ILP><<<148, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * threads * (double)iters * ILP;
    printf("%s ILP=%d threads=%d: %.2f T op/s  (%.1f lane-ops/clk/SM at 1.965 GHz)\n", name, ILP, threads, ops / ms / 1e9,
           ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main() {
    for (int th : {128, 256, 384, 512, 1024}) {
        if (th == 128) { run<0, 1>("DFMA", th); run<0, 4>("DFMA", th); run<0, 8>("DFMA", th); run<1, 8>("DADD", th); run<2, 8>("DMUL", th); }
        if (th == 256) { run<0, 1>("DFMA", th); run<0, 4>("DFMA", th); run<0, 8>("DFMA", th); run<1, 8>("DADD", th); run<2, 8>("DMUL", th); }
        if (th == 384) { run<0, 1>("DFMA", th); run<0, 4>("DFMA", th); run<0, 8>("DFMA", th); run<1, 8>("DADD", th); run<2, 8>("DMUL", th); }
        if (th == 512) { run<0, 1>("DFMA", th); run<0, 4>("DFMA", th); run<0, 8>("DFMA", th); run<1, 8>("DADD", th); }
        if (th == 1024) { run<0, 1>("DFMA", th); run<0, 4>("DFMA", th); run<0, 8>("DFMA", th); run<1, 8>("DADD", th); }
    }
    return 0;
}
