// Micro-benchmark: fp64 throughput of butterfly-like code (many distinct register operands) vs the
// ideal 64 lanes/clk/SM: does register-file banking limit DADD/DMUL/DFMA streams of an FFT codelet?
#include <cstdio>
#include <cuda_runtime.h>
template <int N>
__global__ void k(double* out, int iters, double c, double s) {
    double re[N], im[N];
#pragma unroll
    for (int i = 0; i < N; i++) { re[i] = threadIdx.x + i; im[i] = threadIdx.x - i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int st = 1; st < N; st <<= 1) {
#pragma unroll
            for (int i = 0; i < N; i++) {
                if ((i & st) == 0) {
                    const int j = i | st;
                    // twiddle multiply on the odd input (2 DMUL + 2 DFMA), then butterfly (4 DADD)
                    const double tr = re[j] * c - im[j] * s, ti = re[j] * s + im[j] * c;
                    const double ar = re[i], ai = im[i];
                    re[i] = ar + tr; im[i] = ai + ti; re[j] = ar - tr; im[j] = ai - ti;
                }
            }
        }
    }
    double acc = 0;
#pragma unroll
    for (int i = 0; i < N; i++) acc += re[i] + im[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int N> void run(int threads) {
    double* out; cudaMalloc(&out, 148 * 1024 * 8);
    const int iters = 2000;
    k<N><<<148, threads>>>(out, 10, 0.9999, 0.01);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<N><<<148, threads>>>(out, iters, 0.9999, 0.01);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int stages = 0; for (int st = 1; st < N; st <<= 1) stages++;
    double ops = 148.0 * threads * (double)iters * stages * (N / 2) * 8.0;     // 8 fp64 instructions per butterfly
    printf("N=%d threads=%d: %.2f T fp64 instr-lanes/s  (%.1f lanes/clk/SM at 1.965 GHz)\n", N, threads, ops / ms / 1e9,
           ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main() {
    for (int th : {128, 256, 512}) { run<8>(th); run<16>(th); }
    return 0;
}
