// Micro-benchmark: per-SM bandwidth to an L2-resident private slab (the FFT spectrum kernel's working set):
// 148 CTAs x 256 threads, each CTA streams its own 512 KB slab; row pattern (256-byte runs) and column
// pattern (64-byte segments at 4 KB stride, as the column passes of the four-step FFT touch it).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(double2* base, int mode, int iters, double2* sink) {
    double2* A = base + (size_t)blockIdx.x * 32768;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double2 acc = make_double2(0, 0);
    for (int it = 0; it < iters; it++) {
        if (mode == 0) {            // row reads: 16 x LDG.128 per thread, fully coalesced
            for (int r = warp; r < 64; r += 8) {
                double2 v[16];
#pragma unroll
                for (int a = 0; a < 16; a++) v[a] = A[r * 512 + a * 32 + lane];
#pragma unroll
                for (int a = 0; a < 16; a++) { acc.x += v[a].x; acc.y += v[a].y; }
            }
        } else if (mode == 1) {     // row writes
            for (int r = warp; r < 64; r += 8) {
#pragma unroll
                for (int a = 0; a < 16; a++) A[r * 512 + a * 32 + lane] = make_double2(it + a, lane);
            }
        } else if (mode == 2) {     // column reads: lane = (g: 8 rows) x (seq: 4 columns), 16 rows apart per register
            for (int t = warp; t < 64; t += 8) {
                const int seq = lane & 3, g = lane >> 2;
                double2 v[16];
#pragma unroll
                for (int a = 0; a < 16; a++) v[a] = A[(size_t)(a * 8 + g) * 256 + t * 4 + seq];
#pragma unroll
                for (int a = 0; a < 16; a++) { acc.x += v[a].x; acc.y += v[a].y; }
            }
        } else {                    // column writes
            for (int t = warp; t < 64; t += 8) {
                const int seq = lane & 3, g = lane >> 2;
#pragma unroll
                for (int a = 0; a < 16; a++) A[(size_t)(a * 8 + g) * 256 + t * 4 + seq] = make_double2(it + a, lane);
            }
        }
    }
    if (acc.x == 1.2345) sink[0] = acc;
}
int main() {
    double2* base; cudaMalloc(&base, (size_t)148 * 32768 * 16 + 1024);
    cudaMemset(base, 0, (size_t)148 * 32768 * 16);
    const char* names[4] = {"row reads ", "row writes", "col reads ", "col writes"};
    for (int mode = 0; mode < 4; mode++) {
        k<<<148, 256>>>(base, mode, 5, base);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int iters = 200;
        cudaEventRecord(e0);
        k<<<148, 256>>>(base, mode, iters, base);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double bytes = 148.0 * 32768 * 16 * iters;
        printf("%s: %.2f TB/s chip, %.1f B/clk/SM at 1.965 GHz\n", names[mode], bytes / ms / 1e9, bytes / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
