#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* p, int mode) {
    double* q = p + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 12;   // stride 96 B
    double a = threadIdx.x + 0.25, b = threadIdx.x + 0.5, c = threadIdx.x + 0.75, d = threadIdx.x + 1000.0;
    if (mode == 0) asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(q), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
    else asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(q), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
int main() {
    double* p; cudaMalloc(&p, 64 * 12 * 8); 
    for (int mode = 0; mode < 2; mode++) {
        cudaMemset(p, 0, 64 * 12 * 8);
        k<<<1, 64>>>(p, mode);
        double h[64 * 12];
        cudaMemcpy(h, p, sizeof(h), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int t = 0; t < 64; t++) {
            double e[4] = {t + 0.25, t + 0.5, t + 0.75, t + 1000.0};
            for (int i = 0; i < 4; i++) if (h[t * 12 + i] != e[i]) bad++;
            for (int i = 4; i < 12; i++) if (h[t * 12 + i] != 0.0) bad++;
        }
        printf("mode %d: bad=%d  first: %g %g %g %g | %g\n", mode, bad, h[0], h[1], h[2], h[3], h[4]);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
