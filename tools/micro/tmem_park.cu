// Micro-test: tensor memory (TMEM) as a per-thread parking area -- tcgen05.alloc / st / ld / dealloc round trip.
// 12 warps; warp w uses the lane quarter w % 4; the three warps of a quarter split its 512 columns.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const double (&v)[16]) {
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 16; i++) { r[2 * i] = (uint32_t)__double2loint(v[i]); r[2 * i + 1] = (uint32_t)__double2hiint(v[i]); }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, double (&v)[16]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}

__global__ void __launch_bounds__(384, 1) k(int* bad, int iters) {
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t base = tbase;
    // quarter q = warp % 4; warps q, q+4, q+8 own 6, 5, 5 tiles of 32 columns
    const int q = warp & 3, r = warp >> 2;
    const int ntile = r == 0 ? 6 : 5, col0 = r == 0 ? 0 : (r == 1 ? 192 : 352);
    int nbad = 0;
    for (int it = 0; it < iters; it++) {
        for (int t = 0; t < ntile; t++) {
            double v[16];
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = 1.0 / (1.0 + blockIdx.x) + 1e-3 * tid + 1e-6 * (t * 16 + i) + it;
            tmem_st16(base + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 + t * 32), v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        for (int t = 0; t < ntile; t++) {
            double v[16];
            tmem_ld16(base + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 + t * 32), v);
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (v[i] != 1.0 / (1.0 + blockIdx.x) + 1e-3 * tid + 1e-6 * (t * 16 + i) + it) nbad++;
        }
    }
    if (nbad) atomicAdd(bad, nbad);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(base) : "memory");
}

int main() {
    int* bad; CK(cudaMalloc(&bad, 4)); CK(cudaMemset(bad, 0, 4));
    k<<<148, 384>>>(bad, 2);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 2000;
    CK(cudaEventRecord(e0));
    k<<<148, 384>>>(bad, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    int h; CK(cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost));
    double bytes = 148.0 * 256 * 1024 * 2.0 * iters;   // st + ld of the whole TMEM
    printf("TMEM park: mismatches %d; %.2f TB/s chip (st + ld), %.1f B/clk/SM at 1.965 GHz\n", h, bytes / ms / 1e9, bytes / (ms * 1e-3) / 148 / 1.965e9);
    return h != 0;
}
