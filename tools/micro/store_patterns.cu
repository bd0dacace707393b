// Micro-benchmark: what limits the projection epilogue's store stream?  Persistent grid of 296 CTAs x 320 threads
// (2 per SM, as project_fft_kernel), every "unit" writes one contiguous run of 5120 x 32 B = 160 KB to a fresh
// place in a 2 GB buffer.  Variants: store width / cache operator / a __syncthreads per unit / shared-memory reads
// and shuffles in front of the stores (as the real epilogue has them).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void st16(double2* p, double2 v, int op) {
    if (op == 0) *p = v;
    else if (op == 1) __stcs(p, v);
    else __stcg(p, v);
}
__device__ __forceinline__ void st32(double2* p, double2 a, double2 b, int op) {
    if (op == 0) asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
    else asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}

// mode 0: lane stores 16 B, warp 512 B contiguous (two instructions cover 1 KB)
// mode 1: lane stores 2 x 16 B to its own 32 B (the pre-transposition pattern)
// mode 2: lane stores 32 B with one st.v4.f64
// flags: 1 = __syncthreads per unit, 2 = shared-memory reads + arithmetic in front, 4 = shuffles in front
__global__ void __launch_bounds__(320, 2) k(double2* out, int nunits, int mode, int op, int flags, int unroll8) {
    extern __shared__ double2 W[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int Q = 5120;
    for (int i = tid; i < 5440; i += blockDim.x) W[i] = make_double2(i, -i);
    __syncthreads();
    for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
        double2* o = out + (size_t)unit * Q * 2;
        if (flags & 1) __syncthreads();
        for (int q0 = tid; q0 < Q; q0 += 4 * blockDim.x) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int q = q0 + i * blockDim.x;
                double2 ra = make_double2(q, unit), rb = make_double2(unit, q);
                if (flags & 2) {
                    const double2 C = W[q + (q >> 4)], D = W[5439 - q - (q >> 4)];
                    ra = make_double2(0.5 * (C.x + D.x), 0.5 * (C.y - D.y));
                    rb = make_double2(0.5 * (C.y + D.y), 0.5 * (D.x - C.x));
                }
                if (mode == 0) {
                    const int qbase = q - lane;
                    if (flags & 4) {
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int src = h * 16 + (lane >> 1);
                            const double ax = __shfl_sync(0xffffffffu, ra.x, src), ay = __shfl_sync(0xffffffffu, ra.y, src);
                            const double bx = __shfl_sync(0xffffffffu, rb.x, src), by = __shfl_sync(0xffffffffu, rb.y, src);
                            st16(o + (size_t)(qbase + h * 16) * 2 + lane, (lane & 1) ? make_double2(bx, by) : make_double2(ax, ay), op);
                        }
                    } else {
                        st16(o + (size_t)qbase * 2 + lane, ra, op);
                        st16(o + (size_t)qbase * 2 + 32 + lane, rb, op);
                    }
                } else if (mode == 1) {
                    st16(o + (size_t)q * 2, ra, op);
                    st16(o + (size_t)q * 2 + 1, rb, op);
                } else {
                    st32(o + (size_t)q * 2, ra, rb, op);
                }
            }
        }
    }
    (void)unroll8;
}

int main() {
    const int nunits = 4096 * 3;
    const size_t bytes = (size_t)nunits * 5120 * 32;
    double2* out; cudaMalloc(&out, bytes);
    cudaMemset(out, 0, bytes);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 107 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    {   // reference points: cudaMemset and a device-to-device copy of the same size
        float ms;
        cudaEventRecord(e0); cudaMemsetAsync(out, 1, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("cudaMemset %.2f GB: %.3f ms = %.2f TB/s\n", bytes / 1e9, ms, bytes / ms / 1e9);
    }
    const char* mname[3] = {"16B/lane coalesced", "2x16B/lane own 32B", "32B/lane st.v4.f64 "};
    const char* oname[3] = {"plain", ".cs  ", ".cg  "};
    for (int smem_kb = 0; smem_kb < 2; smem_kb++)       // 0: 87 KB (2 CTAs/SM as the kernel), 1: same (placeholder)
        for (int mode = 0; mode < 3; mode++)
            for (int op = 0; op < (mode == 2 ? 2 : 3); op++)
                for (int flags = 0; flags < 8; flags++) {
                    if (smem_kb) continue;
                    if ((flags & 4) && mode != 0) continue;
                    k<<<296, 320, 107 * 1024>>>(out, nunits, mode, op, flags, 0);
                    cudaEventRecord(e0);
                    k<<<296, 320, 107 * 1024>>>(out, nunits, mode, op, flags, 0);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    printf("%s %s sync=%d smem=%d shfl=%d : %.3f ms = %.2f TB/s\n", mname[mode], oname[op], flags & 1,
                           (flags >> 1) & 1, (flags >> 2) & 1, ms, bytes / ms / 1e9);
                }
    // occupancy variant: 148 x 1024 threads plain coalesced
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
