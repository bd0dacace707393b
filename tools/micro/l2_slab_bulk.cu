// Micro-benchmark: per-SM bandwidth to an L2-resident private slab through the bulk-copy (TMA) path,
// beside the LSU path of l2_slab_bw.cu.  148 CTAs, each owns a 512 KB slab (128 rows x 256 complex128).
//   mode 0  row reads   cp.async.bulk global -> shared, 4 KB per copy, NBUF copies in flight per warp group
//   mode 1  row writes  cp.async.bulk shared -> global, 4 KB per copy
//   mode 2  col reads   cp.async.bulk.tensor.2d {8 doubles x 128 rows} boxes (64-byte segments at 4 KB stride)
//   mode 3  col writes  cp.async.bulk.tensor.2d store of the same boxes
//   mode 4  row reads + row writes concurrently (two warps issue)
//   mode 5/6  LSU row writes / column writes with `threads` threads (occupancy check of the 27 B/clk figure)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_slab_bulk l2_slab_bulk.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* b) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, const void* src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}

constexpr int NBUF = 8;          // 8 x 8 KB staging buffers
constexpr int TILE = 8192;       // bytes per copy (two 4 KB rows, or one 8-double x 128-row box)

__global__ void __launch_bounds__(256) kbulk(double2* base, const __grid_constant__ CUtensorMap tm, int mode, int iters) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[NBUF];
    double2* A = base + (size_t)blockIdx.x * 32768;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NBUF; i++) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int ntile = 32768 * 16 / TILE;     // 64 tiles per slab sweep
    if (mode == 0 || mode == 2) {
        // one thread drives NBUF loads in flight; a consumer is not modelled (pure transport rate)
        if (tid == 0) {
            unsigned phase = 0;
            int issued = 0, total = ntile * iters;
            for (int i = 0; i < NBUF && issued < total; i++, issued++) {
                mbar_expect(&bars[i], TILE);
                const int t = issued % ntile;
                if (mode == 0) bulk_g2s(smem + i * TILE, (const char*)A + (size_t)t * TILE, TILE, &bars[i]);
                else tma_load_2d(smem + i * TILE, &tm, t * 8, blockIdx.x * 128, &bars[i]);
            }
            for (int done = 0; done < total; done++) {
                const int i = done % NBUF;
                mbar_wait(&bars[i], phase);
                if (i == NBUF - 1) phase ^= 1;
                if (issued < total) {
                    mbar_expect(&bars[i], TILE);
                    const int t = issued % ntile;
                    if (mode == 0) bulk_g2s(smem + i * TILE, (const char*)A + (size_t)t * TILE, TILE, &bars[i]);
                    else tma_load_2d(smem + i * TILE, &tm, t * 8, blockIdx.x * 128, &bars[i]);
                    issued++;
                }
            }
        }
    } else if (mode == 1 || mode == 3) {
        for (int i = tid; i < NBUF * TILE / 16; i += blockDim.x) ((double2*)smem)[i] = make_double2(i, tid);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const int total = ntile * iters;
            for (int n = 0; n < total; n++) {
                const int i = n % NBUF, t = n % ntile;
                if (mode == 1) bulk_s2g((char*)A + (size_t)t * TILE, smem + i * TILE, TILE);
                else tma_store_2d(&tm, t * 8, blockIdx.x * 128, smem + i * TILE);
                bulk_commit();
                bulk_wait_read<NBUF - 1>();       // the staging buffer may be refilled once its reads are done
            }
            bulk_wait<0>();
        }
    } else if (mode == 4) {
        for (int i = tid; i < NBUF * TILE / 16; i += blockDim.x) ((double2*)smem)[i] = make_double2(i, tid);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();
        const int total = ntile * iters;
        if (tid == 0) {             // loads into buffers 0..3
            unsigned phase = 0;
            int issued = 0;
            for (int i = 0; i < 4 && issued < total; i++, issued++) {
                mbar_expect(&bars[i], TILE);
                bulk_g2s(smem + i * TILE, (const char*)A + (size_t)(issued % ntile) * TILE, TILE, &bars[i]);
            }
            for (int done = 0; done < total; done++) {
                const int i = done % 4;
                mbar_wait(&bars[i], phase);
                if (i == 3) phase ^= 1;
                if (issued < total) {
                    mbar_expect(&bars[i], TILE);
                    bulk_g2s(smem + i * TILE, (const char*)A + (size_t)(issued % ntile) * TILE, TILE, &bars[i]);
                    issued++;
                }
            }
        } else if (tid == 32) {     // stores from buffers 4..7 into the other half of the sweep
            for (int n = 0; n < total; n++) {
                const int i = 4 + n % 4, t = (n + ntile / 2) % ntile;
                bulk_s2g((char*)A + (size_t)t * TILE, smem + i * TILE, TILE);
                bulk_commit();
                bulk_wait_read<3>();
            }
            bulk_wait<0>();
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024) klsu(double2* base, int mode, int iters) {
    double2* A = base + (size_t)blockIdx.x * 32768;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    for (int it = 0; it < iters; it++) {
        if (mode == 5) {
            for (int r = warp; r < 64; r += nwarps) {
#pragma unroll
                for (int a = 0; a < 16; a++) A[r * 512 + a * 32 + lane] = make_double2(it + a, lane);
            }
        } else {
            for (int t = warp; t < 64; t += nwarps) {
                const int seq = lane & 3, g = lane >> 2;
#pragma unroll
                for (int a = 0; a < 16; a++) A[(size_t)(a * 8 + g) * 256 + t * 4 + seq] = make_double2(it + a, lane);
            }
        }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    double2* base;
    CK(cudaMalloc(&base, (size_t)148 * 32768 * 16 + 1024));
    CK(cudaMemset(base, 0, (size_t)148 * 32768 * 16));
    // tensor map: all slabs as one (148*128 rows) x (512 doubles) matrix, box = 8 doubles x 128 rows
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap tm;
    cuuint64_t gdim[2] = {512, (cuuint64_t)148 * 128};
    cuuint64_t gstr[1] = {4096};
    cuuint32_t box[2] = {8, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
    const size_t smem = NBUF * TILE;
    CK(cudaFuncSetAttribute(kbulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const char* names[5] = {"bulk row reads        ", "bulk row writes       ", "tensor col reads      ", "tensor col writes     ", "bulk row reads+writes "};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int mode = 0; mode < 5; mode++) {
        kbulk<<<148, 256, smem>>>(base, tm, mode, 3);
        CK(cudaDeviceSynchronize());
        const int iters = 100;
        CK(cudaEventRecord(e0));
        kbulk<<<148, 256, smem>>>(base, tm, mode, iters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double bytes = 148.0 * 32768 * 16 * iters * (mode == 4 ? 2 : 1);
        printf("%s: %.2f TB/s chip, %.1f B/clk/SM at 1.965 GHz\n", names[mode], bytes / ms / 1e9, bytes / (ms * 1e-3) / 148 / 1.965e9);
    }
    for (int mode = 5; mode < 7; mode++)
        for (int threads = 256; threads <= 1024; threads *= 2) {
            klsu<<<148, threads>>>(base, mode, 3);
            CK(cudaDeviceSynchronize());
            const int iters = 100;
            CK(cudaEventRecord(e0));
            klsu<<<148, threads>>>(base, mode, iters);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            double bytes = 148.0 * 32768 * 16 * iters;
            printf("LSU %s writes, %4d threads: %.2f TB/s chip, %.1f B/clk/SM at 1.965 GHz\n", mode == 5 ? "row" : "col", threads,
                   bytes / ms / 1e9, bytes / (ms * 1e-3) / 148 / 1.965e9);
        }
    return 0;
}
