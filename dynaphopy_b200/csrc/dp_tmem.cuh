// Tensor memory (TMEM, 256 KB per SM: 128 lanes x 512 columns of 32 bits) used as a per-thread parking area for
// fp64 values: tcgen05.alloc / tcgen05.st / tcgen05.ld / tcgen05.dealloc (SASS: STTM / LDTM).  A thread can only get
// back what it stored itself (shape 32x32b: lane i of a warp addresses TMEM lane 32 * (warp % 4) + i), which is all the
// FFT spectrum needs: |X_a|^2 of a pair's first window waits here for the second window instead of travelling to L2
// and back (dp_spectrum_fft2.cuh).  Under -DDP_EMUL (tests only) a per-CTA array stands in.
#pragma once

#include "dp_common.cuh"

namespace dp {

#ifdef DP_EMUL
inline double* emul_tmem() { static thread_local double t[128 * 256]; return t; }
#endif

// one warp (all lanes) allocates `cols` columns (power of two, 32..512) and publishes the base address in *slot
// (shared memory); tmem_alloc_done() by ALL threads afterwards (includes the CTA barrier)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
#ifdef DP_EMUL
    *slot = 0; (void)cols;
#else
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                 ::"r"((uint32_t)__cvta_generic_to_shared(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
#endif
}
__device__ __forceinline__ void tmem_alloc_done() {
#ifndef DP_EMUL
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
#endif
    __syncthreads();
#ifndef DP_EMUL
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#endif
}
// the allocating warp frees the columns; every thread's TMEM accesses must be complete (a CTA barrier before)
__device__ __forceinline__ void tmem_free(uint32_t base, uint32_t cols) {
#ifndef DP_EMUL
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base), "r"(cols) : "memory");
#else
    (void)base; (void)cols;
#endif
}

// 16 doubles of this thread <-> 32 consecutive columns starting at `col` of the thread's own TMEM lane
__device__ __forceinline__ void tmem_park16(uint32_t base, int warp, int lane, int col, const double (&v)[16]) {
#ifdef DP_EMUL
    double* t = emul_tmem() + ((size_t)((warp & 3) * 32 + lane) * 256 + col / 2);
    (void)base;
    for (int i = 0; i < 16; i++) t[i] = v[i];
#else
    (void)lane;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 16; i++) { r[2 * i] = (uint32_t)__double2loint(v[i]); r[2 * i + 1] = (uint32_t)__double2hiint(v[i]); }
    const uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31]) : "memory");
#endif
}
__device__ __forceinline__ void tmem_park_wait() {         // the thread's parked values are in place
#ifndef DP_EMUL
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
#endif
}
// issue the read-back; the values may be used after tmem_fetch_wait()
__device__ __forceinline__ void tmem_fetch16(uint32_t base, int warp, int lane, int col, uint32_t (&r)[32]) {
#ifdef DP_EMUL
    const double* t = emul_tmem() + ((size_t)((warp & 3) * 32 + lane) * 256 + col / 2);
    (void)base;
    memcpy(r, t, 16 * sizeof(double));
#else
    (void)lane;
    const uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
#endif
}
__device__ __forceinline__ void tmem_fetch_wait() {
#ifndef DP_EMUL
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#endif
}
__device__ __forceinline__ double tmem_value(const uint32_t (&r)[32], int i) {
#ifdef DP_EMUL
    double d;
    memcpy(&d, &r[2 * i], sizeof(double));
    return d;
#else
    return __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
#endif
}

}  // namespace dp
