// Bulk asynchronous copies (TMA) and mbarriers for sm_100a: thin wrappers over the PTX instructions
//   cp.async.bulk            (1-D, contiguous runs; SASS UBLKCP)
//   cp.async.bulk.tensor.2d  (boxes of a 2-D tensor described by a CUtensorMap; SASS UTMALDG / UTMASTG)
//   mbarrier.*               (transaction barriers the loads complete on; SASS SYNCS)
// The copies run on the asynchronous proxy: a thread only issues them, the data moves while the warps compute,
// and neither direction occupies the load/store unit.  Ordering rules used by the callers:
//   shared memory written by ordinary stores, then read by a bulk store:  fence_proxy_async() by the writers,
//     a barrier, then the issue;
//   a staging buffer may be refilled after bulk_wait_read<0>() of the issuing thread (+ a barrier);
//   global data written by bulk stores is complete after bulk_wait<0>() of the issuing thread.
// Under -DDP_EMUL (tests only) the copies are carried out immediately by the issuing fiber.
#pragma once

#include "dp_common.cuh"

#ifndef DP_EMUL
#include <cuda.h>
#endif

namespace dp {

#ifdef DP_EMUL
struct TensorMap2D {            // description of the same 2-D tensor / box for the emulated copies
    double* base;
    uint64_t inner, rows, pitch_bytes;
    uint32_t box_inner, box_rows;
};
typedef uint64_t mbar_t;
#else
typedef CUtensorMap TensorMap2D;
typedef uint64_t mbar_t;
#endif

// 2-D tensor of doubles: `rows` rows of `inner` doubles, row pitch `pitch_bytes` (multiple of 16), box = box_inner x
// box_rows (box_inner * 8 a multiple of 16, both <= 256).  Host only; DP_ERR_UNSUPPORTED when the driver cannot encode it.
int make_tensor_map_2d(TensorMap2D* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t pitch_bytes,
                       uint32_t box_inner, uint32_t box_rows);

#ifndef DP_EMUL
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
#endif

__device__ __forceinline__ void mbar_init(mbar_t* b, int count) {
#ifdef DP_EMUL
    *b = 0; (void)count;
#else
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(b)), "r"(count) : "memory");
#endif
}
__device__ __forceinline__ void mbar_fence_init() {
#ifndef DP_EMUL
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
#endif
}
// arrive (count 1) and announce `bytes` of bulk-copy traffic that will complete on the barrier
__device__ __forceinline__ void mbar_expect_tx(mbar_t* b, unsigned bytes) {
#ifdef DP_EMUL
    (void)b; (void)bytes;
#else
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_addr(b)), "r"(bytes) : "memory");
#endif
}
// wait for the phase of the given parity to complete; every lane of the warp calls it (emulation: a warp barrier,
// the issuing fiber has already done the copy)
__device__ __forceinline__ void mbar_wait(mbar_t* b, unsigned parity) {
#ifdef DP_EMUL
    (void)b; (void)parity;
    __syncwarp();
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "DP_MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DP_MBAR_DONE_%=;\n"
        "bra DP_MBAR_WAIT_%=;\n"
        "DP_MBAR_DONE_%=:\n"
        "}\n" ::"r"(smem_addr(b)), "r"(parity) : "memory");
#endif
}

__device__ __forceinline__ void fence_proxy_async() {            // all state spaces
#ifndef DP_EMUL
    asm volatile("fence.proxy.async;\n" ::: "memory");
#endif
}
__device__ __forceinline__ void fence_proxy_async_smem() {       // shared memory of this CTA only
#ifndef DP_EMUL
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
#endif
}

// ---- 1-D bulk copies (16-byte aligned, size a multiple of 16) ----------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, unsigned bytes, mbar_t* b) {
#ifdef DP_EMUL
    memcpy(dst_smem, src, bytes); (void)b;
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_addr(dst_smem)), "l"(src), "r"(bytes), "r"(smem_addr(b)) : "memory");
#endif
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, unsigned bytes) {
#ifdef DP_EMUL
    memcpy(dst, src_smem, bytes);
#else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(smem_addr(src_smem)), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void bulk_commit() {
#ifndef DP_EMUL
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
#endif
}
// the issuing thread's bulk stores have finished READING shared memory, all but the newest N groups
template <int N> __device__ __forceinline__ void bulk_wait_read() {
#ifndef DP_EMUL
    asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
#endif
}
// ... have completed (their global writes are done)
template <int N> __device__ __forceinline__ void bulk_wait() {
#ifndef DP_EMUL
    asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory");
#endif
}

// bring [src, src + bytes) into L2 ahead of its use (bytes a multiple of 16) with evict-first priority: the lines are
// used once, soon, and must not displace longer-lived data
__device__ __forceinline__ void bulk_prefetch_l2_evict_first(const void* src, unsigned bytes) {
#ifndef DP_EMUL
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;\n" ::"l"(src), "r"(bytes), "l"(pol) : "memory");
#else
    (void)src; (void)bytes;
#endif
}

// bring [src, src + bytes) into L2 ahead of its use (bytes a multiple of 16), normal priority
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
#ifndef DP_EMUL
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src), "r"(bytes) : "memory");
#else
    (void)src; (void)bytes;
#endif
}

// ---- 2-D tensor boxes: shared-memory tile is dense, [box_rows][box_inner] doubles ----------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const TensorMap2D* tm, int c_inner, int c_row, mbar_t* b) {
#ifdef DP_EMUL
    (void)b;
    double* d = (double*)dst_smem;
    for (uint32_t r = 0; r < tm->box_rows; r++)
        for (uint32_t i = 0; i < tm->box_inner; i++) {
            const uint64_t rr = (uint64_t)c_row + r, ii = (uint64_t)c_inner + i;
            d[(size_t)r * tm->box_inner + i] = (rr < tm->rows && ii < tm->inner)
                ? *(const double*)((const char*)tm->base + rr * tm->pitch_bytes + ii * 8) : 0.0;
        }
#else
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 ::"r"(smem_addr(dst_smem)), "l"(tm), "r"(c_inner), "r"(c_row), "r"(smem_addr(b)) : "memory");
#endif
}
__device__ __forceinline__ void tma_store_2d(const TensorMap2D* tm, int c_inner, int c_row, const void* src_smem) {
#ifdef DP_EMUL
    const double* s = (const double*)src_smem;
    for (uint32_t r = 0; r < tm->box_rows; r++)
        for (uint32_t i = 0; i < tm->box_inner; i++) {
            const uint64_t rr = (uint64_t)c_row + r, ii = (uint64_t)c_inner + i;
            if (rr < tm->rows && ii < tm->inner)
                *(double*)((char*)tm->base + rr * tm->pitch_bytes + ii * 8) = s[(size_t)r * tm->box_inner + i];
        }
#else
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n"
                 ::"l"(tm), "r"(c_inner), "r"(c_row), "r"(smem_addr(src_smem)) : "memory");
#endif
}

}  // namespace dp
