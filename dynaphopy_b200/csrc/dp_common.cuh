// Shared helpers for the dynaphopy_b200 CUDA kernels (sm_100a).
#pragma once

#ifdef DP_EMUL
#include "dp_emul.h"
#else
#include <cuda_runtime.h>
#endif

#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>

#include "../../include/dynaphopy_b200.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace dp {

// ---- error plumbing ------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define DP_CUDA_OK(expr)                                            \
    do {                                                            \
        cudaError_t dp_e_ = (expr);                                 \
        if (dp_e_ != cudaSuccess) return dp::cuda_fail(dp_e_, #expr); \
    } while (0)

#define DP_REQUIRE(cond, code, ...)          \
    do {                                     \
        if (!(cond)) {                       \
            dp::set_error(__VA_ARGS__);      \
            return (code);                   \
        }                                    \
    } while (0)

// ---- launch / dynamic shared memory, identical source for nvcc and the test emulator ---
// every kernel launch of the library is counted (dp_launch_count), so benchmarks can report how
// many of OUR kernels ran inside a timed region
void count_launch();

#ifdef DP_EMUL
#define DP_LAUNCH(kfn, grid, block, smem, stream, ...) \
    (dp::count_launch(), dp_emul::launch((grid), (block), (size_t)(smem), [=]() { kfn(__VA_ARGS__); }))
#define DP_DYNSMEM(type, name) type* name = reinterpret_cast<type*>(dp_emul::tls_worker()->dyn_smem)
#define DP_SET_SMEM(kfn, bytes) cudaSuccess
#define DP_LAUNCH_PDL(kfn, grid, block, smem, stream, ...) (DP_LAUNCH(kfn, grid, block, smem, stream, __VA_ARGS__), cudaSuccess)
#else
#define DP_LAUNCH(kfn, grid, block, smem, stream, ...) \
    (dp::count_launch(), kfn<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__))
#define DP_DYNSMEM(type, name)                                   \
    extern __shared__ __align__(128) unsigned char name##_raw_[]; \
    type* name = reinterpret_cast<type*>(name##_raw_)
#define DP_SET_SMEM(kfn, bytes) \
    cudaFuncSetAttribute((const void*)(kfn), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
// Programmatic dependent launch: the grid may become resident while the tail of the previous kernel of the stream is
// still running; the kernel calls pdl_wait() before it touches anything an earlier kernel wrote.
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, std::forward<Args>(args)...);
}
#define DP_LAUNCH_PDL(kfn, grid, block, smem, stream, ...) \
    (dp::count_launch(), dp::launch_pdl(kfn, (grid), (block), (size_t)(smem), (stream), __VA_ARGS__))
#endif

// cudaFuncSetAttribute once per kernel and device instead of once per launch (flags: one static array per call site)
#define DP_SET_SMEM_ONCE(kfn, bytes)                                                    \
    do {                                                                                \
        static bool dp_once_[64] = {};                                                  \
        int dp_dev_ = 0;                                                                \
        DP_CUDA_OK(cudaGetDevice(&dp_dev_));                                            \
        if (dp_dev_ < 0 || dp_dev_ >= 64 || !dp_once_[dp_dev_]) {                       \
            DP_CUDA_OK(DP_SET_SMEM(kfn, bytes));                                        \
            if (dp_dev_ >= 0 && dp_dev_ < 64) dp_once_[dp_dev_] = true;                 \
        }                                                                               \
    } while (0)

// Inside a kernel launched with DP_LAUNCH_PDL: wait until every earlier kernel of the stream has completed and its
// writes are visible, then let the next kernel of the stream become resident (it waits in turn, so by induction a grid
// never starts before the grid two launches earlier has completed).  No-ops under a plain launch.
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// ---- small device helpers ---------------------------------------------------------------
template <typename T> __host__ __device__ __forceinline__ T imin(T a, T b) { return a < b ? a : b; }
template <typename T> __host__ __device__ __forceinline__ T imax(T a, T b) { return a > b ? a : b; }

typedef double2 cplx;   // (re, im)

__host__ __device__ __forceinline__ cplx cmake(double r, double i) { return make_double2(r, i); }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx cmulc(cplx a, cplx b) {   // a * conj(b)
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ cplx cmul_i(cplx a) { return make_double2(-a.y, a.x); }      //  i*a
__host__ __device__ __forceinline__ cplx cmul_mi(cplx a) { return make_double2(a.y, -a.x); }     // -i*a
__host__ __device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }

// exp(-2*pi*i * num/den), sign = -1 forward
__device__ __forceinline__ cplx root_of_unity(long long num, long long den, int sign) {
    double s, c;
    long long r = num % den;
    sincospi(2.0 * (double)r / (double)den, &s, &c);
    return make_double2(c, sign < 0 ? -s : s);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// A set of complex128 time series in device memory: sample t of series s is at
//   base[s*stride_s + (t / tb_len)*tb_stride + (t % tb_len)*stride_t]   (tb_len > 0: the time axis arrives in blocks,
//   e.g. one block per source rank and chunk after the time-sharded -> series-sharded exchange), or
//   base[s*stride_s + t*stride_t]                                       (tb_len == 0).
// (T, ld) time-major arrays of the reference are stride_s = 1, stride_t = ld; series-major is stride_t = 1.
struct SeriesView {
    const cplx* base;
    int64_t stride_s, stride_t, tb_stride;
    int tb_len;
    __host__ __device__ __forceinline__ const cplx* at(int64_t s, int64_t t) const {
        if (tb_len == 0) return base + s * stride_s + t * stride_t;
        const int64_t b = t / tb_len;
        return base + s * stride_s + b * tb_stride + (t - b * tb_len) * stride_t;
    }
};

int sm_count();
// small host array -> device through kernel parameters (never blocks the host; see dp_core.cu)
int upload_small(void* dst, const void* src, size_t bytes, cudaStream_t st);
// host table of any size -> device without blocking the host when it is small (<= 64 KB: kernel parameters); larger
// tables fall back to one pageable cudaMemcpyAsync (staged before it returns, may wait for earlier work of the stream)
int upload_table(void* dst, const void* src, size_t bytes, cudaStream_t st);
// Development knobs (skip masks, forced engines, grid overrides) read the environment only in -DDP_DEVELOP builds:
// a stray variable can never change what the shipped library computes.
inline const char* dev_env(const char* name) {
#ifdef DP_DEVELOP
    return getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}

}  // namespace dp
