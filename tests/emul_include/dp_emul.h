// TEST INFRASTRUCTURE ONLY -- host emulation of the CUDA execution model.
//
// The product is the nvcc/sm_100a build of these sources (libdynaphopy_b200.so); it never
// includes this header and has no CPU fallback.  tests/ additionally compile the SAME
// kernel sources with g++ -DDP_EMUL so that index arithmetic, barrier placement and
// warp-shuffle logic can be checked in the GPU-less dev container before spending GPU
// minutes.  Each CUDA thread is a ucontext fiber; a CTA's fibers run round-robin on one OS
// thread and yield at __syncthreads()/__syncwarp()/shuffles; CTAs are distributed over a
// few OS threads.  Deterministic inside a CTA, deadlocks are detected and reported.
#pragma once
#ifndef DP_EMUL
#error "dp_emul.h is for the -DDP_EMUL test build only"
#endif

#include <ucontext.h>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static thread_local
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(16))) double4_emul { double x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost,
                      cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }

namespace dp_emul {

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    int state = 0;             // 0 runnable, 1 waiting block barrier, 2 waiting warp barrier, 3 done
    uint3 tid;
    int linear = 0;
};

struct Worker {
    ucontext_t sched;
    std::vector<Fiber> fibers;
    Fiber* cur = nullptr;
    uint3 bid; dim3 bdim, gdim;
    unsigned char* dyn_smem = nullptr;
    size_t dyn_smem_cap = 0;
    const std::function<void()>* body = nullptr;
    std::vector<uint64_t> warp_buf;   // 32 slots per warp
    int block_wait = 0;
    std::vector<int> warp_wait, warp_alive;
    int alive = 0;
};

inline Worker*& tls_worker() { static thread_local Worker* w = nullptr; return w; }
static const size_t kStack = 256 * 1024;

inline void fiber_entry() {
    Worker* w = tls_worker();
    (*w->body)();
    Fiber* f = w->cur;
    f->state = 3;
    w->alive--;
    w->warp_alive[f->linear / 32]--;
    swapcontext(&f->ctx, &w->sched);
}

inline void yield_wait(int state) {
    Worker* w = tls_worker();
    Fiber* f = w->cur;
    f->state = state;
    if (state == 1) w->block_wait++; else w->warp_wait[f->linear / 32]++;
    swapcontext(&f->ctx, &w->sched);
}

inline void run_block(Worker* w, unsigned bx, unsigned by, unsigned bz, dim3 block) {
    int n = (int)(block.x * block.y * block.z);
    if ((int)w->fibers.size() < n) {
        size_t old = w->fibers.size();
        w->fibers.resize(n);
        for (size_t i = old; i < (size_t)n; i++) w->fibers[i].stack = (char*)std::malloc(kStack);
    }
    int nwarps = (n + 31) / 32;
    w->warp_buf.assign((size_t)nwarps * 32, 0);
    w->warp_wait.assign(nwarps, 0);
    w->warp_alive.assign(nwarps, 0);
    w->block_wait = 0;
    w->alive = n;
    w->bid.x = bx; w->bid.y = by; w->bid.z = bz;
    w->bdim = block;
    for (int i = 0; i < n; i++) {
        Fiber& f = w->fibers[i];
        f.state = 0; f.linear = i;
        f.tid.x = i % block.x; f.tid.y = (i / block.x) % block.y; f.tid.z = i / (block.x * block.y);
        w->warp_alive[i / 32]++;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &w->sched;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    while (w->alive > 0) {
        bool progressed = false;
        for (int i = 0; i < n; i++) {
            Fiber& f = w->fibers[i];
            if (f.state != 0) continue;
            w->cur = &f;
            swapcontext(&w->sched, &f.ctx);
            progressed = true;
        }
        bool released = false;
        for (int wp = 0; wp < nwarps; wp++) {
            if (w->warp_wait[wp] > 0 && w->warp_wait[wp] == w->warp_alive[wp]) {
                for (int i = wp * 32; i < n && i < wp * 32 + 32; i++)
                    if (w->fibers[i].state == 2) w->fibers[i].state = 0;
                w->warp_wait[wp] = 0;
                released = true;
            }
        }
        if (w->block_wait > 0 && w->block_wait == w->alive) {
            for (int i = 0; i < n; i++) if (w->fibers[i].state == 1) w->fibers[i].state = 0;
            w->block_wait = 0;
            released = true;
        }
        if (!progressed && !released && w->alive > 0) {
            std::fprintf(stderr, "dp_emul: DEADLOCK in block (%u,%u,%u): alive=%d block_wait=%d\n",
                         bx, by, bz, w->alive, w->block_wait);
            std::abort();
        }
    }
}

inline int worker_count() {
    const char* e = std::getenv("DP_EMUL_THREADS");
    int n = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
    return n < 1 ? 1 : (n > 16 ? 16 : n);
}

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    if (nblocks == 0) return;
    int nw = (int)std::min<size_t>(worker_count(), nblocks);
    std::atomic<size_t> next(0);
    auto work = [&]() {
        Worker w;
        w.gdim = grid;
        w.body = &body;
        w.dyn_smem_cap = smem + 64;
        w.dyn_smem = (unsigned char*)std::aligned_alloc(64, (w.dyn_smem_cap + 63) / 64 * 64);
        tls_worker() = &w;
        for (;;) {
            size_t b = next.fetch_add(1);
            if (b >= nblocks) break;
            run_block(&w, (unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y),
                      (unsigned)(b / ((size_t)grid.x * grid.y)), block);
        }
        for (auto& f : w.fibers) std::free(f.stack);
        std::free(w.dyn_smem);
        tls_worker() = nullptr;
    };
    if (nw == 1) { work(); return; }
    std::vector<std::thread> th;
    for (int i = 0; i < nw; i++) th.emplace_back(work);
    for (auto& t : th) t.join();
}

inline void warp_barrier() { yield_wait(2); }

template <typename T> inline T shfl_any(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of <= 64-bit values");
    Worker* w = tls_worker();
    Fiber* f = w->cur;
    int warp = f->linear / 32, lane = f->linear % 32;
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    w->warp_buf[(size_t)warp * 32 + lane] = bits;
    warp_barrier();
    int nl = w->warp_alive[warp];
    int s = ((src_lane % 32) + 32) % 32;
    if (s >= nl) s = lane;
    uint64_t got = w->warp_buf[(size_t)warp * 32 + s];
    warp_barrier();
    T r;
    std::memcpy(&r, &got, sizeof(T));
    return r;
}

}  // namespace dp_emul

#define threadIdx (dp_emul::tls_worker()->cur->tid)
#define blockIdx (dp_emul::tls_worker()->bid)
#define blockDim (dp_emul::tls_worker()->bdim)
#define gridDim (dp_emul::tls_worker()->gdim)
static const int warpSize = 32;

static inline void __syncthreads() { dp_emul::yield_wait(1); }
static inline void __syncwarp(unsigned = 0xffffffffu) { dp_emul::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return dp_emul::shfl_any(v, src); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    return dp_emul::shfl_any(v, (dp_emul::tls_worker()->cur->linear % 32) ^ m);
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    int lane = dp_emul::tls_worker()->cur->linear % 32;
    return dp_emul::shfl_any(v, lane + (int)d < 32 ? lane + (int)d : lane);
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    int lane = dp_emul::tls_worker()->cur->linear % 32;
    return dp_emul::shfl_any(v, lane - (int)d >= 0 ? lane - (int)d : lane);
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
template <typename T> static inline void __stcg(T* p, T v) { *p = v; }

static inline double atomicAdd(double* p, double v) {
    static std::mutex m; std::lock_guard<std::mutex> g(m);
    double o = *p; *p = o + v; return o;
}
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicCAS(int* p, int expected, int desired) {
    __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return expected;
}
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

static inline void sincospi(double x, double* s, double* c) {
    // exact quadrant reduction so that emulated twiddles are as accurate as the device's
    double r = std::fmod(x, 2.0);
    if (r < 0) r += 2.0;
    int q = (int)std::floor(r * 2.0 + 0.5);          // nearest multiple of 1/2
    double t = (r - 0.5 * q) * M_PI;
    double st = std::sin(t), ct = std::cos(t);
    switch (q & 3) {
        case 0: *s = st; *c = ct; break;
        case 1: *s = ct; *c = -st; break;
        case 2: *s = -st; *c = -ct; break;
        default: *s = -ct; *c = st; break;
    }
}
static inline void sincos_emul(double x, double* s, double* c) { *s = std::sin(x); *c = std::cos(x); }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
