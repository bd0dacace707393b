// Error plumbing and device queries of the dynaphopy_b200 C ABI.
#include <atomic>
#include <cstdarg>

#include "dp_common.cuh"

namespace dp {

long long launches();

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return DP_ERR_CUDA;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
#ifdef DP_EMUL
    return 4;
#else
    static int cached = 0;
    if (cached > 0) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached = n;
    return n;
#endif
}

// Small host arrays -> device WITHOUT blocking the host: the bytes travel as kernel parameters (1 KB per launch).
// cudaMemcpyAsync from pageable memory synchronises the host with the stream before it starts, which serialises a
// caller that wants to queue work ahead of the device (MeshPipeline.run_cyclic, run_from_host).
struct UploadChunk { unsigned int w[256]; };

__global__ void __launch_bounds__(256) upload_chunk_kernel(unsigned int* __restrict__ dst, UploadChunk c, int nwords) {
    if ((int)threadIdx.x < nwords) dst[threadIdx.x] = c.w[threadIdx.x];
}

int upload_small(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    DP_REQUIRE(bytes % 4 == 0 && ((size_t)dst & 3) == 0, DP_ERR_INVALID, "upload_small: 4-byte granularity");
    const unsigned int* s = (const unsigned int*)src;
    size_t words = bytes / 4;
    for (size_t o = 0; o < words; o += 256) {
        UploadChunk c;
        const int n = (int)imin<size_t>(256, words - o);
        memcpy(c.w, s + o, (size_t)n * 4);
        auto k = upload_chunk_kernel;
        DP_LAUNCH(k, dim3(1), dim3(256), 0, st, (unsigned int*)dst + o, c, n);
        DP_CUDA_OK(cudaGetLastError());
    }
    return DP_OK;
}

int upload_table(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return DP_OK;
    if (bytes <= 64 * 1024 && bytes % 4 == 0 && ((size_t)dst & 3) == 0) return upload_small(dst, src, bytes, st);
    DP_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));   // pageable source: staged before return
    return DP_OK;
}

}  // namespace dp

extern "C" {

const char* dp_version(void) { return "dynaphopy_b200 0.1.0 (sm_100a)"; }

const char* dp_last_error(void) { return dp::g_err; }

long long dp_launch_count(void) { return dp::launches(); }

int dp_sm_count(void) {
    int n = dp::sm_count();
    if (n <= 0) {
        dp::set_error("no CUDA device available (dynaphopy_b200 has no CPU fallback)");
        return DP_ERR_CUDA;
    }
    return n;
}

}  // extern "C"
