// Error plumbing and device queries of the dynaphopy_b200 C ABI.
#include <atomic>
#include <cstdarg>

#include "dp_common.cuh"
#include "dp_tma.cuh"

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

int make_tensor_map_2d(TensorMap2D* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t pitch_bytes,
                       uint32_t box_inner, uint32_t box_rows) {
    DP_REQUIRE(tm && base && inner > 0 && rows > 0 && pitch_bytes % 16 == 0 && ((size_t)base & 15) == 0 &&
                   box_inner >= 2 && box_inner <= 256 && box_inner % 2 == 0 && box_rows >= 1 && box_rows <= 256,
               DP_ERR_INVALID, "make_tensor_map_2d: bad tensor description");
#ifdef DP_EMUL
    tm->base = (double*)base; tm->inner = inner; tm->rows = rows; tm->pitch_bytes = pitch_bytes;
    tm->box_inner = box_inner; tm->box_rows = box_rows;
    return DP_OK;
#else
    // cuTensorMapEncodeTiled comes from the driver; fetched through the runtime so that the library does not link libcuda
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || !fn) {
            (void)cudaGetLastError();
            set_error("make_tensor_map_2d: the driver does not provide cuTensorMapEncodeTiled");
            return DP_ERR_UNSUPPORTED;
        }
        encode = (EncodeFn)fn;
    }
    cuuint64_t gdim[2] = {inner, rows};
    cuuint64_t gstr[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("make_tensor_map_2d: cuTensorMapEncodeTiled failed with %d", (int)r);
        return DP_ERR_UNSUPPORTED;
    }
    return DP_OK;
#endif
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
