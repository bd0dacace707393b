// A6: maximum-entropy (Burg) power spectrum, selector key 1.
// Reference: c/mem.c:102-252 (MaximumEntropyMethod, GetCoefficients, FrequencyEvaluation) and
// dynaphopy/power_spectrum/__init__.py:81-103, with the canonical Data[N] := 0 semantics for the
// reference's out-of-bounds read.
//
// Stage 1 (burg_kernel): one CTA per (series, part) with part = Re or Im.  The forward and
// backward prediction-error arrays live in shared memory when 16*(N-1) bytes fit (on-chip,
// latency/fp64 bound), otherwise in a workspace slab (L2/HBM bound, 32*(N-1) bytes per order).
// Each order is ONE sweep: the update of order k and the inner products of order k+1 are fused,
// elements are read once and written once, neighbours travel by warp shuffle.  All reductions
// are fixed-order trees (no atomics) so results are run-to-run reproducible.
// Stage 2 (mem_eval_kernel): P(f) = dt * sum_parts xms / |1 - sum_k d_k z^k|^2 per (series, f).
#include <cmath>
#include <cstdlib>

#include "dp_common.cuh"
#ifndef DP_EMUL
#include <cooperative_groups.h>
#endif

namespace dp {

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int BURG_THREADS = 512;
constexpr int BURG_WARPS = BURG_THREADS / 32;

struct BurgParams {
    SeriesView x;           // the series (any layout, see SeriesView)
    int N, m;
    int unit0, nunits;      // units = 2*series + part
    double* slabs;          // global-mode storage: per CTA 2*len doubles (nullptr in smem mode)
    double* coef;           // [2S][m]
    double* xms;            // [2S]
};

__device__ __forceinline__ void block_sum2(double& a, double& b, double* red /*[2*BURG_WARPS]*/) {
    a = warp_sum(a);
    b = warp_sum(b);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();                 // protect `red` from the previous use
    if (lane == 0) { red[warp] = a; red[BURG_WARPS + warp] = b; }
    __syncthreads();
    double sa = 0.0, sb = 0.0;
#pragma unroll
    for (int w = 0; w < BURG_WARPS; w++) { sa += red[w]; sb += red[BURG_WARPS + w]; }
    a = sa; b = sb;
}

template <bool SMEM>
__global__ void __launch_bounds__(BURG_THREADS) burg_kernel(BurgParams p) {
    DP_DYNSMEM(double, sm);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, m = p.m, len = N - 1;
    double* red = sm;                         // 2*BURG_WARPS
    double* d_a = sm + 2 * BURG_WARPS;        // m
    double* d_b = d_a + m;                    // m
    double* f;
    double* b;
    if (SMEM) { f = d_b + m; b = f + len; }
    else { f = p.slabs + (size_t)blockIdx.x * 2 * (size_t)len; b = f + len; }

    for (int unit = blockIdx.x; unit < p.nunits; unit += gridDim.x) {
        const int gu = p.unit0 + unit;
        const int xser = gu >> 1, xpart = gu & 1;           // sample j, part Re / Im: xval(j)
        auto xval = [&](int j) { return __ldg(reinterpret_cast<const double*>(p.x.at(xser, j)) + xpart); };
        __syncthreads();
        // ---- init: f[j] = x[j+1], b[j] = x[j+2] (x[N] := 0); power and first inner products ----
        double pw = 0.0, num = 0.0, den = 0.0;
        for (int j = tid; j < len; j += BURG_THREADS) {
            double fj = xval(j + 1);
            double bj = (j + 2 < N) ? xval(j + 2) : 0.0;
            f[j] = fj; b[j] = bj;
            pw += __dmul_rn(fj, fj);
            num += __dmul_rn(fj, bj);
            den += __dadd_rn(__dmul_rn(fj, fj), __dmul_rn(bj, bj));
        }
        double dummy = 0.0;
        block_sum2(pw, dummy, red);
        double xms = pw / (double)N;
        double* dprev = d_a;
        double* dcur = d_b;
        bool dead = false;
        for (int k = 1; k <= m; k++) {
            block_sum2(num, den, red);
            if (fabs(den) < 1.0e-6) { dead = true; break; }      // c/mem.c:229
            const double c = __ddiv_rn(__dmul_rn(2.0, num), den);
            xms = __dmul_rn(xms, __dsub_rn(1.0, __dmul_rn(c, c)));
            for (int a = tid; a < k - 1; a += BURG_THREADS)
                dcur[a] = __dsub_rn(dprev[a], __dmul_rn(c, dprev[k - 2 - a]));
            if (tid == 0) dcur[k - 1] = c;
            { double* t = dprev; dprev = dcur; dcur = t; }          // dprev now holds order-k coefficients
            if (k == m) break;
            // ---- fused sweep: update to order k and accumulate the order k+1 inner products ----
            const int cnt_new = N - k - 1;                            // elements j = 0 .. cnt_new-1
            int seg = (cnt_new + BURG_WARPS - 1) / BURG_WARPS;
            seg = (seg + 31) / 32 * 32;
            const int seg0 = warp * seg;
            const int seg1 = imin(seg0 + seg, cnt_new);
            double edge_f = 0.0, edge_b = 0.0;                         // old values at index seg1
            if (seg0 < cnt_new) { edge_f = f[seg1]; edge_b = b[seg1]; }
            __syncthreads();                                           // all edges read before any write
            num = 0.0; den = 0.0;
            for (int base = seg0; base < seg1; base += 32) {
                const int j = base + lane;
                double fo = 0.0, bo = 0.0;
                if (j < seg1) { fo = f[j]; bo = b[j]; }
                double fn = __shfl_down_sync(0xffffffffu, fo, 1);
                double bn = __shfl_down_sync(0xffffffffu, bo, 1);
                if (j + 1 == seg1) { fn = edge_f; bn = edge_b; }
                else if (lane == 31 && j + 1 < seg1) { fn = f[j + 1]; bn = b[j + 1]; }
                if (j < seg1) {
                    double f2 = __dsub_rn(fo, __dmul_rn(c, bo));
                    double b2 = __dsub_rn(bn, __dmul_rn(c, fn));
                    f[j] = f2; b[j] = b2;
                    num += __dmul_rn(f2, b2);
                    den += __dadd_rn(__dmul_rn(f2, f2), __dmul_rn(b2, b2));
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (dead) xms = 0.0;
        for (int a = tid; a < m; a += BURG_THREADS) p.coef[(size_t)gu * m + a] = dead ? 0.0 : dprev[a];
        if (tid == 0) p.xms[gu] = xms;
    }
}

#ifndef DP_EMUL
// ---- long series: one thread-block CLUSTER per (series, part), the prediction-error arrays never leave the SMs ----
// When 16*(N-1) bytes do not fit one CTA's shared memory the single-CTA kernel streams both arrays through L2/HBM
// once per order (32*(N-1) bytes x m orders per part: 8.4 GB per series at N = 2^17, m = 1000).  Here a cluster of
// BC_CTAS CTAs splits the time axis into contiguous blocks: a thread owns E consecutive samples, the forward errors
// f live in its REGISTERS, the backward errors b in shared memory (row pitch E+1: conflict-free), and one order is
//   sweep (registers + shared memory; the one neighbour sample a thread needs comes from a parity-double-buffered
//   edge array, across CTAs through distributed shared memory) -> warp/CTA partial sums -> partials pushed into
//   every peer's shared memory -> ONE cluster barrier -> every CTA adds the BC_CTAS partials in rank order.
// Element arithmetic is the single-CTA kernel's (explicit round-to-nearest mul/sub, reference operation order).
namespace cg = cooperative_groups;
constexpr int BC_CTAS = 8;                       // portable cluster size
constexpr int BC_THREADS = 512;
constexpr int BC_WARPS = BC_THREADS / 32;

struct BurgClusterSmem {
    double red_w[2 * BC_WARPS];                  // per-warp partial (num | den)
    double cl_red[2][BC_CTAS][2];                // [order parity][source rank] (num, den), written by the peers
    double edge_in[2][2];                        // [order parity] (f, b) of the first sample of the NEXT rank
    double fb0[2][BC_THREADS][2];                // [order parity][thread] (f, b) of the thread's first sample
};

template <int E>
__global__ void __launch_bounds__(BC_THREADS, 1) burg_cluster_kernel(BurgParams p) {
    DP_DYNSMEM(double, sm);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, m = p.m, len = N - 1;
    BurgClusterSmem* cs = reinterpret_cast<BurgClusterSmem*>(sm);
    double* d_a = sm + (sizeof(BurgClusterSmem) + 7) / 8;
    double* d_b = d_a + m;
    double* bs = d_b + m + (size_t)tid * (E + 1);          // this thread's row of backward errors
    const int per_cta = BC_THREADS * E;
    const int j0 = rank * per_cta + tid * E;               // first sample owned by this thread
    BurgClusterSmem* peer[BC_CTAS];
#pragma unroll
    for (int r = 0; r < BC_CTAS; r++) peer[r] = cluster.map_shared_rank(cs, r);
    const int ncl = gridDim.x / BC_CTAS, cl = blockIdx.x / BC_CTAS;

    // sum of (a, b) over the whole cluster for the reduction round `par`; also delivers the first sample of this
    // CTA (e0f, e0b) to the previous rank.  One cluster barrier.
    auto cluster_sum2 = [&](double& a, double& b, int par, double e0f, double e0b) {
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) { cs->red_w[warp] = a; cs->red_w[BC_WARPS + warp] = b; }
        __syncthreads();
        if (tid < BC_CTAS) {
            double sa = 0.0, sb = 0.0;
#pragma unroll
            for (int w = 0; w < BC_WARPS; w++) { sa += cs->red_w[w]; sb += cs->red_w[BC_WARPS + w]; }
            peer[tid]->cl_red[par][rank][0] = sa;
            peer[tid]->cl_red[par][rank][1] = sb;
        }
        if (tid == 0 && rank > 0) { peer[rank - 1]->edge_in[par][0] = e0f; peer[rank - 1]->edge_in[par][1] = e0b; }
        cluster.sync();
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int r = 0; r < BC_CTAS; r++) { sa += cs->cl_red[par][r][0]; sb += cs->cl_red[par][r][1]; }
        a = sa; b = sb;
    };

    for (int unit = cl; unit < p.nunits; unit += ncl) {
        const int gu = p.unit0 + unit;
        const int xser = gu >> 1, xpart = gu & 1;
        auto xval = [&](int j) { return __ldg(reinterpret_cast<const double*>(p.x.at(xser, j)) + xpart); };
        // ---- init: f[j] = x[j+1], b[j] = x[j+2] (x[N] := 0) ----
        double f[E];
        double pw = 0.0, num = 0.0, den = 0.0;
#pragma unroll
        for (int e = 0; e < E; e++) {
            const int j = j0 + e;
            double fj = 0.0, bj = 0.0;
            if (j < len) {
                fj = xval(j + 1);
                bj = (j + 2 < N) ? xval(j + 2) : 0.0;
                pw = fma(fj, fj, pw);
                num = fma(fj, bj, num);
                den = fma(fj, fj, fma(bj, bj, den));
            }
            f[e] = fj; bs[e] = bj;
        }
        double dummy = 0.0;
        cluster_sum2(pw, dummy, 0, 0.0, 0.0);
        double xms = pw / (double)N;
        cs->fb0[1][tid][0] = f[0]; cs->fb0[1][tid][1] = bs[0];        // edges for order k = 1 (parity 1)
        double* dprev = d_a;
        double* dcur = d_b;
        bool dead = false;
        for (int k = 1; k <= m; k++) {
            const int par = k & 1;
            cluster_sum2(num, den, par, f[0], bs[0]);
            if (fabs(den) < 1.0e-6) { dead = true; break; }          // c/mem.c:229 (uniform over the cluster)
            const double c = __ddiv_rn(__dmul_rn(2.0, num), den);
            xms = __dmul_rn(xms, __dsub_rn(1.0, __dmul_rn(c, c)));
            for (int a = tid; a < k - 1; a += BC_THREADS)
                dcur[a] = __dsub_rn(dprev[a], __dmul_rn(c, dprev[k - 2 - a]));
            if (tid == 0) dcur[k - 1] = c;
            { double* t = dprev; dprev = dcur; dcur = t; }
            if (k == m) break;
            // ---- sweep: update to order k, accumulate the order k+1 inner products ----
            const int cnt_new = N - k - 1;
            double fnx, bnx;                                          // old (f, b) of sample j0 + E
            if (tid + 1 < BC_THREADS) { fnx = cs->fb0[par][tid + 1][0]; bnx = cs->fb0[par][tid + 1][1]; }
            else if (rank + 1 < BC_CTAS) { fnx = cs->edge_in[par][0]; bnx = cs->edge_in[par][1]; }
            else { fnx = 0.0; bnx = 0.0; }
            num = 0.0; den = 0.0;
            // the element updates keep the reference's two roundings (mul, then sub); the inner products are summed in
            // a different order than the reference's serial loop anyway, so they use fused multiply-adds
            if (j0 + E <= cnt_new) {                                  // whole block valid (all but one warp per order)
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const double fo = f[e], bo = bs[e];
                    const double fn = (e + 1 < E) ? f[e + 1] : fnx;
                    const double bn = (e + 1 < E) ? bs[e + 1] : bnx;
                    const double f2 = __dsub_rn(fo, __dmul_rn(c, bo));
                    const double b2 = __dsub_rn(bn, __dmul_rn(c, fn));
                    f[e] = f2; bs[e] = b2;
                    num = fma(f2, b2, num);
                    den = fma(f2, f2, fma(b2, b2, den));
                }
            } else if (j0 < cnt_new + 1) {                             // block straddles the end of the valid range
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const double fo = f[e], bo = bs[e];
                    const double fn = (e + 1 < E) ? f[e + 1] : fnx;
                    const double bn = (e + 1 < E) ? bs[e + 1] : bnx;
                    double f2 = __dsub_rn(fo, __dmul_rn(c, bo));
                    double b2 = __dsub_rn(bn, __dmul_rn(c, fn));
                    if (j0 + e >= cnt_new) { f2 = 0.0; b2 = 0.0; }
                    f[e] = f2; bs[e] = b2;
                    num = fma(f2, b2, num);
                    den = fma(f2, f2, fma(b2, b2, den));
                }
            }                                                          // else: block already retired (all zeros)
            cs->fb0[par ^ 1][tid][0] = f[0]; cs->fb0[par ^ 1][tid][1] = bs[0];
        }
        __syncthreads();
        if (dead) xms = 0.0;
        if (rank == 0) {
            for (int a = tid; a < m; a += BC_THREADS) p.coef[(size_t)gu * m + a] = dead ? 0.0 : dprev[a];
            if (tid == 0) p.xms[gu] = xms;
        }
        cluster.sync();                           // nobody starts the next unit while a peer still reads this one's edges
    }
}

static size_t burg_cluster_smem(int E, int m) {
    return (sizeof(BurgClusterSmem) + 7) / 8 * 8 + (2 * (size_t)m + (size_t)BC_THREADS * (E + 1)) * sizeof(double);
}

// smallest per-thread block that covers the series, or 0 when the cluster form does not apply
static int burg_cluster_E(int64_t T, int m) {
    const int64_t len = T - 1;
    for (int E : {8, 16, 32})
        if ((int64_t)BC_CTAS * BC_THREADS * E >= len && burg_cluster_smem(E, m) <= 220 * 1024) return E;
    return 0;
}

template <int E>
static cudaError_t launch_burg_cluster(const BurgParams& p, cudaStream_t st) {
    auto k = burg_cluster_kernel<E>;
    const size_t smem = burg_cluster_smem(E, p.m);
    cudaError_t e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = BC_CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(BC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(BC_CTAS);
    int ncl = 0;
    e = cudaOccupancyMaxActiveClusters(&ncl, (const void*)k, &cfg);
    if (e != cudaSuccess) return e;
    if (ncl < 1) return cudaErrorLaunchOutOfResources;
    ncl = imin(ncl, p.nunits);
    cfg.gridDim = dim3((unsigned)(ncl * BC_CTAS));
    count_launch();
    return cudaLaunchKernelEx(&cfg, k, p);
}
#endif  // !DP_EMUL

// P(f) for every (series, frequency).  z^k through sincos(k*theta) with theta = arg(exp(i delta)),
// the quantity glibc's cpow(z, k) = cexp(k * clog(z)) uses (c/mem.c:186-195).
__global__ void __launch_bounds__(256)
mem_eval_kernel(const double* __restrict__ coef, const double* __restrict__ xms, const double* __restrict__ freq,
                int S, int F, int m, double dt, double scale, int nan_to_num, double* __restrict__ out) {
    int64_t total = (int64_t)S * F;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int s = (int)(e / F), fi = (int)(e % F);
        const double delta = __dmul_rn(__dmul_rn(__dmul_rn(freq[fi], 2.0), M_PI), dt);
        double sd, cd;
        sincos(delta, &sd, &cd);
        const double theta = atan2(sd, cd);
        const double x0 = xms[2 * s], x1 = xms[2 * s + 1];
        const double* d0 = coef + (size_t)(2 * s) * m;
        const double* d1 = d0 + m;
        double a0r = 1.0, a0i = 0.0, a1r = 1.0, a1i = 0.0;
        for (int k = 1; k <= m; k++) {
            double sk, ck;
            sincos((double)k * theta, &sk, &ck);
            const double c0 = __ldg(d0 + k - 1), c1 = __ldg(d1 + k - 1);
            a0r -= c0 * ck; a0i -= c0 * sk;
            a1r -= c1 * ck; a1i -= c1 * sk;
        }
        double acc = 0.0;
        if (x0 != 0.0) acc += x0 / (a0r * a0r + a0i * a0i);
        if (x1 != 0.0) acc += x1 / (a1r * a1r + a1i * a1i);
        double val = acc * dt;
        if (nan_to_num) {
            if (val != val) val = 0.0;
            else if (fabs(val) > 1.7976931348623157e308) val = val > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
        }
        out[(size_t)fi * S + s] = val * scale;
    }
}

struct MemLayout {
    size_t off_freq, off_coef, off_xms, off_slabs, total;
    int slots;
    bool smem_mode;
    size_t smem_bytes;
};

static MemLayout mem_layout(int64_t T, int S, int F, int m) {
    MemLayout l;
    size_t o = 0;
    l.off_freq = o; o += align_up((size_t)F * sizeof(double), 256);
    l.off_coef = o; o += align_up((size_t)2 * S * m * sizeof(double), 256);
    l.off_xms = o;  o += align_up((size_t)2 * S * sizeof(double), 256);
    size_t len = (size_t)(T - 1);
    size_t fixed = (2 * BURG_WARPS + 2 * (size_t)m) * sizeof(double);
    l.smem_mode = fixed + 2 * len * sizeof(double) <= 200 * 1024;
    l.smem_bytes = fixed + (l.smem_mode ? 2 * len * sizeof(double) : 0);
    int units = 2 * S;
    l.slots = imin(units, imax(sm_count(), 1) * (l.smem_mode ? 4 : 2));
    l.off_slabs = o;
    if (!l.smem_mode) o += align_up((size_t)l.slots * 2 * len * sizeof(double), 256);
    l.total = o;
    return l;
}

}  // namespace dp

extern "C" {

size_t dp_power_mem_workspace_bytes(int64_t T, int S, int F, int coefficients) {
    if (T < 3 || S <= 0 || F <= 0 || coefficients < 1) return 0;
    return dp::mem_layout(T, S, F, coefficients).total;
}

int dp_power_mem(const double* series, int64_t T, int S, int64_t ld, double dt, const double* freq, int F,
                 int coefficients, double scale, int nan_to_num, double* out, void* workspace,
                 size_t workspace_bytes, void* stream) {
    DP_REQUIRE(S > 0 && ld >= S, DP_ERR_INVALID, "dp_power_mem: bad sizes");
    return dp_power_mem_strided(series, T, S, 1, ld, 0, 0, dt, freq, F, coefficients, scale, nan_to_num, out, workspace,
                                workspace_bytes, stream);
}

int dp_power_mem_strided(const double* series, int64_t T, int S, int64_t stride_s, int64_t stride_t, int64_t tb_len,
                         int64_t tb_stride, double dt, const double* freq, int F, int coefficients, double scale,
                         int nan_to_num, double* out, void* workspace, size_t workspace_bytes, void* stream) {
    DP_REQUIRE(series && freq && out && workspace, DP_ERR_INVALID, "dp_power_mem: null pointer");
    DP_REQUIRE(S > 0 && F > 0 && coefficients >= 1 && stride_t != 0 && tb_len >= 0 && tb_len < (1LL << 31), DP_ERR_INVALID,
               "dp_power_mem: bad sizes");
    if (tb_len >= T) tb_len = 0;
    // reference guard, dynaphopy/power_spectrum/__init__.py:85-87 (prints and exit()s there)
    DP_REQUIRE(T > (int64_t)coefficients + 1, DP_ERR_INVALID,
               "Number of coefficients should be smaller than the number of time steps (T=%lld, m=%d)",
               (long long)T, coefficients);
    DP_REQUIRE(T < (1LL << 30), DP_ERR_UNSUPPORTED, "dp_power_mem: T too large");
    DP_REQUIRE((size_t)coefficients * 16 <= 150 * 1024, DP_ERR_UNSUPPORTED, "dp_power_mem: more than 9600 coefficients unsupported");
    dp::MemLayout lay = dp::mem_layout(T, S, F, coefficients);
    DP_REQUIRE(workspace_bytes >= lay.total, DP_ERR_WORKSPACE, "dp_power_mem: workspace %zu < %zu", workspace_bytes, lay.total);
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    {
        int rc_up = dp::upload_table(ws + lay.off_freq, freq, (size_t)F * sizeof(double), st);   // kernel parameters: no host sync
        if (rc_up) return rc_up;
    }
    dp::BurgParams p;
    p.x.base = (const dp::cplx*)series; p.x.stride_s = stride_s; p.x.stride_t = stride_t; p.x.tb_len = (int)tb_len;
    p.x.tb_stride = tb_stride;
    p.N = (int)T; p.m = coefficients;
    p.unit0 = 0; p.nunits = 2 * S;
    p.slabs = lay.smem_mode ? nullptr : (double*)(ws + lay.off_slabs);
    p.coef = (double*)(ws + lay.off_coef);
    p.xms = (double*)(ws + lay.off_xms);
#ifndef DP_EMUL
    const int cluster_E = (lay.smem_mode || dp::dev_env("DP_MEM_NO_CLUSTER")) ? 0 : dp::burg_cluster_E(T, coefficients);
#else
    const int cluster_E = 0;
#endif
    if (cluster_E) {
#ifndef DP_EMUL
        cudaError_t e = cluster_E == 8 ? dp::launch_burg_cluster<8>(p, st)
                        : cluster_E == 16 ? dp::launch_burg_cluster<16>(p, st) : dp::launch_burg_cluster<32>(p, st);
        DP_CUDA_OK(e);
#endif
    } else if (lay.smem_mode) {
        auto k = dp::burg_kernel<true>;
        DP_CUDA_OK(DP_SET_SMEM(k, lay.smem_bytes));
        DP_LAUNCH(k, dim3(lay.slots), dim3(dp::BURG_THREADS), lay.smem_bytes, st, p);
    } else {
        auto k = dp::burg_kernel<false>;
        DP_CUDA_OK(DP_SET_SMEM(k, lay.smem_bytes));
        DP_LAUNCH(k, dim3(lay.slots), dim3(dp::BURG_THREADS), lay.smem_bytes, st, p);
    }
    DP_CUDA_OK(cudaGetLastError());
    auto ke = dp::mem_eval_kernel;
    int64_t total = (int64_t)S * F;
    DP_LAUNCH(ke, dim3((unsigned)dp::imin<int64_t>((total + 255) / 256, 65535)), dim3(256), 0, st,
              (const double*)p.coef, (const double*)p.xms, (const double*)(ws + lay.off_freq), S, F, coefficients, dt,
              scale, nan_to_num, out);
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

}  // extern "C"
