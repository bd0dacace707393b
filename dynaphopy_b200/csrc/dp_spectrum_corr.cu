// A5: direct ("Fourier transform of the velocity autocorrelation") power spectrum, selector key 0.
// Reference: c/correlation.c:101-178 (correlation_par / EvaluateCorrelation) and
// dynaphopy/power_spectrum/__init__.py:57-75.
//
// The reference recomputes every lag product for every output frequency: O(F * N^2 / step).
// The products do not depend on the frequency, so this implementation factors the sum:
//   stage 1  C(i) = sum_{j < N-i-step} conj(v_j) v_{j+i}   for lags i = 0, step, 2 step, ...
//            -- the compute-bound fp64 FMA kernel (4 N^2 / step flop per series);
//            D(i) = sum_{j < N-i-step} conj(v_j) v_{j+i+step} when integration_method == 0
//   stage 2  P(f) = Re[ sum_i weight terms ] * dt / (N div step), lag-major like the reference.
// weight_mode REFERENCE reproduces the shipped weight exp(w*tau + 1i) (real exponential growth,
// c/correlation.c:158-170) including glibc cexp's staged overflow; INTENDED uses exp(i w tau).
#include <cmath>
#include <algorithm>
#include <vector>

#include "dp_common.cuh"

namespace dp {

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// series[t*ld + s] -> vt[(s - s0)*N + t]  (32 x 32 tile through shared memory)
__global__ void __launch_bounds__(256)
transpose_series_kernel(const cplx* __restrict__ series, int64_t ld, int N, int s0, int ns, cplx* __restrict__ vt) {
    __shared__ cplx tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    const int tb = blockIdx.x * 32, sb = blockIdx.y * 32;
    for (int r = ty; r < 32; r += 8) {
        int t = tb + r, s = sb + tx;
        if (t < N && s < ns) tile[r][tx] = series[(int64_t)t * ld + s0 + s];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int s = sb + r, t = tb + tx;
        if (t < N && s < ns) vt[(size_t)s * N + t] = tile[tx][r];
    }
}

constexpr int CR_LAGS = 64;       // lags per CTA
constexpr int CR_JP = 4;          // origin phases (threads per lag)
constexpr int CR_CHUNK = 1024;    // origins per shared-memory window

// C[s][a] for a in [a0, a0 + CR_LAGS): window of v staged in shared memory.
__global__ void __launch_bounds__(CR_LAGS* CR_JP)
lag_sums_kernel(const cplx* __restrict__ vt, int N, int step, int nl_valid, cplx* __restrict__ C, int c_ld) {
    DP_DYNSMEM(cplx, win);
    __shared__ cplx part[CR_JP][CR_LAGS];
    const int s = blockIdx.y;
    const int a0 = blockIdx.x * CR_LAGS;
    const int la = threadIdx.x % CR_LAGS, jp = threadIdx.x / CR_LAGS;
    const int a = a0 + la;
    const int lag = a * step;
    const int cnt = (a < nl_valid) ? N - lag - step : 0;          // origins j in [0, cnt)
    const int lag_min = a0 * step;
    const int span = CR_CHUNK + (CR_LAGS - 1) * step;             // window length in elements
    const int cnt_max = N - lag_min - step;                       // largest origin count in this tile
    const cplx* v = vt + (size_t)s * N;
    double ar = 0.0, ai = 0.0;
    for (int j0 = 0; j0 < cnt_max; j0 += CR_CHUNK) {
        __syncthreads();
        // origins v[j0 .. j0+CHUNK) and partners v[j0+lag_min .. j0+lag_min+span)
        for (int e = threadIdx.x; e < CR_CHUNK; e += blockDim.x) {
            int j = j0 + e;
            win[e] = j < N ? v[j] : make_double2(0.0, 0.0);
        }
        for (int e = threadIdx.x; e < span; e += blockDim.x) {
            int j = j0 + lag_min + e;
            win[CR_CHUNK + e] = j < N ? v[j] : make_double2(0.0, 0.0);
        }
        __syncthreads();
        const int jend = imin(CR_CHUNK, cnt - j0);
        const cplx* pw = win + CR_CHUNK + la * step;
        for (int e = jp; e < jend; e += CR_JP) {
            cplx x = win[e], y = pw[e];
            ar = fma(x.x, y.x, fma(x.y, y.y, ar));                // conj(x) * y
            ai = fma(x.x, y.y, fma(-x.y, y.x, ai));
        }
    }
    part[jp][la] = make_double2(ar, ai);
    __syncthreads();
    if (jp == 0 && a < nl_valid) {
        cplx acc = part[0][la];
#pragma unroll
        for (int q = 1; q < CR_JP; q++) acc = cadd(acc, part[q][la]);
        C[(size_t)s * c_ld + a] = acc;
    }
}

// D(a) = sum_{j < cnt_a} conj(v_j) v_{j + lag_a + step} = C(a+1) + the `step` trailing origins
__global__ void __launch_bounds__(256)
lag_sums_shifted_kernel(const cplx* __restrict__ vt, int N, int step, int nl_valid, int ns,
                        const cplx* __restrict__ C, cplx* __restrict__ D, int c_ld) {
    int64_t total = (int64_t)ns * nl_valid;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int s = (int)(e / nl_valid), a = (int)(e % nl_valid);
        const cplx* v = vt + (size_t)s * N;
        int lag1 = (a + 1) * step;
        int cnt = N - a * step - step;
        int cnt1 = imax(cnt - step, 0);
        cplx acc = (a + 1 < nl_valid) ? C[(size_t)s * c_ld + a + 1] : make_double2(0.0, 0.0);
        for (int j = cnt1; j < cnt; j++) acc = cadd(acc, cmul(cconj(v[j]), v[j + lag1]));
        D[(size_t)s * c_ld + a] = acc;
    }
}

// glibc cexp(x + 1i): exp(x) * (cos 1, sin 1) with the staged scaling it applies for x > 709
__device__ __forceinline__ cplx cexp_plus_i(double x) {
    const double c1 = 0.54030230586813971740, s1 = 0.84147098480789650665;
    double c = c1, s = s1;
    const double t = 709.0;
    if (x > t) {
        const double et = exp(t);
        x -= t; c *= et; s *= et;
        if (x > t) { x -= t; c *= et; s *= et; }
        if (x > t) return make_double2(1.7976931348623157e308 * c, 1.7976931348623157e308 * s);
    }
    const double ev = exp(x);
    return make_double2(ev * c, ev * s);
}

__device__ __forceinline__ cplx corr_weight(double arg, int weight_mode) {
    if (weight_mode == DP_CORR_WEIGHT_REFERENCE) return cexp_plus_i(arg);
    double s, c;
    sincos(arg, &s, &c);
    return make_double2(c, s);
}

// W[k][a][f], k = 0: w*(lag*dt); 1: w*((lag+step)*dt); 2: w*(lag+step)   (c/correlation.c:158,162,163)
__global__ void __launch_bounds__(256)
corr_weights_kernel(const double* __restrict__ freq, int F, int nl_valid, int step, double dt, int method,
                    int weight_mode, cplx* __restrict__ W) {
    int64_t total = (int64_t)nl_valid * F;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int a = (int)(e / F), f = (int)(e % F);
        double w = __dmul_rn(__dmul_rn(freq[f], 2.0), M_PI);
        int lag = a * step;
        W[e] = corr_weight(__dmul_rn(w, __dmul_rn((double)lag, dt)), weight_mode);
        if (method == 0) {
            W[total + e] = corr_weight(__dmul_rn(w, __dmul_rn((double)(lag + step), dt)), weight_mode);
            W[2 * total + e] = corr_weight(__dmul_rn(w, (double)(lag + step)), weight_mode);
        }
    }
}

__global__ void __launch_bounds__(256)
corr_spectrum_kernel(const cplx* __restrict__ C, const cplx* __restrict__ D, const cplx* __restrict__ W,
                     int s0, int ns, int S, int F, int nl_valid, int c_ld, int method, double dt, double ndiv, double scale,
                     double* __restrict__ out) {
    int64_t total = (int64_t)ns * F;
    const int64_t wsz = (int64_t)nl_valid * F;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int s = (int)(e / F), f = (int)(e % F);
        double ar = 0.0, ai = 0.0;
        for (int a = 0; a < nl_valid; a++) {
            cplx c = C[(size_t)s * c_ld + a];
            cplx t = cmul(c, W[(int64_t)a * F + f]);
            ar += t.x; ai += t.y;
            if (method == 0) {
                cplx ti = cmul(D[(size_t)s * c_ld + a], W[wsz + (int64_t)a * F + f]);
                cplx tj = cmul(c, W[2 * wsz + (int64_t)a * F + f]);
                ar = ar + ti.x + tj.x; ai = ai + ti.y + tj.y;
            } else {
                ar += t.x; ai += t.y;
            }
        }
        // creal(Correl) * TimeStep / (NumberOfData / Increment)   (c/correlation.c:177)
        out[(size_t)f * S + s0 + s] = __dmul_rn(__ddiv_rn(__dmul_rn(ar, dt), ndiv), scale);
        (void)ai;
    }
}

struct CorrLayout {
    size_t off_freq, off_W, off_C, off_D, off_vt, total;
    int nl_valid, c_ld, chunk;
};

static CorrLayout corr_layout(int64_t T, int S, int F, int step, int method) {
    CorrLayout l;
    int N = (int)T;
    // valid lags: i = a*step with N - i - step > 0
    l.nl_valid = (N - step > 0) ? (N - step + step - 1) / step : 0;
    l.c_ld = imax(l.nl_valid, 1);
    size_t o = 0;
    l.off_freq = o; o += align_up((size_t)F * sizeof(double), 256);
    l.off_W = o;    o += align_up((size_t)(method == 0 ? 3 : 1) * l.c_ld * F * sizeof(cplx), 256);
    l.off_C = o;    o += align_up((size_t)S * l.c_ld * sizeof(cplx), 256);
    l.off_D = o;    o += align_up((size_t)S * l.c_ld * sizeof(cplx), 256);
    // series are transposed chunk-wise into a contiguous [chunk][N] slab of at most 1 GiB
    size_t per = (size_t)N * sizeof(cplx);
    l.chunk = (int)imin<size_t>((size_t)S, imax<size_t>(1, ((size_t)1 << 30) / per));
    l.off_vt = o;   o += align_up((size_t)l.chunk * per, 256);
    l.total = o;
    return l;
}

}  // namespace dp

extern "C" {

size_t dp_power_correlation_workspace_bytes(int64_t T, int S, int F, int step) {
    if (T < 1 || S <= 0 || F <= 0 || step < 1) return 0;
    return dp::corr_layout(T, S, F, step, 0).total;
}

int dp_power_correlation(const double* series, int64_t T, int S, int64_t ld, double dt, const double* freq, int F,
                         int step, int integration_method, int weight_mode, double scale, double* out,
                         void* workspace, size_t workspace_bytes, void* stream) {
    DP_REQUIRE(series && freq && out && workspace, DP_ERR_INVALID, "dp_power_correlation: null pointer");
    DP_REQUIRE(S > 0 && ld >= S && F > 0 && T >= 1 && step >= 1, DP_ERR_INVALID, "dp_power_correlation: bad sizes");
    DP_REQUIRE(T < (1LL << 30), DP_ERR_UNSUPPORTED, "dp_power_correlation: T too large");
    // c/correlation.c:138-141: "Integration method selected does not exist" (returns NULL)
    DP_REQUIRE(integration_method == 0 || integration_method == 1, DP_ERR_INVALID,
               "Integration method selected does not exist (%d)", integration_method);
    DP_REQUIRE(weight_mode == DP_CORR_WEIGHT_REFERENCE || weight_mode == DP_CORR_WEIGHT_INTENDED, DP_ERR_INVALID,
               "dp_power_correlation: bad weight_mode %d", weight_mode);
    dp::CorrLayout lay = dp::corr_layout(T, S, F, step, integration_method);
    DP_REQUIRE(workspace_bytes >= lay.total, DP_ERR_WORKSPACE, "dp_power_correlation: workspace %zu < %zu", workspace_bytes, lay.total);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = (int)T;
    char* ws = (char*)workspace;
    double* d_freq = (double*)(ws + lay.off_freq);
    dp::cplx* W = (dp::cplx*)(ws + lay.off_W);
    dp::cplx* C = (dp::cplx*)(ws + lay.off_C);
    dp::cplx* D = (dp::cplx*)(ws + lay.off_D);
    dp::cplx* vt = (dp::cplx*)(ws + lay.off_vt);
    const double ndiv = (double)(N / step);               // integer division, c/correlation.c:177
    if (lay.nl_valid == 0) {
        // no lag has any origin: the reference returns 0 * dt / (N div step) (NaN when N < step)
        std::vector<double> host((size_t)F * S, (0.0 * dt / ndiv) * scale);
        DP_CUDA_OK(cudaMemcpyAsync(out, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        DP_CUDA_OK(cudaStreamSynchronize(st));
        return DP_OK;
    }
    DP_CUDA_OK(cudaMemcpyAsync(d_freq, freq, (size_t)F * sizeof(double), cudaMemcpyHostToDevice, st));
    {
        auto k = dp::corr_weights_kernel;
        int64_t total = (int64_t)lay.nl_valid * F;
        DP_LAUNCH(k, dim3((unsigned)dp::imin<int64_t>((total + 255) / 256, 65535)), dim3(256), 0, st,
                  (const double*)d_freq, F, lay.nl_valid, step, dt, integration_method, weight_mode, W);
        DP_CUDA_OK(cudaGetLastError());
    }
    size_t smem = ((size_t)2 * dp::CR_CHUNK + (size_t)(dp::CR_LAGS - 1) * step) * sizeof(dp::cplx);
    DP_REQUIRE(smem <= 200 * 1024, DP_ERR_UNSUPPORTED, "dp_power_correlation: step=%d too large", step);
    for (int s0 = 0; s0 < S; s0 += lay.chunk) {
        const int ns = dp::imin(lay.chunk, S - s0);
        {
            auto k = dp::transpose_series_kernel;
            dim3 grid((unsigned)((N + 31) / 32), (unsigned)((ns + 31) / 32));
            DP_LAUNCH(k, grid, dim3(256), 0, st, (const dp::cplx*)series, ld, N, s0, ns, vt);
            DP_CUDA_OK(cudaGetLastError());
        }
        {
            auto k = dp::lag_sums_kernel;
            DP_CUDA_OK(DP_SET_SMEM(k, smem));
            dim3 grid((unsigned)((lay.nl_valid + dp::CR_LAGS - 1) / dp::CR_LAGS), (unsigned)ns);
            DP_LAUNCH(k, grid, dim3(dp::CR_LAGS * dp::CR_JP), smem, st, (const dp::cplx*)vt, N, step, lay.nl_valid,
                      C + (size_t)s0 * lay.c_ld, lay.c_ld);
            DP_CUDA_OK(cudaGetLastError());
        }
        if (integration_method == 0) {
            auto k = dp::lag_sums_shifted_kernel;
            int64_t total = (int64_t)ns * lay.nl_valid;
            DP_LAUNCH(k, dim3((unsigned)dp::imin<int64_t>((total + 255) / 256, 65535)), dim3(256), 0, st,
                      (const dp::cplx*)vt, N, step, lay.nl_valid, ns, (const dp::cplx*)(C + (size_t)s0 * lay.c_ld),
                      D + (size_t)s0 * lay.c_ld, lay.c_ld);
            DP_CUDA_OK(cudaGetLastError());
        }
        {
            auto k = dp::corr_spectrum_kernel;
            int64_t total = (int64_t)ns * F;
            DP_LAUNCH(k, dim3((unsigned)dp::imin<int64_t>((total + 255) / 256, 65535)), dim3(256), 0, st,
                      (const dp::cplx*)(C + (size_t)s0 * lay.c_ld), (const dp::cplx*)(D + (size_t)s0 * lay.c_ld),
                      (const dp::cplx*)W, s0, ns, S, F, lay.nl_valid, lay.c_ld, integration_method, dt, ndiv, scale, out);
            DP_CUDA_OK(cudaGetLastError());
        }
    }
    return DP_OK;
}

}  // extern "C"
