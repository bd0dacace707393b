// A5: direct ("Fourier transform of the velocity autocorrelation") power spectrum, selector key 0.
// Reference: c/correlation.c:101-178 (correlation_par / EvaluateCorrelation) and
// dynaphopy/power_spectrum/__init__.py:57-75.
//
// The reference recomputes every lag product for every output frequency: O(F * N^2 / step).
// The products do not depend on the frequency, so this implementation factors the sum:
//   stage 1  C(i) = sum_{j < N-i-step} conj(v_j) v_{j+i}   for lags i = 0, step, 2 step, ...
//            -- the compute-bound fp64 FMA kernel (4 N^2 / step flop per series): `step` independent
//            autocorrelations of the decimated sequences v[r + n*step], register-tiled (lag_sums_kernel);
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

// Decimated layout of one series: u_r[n] = v[r + n*step], r in [0, step), stored as ud[s][r][n] with row pitch NP.
// Time-major input (stride_s == 1, the reference's (T, S) array): 32 x 32 tiles through shared memory so that both the
// read (consecutive series) and the write (consecutive n) are coalesced.
__global__ void __launch_bounds__(256)
decimate_series_kernel(SeriesView x, int N, int s0, int ns, int step, int NP, cplx* __restrict__ ud) {
    __shared__ cplx tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    const int ntb = (N + 31) / 32;
    const int tb = (int)(blockIdx.x % ntb) * 32, sb = (int)(blockIdx.x / ntb) * 32;
    for (int r = ty; r < 32; r += 8) {
        int t = tb + r, s = sb + tx;
        if (t < N && s < ns) tile[r][tx] = *x.at(s0 + s, t);
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int s = sb + r, t = tb + tx;
        if (t < N && s < ns) ud[((size_t)s * step + t % step) * NP + t / step] = tile[tx][r];
    }
}
// Any other layout (series-major, blocked time axis): thread n of a series gathers u_r[n] for every residue r -- the
// `step` reads of a warp cover one contiguous span of the series (each line is fetched from DRAM once), the writes are
// contiguous in n.
__global__ void __launch_bounds__(256)
decimate_series_gather_kernel(SeriesView x, int N, int s0, int ns, int step, int NP, cplx* __restrict__ ud) {
    const int nn = (N + step - 1) / step;
    const int nchunks = (nn + 255) / 256;
    for (int64_t blk = blockIdx.x; blk < (int64_t)ns * nchunks; blk += gridDim.x) {
        const int s = (int)(blk / nchunks), n = (int)(blk % nchunks) * 256 + threadIdx.x;
        if (n >= nn) continue;
        for (int r = 0; r < step; r++) {
            const int t = r + n * step;
            if (t < N) ud[((size_t)s * step + r) * NP + n] = __ldg(x.at(s0 + s, t));
        }
    }
}

constexpr int CR_R = 8;           // lags per thread (register tile)
constexpr int CR_CH = 512;        // origins per staged chunk
constexpr int CR_MAXW = 4;        // warps per CTA (a CTA covers 32 * CR_R lags per warp)

// C(a) = sum_j conj(v_j) v_{j + a*step}, j < N - a*step - step  (c/correlation.c:153-157 summed once instead of
// once per frequency).  With j = r + n*step the sum is, for every residue r, the plain autocorrelation at lag a of
// the decimated sequence u_r WITHOUT its last element (partner index n + a <= len(u_r) - 2), so
//   C(a) = sum_r sum_n conj(u'_r[n]) u'_r[n + a],   u'_r zero beyond len_r = (N - 1 - r) / step.
// Register tiling: a thread owns CR_R consecutive lags and slides a CR_R-deep window of partners through registers:
// per origin one broadcast LDS.128 (the origin) + one LDS.128 (the new partner; 9/8 padding makes the 128-byte lane
// stride conflict-free) feed 4*CR_R = 32 DFMA, so the loop is bound by the fp64 pipe, not by shared memory.
__global__ void __launch_bounds__(32 * CR_MAXW)
lag_sums_kernel(const cplx* __restrict__ ud, int N, int step, int NP, int nl_valid, int nblk,
                cplx* __restrict__ C, int c_ld) {
    DP_DYNSMEM(cplx, sm);
    const int nthreads = blockDim.x, tid = threadIdx.x;
    const int LB = nthreads * CR_R;                 // lags per CTA
    cplx* xs = sm;                                  // origins  [CR_CH]
    cplx* ys = sm + CR_CH;                          // partners [CR_CH + LB], index k stored at k + k/8
    const int s = (int)(blockIdx.x / nblk);
    const int A = (int)(blockIdx.x % nblk) * LB;    // first lag of this CTA
    const int W = CR_CH + LB;
    const cplx* us = ud + (size_t)s * step * NP;
    const cplx zero = make_double2(0.0, 0.0);
    double ar[CR_R], ai[CR_R];
#pragma unroll
    for (int i = 0; i < CR_R; i++) { ar[i] = 0.0; ai[i] = 0.0; }
    for (int r = 0; r < step && r < N; r++) {
        const int len = (N - 1 - r) / step;         // u'_r
        const int cnt = len - A;                    // origins with a partner inside this CTA's lag range
        const cplx* u = us + (size_t)r * NP;
        for (int n0 = 0; n0 < cnt; n0 += CR_CH) {
            __syncthreads();
            for (int e = tid; e < CR_CH; e += nthreads) {
                const int j = n0 + e;
                xs[e] = j < len ? u[j] : zero;
            }
            for (int k = tid; k < W; k += nthreads) {
                const int j = n0 + A + k;
                ys[k + (k >> 3)] = j < len ? u[j] : zero;
            }
            __syncthreads();
            const int mend = (imin(CR_CH, cnt - n0) + 7) >> 3;
            const cplx* yp = ys + 9 * tid;
            cplx w[CR_R];
#pragma unroll
            for (int j = 0; j < CR_R; j++) w[j] = yp[j];
            for (int m = 0; m < mend; m++) {
                const cplx* xp = xs + 8 * m;
                const cplx* yn = yp + 9 * (m + 1);
#pragma unroll
                for (int kk = 0; kk < CR_R; kk++) {
                    const cplx x = xp[kk];
                    const double nxy = -x.y;
#pragma unroll
                    for (int i = 0; i < CR_R; i++) {
                        const cplx y = w[(kk + i) & (CR_R - 1)];
                        ar[i] = fma(x.x, y.x, fma(x.y, y.y, ar[i]));          // conj(x) * y
                        ai[i] = fma(x.x, y.y, fma(nxy, y.x, ai[i]));
                    }
                    w[kk] = yn[kk];
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < CR_R; i++) {
        const int a = A + tid * CR_R + i;
        if (a < nl_valid) C[(size_t)s * c_ld + a] = make_double2(ar[i], ai[i]);
    }
}

// D(a) = sum_{j < cnt_a} conj(v_j) v_{j + lag_a + step} = C(a+1) + the `step` trailing origins
__global__ void __launch_bounds__(256)
lag_sums_shifted_kernel(const cplx* __restrict__ ud, int N, int step, int NP, int nl_valid, int ns,
                        const cplx* __restrict__ C, cplx* __restrict__ D, int c_ld) {
    int64_t total = (int64_t)ns * nl_valid;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int s = (int)(e / nl_valid), a = (int)(e % nl_valid);
        const cplx* us = ud + (size_t)s * step * NP;
        int lag1 = (a + 1) * step;
        int cnt = N - a * step - step;
        int cnt1 = imax(cnt - step, 0);
        cplx acc = (a + 1 < nl_valid) ? C[(size_t)s * c_ld + a + 1] : make_double2(0.0, 0.0);
        for (int j = cnt1; j < cnt; j++) {
            const int k = j + lag1;
            acc = cadd(acc, cmul(cconj(us[(size_t)(j % step) * NP + j / step]), us[(size_t)(k % step) * NP + k / step]));
        }
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

constexpr int CS_SB = 4;          // series per thread of the weighting pass (one weight load serves CS_SB products)

// out[f][s] = Re sum_a (weight terms) * dt / (N div step): thread = (frequency, block of CS_SB series)
__global__ void __launch_bounds__(256)
corr_spectrum_kernel(const cplx* __restrict__ C, const cplx* __restrict__ D, const cplx* __restrict__ W,
                     int s0, int ns, int S, int F, int nl_valid, int c_ld, int method, double dt, double ndiv, double scale,
                     double* __restrict__ out) {
    const int nsb = (ns + CS_SB - 1) / CS_SB;
    int64_t total = (int64_t)nsb * F;
    const int64_t wsz = (int64_t)nl_valid * F;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int sb = (int)(e / F) * CS_SB, f = (int)(e % F);
        const cplx* c[CS_SB];
        const cplx* d[CS_SB];
        double ar[CS_SB];
#pragma unroll
        for (int i = 0; i < CS_SB; i++) {
            const int s = imin(sb + i, ns - 1);
            c[i] = C + (size_t)s * c_ld;
            d[i] = D + (size_t)s * c_ld;
            ar[i] = 0.0;
        }
        if (method == 0) {
            for (int a = 0; a < nl_valid; a++) {
                const cplx w0 = W[(int64_t)a * F + f], w1 = W[wsz + (int64_t)a * F + f], w2 = W[2 * wsz + (int64_t)a * F + f];
#pragma unroll
                for (int i = 0; i < CS_SB; i++) {
                    const cplx ci = c[i][a], di = d[i][a];
                    ar[i] += ci.x * w0.x - ci.y * w0.y;
                    ar[i] = ar[i] + (di.x * w1.x - di.y * w1.y) + (ci.x * w2.x - ci.y * w2.y);
                }
            }
        } else {
            for (int a = 0; a < nl_valid; a++) {
                const cplx w0 = W[(int64_t)a * F + f];
#pragma unroll
                for (int i = 0; i < CS_SB; i++) {
                    const cplx ci = c[i][a];
                    const double t = ci.x * w0.x - ci.y * w0.y;
                    ar[i] += t;                                    // the reference adds Correl_i twice (:158-159, :170-171)
                    ar[i] += t;
                }
            }
        }
        // creal(Correl) * TimeStep / (NumberOfData / Increment)   (c/correlation.c:177)
#pragma unroll
        for (int i = 0; i < CS_SB; i++)
            if (sb + i < ns) out[(size_t)f * S + s0 + sb + i] = __dmul_rn(__ddiv_rn(__dmul_rn(ar[i], dt), ndiv), scale);
    }
}

struct CorrLayout {
    size_t off_freq, off_W, off_C, off_D, off_ud, total;
    int nl_valid, c_ld, chunk, NP;
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
    // series are decimated chunk-wise into a [chunk][step][NP] slab of at most 1 GiB
    l.NP = ((N + step - 1) / step + 7) / 8 * 8;
    size_t per = (size_t)imin(step, N) * l.NP * sizeof(cplx);     // step >= N: no valid lag, the slab is not used
    l.chunk = (int)imin<size_t>((size_t)S, imax<size_t>(1, ((size_t)1 << 30) / per));
    l.off_ud = o;   o += align_up((size_t)l.chunk * per, 256);
    l.total = o;
    return l;
}

__global__ void __launch_bounds__(256) corr_fill_kernel(double* __restrict__ out, int64_t n, double value) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) out[e] = value;
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
    DP_REQUIRE(S > 0 && ld >= S, DP_ERR_INVALID, "dp_power_correlation: bad sizes");
    return dp_power_correlation_strided(series, T, S, 1, ld, 0, 0, dt, freq, F, step, integration_method, weight_mode, scale,
                                        out, workspace, workspace_bytes, stream);
}

int dp_power_correlation_strided(const double* series, int64_t T, int S, int64_t stride_s, int64_t stride_t,
                                 int64_t tb_len, int64_t tb_stride, double dt, const double* freq, int F, int step,
                                 int integration_method, int weight_mode, double scale, double* out, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    DP_REQUIRE(series && freq && out && workspace, DP_ERR_INVALID, "dp_power_correlation: null pointer");
    DP_REQUIRE(S > 0 && F > 0 && T >= 1 && step >= 1 && stride_t != 0 && tb_len >= 0 && tb_len < (1LL << 31), DP_ERR_INVALID,
               "dp_power_correlation: bad sizes");
    if (tb_len >= T) tb_len = 0;
    dp::SeriesView view;
    view.base = (const dp::cplx*)series; view.stride_s = stride_s; view.stride_t = stride_t; view.tb_len = (int)tb_len;
    view.tb_stride = tb_stride;
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
    dp::cplx* ud = (dp::cplx*)(ws + lay.off_ud);
    const double ndiv = (double)(N / step);               // integer division, c/correlation.c:177
    if (lay.nl_valid == 0) {
        // no lag has any origin: the reference returns 0 * dt / (N div step) (NaN when N < step)
        auto kf = dp::corr_fill_kernel;
        const int64_t total = (int64_t)F * S;
        DP_LAUNCH(kf, dim3((unsigned)dp::imin<int64_t>((total + 255) / 256, 1024)), dim3(256), 0, st, out, total,
                  (0.0 * dt / ndiv) * scale);
        DP_CUDA_OK(cudaGetLastError());
        return DP_OK;
    }
    {
        int rc = dp::upload_table(d_freq, freq, (size_t)F * sizeof(double), st);    // kernel parameters: no host sync
        if (rc) return rc;
    }
    {
        auto k = dp::corr_weights_kernel;
        int64_t total = (int64_t)lay.nl_valid * F;
        DP_LAUNCH(k, dim3((unsigned)dp::imin<int64_t>((total + 255) / 256, 65535)), dim3(256), 0, st,
                  (const double*)d_freq, F, lay.nl_valid, step, dt, integration_method, weight_mode, W);
        DP_CUDA_OK(cudaGetLastError());
    }
    // CTA = nw warps x 32 lanes x CR_R lags; small problems use fewer warps per CTA
    const int nw = dp::imax(1, dp::imin(dp::CR_MAXW, (lay.nl_valid + 32 * dp::CR_R - 1) / (32 * dp::CR_R)));
    const int lags_per_cta = nw * 32 * dp::CR_R;
    const int nblk = (lay.nl_valid + lags_per_cta - 1) / lags_per_cta;
    const size_t smem = ((size_t)dp::CR_CH + 9 * ((size_t)dp::CR_CH / 8 + nw * 32)) * sizeof(dp::cplx);
    for (int s0 = 0; s0 < S; s0 += lay.chunk) {
        const int ns = dp::imin(lay.chunk, S - s0);
        if (stride_s == 1) {
            auto k = dp::decimate_series_kernel;
            const int64_t nblocks = (int64_t)((N + 31) / 32) * ((ns + 31) / 32);
            DP_REQUIRE(nblocks < (1LL << 31), DP_ERR_UNSUPPORTED, "dp_power_correlation: chunk too large");
            DP_LAUNCH(k, dim3((unsigned)nblocks), dim3(256), 0, st, view, N, s0, ns, step, lay.NP, ud);
            DP_CUDA_OK(cudaGetLastError());
        } else {
            auto k = dp::decimate_series_gather_kernel;
            const int64_t nblocks = (int64_t)ns * (((N + step - 1) / step + 255) / 256);
            DP_LAUNCH(k, dim3((unsigned)dp::imin<int64_t>(nblocks, 1 << 20)), dim3(256), 0, st, view, N, s0, ns, step, lay.NP, ud);
            DP_CUDA_OK(cudaGetLastError());
        }
        {
            auto k = dp::lag_sums_kernel;
            DP_CUDA_OK(DP_SET_SMEM(k, smem));
            const int64_t nblocks = (int64_t)ns * nblk;
            DP_REQUIRE(nblocks < (1LL << 31), DP_ERR_UNSUPPORTED, "dp_power_correlation: chunk too large");
            DP_LAUNCH(k, dim3((unsigned)nblocks), dim3(nw * 32), smem, st, (const dp::cplx*)ud, N, step, lay.NP,
                      lay.nl_valid, nblk, C + (size_t)s0 * lay.c_ld, lay.c_ld);
            DP_CUDA_OK(cudaGetLastError());
        }
        if (integration_method == 0) {
            auto k = dp::lag_sums_shifted_kernel;
            int64_t total = (int64_t)ns * lay.nl_valid;
            DP_LAUNCH(k, dim3((unsigned)dp::imin<int64_t>((total + 255) / 256, 65535)), dim3(256), 0, st,
                      (const dp::cplx*)ud, N, step, lay.NP, lay.nl_valid, ns, (const dp::cplx*)(C + (size_t)s0 * lay.c_ld),
                      D + (size_t)s0 * lay.c_ld, lay.c_ld);
            DP_CUDA_OK(cudaGetLastError());
        }
        {
            auto k = dp::corr_spectrum_kernel;
            int64_t total = (int64_t)((ns + dp::CS_SB - 1) / dp::CS_SB) * F;
            DP_LAUNCH(k, dim3((unsigned)dp::imin<int64_t>((total + 255) / 256, 65535)), dim3(256), 0, st,
                      (const dp::cplx*)(C + (size_t)s0 * lay.c_ld), (const dp::cplx*)(D + (size_t)s0 * lay.c_ld),
                      (const dp::cplx*)W, s0, ns, S, F, lay.nl_valid, lay.c_ld, integration_method, dt, ndiv, scale, out);
            DP_CUDA_OK(cudaGetLastError());
        }
    }
    return DP_OK;
}

}  // extern "C"
