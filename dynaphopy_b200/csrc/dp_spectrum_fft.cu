// A7: FFT power spectrum (selector keys 2/3) -- reference
// dynaphopy/power_spectrum/__init__.py:226-270 (_numpy_power, get_fft_numpy_spectra) and :33-51
// (_division_of_data).
//
// Per window ("piece") of length L and per series the reference evaluates
//     a = np.correlate(x, x, 'same') / L          O(L^2)
//     S = |FFT_L(a)| * dt ; mean over windows ; np.interp onto the requested grid.
// Here one persistent CTA owns one (series, window) item and runs, with the hand-written
// engine of dp_fft.cuh,
//     X = FFT_P(x zero-padded to P >= L + L//2) ; |X|^2 ; IFFT_P      (linear autocorrelation)
//     gather lags [-L//2, L - L//2) ; FFT_L ; |.| * dt
// through a private scratch slab (P + L complex) that stays L2-resident; |X|^2 and the inverse
// row transforms are fused in shared memory.  A second kernel averages the windows in the
// reference's order and applies np.interp's exact arithmetic.
#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "dp_common.cuh"
#include "dp_fft.cuh"
#include "dp_spectrum_fft2.cuh"
#include "dp_spectrum_fft3.cuh"

namespace dp {

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------- host planning ---------------------------------------------------------------
struct SpecHostPlan {
    int L = 0, P = 0, half = 0;
    std::vector<int> starts;          // distinct windows actually transformed
    int n_avg = 0;                    // number of windows the reference averages (npieces + 1)
    bool duplicate_single = false;    // reference list is [[0,T],[0,T]]
    FftSplit sp, sl;
    int lo = 0, nkeep = 0;
    std::vector<int> idx0, mode;      // np.interp bracket (sorted-frequency index), 1 = copy fp[idx0]
    std::vector<double> dx, den;
    bool use_v2 = false;              // register-codelet fast path (dp_spectrum_fft2.cuh)
    S2Plan s2;
    bool use_v3 = false;              // long windows / wide bin ranges: three-factor batched passes (dp_spectrum_fft3.cuh)
    S3Plan s3;
};

// _division_of_data (power_spectrum/__init__.py:33-51), same floating-point expressions.
static int plan_pieces(int64_t T, double dt, double resolution, int* piece_len, std::vector<int>* starts,
                       bool* duplicate_single) {
    double inv = 1.0 / (dt * resolution);
    if (!(inv > 0.0) || !std::isfinite(inv) || inv > 2.0e9) {
        set_error("power_fft: bad resolution %g / time step %g", resolution, dt);
        return DP_ERR_INVALID;
    }
    long long piece = (long long)std::nearbyint(inv);      // Python round(): half to even
    if (piece < 1) { set_error("power_fft: window length rounds to 0"); return DP_ERR_INVALID; }
    long long npieces = (long long)((double)(T - 1) / (double)piece);
    double interval;
    *duplicate_single = false;
    if (npieces > 0) {
        interval = (double)(T - piece) / (double)npieces;
    } else {
        interval = 0.0;
        npieces = 1;
        piece = T;
        *duplicate_single = true;
    }
    starts->clear();
    for (long long i = 0; i <= npieces; i++) {
        double mid = (double)piece / 2.0 + (double)i * interval;
        long long ini = (long long)(mid - (double)piece / 2.0);
        long long fin = (long long)(mid + (double)piece / 2.0);
        if (fin - ini != piece || ini < 0 || fin > T) {
            set_error("power_fft: ragged window [%lld,%lld) for length %lld (the reference fails here too)",
                      ini, fin, piece);
            return DP_ERR_UNSUPPORTED;
        }
        starts->push_back((int)ini);
    }
    *piece_len = (int)piece;
    return DP_OK;
}

static int make_spec_plan(int64_t T, double dt, const double* freq, int F, SpecHostPlan* pl) {
    if (T < 2 || F < 2 || !(dt > 0)) { set_error("power_fft: need T >= 2, F >= 2, dt > 0"); return DP_ERR_INVALID; }
    if (T > (1 << 22)) { set_error("power_fft: T=%lld exceeds the supported 2^22", (long long)T); return DP_ERR_UNSUPPORTED; }
    int rc = plan_pieces(T, dt, freq[1] - freq[0], &pl->L, &pl->starts, &pl->duplicate_single);
    if (rc) return rc;
    pl->n_avg = (int)pl->starts.size();
    if (pl->duplicate_single) pl->starts.resize(1);     // (a + a) / 2 == a exactly
    const int L = pl->L;
    pl->half = L / 2;
    // np.interp plan on the sorted fftfreq axis xp[i] = (i - half) * val, val = 1/(L*dt)
    const double val = 1.0 / ((double)L * dt);
    auto xp = [&](long long i) { return (double)(i - pl->half) * val; };
    pl->idx0.resize(F); pl->mode.resize(F); pl->dx.resize(F); pl->den.resize(F);
    long long lo = L, hi = -1;
    for (int f = 0; f < F; f++) {
        double x = freq[f];
        long long j;
        if (std::isnan(x)) { set_error("power_fft: NaN in frequency grid"); return DP_ERR_INVALID; }
        if (x < xp(0)) j = -1;
        else if (x >= xp(L - 1)) j = L - 1;
        else {
            double g = std::floor(x / val) + (double)pl->half;
            j = (long long)g;
            if (j < 0) j = 0;
            if (j > L - 2) j = L - 2;
            while (j > 0 && xp(j) > x) j--;
            while (j < L - 2 && xp(j + 1) <= x) j++;
        }
        int mode = 0; long long i0 = j;
        if (j == -1) { i0 = 0; mode = 1; }
        else if (j >= L - 1) { i0 = L - 1; mode = 1; }
        else if (xp(j) == x) mode = 1;
        pl->idx0[f] = (int)i0; pl->mode[f] = mode;
        pl->dx[f] = mode ? 0.0 : x - xp(j);
        pl->den[f] = mode ? 1.0 : xp(j + 1) - xp(j);
        lo = std::min(lo, i0);
        hi = std::max(hi, mode ? i0 : i0 + 1);
    }
    pl->lo = (int)lo;
    pl->nkeep = (int)(hi - lo + 1);
    const char* force = dev_env("DP_SPECTRUM_ENGINE");      // "v1" forces the general engine (validation)
    pl->use_v2 = !(force && force[0] == 'v' && force[1] == '1') && make_s2_plan(L, pl->nkeep, &pl->s2);
    if (pl->use_v2) { pl->P = pl->s2.P; return DP_OK; }
    const bool force_v1 = force && force[0] == 'v' && force[1] == '1';
    pl->use_v3 = !force_v1 && make_s3_plan(L, pl->nkeep, &pl->s3);
    if (pl->use_v3) { pl->P = pl->s3.P; return DP_OK; }
    if (!make_fft_split(L, &pl->sl)) {
        set_error("power_fft: window length %d has a prime factor > %d or no usable split", L, FFT_MAX_RADIX);
        return DP_ERR_UNSUPPORTED;
    }
    pl->P = next_fast_length(L + L / 2);
    if (!make_fft_split(pl->P, &pl->sp)) { set_error("power_fft: internal: no split for P=%d", pl->P); return DP_ERR_UNSUPPORTED; }
    return DP_OK;
}

// ---------------- device side -------------------------------------------------------------------
struct SpecParams {
    const cplx* series; int64_t stride_s, stride_t; int tb_len; int64_t tb_stride; int S;
    int L, P, half, npieces;
    int w_first, w_count;     // windows [w_first, w_first + w_count) are transformed by this launch
    const int* starts;
    FftSplit sp, sl;
    const cplx *twP1, *twP2, *twPN, *twL1, *twL2, *twLN;
    cplx* scratch;            // gridDim.x slabs of (P + L)
    double* spec;             // [S][npieces][nkeep]
    int lo, nkeep;
    double dt, acf_scale;     // acf_scale = 1 / (P * L)
};

// |X|^2 then inverse transform of the same rows, all in shared memory
struct MidPowerInverse {
    const FftPlan* plan;
    const cplx* tw;
    int n2;
    __device__ __forceinline__ cplx* operator()(cplx* r, cplx* other, int nr, int tid, int nthreads) const {
        for (int e = tid; e < n2 * nr; e += nthreads) {
            cplx v = r[e];
            r[e] = make_double2(v.x * v.x + v.y * v.y, 0.0);
        }
        __syncthreads();
        return stockham<+1>(r, other, *plan, nr, tw, tid, nthreads);
    }
};

__global__ void __launch_bounds__(FFT_THREADS)
spectrum_items_kernel(SpecParams p) {
    DP_DYNSMEM(cplx, smem);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    cplx* A = p.scratch + (size_t)blockIdx.x * ((size_t)p.P + p.L);
    cplx* B = A + p.P;
    const int64_t nitems = (int64_t)p.S * p.w_count;
    for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int pc = p.w_first + (int)(item / p.S), s = (int)(item % p.S);
        const cplx* x0 = p.series + (int64_t)s * p.stride_s;
        const int t0 = p.starts[pc];
        const int64_t stride_t = p.stride_t, tb_stride = p.tb_stride;
        const int tb_len = p.tb_len;
        const int L = p.L, P = p.P, half = p.half;
        auto load_x = [=](int idx) -> cplx {
            if (idx >= L) return make_double2(0.0, 0.0);
            const int t = t0 + idx;
            if (tb_len == 0) return __ldg(x0 + (int64_t)t * stride_t);
            const int b = t / tb_len;
            return __ldg(x0 + (int64_t)b * tb_stride + (int64_t)(t - b * tb_len) * stride_t);
        };
        MidPowerInverse mid{&p.sp.p2, p.twP2, p.sp.n2};
        // ---- linear autocorrelation through FFT_P ----
        if (p.sp.n1 > 1) {
            const int n2 = p.sp.n2;
            fft_column_pass<-1, false>(p.sp, p.twP1, p.twPN, smem, load_x,
                                       [=](int k1, int c, cplx v) { A[(size_t)k1 * n2 + c] = v; }, tid, nthreads);
            __syncthreads();
            fft_row_pass<-1>(p.sp, p.twP2, smem, [=](int row, int j2) { return A[(size_t)row * n2 + j2]; },
                             [=](int row, int k2, cplx v) { A[(size_t)row * n2 + k2] = v; }, mid, tid, nthreads);
            __syncthreads();
            fft_column_pass<+1, true>(p.sp, p.twP1, p.twPN, smem, [=](int idx) { return A[idx]; },
                                      [=](int j1, int c, cplx v) { A[(size_t)j1 * n2 + c] = v; }, tid, nthreads);
        } else {
            fft_row_pass<-1>(p.sp, p.twP2, smem, [=](int, int j2) { return load_x(j2); },
                             [=](int, int k2, cplx v) { A[k2] = v; }, mid, tid, nthreads);
        }
        __syncthreads();
        // ---- FFT_L of the gathered lags, |.| * dt into the kept sorted-frequency range ----
        const double sc = p.acf_scale;
        auto load_acf = [=](int m) -> cplx {
            int k = m - half;
            if (k < 0) k += P;
            cplx v = A[k];
            return make_double2(v.x * sc, v.y * sc);
        };
        double* spec = p.spec + ((size_t)s * p.npieces + pc) * p.nkeep;
        const int n1L = p.sl.n1, lo = p.lo, nkeep = p.nkeep;
        const double dt = p.dt;
        auto store_bin = [=](int k1, int k2, cplx v) {
            int bin = k1 + n1L * k2;
            int i = bin + half;
            if (i >= L) i -= L;
            i -= lo;
            if (i >= 0 && i < nkeep) spec[i] = __dmul_rn(hypot(v.x, v.y), dt);
        };
        if (p.sl.n1 > 1) {
            const int n2 = p.sl.n2;
            fft_column_pass<-1, false>(p.sl, p.twL1, p.twLN, smem, load_acf,
                                       [=](int k1, int c, cplx v) { B[(size_t)k1 * n2 + c] = v; }, tid, nthreads);
            __syncthreads();
            fft_row_pass<-1>(p.sl, p.twL2, smem, [=](int row, int j2) { return B[(size_t)row * n2 + j2]; },
                             store_bin, MidIdentity(), tid, nthreads);
        } else {
            fft_row_pass<-1>(p.sl, p.twL2, smem, [=](int, int j2) { return load_acf(j2); }, store_bin,
                             MidIdentity(), tid, nthreads);
        }
        __syncthreads();
    }
}

// mean over windows (sequential, np.add.reduce order) + np.interp arithmetic + unit scale.
// The window spectra are series-major ([S][npieces][nkeep]), the result is frequency-major ((F, S), the reference's
// layout): a CTA works on tiles of 32 series x 32 frequencies -- lanes along the frequency axis while the window
// spectra are read (contiguous bins), along the series axis while the result is stored; the transpose goes through
// shared memory.  (One thread per element with the lanes along S read one 45 KB-strided value per lane: 2.7 ms at
// cfg 5-A instead of the ~0.3 ms the 1.6 GB of traffic need.)
__global__ void __launch_bounds__(256)
spectrum_finalize_kernel(const double* __restrict__ spec, int S, int npieces, int n_avg, int nkeep, int lo,
                         const int* __restrict__ idx0, const int* __restrict__ mode,
                         const double* __restrict__ dx, const double* __restrict__ den, int F, double scale,
                         double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;             // 8 rows of 32 lanes
    const int ts = (S + 31) / 32, tf = (F + 31) / 32;
    for (int64_t t = blockIdx.x; t < (int64_t)ts * tf; t += gridDim.x) {
        const int s0 = (int)(t / tf) * 32, f0 = (int)(t % tf) * 32;
        const int f = f0 + lane;
        int i = 0, md = 1;
        double dxf = 0.0, denf = 1.0;
        if (f < F) { i = idx0[f] - lo; md = mode[f]; dxf = dx[f]; denf = den[f]; }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int s = s0 + row + 8 * k;
            if (s < S && f < F) {
                const double* base = spec + (size_t)s * npieces * nkeep;
                auto mean_at = [&](int ii) {
                    double acc = base[ii];
                    if (n_avg != npieces) {                 // duplicated single window: (a + a) / 2
                        acc = __dadd_rn(acc, acc);
                    } else {
                        for (int pc = 1; pc < npieces; pc++) acc = __dadd_rn(acc, base[(size_t)pc * nkeep + ii]);
                    }
                    return __ddiv_rn(acc, (double)n_avg);
                };
                double f0v = mean_at(i), res;
                if (md) res = f0v;
                else {
                    double f1v = mean_at(i + 1);
                    double slope = __ddiv_rn(__dsub_rn(f1v, f0v), denf);
                    res = __dadd_rn(__dmul_rn(slope, dxf), f0v);
                }
                tile[row + 8 * k][lane] = __dmul_rn(res, scale);
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int ff = f0 + row + 8 * k, s = s0 + lane;
            if (ff < F && s < S) out[(size_t)ff * S + s] = tile[lane][row + 8 * k];
        }
        __syncthreads();
    }
}

struct SpecLayout {
    size_t off_tw[6], off_starts, off_idx0, off_mode, off_dx, off_den, off_spec, off_scratch, off_claim, total;
    size_t slab_elems;
    int nslots;
};

static SpecLayout spec_layout(const SpecHostPlan& pl, int S, int F) {
    SpecLayout l;
    size_t o = 0;
    // v2: w256, wsub, fine, coarse (256 each), unused, w_L table
    int lens_v1[6] = {pl.sp.n1, pl.sp.n2, pl.P, pl.sl.n1, pl.sl.n2, pl.L};
    int lens_v2[6] = {6 * 256, 1, 1, 1, 1, pl.L};
    int lens_v3[6] = {S3_TAB_BLOCKS * 256, 1, 1, 1, 1, pl.L};
    const int* lens = pl.use_v2 ? lens_v2 : pl.use_v3 ? lens_v3 : lens_v1;
    for (int i = 0; i < 6; i++) { l.off_tw[i] = o; o += align_up((size_t)lens[i] * sizeof(cplx), 256); }
    l.off_starts = o; o += align_up(pl.starts.size() * sizeof(int), 256);
    l.off_idx0 = o;   o += align_up((size_t)F * sizeof(int), 256);
    l.off_mode = o;   o += align_up((size_t)F * sizeof(int), 256);
    l.off_dx = o;     o += align_up((size_t)F * sizeof(double), 256);
    l.off_den = o;    o += align_up((size_t)F * sizeof(double), 256);
    l.off_spec = o;   o += align_up((size_t)S * pl.starts.size() * pl.nkeep * sizeof(double), 256);
    int64_t items = (int64_t)S * (int64_t)(pl.use_v2 ? (pl.starts.size() + 1) / 2 : pl.starts.size());
    l.nslots = (int)imin<int64_t>(items, (int64_t)imax(sm_count(), 1));
    if (dev_env("DP_S2_GRID") && atoi(dev_env("DP_S2_GRID")) > 0)      // development aid: fewer persistent CTAs
        l.nslots = (int)imin<int64_t>(l.nslots, atoi(dev_env("DP_S2_GRID")));
    // v2: A[P] complex (+ Sr[P] real when |X_a|^2 of a window pair cannot be parked in the SM's 256 KB of tensor memory)
    l.slab_elems = pl.use_v2 ? (size_t)pl.P + (s2_parks_in_tmem(pl.s2) ? 0 : (size_t)pl.P / 2) : (size_t)pl.P + pl.L;
    if (pl.use_v3) {
        // a batch of series passes through the seven grid-wide passes together; its slabs (P complex per series) are
        // sized to stay in L2 between the passes, the C arrays of pass 6 (L complex per series) follow them
        int64_t batch = imax<int64_t>(1, (int64_t)(64u << 20) / ((int64_t)pl.P * (int64_t)sizeof(cplx)));
        if (dev_env("DP_S3_BATCH") && atoi(dev_env("DP_S3_BATCH")) > 0) batch = atoi(dev_env("DP_S3_BATCH"));
        l.nslots = (int)imin<int64_t>(S, batch);
    }
    l.off_scratch = o; o += align_up((size_t)l.nslots * l.slab_elems * sizeof(cplx), 256);
    l.off_claim = o; o += align_up((size_t)imax(l.nslots, 1) * sizeof(int), 256);      // slab free list (preemptible form)
    l.total = o;
    return l;
}

static int launch_twiddles(cplx* tw, int n, cudaStream_t st) {
    auto k = twiddle_table_kernel;
    DP_LAUNCH(k, dim3((unsigned)imin((n + 255) / 256, 1024)), dim3(256), 0, st, tw, n);
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

}  // namespace dp

extern "C" {

size_t dp_power_fft_workspace_bytes(int64_t T, int S, double dt, const double* freq, int F) {
    dp::SpecHostPlan pl;
    if (!freq || S <= 0 || dp::make_spec_plan(T, dt, freq, F, &pl) != DP_OK) return 0;
    return dp::spec_layout(pl, S, F).total;
}

int dp_power_fft_pieces(int64_t T, double dt, const double* freq, int F, int* piece_len, int* n_pieces,
                        int* starts, int max_pieces) {
    DP_REQUIRE(freq && piece_len && n_pieces, DP_ERR_INVALID, "dp_power_fft_pieces: null pointer");
    DP_REQUIRE(F >= 2, DP_ERR_INVALID, "dp_power_fft_pieces: need F >= 2");
    std::vector<int> st;
    bool dup;
    int rc = dp::plan_pieces(T, dt, freq[1] - freq[0], piece_len, &st, &dup);
    if (rc) return rc;
    *n_pieces = (int)st.size();
    for (int i = 0; i < (int)st.size() && i < max_pieces && starts; i++) starts[i] = st[i];
    return DP_OK;
}

int dp_power_fft_engine(int64_t T, double dt, const double* freq, int F) {
    dp::SpecHostPlan pl;
    if (!freq) return -DP_ERR_INVALID;
    int rc = dp::make_spec_plan(T, dt, freq, F, &pl);
    if (rc) return -rc;
    return pl.use_v2 ? 2 : pl.use_v3 ? 3 : 1;
}

int dp_power_fft_windows(const double* series, int64_t T, int S, int64_t stride_s, int64_t stride_t,
                         int64_t tb_len, int64_t tb_stride, double dt, const double* freq, int F, int first_window,
                         int n_windows, void* workspace, size_t workspace_bytes, void* stream) {
    return dp_power_fft_windows_ex(series, T, S, stride_s, stride_t, tb_len, tb_stride, dt, freq, F, first_window, n_windows,
                                   0, workspace, workspace_bytes, stream);
}

static int power_fft_windows_impl(const double* series, int64_t T, int S, int64_t stride_s, int64_t stride_t,
                                  int64_t tb_len, int64_t tb_stride, double dt, const double* freq, int F, int first_window,
                                  int n_windows, int max_ctas, void* workspace, size_t workspace_bytes, void* stream,
                                  bool preemptible);

int dp_power_fft_windows_ex(const double* series, int64_t T, int S, int64_t stride_s, int64_t stride_t,
                            int64_t tb_len, int64_t tb_stride, double dt, const double* freq, int F, int first_window,
                            int n_windows, int max_ctas, void* workspace, size_t workspace_bytes, void* stream) {
    return power_fft_windows_impl(series, T, S, stride_s, stride_t, tb_len, tb_stride, dt, freq, F, first_window, n_windows,
                                  max_ctas, workspace, workspace_bytes, stream, false);
}

int dp_power_fft_windows_preemptible(const double* series, int64_t T, int S, int64_t stride_s, int64_t stride_t,
                                     int64_t tb_len, int64_t tb_stride, double dt, const double* freq, int F,
                                     int first_window, int n_windows, int max_ctas, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    return power_fft_windows_impl(series, T, S, stride_s, stride_t, tb_len, tb_stride, dt, freq, F, first_window, n_windows,
                                  max_ctas, workspace, workspace_bytes, stream, true);
}

static int power_fft_windows_impl(const double* series, int64_t T, int S, int64_t stride_s, int64_t stride_t,
                                  int64_t tb_len, int64_t tb_stride, double dt, const double* freq, int F, int first_window,
                                  int n_windows, int max_ctas, void* workspace, size_t workspace_bytes, void* stream,
                                  bool preemptible) {
    DP_REQUIRE(series && freq && workspace, DP_ERR_INVALID, "dp_power_fft: null pointer");
    DP_REQUIRE(S > 0 && stride_t != 0, DP_ERR_INVALID, "dp_power_fft: bad S=%d stride_t=%lld", S, (long long)stride_t);
    DP_REQUIRE(tb_len >= 0 && tb_len < (1LL << 31), DP_ERR_INVALID, "dp_power_fft: bad time-block length %lld", (long long)tb_len);
    if (tb_len >= T) tb_len = 0;
    dp::SpecHostPlan pl;
    int rc = dp::make_spec_plan(T, dt, freq, F, &pl);
    if (rc) return rc;
    const int npieces = (int)pl.starts.size();
    if (n_windows < 0) n_windows = npieces - first_window;
    DP_REQUIRE(first_window >= 0 && n_windows >= 0 && first_window + n_windows <= npieces, DP_ERR_INVALID,
               "dp_power_fft_windows: windows [%d, %d) outside the %d distinct windows", first_window,
               first_window + n_windows, npieces);
    dp::SpecLayout lay = dp::spec_layout(pl, S, F);
    DP_REQUIRE(workspace_bytes >= lay.total, DP_ERR_WORKSPACE, "dp_power_fft: workspace %zu < %zu", workspace_bytes, lay.total);
    if (n_windows == 0) return DP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    dp::cplx* tw[6];
    for (int i = 0; i < 6; i++) tw[i] = (dp::cplx*)(ws + lay.off_tw[i]);
    // tables and window starts are (re)written by every call: deterministic, and stream-ordered with the kernels
    if (pl.use_v3) {
        auto kt = dp::s3_tables_kernel;
        DP_LAUNCH(kt, dim3(1), dim3(256), 0, st, tw[0], pl.s3.m1, pl.s3.m2, pl.s3.P);
        DP_CUDA_OK(cudaGetLastError());
        rc = dp::launch_twiddles(tw[5], pl.L, st);
        if (rc) return rc;
        dp::S3Params p;
        p.series = (const dp::cplx*)series; p.stride_s = stride_s; p.stride_t = stride_t;
        p.tb_len = (int)tb_len; p.tb_stride = tb_stride;
        p.tb_shift = -1;
        if (tb_len > 0 && (tb_len & (tb_len - 1)) == 0) { p.tb_shift = 0; while ((1LL << p.tb_shift) < tb_len) p.tb_shift++; }
        p.pl = pl.s3;
        p.g_tab = tw[0]; p.g_twL = tw[5];
        p.slab = (dp::cplx*)(ws + lay.off_scratch);
        p.cbuf = p.slab + (size_t)lay.nslots * pl.P;
        p.spec_stride = (int64_t)npieces * pl.nkeep;
        p.lo = pl.lo; p.nkeep = pl.nkeep;
        p.dt = dt; p.acf_scale = 1.0 / ((double)pl.P * (double)pl.L);
        for (int s0 = 0; s0 < S; s0 += lay.nslots) {
            p.s0 = s0; p.bs = dp::imin(lay.nslots, S - s0);
            for (int w = first_window; w < first_window + n_windows; w++) {
                p.t0 = pl.starts[w];
                p.spec = (double*)(ws + lay.off_spec) + ((size_t)s0 * npieces + w) * pl.nkeep;
                if ((rc = dp::s3_run_window(p, max_ctas, st))) return rc;
            }
        }
        return DP_OK;
    }
    if (pl.use_v2) {
        auto kt = dp::s2_tables_kernel;
        int a1, a2, b1, b2;
        dp::s2_radices(pl.s2.n1, &a1, &a2);
        dp::s2_radices(pl.s2.n2, &b1, &b2);
        DP_LAUNCH(kt, dim3(1), dim3(256), 0, st, tw[0], a1, a2, b1, b2, pl.s2.m1, pl.s2.m2, pl.s2.P);
        DP_CUDA_OK(cudaGetLastError());
        rc = dp::launch_twiddles(tw[5], pl.L, st);
        if (rc) return rc;
    } else {
        int lens[6] = {pl.sp.n1, pl.sp.n2, pl.P, pl.sl.n1, pl.sl.n2, pl.L};
        for (int i = 0; i < 6; i++) {
            rc = dp::launch_twiddles(tw[i], lens[i], st);
            if (rc) return rc;
        }
    }
    // window starts travel as kernel parameters: this entry point never blocks the host, so a caller can queue the
    // spectra of finished windows behind work that has not run yet
    rc = dp::upload_small(ws + lay.off_starts, pl.starts.data(), pl.starts.size() * sizeof(int), st);
    if (rc) return rc;
    double* spec = (double*)(ws + lay.off_spec);
    const int64_t items = (int64_t)S * (pl.use_v2 ? (n_windows + 1) / 2 : n_windows);
    int grid = (int)dp::imin<int64_t>(items, lay.nslots);
    // leave SMs to work that runs concurrently on other streams: a fixed share of the grid, or (slab engine, preemptible
    // form) one short-lived CTA per item, which lets the block scheduler honour stream priorities between items
    // (only when every CTA that can be resident at once finds a slab: one CTA fits an SM, so min(items, SMs) slabs)
    const bool oneshot = preemptible && pl.use_v2 && items <= 0x7fffffff &&
                         (int64_t)lay.nslots >= dp::imin<int64_t>(items, (int64_t)dp::imax(dp::sm_count(), 1));
    if (max_ctas > 0 && !oneshot) grid = dp::imin(grid, max_ctas);
    const int n_slabs = grid;                                   // slabs (and rows of the tensor maps) behind the launch
    if (pl.use_v2) {
        dp::S2Params p;
        p.series = (const dp::cplx*)series; p.stride_s = stride_s; p.stride_t = stride_t;
        p.tb_len = (int)tb_len; p.tb_stride = tb_stride; p.S = S; p.npieces = npieces;
        p.w_first = first_window; p.w_count = n_windows;
        p.tb_shift = -1;
        if (tb_len > 0 && (tb_len & (tb_len - 1)) == 0) { p.tb_shift = 0; while ((1LL << p.tb_shift) < tb_len) p.tb_shift++; }
        p.starts = (const int*)(ws + lay.off_starts);
        p.pl = pl.s2;
        p.g_tables = tw[0]; p.g_twL = tw[5];
        p.scratch = (dp::cplx*)(ws + lay.off_scratch); p.slab_elems = lay.slab_elems;
        p.spec = spec; p.lo = pl.lo; p.nkeep = pl.nkeep;
        p.dt = dt; p.acf_scale = 1.0 / ((double)pl.P * (double)pl.L);
        // bulk stores of passes 1-3 (dp_spectrum_fft2.cuh): the slabs as one 2-D tensor, one box per warp tile
        {
            const int n1 = pl.s2.n1, n2 = pl.s2.n2;
            const int tg = n1 == 16 ? 1 : (n1 == 32 || n1 == 64) ? 4 : n1 == 128 ? 8 : 16;   // SubCfg::TG of C16 .. C256
            const int ns = 32 / tg;                            // columns per warp tile
            p.slab_rows = (int)(lay.slab_elems / (size_t)n2);
            p.r_lo = dp::imin(n1, pl.s2.half / n2 + 1);
            p.r_hi = dp::imax(p.r_lo, (pl.s2.P - pl.s2.half) / n2);
            const uint64_t rows = (uint64_t)n_slabs * (uint64_t)p.slab_rows;
            const uint64_t pitch = (uint64_t)n2 * sizeof(dp::cplx);
            rc = dp::make_tensor_map_2d(&p.tm_col, p.scratch, 2ull * n2, rows, pitch, 2u * ns, (uint32_t)n1);
            if (rc) return rc;
            rc = dp::make_tensor_map_2d(&p.tm_lo, p.scratch, 2ull * n2, rows, pitch, 2u * ns, (uint32_t)p.r_lo);
            if (rc) return rc;
            rc = dp::make_tensor_map_2d(&p.tm_hi, p.scratch, 2ull * n2, rows, pitch, 2u * ns,
                                        (uint32_t)dp::imax(1, n1 - p.r_hi));
            if (rc) return rc;
        }
        p.park = dp::s2_parks_in_tmem(pl.s2) ? 1 : 0;
        p.park_cols = 32;
        while (p.park_cols < pl.s2.P / 64) p.park_cols *= 2;       // P doubles over 128 lanes: P / 128 * 2 columns
        p.skip_mask = dp::dev_env("DP_S2_SKIP") ? atoi(dp::dev_env("DP_S2_SKIP")) : 0;
        p.pair = (n_windows >= 2 && !(dp::dev_env("DP_S2_PAIR") && atoi(dp::dev_env("DP_S2_PAIR")) == 0)) ? 1 : 0;
        p.slab_busy = nullptr; p.n_slabs = n_slabs;
        if (oneshot) {
            p.slab_busy = (int*)(ws + lay.off_claim);
            DP_CUDA_OK(cudaMemsetAsync(p.slab_busy, 0, (size_t)n_slabs * sizeof(int), st));
            // a CTA works on S2_PREEMPT_ITEMS items (stride = grid) before it leaves its SM.  Measured on an idle B200
            // (tools/preempt_overhead.py, 1480 series x 7 windows; persistent form 4.74 ms): 1 item per CTA 4.87 ms, 4 items
            // 4.89 ms -- with the slab chosen by SM number; with slabs handed out by block index (an SM keeps changing
            // slabs, i.e. L2 lines) 5.07 / 5.11-5.21 ms.
            grid = (int)dp::imax<int64_t>(n_slabs, (items + dp::S2_PREEMPT_ITEMS - 1) / dp::S2_PREEMPT_ITEMS);
            grid = (int)dp::imin<int64_t>(grid, items);
        }
        auto k = dp::spectrum2_kernel;
        DP_CUDA_OK(DP_SET_SMEM(k, dp::S2_SMEM_BYTES));
        DP_LAUNCH(k, dim3(grid), dim3(dp::S2_THREADS), dp::S2_SMEM_BYTES, st, p);
        DP_CUDA_OK(cudaGetLastError());
    } else {
        dp::SpecParams p;
        p.series = (const dp::cplx*)series; p.stride_s = stride_s; p.stride_t = stride_t;
        p.tb_len = (int)tb_len; p.tb_stride = tb_stride; p.S = S;
        p.L = pl.L; p.P = pl.P; p.half = pl.half; p.npieces = npieces;
        p.w_first = first_window; p.w_count = n_windows;
        p.starts = (const int*)(ws + lay.off_starts);
        p.sp = pl.sp; p.sl = pl.sl;
        p.twP1 = tw[0]; p.twP2 = tw[1]; p.twPN = tw[2]; p.twL1 = tw[3]; p.twL2 = tw[4]; p.twLN = tw[5];
        p.scratch = (dp::cplx*)(ws + lay.off_scratch);
        p.spec = spec;
        p.lo = pl.lo; p.nkeep = pl.nkeep;
        p.dt = dt; p.acf_scale = 1.0 / ((double)pl.P * (double)pl.L);
        size_t smem = std::max(dp::fft_split_smem_elems(pl.sp), dp::fft_split_smem_elems(pl.sl)) * sizeof(dp::cplx);
        auto k = dp::spectrum_items_kernel;
        DP_CUDA_OK(DP_SET_SMEM(k, smem));
        DP_LAUNCH(k, dim3(grid), dim3(dp::FFT_THREADS), smem, st, p);
        DP_CUDA_OK(cudaGetLastError());
    }
    return DP_OK;
}

int dp_power_fft_finish(int64_t T, int S, double dt, const double* freq, int F, double scale, double* out,
                        void* workspace, size_t workspace_bytes, void* stream) {
    DP_REQUIRE(freq && out && workspace && S > 0, DP_ERR_INVALID, "dp_power_fft_finish: null pointer");
    dp::SpecHostPlan pl;
    int rc = dp::make_spec_plan(T, dt, freq, F, &pl);
    if (rc) return rc;
    dp::SpecLayout lay = dp::spec_layout(pl, S, F);
    DP_REQUIRE(workspace_bytes >= lay.total, DP_ERR_WORKSPACE, "dp_power_fft: workspace %zu < %zu", workspace_bytes, lay.total);
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    // interpolation plan: small tables as kernel parameters -- no host synchronisation
    if ((rc = dp::upload_table(ws + lay.off_idx0, pl.idx0.data(), (size_t)F * sizeof(int), st))) return rc;
    if ((rc = dp::upload_table(ws + lay.off_mode, pl.mode.data(), (size_t)F * sizeof(int), st))) return rc;
    if ((rc = dp::upload_table(ws + lay.off_dx, pl.dx.data(), (size_t)F * sizeof(double), st))) return rc;
    if ((rc = dp::upload_table(ws + lay.off_den, pl.den.data(), (size_t)F * sizeof(double), st))) return rc;
    const int npieces = (int)pl.starts.size();
    auto kf = dp::spectrum_finalize_kernel;
    const int64_t tiles = (int64_t)((S + 31) / 32) * ((F + 31) / 32);
    DP_LAUNCH(kf, dim3((unsigned)dp::imin<int64_t>(tiles, 16 * 148)), dim3(256), 0, st,
              (const double*)(ws + lay.off_spec), S, npieces, pl.n_avg, pl.nkeep, pl.lo, (const int*)(ws + lay.off_idx0),
              (const int*)(ws + lay.off_mode), (const double*)(ws + lay.off_dx), (const double*)(ws + lay.off_den), F,
              scale, out);
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

int dp_power_fft_strided(const double* series, int64_t T, int S, int64_t stride_s, int64_t stride_t,
                         int64_t tb_len, int64_t tb_stride, double dt, const double* freq, int F, double scale,
                         double* out, void* workspace, size_t workspace_bytes, void* stream) {
    DP_REQUIRE(out, DP_ERR_INVALID, "dp_power_fft: null pointer");
    int rc = dp_power_fft_windows(series, T, S, stride_s, stride_t, tb_len, tb_stride, dt, freq, F, 0, -1, workspace,
                                  workspace_bytes, stream);
    if (rc) return rc;
    return dp_power_fft_finish(T, S, dt, freq, F, scale, out, workspace, workspace_bytes, stream);
}

int dp_power_fft(const double* series, int64_t T, int S, int64_t ld, double dt, const double* freq, int F,
                 double scale, double* out, void* workspace, size_t workspace_bytes, void* stream) {
    DP_REQUIRE(S > 0 && ld >= S, DP_ERR_INVALID, "dp_power_fft: bad S=%d ld=%lld", S, (long long)ld);
    return dp_power_fft_strided(series, T, S, 1, ld, 0, 0, dt, freq, F, scale, out, workspace, workspace_bytes, stream);
}

}  // extern "C"

// ---- plain batched transform through the same engine (validation / benchmarking) ---------------
namespace dp {
struct C2cParams {
    const cplx* in; cplx* out; cplx* scratch;
    FftSplit sp; const cplx *tw1, *tw2, *twN; int batch;
};

template <int DIR>
__global__ void __launch_bounds__(FFT_THREADS) c2c_kernel(C2cParams p) {
    DP_DYNSMEM(cplx, smem);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int N = p.sp.N, n1 = p.sp.n1, n2 = p.sp.n2;
    for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
        const cplx* x = p.in + (size_t)b * N;
        cplx* y = p.out + (size_t)b * N;
        cplx* A = p.scratch + (size_t)blockIdx.x * N;
        if (n1 > 1) {
            fft_column_pass<DIR, false>(p.sp, p.tw1, p.twN, smem, [=](int idx) { return x[idx]; },
                                        [=](int k1, int c, cplx v) { A[(size_t)k1 * n2 + c] = v; }, tid, nthreads);
            __syncthreads();
            fft_row_pass<DIR>(p.sp, p.tw2, smem, [=](int row, int j2) { return A[(size_t)row * n2 + j2]; },
                              [=](int k1, int k2, cplx v) { y[k1 + (size_t)n1 * k2] = v; }, MidIdentity(), tid, nthreads);
        } else {
            fft_row_pass<DIR>(p.sp, p.tw2, smem, [=](int, int j2) { return x[j2]; },
                              [=](int, int k2, cplx v) { y[k2] = v; }, MidIdentity(), tid, nthreads);
        }
        __syncthreads();
    }
}
}  // namespace dp

extern "C" {

size_t dp_fft_c2c_workspace_bytes(int n, int batch) {
    dp::FftSplit sp;
    if (n < 1 || batch < 1 || !dp::make_fft_split(n, &sp)) return 0;
    size_t tw = dp::align_up((size_t)sp.n1 * 16, 256) + dp::align_up((size_t)sp.n2 * 16, 256) + dp::align_up((size_t)n * 16, 256);
    int slots = dp::imin(batch, dp::imax(dp::sm_count(), 1));
    return tw + dp::align_up((size_t)slots * n * 16, 256);
}

int dp_fft_c2c(const double* in, double* out, int n, int batch, int direction, void* workspace,
               size_t workspace_bytes, void* stream) {
    DP_REQUIRE(in && out && workspace && in != out, DP_ERR_INVALID, "dp_fft_c2c: null or aliased pointer");
    DP_REQUIRE(n >= 1 && batch >= 1 && (direction == 1 || direction == -1), DP_ERR_INVALID, "dp_fft_c2c: bad arguments");
    dp::FftSplit sp;
    DP_REQUIRE(dp::make_fft_split(n, &sp), DP_ERR_UNSUPPORTED, "dp_fft_c2c: length %d unsupported (prime factor > %d?)", n, dp::FFT_MAX_RADIX);
    DP_REQUIRE(workspace_bytes >= dp_fft_c2c_workspace_bytes(n, batch), DP_ERR_WORKSPACE, "dp_fft_c2c: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    dp::C2cParams p;
    p.in = (const dp::cplx*)in; p.out = (dp::cplx*)out; p.sp = sp; p.batch = batch;
    dp::cplx* t1 = (dp::cplx*)ws; ws += dp::align_up((size_t)sp.n1 * 16, 256);
    dp::cplx* t2 = (dp::cplx*)ws; ws += dp::align_up((size_t)sp.n2 * 16, 256);
    dp::cplx* tN = (dp::cplx*)ws; ws += dp::align_up((size_t)n * 16, 256);
    p.tw1 = t1; p.tw2 = t2; p.twN = tN; p.scratch = (dp::cplx*)ws;
    int rc;
    if ((rc = dp::launch_twiddles(t1, sp.n1, st))) return rc;
    if ((rc = dp::launch_twiddles(t2, sp.n2, st))) return rc;
    if ((rc = dp::launch_twiddles(tN, n, st))) return rc;
    size_t smem = dp::fft_split_smem_elems(sp) * sizeof(dp::cplx);
    int slots = dp::imin(batch, dp::imax(dp::sm_count(), 1));
    if (direction < 0) {
        auto k = dp::c2c_kernel<-1>;
        DP_CUDA_OK(DP_SET_SMEM(k, smem));
        DP_LAUNCH(k, dim3(slots), dim3(dp::FFT_THREADS), smem, st, p);
    } else {
        auto k = dp::c2c_kernel<+1>;
        DP_CUDA_OK(DP_SET_SMEM(k, smem));
        DP_LAUNCH(k, dim3(slots), dim3(dp::FFT_THREADS), smem, st, p);
    }
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

}  // extern "C"
