// Hand-written mixed-radix FFT engine (complex fp64) used by the FFT power spectrum.
//
// Building blocks, all executed cooperatively by one CTA:
//   stockham()         autosort Stockham FFT of `batch` interleaved sequences resident in shared
//                      memory (radix 4/2/3/5 butterflies, generic prime radix up to 61)
//   fft_column_pass()  step 1 of the four-step factorisation N = n1*n2: n2/C tiles of C columns,
//                      each column an n1-point transform at stride n2, fused inter-step twiddle
//   fft_row_pass()     step 2: n1 contiguous rows of length n2, with a caller-supplied
//                      per-row functor so |X|^2 + inverse transform can stay in shared memory
// Index algebra (n = j1*n2 + j2, k = k1 + n1*k2):
//   X[k1 + n1*k2] = sum_j2 w_n2^(j2 k2) [ w_N^(j2 k1) sum_j1 x[j1*n2 + j2] w_n1^(j1 k1) ]
#pragma once

#include "dp_common.cuh"

namespace dp {

constexpr int FFT_MAX_STAGES = 20;
constexpr int FFT_MAX_RADIX = 61;
constexpr int FFT_ROW_MAX = 4096;     // longest transform held in shared memory
constexpr int FFT_COL_MAX = 512;      // longest column transform of the four-step split
constexpr int FFT_COL_TILE = 8;       // columns per tile: 8 * 16 B = one 128 B line per row
constexpr int FFT_THREADS = 512;

struct FftPlan {
    int n;
    int nstage;
    unsigned char radix[FFT_MAX_STAGES];
};

struct FftSplit {       // N = n1 * n2, n1 == 1 means single shared-memory transform
    int N, n1, n2;
    FftPlan p1, p2;
};

// ---------------- host-side planning ----------------------------------------------------------
inline bool make_fft_plan(int n, FftPlan* p) {
    p->n = n;
    p->nstage = 0;
    if (n < 1) return false;
    int rem = n;
    while (rem % 4 == 0) { if (p->nstage >= FFT_MAX_STAGES) return false; p->radix[p->nstage++] = 4; rem /= 4; }
    for (int f = 2; f <= FFT_MAX_RADIX && rem > 1; f++) {
        while (rem % f == 0) {
            if (p->nstage >= FFT_MAX_STAGES) return false;
            p->radix[p->nstage++] = (unsigned char)f;
            rem /= f;
        }
    }
    return rem == 1;
}

inline bool make_fft_split(int N, FftSplit* s) {
    s->N = N;
    if (N <= FFT_ROW_MAX) {
        s->n1 = 1; s->n2 = N;
        if (!make_fft_plan(1, &s->p1)) return false;
        return make_fft_plan(N, &s->p2);
    }
    int best = 0;
    double target = sqrt((double)N);
    for (int n1 = 2; n1 <= FFT_COL_MAX; n1++) {
        if (N % n1) continue;
        int n2 = N / n1;
        if (n2 > FFT_ROW_MAX || n2 < n1) continue;
        FftPlan a, b;
        if (!make_fft_plan(n1, &a) || !make_fft_plan(n2, &b)) continue;
        if (!best || fabs(n1 - target) < fabs(best - target)) best = n1;
    }
    if (!best) {
        for (int n1 = 2; n1 <= FFT_COL_MAX; n1++) {     // allow n2 < n1 as a last resort
            if (N % n1) continue;
            int n2 = N / n1;
            FftPlan a, b;
            if (n2 > FFT_ROW_MAX || !make_fft_plan(n1, &a) || !make_fft_plan(n2, &b)) continue;
            best = n1;
            break;
        }
    }
    if (!best) return false;
    s->n1 = best; s->n2 = N / best;
    return make_fft_plan(s->n1, &s->p1) && make_fft_plan(s->n2, &s->p2);
}

// shared-memory elements (cplx) one CTA needs for the passes of this split
inline size_t fft_split_smem_elems(const FftSplit& s) {
    size_t col = s.n1 > 1 ? 2 * (size_t)s.n1 * FFT_COL_TILE : 0;
    int rows = s.n2 >= 1024 ? 1 : imax(1, 1024 / s.n2);
    size_t row = 2 * (size_t)s.n2 * rows;
    return imax(col, row);
}
inline int fft_rows_per_tile(int n1, int n2) {
    int rows = n2 >= 1024 ? 1 : imax(1, 1024 / n2);
    return imin(rows, n1);
}

// smallest length >= minimum whose only prime factors are 2, 3, 5 and that splits
inline int next_fast_length(int minimum) {
    for (int n = imax(minimum, 1);; n++) {
        int r = n;
        while (r % 2 == 0) r /= 2;
        while (r % 3 == 0) r /= 3;
        while (r % 5 == 0) r /= 5;
        if (r != 1) continue;
        FftSplit s;
        if (make_fft_split(n, &s)) return n;
    }
}

// ---------------- device: twiddle tables ------------------------------------------------------
// tw[k] = exp(-2 pi i k / n), k in [0, n)
static __global__ void __launch_bounds__(256) twiddle_table_kernel(cplx* __restrict__ tw, int n) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        double s, c;
        sincospi(2.0 * (double)k / (double)n, &s, &c);
        tw[k] = make_double2(c, -s);
    }
}

// ---------------- device: butterflies -----------------------------------------------------------
// forward (DIR = -1): X[u] = sum_t a[t] exp(-2 pi i t u / R); inverse (DIR = +1): conjugate roots.
template <int DIR> __device__ __forceinline__ cplx mul_mi_dir(cplx a) {   // a * (DIR<0 ? -i : +i)
    return DIR < 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
}

template <int DIR> __device__ __forceinline__ void bfly2(cplx* a) {
    cplx t = a[0];
    a[0] = cadd(t, a[1]);
    a[1] = csub(t, a[1]);
}

template <int DIR> __device__ __forceinline__ void bfly4(cplx* a) {
    cplx t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
    cplx t2 = cadd(a[1], a[3]), t3 = mul_mi_dir<DIR>(csub(a[1], a[3]));
    a[0] = cadd(t0, t2);
    a[2] = csub(t0, t2);
    a[1] = cadd(t1, t3);
    a[3] = csub(t1, t3);
}

template <int DIR> __device__ __forceinline__ void bfly3(cplx* a) {
    const double s3 = 0.86602540378443864676;   // sin(2 pi / 3)
    cplx t = cadd(a[1], a[2]);
    cplx d = mul_mi_dir<DIR>(cscale(csub(a[1], a[2]), s3));
    cplx m = make_double2(a[0].x - 0.5 * t.x, a[0].y - 0.5 * t.y);
    a[0] = cadd(a[0], t);
    a[1] = cadd(m, d);
    a[2] = csub(m, d);
}

template <int DIR> __device__ __forceinline__ void bfly5(cplx* a) {
    const double c1 = 0.30901699437494742410;    // cos(2 pi / 5)
    const double c2 = -0.80901699437494742410;   // cos(4 pi / 5)
    const double s1 = 0.95105651629515357212;    // sin(2 pi / 5)
    const double s2 = 0.58778525229247312917;    // sin(4 pi / 5)
    cplx t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]);
    cplx d1 = csub(a[1], a[4]), d2 = csub(a[2], a[3]);
    cplx m1 = make_double2(a[0].x + c1 * t1.x + c2 * t2.x, a[0].y + c1 * t1.y + c2 * t2.y);
    cplx m2 = make_double2(a[0].x + c2 * t1.x + c1 * t2.x, a[0].y + c2 * t1.y + c1 * t2.y);
    cplx n1 = mul_mi_dir<DIR>(make_double2(s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y));
    cplx n2 = mul_mi_dir<DIR>(make_double2(s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y));
    a[0] = cadd(a[0], cadd(t1, t2));
    a[1] = cadd(m1, n1);
    a[4] = csub(m1, n1);
    a[2] = cadd(m2, n2);
    a[3] = csub(m2, n2);
}

template <int DIR> __device__ __forceinline__ cplx tw_dir(cplx w) {
    return DIR < 0 ? w : make_double2(w.x, -w.y);
}

// One Stockham stage of radix R over `batch` interleaved sequences.
//   x[(j + t*m)*batch + b]  ->  y[((j - q)*R + q + s*u)*batch + b],  q = j mod s, m = n / R
template <int R, int DIR>
__device__ __forceinline__ void stockham_stage(const cplx* __restrict__ x, cplx* __restrict__ y, int n, int s,
                                               int batch, const cplx* __restrict__ tw, int tid, int nthreads) {
    const int m = n / R;
    const int work = m * batch;
    for (int w = tid; w < work; w += nthreads) {
        const int b = w % batch, j = w / batch;
        const int q = j % s, base = j - q;
        cplx a[R];
#pragma unroll
        for (int t = 0; t < R; t++) a[t] = x[(j + t * m) * batch + b];
        if (R == 2) bfly2<DIR>(a);
        else if (R == 3) bfly3<DIR>(a);
        else if (R == 4) bfly4<DIR>(a);
        else bfly5<DIR>(a);
        const int o = (base * R + q) * batch + b;
        y[o] = a[0];
#pragma unroll
        for (int u = 1; u < R; u++) {
            cplx wv = tw_dir<DIR>(__ldg(tw + base * u));
            y[o + s * u * batch] = cmul(a[u], wv);
        }
    }
}

// generic prime radix (7 .. 61): O(R^2) butterfly with roots taken from the same table
template <int DIR>
__device__ __noinline__ void stockham_stage_generic(const cplx* __restrict__ x, cplx* __restrict__ y, int n,
                                                    int s, int R, int batch, const cplx* __restrict__ tw,
                                                    int tid, int nthreads) {
    const int m = n / R;
    const int work = m * batch;
    for (int w = tid; w < work; w += nthreads) {
        const int b = w % batch, j = w / batch;
        const int q = j % s, base = j - q;
        cplx a[FFT_MAX_RADIX];
        for (int t = 0; t < R; t++) a[t] = x[(j + t * m) * batch + b];
        const int o = (base * R + q) * batch + b;
        for (int u = 0; u < R; u++) {
            cplx acc = a[0];
            for (int t = 1; t < R; t++) {
                cplx wr = tw_dir<DIR>(__ldg(tw + ((t * u) % R) * m));
                acc = cadd(acc, cmul(a[t], wr));
            }
            if (u) acc = cmul(acc, tw_dir<DIR>(__ldg(tw + base * u)));
            y[o + s * u * batch] = acc;
        }
    }
}

// Full transform of `batch` interleaved sequences in shared memory; result pointer returned
// (x or y).  Every thread of the CTA must call it; ends with a barrier.
template <int DIR>
__device__ __forceinline__ cplx* stockham(cplx* x, cplx* y, const FftPlan& plan, int batch,
                                          const cplx* __restrict__ tw, int tid, int nthreads) {
    int s = 1;
    const int n = plan.n;
    for (int st = 0; st < plan.nstage; st++) {
        const int r = plan.radix[st];
        switch (r) {
            case 2: stockham_stage<2, DIR>(x, y, n, s, batch, tw, tid, nthreads); break;
            case 3: stockham_stage<3, DIR>(x, y, n, s, batch, tw, tid, nthreads); break;
            case 4: stockham_stage<4, DIR>(x, y, n, s, batch, tw, tid, nthreads); break;
            case 5: stockham_stage<5, DIR>(x, y, n, s, batch, tw, tid, nthreads); break;
            default: stockham_stage_generic<DIR>(x, y, n, s, r, batch, tw, tid, nthreads); break;
        }
        __syncthreads();
        cplx* t = x; x = y; y = t;
        s *= r;
    }
    return x;
}

// Column pass.  load(idx) returns element idx = j1*n2 + c of the sequence; store(k1, c, val)
// receives the transformed, twiddled value.  TW_BEFORE: multiply by w_N^(DIR*c*k1) before the
// transform (inverse path), otherwise after it (forward path).
template <int DIR, bool TW_BEFORE, class Load, class Store>
__device__ __forceinline__ void fft_column_pass(const FftSplit& sp, const cplx* __restrict__ tw1,
                                                const cplx* __restrict__ twN, cplx* smem, Load load,
                                                Store store, int tid, int nthreads) {
    const int n1 = sp.n1, n2 = sp.n2;
    for (int c0 = 0; c0 < n2; c0 += FFT_COL_TILE) {
        const int nc = imin(FFT_COL_TILE, n2 - c0);
        cplx* x = smem;
        cplx* y = smem + (size_t)n1 * nc;
        for (int e = tid; e < n1 * nc; e += nthreads) {
            const int j1 = e / nc, c = e % nc;
            cplx val = load(j1 * n2 + c0 + c);
            if (TW_BEFORE) {
                long long idx = ((long long)j1 * (c0 + c)) % sp.N;
                val = cmul(val, tw_dir<DIR>(__ldg(twN + idx)));
            }
            x[e] = val;
        }
        __syncthreads();
        cplx* r = stockham<DIR>(x, y, sp.p1, nc, tw1, tid, nthreads);
        for (int e = tid; e < n1 * nc; e += nthreads) {
            const int k1 = e / nc, c = e % nc;
            cplx val = r[e];
            if (!TW_BEFORE) {
                long long idx = ((long long)k1 * (c0 + c)) % sp.N;
                val = cmul(val, tw_dir<DIR>(__ldg(twN + idx)));
            }
            store(k1, c0 + c, val);
        }
        __syncthreads();
    }
}

// Row pass over rows [0, n1).  load(row, j2) / store(row, k2, val).  `mid(buf, rows, tid, nthreads)`
// is called with the transformed rows still in shared memory (element (k2, b) at buf[k2*rows + b])
// and returns the buffer to store from -- identity for a plain transform, or
// |X|^2 followed by the inverse transform for the autocorrelation.
template <int DIR, class Load, class Store, class Mid>
__device__ __forceinline__ void fft_row_pass(const FftSplit& sp, const cplx* __restrict__ tw2, cplx* smem,
                                             Load load, Store store, Mid mid, int tid, int nthreads) {
    const int n1 = sp.n1, n2 = sp.n2;
    const int rows_per_tile = n2 >= 1024 ? 1 : imin(imax(1, 1024 / n2), n1);
    for (int r0 = 0; r0 < n1; r0 += rows_per_tile) {
        const int nr = imin(rows_per_tile, n1 - r0);
        cplx* x = smem;
        cplx* y = smem + (size_t)n2 * nr;
        for (int e = tid; e < n2 * nr; e += nthreads) {
            const int b = e / n2, j2 = e % n2;          // coalesced along the row
            x[j2 * nr + b] = load(r0 + b, j2);
        }
        __syncthreads();
        cplx* r = stockham<DIR>(x, y, sp.p2, nr, tw2, tid, nthreads);
        cplx* other = (r == x) ? y : x;
        r = mid(r, other, nr, tid, nthreads);
        for (int e = tid; e < n2 * nr; e += nthreads) {
            const int b = e / n2, k2 = e % n2;
            store(r0 + b, k2, r[k2 * nr + b]);
        }
        __syncthreads();
    }
}

struct MidIdentity {
    __device__ __forceinline__ cplx* operator()(cplx* r, cplx*, int, int, int) const { return r; }
};

}  // namespace dp
