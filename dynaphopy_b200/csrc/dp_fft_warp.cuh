// Warp-private FFT engine (v2): every sub-transform of the four-step factorisation N = n1*n2
// (n1, n2 <= 512) is carried out by ONE warp in its own shared-memory slab, synchronised with
// __syncwarp only, so the warps of a CTA stream through the rows / column tiles of an item
// independently and hide each other's latency -- no block-wide barrier inside a pass.
//   * Stockham autosort stages, radix 8/4/2/5/3/... butterflies computed in registers by the
//     compile-time codelets (PencilDft), all forward; inverse transforms conjugate on load/store.
//   * stage twiddles w_n^k come from shared-memory tables; the inter-step twiddle w_N^(k1*c) is
//     the product of a coarse and a fine table entry (N <= 512*512), also in shared memory.
// The block-level engine of dp_fft.cuh remains the fallback for lengths this one cannot split.
#pragma once

#include "dp_fft.cuh"
#include "dp_pencil.cuh"

namespace dp {

constexpr int WF_MAX = 512;        // longest sub-transform a warp handles
constexpr int WF_TILE = 256;       // complex elements per warp slab (two slabs per warp: ping-pong)
constexpr int WF_FINE = 512;       // fine table length of the two-level inter-step twiddle

struct WSplit {
    int N, n1, n2;                 // n1 == 1: single pass of rows
    int c1, r2;                    // columns per column tile, rows per row tile
    FftPlan p1, p2;
};

inline bool make_warp_plan(int n, FftPlan* p) {
    p->n = n;
    p->nstage = 0;
    int rem = n;
    const int pref[] = {8, 4, 2, 5, 3, 7, 11, 13, 17, 19, 23, 29, 31};
    for (int f : pref) {
        while (rem % f == 0) {
            if (p->nstage >= FFT_MAX_STAGES) return false;
            p->radix[p->nstage++] = (unsigned char)f;
            rem /= f;
        }
    }
    return rem == 1;
}

inline bool make_wsplit(int N, WSplit* s) {
    s->N = N;
    int best = 0;
    if (N <= WF_TILE) {
        best = 1;
    } else {
        double target = sqrt((double)N);
        for (int n1 = 2; n1 <= WF_MAX; n1++) {
            if (N % n1) continue;
            int n2 = N / n1;
            if (n2 > WF_MAX) continue;
            FftPlan a, b;
            if (!make_warp_plan(n1, &a) || !make_warp_plan(n2, &b)) continue;
            if (!best || fabs(n1 - target) < fabs(best - target) - 1e-9 ||
                (fabs(fabs(n1 - target) - fabs(best - target)) < 1e-9 && n1 < best))
                best = n1;
        }
    }
    if (!best) return false;
    s->n1 = best;
    s->n2 = N / best;
    if (!make_warp_plan(s->n1, &s->p1) || !make_warp_plan(s->n2, &s->p2)) return false;
    s->c1 = imax(1, imin(8, WF_TILE / s->n1));
    s->r2 = imax(1, imin(8, WF_TILE / s->n2));
    if (s->n1 * s->c1 > 2 * WF_TILE || s->n2 * s->r2 > 2 * WF_TILE) return false;   // must fit the slabs
    return s->n1 <= 2 * WF_TILE && s->n2 <= 2 * WF_TILE;
}

// shared-memory layout of one CTA (complex elements)
struct WSmem {
    cplx* tw1;       // [WF_MAX]   w_n1^k
    cplx* tw2;       // [WF_MAX]   w_n2^k
    cplx* fine;      // [WF_FINE]  w_N^k,            k < 512
    cplx* coarse;    // [WF_MAX]   w_N^(512 k),      k < N/512
    cplx* slabs;     // nwarps * 2 * slab
    int slab;        // elements per slab
};
constexpr int WF_TABLE_ELEMS = 3 * WF_MAX + WF_FINE;

__device__ __forceinline__ WSmem wsmem_carve(cplx* base, int slab) {
    WSmem m;
    m.tw1 = base;
    m.tw2 = base + WF_MAX;
    m.fine = base + 2 * WF_MAX;
    m.coarse = base + 2 * WF_MAX + WF_FINE;
    m.slabs = base + WF_TABLE_ELEMS;
    m.slab = slab;
    return m;
}

// all threads of the CTA: stage the tables of this split from the global tables
// (g1[k] = w_n1^k, g2[k] = w_n2^k, gN[k] = w_N^k); starts and ends with __syncthreads
__device__ __forceinline__ void wtables_load(const WSplit& sp, const WSmem& m, const cplx* __restrict__ g1,
                                             const cplx* __restrict__ g2, const cplx* __restrict__ gN, int tid,
                                             int nthreads) {
    __syncthreads();
    for (int k = tid; k < sp.n1; k += nthreads) m.tw1[k] = g1[k];
    for (int k = tid; k < sp.n2; k += nthreads) m.tw2[k] = g2[k];
    for (int k = tid; k < WF_FINE && k < sp.N; k += nthreads) m.fine[k] = gN[k];
    const int nco = (sp.N + WF_FINE - 1) / WF_FINE;
    for (int k = tid; k < nco; k += nthreads) m.coarse[k] = gN[(size_t)k * WF_FINE];
    __syncthreads();
}

__device__ __forceinline__ cplx wtwiddle(const WSmem& m, int idx) {     // w_N^idx, idx < N
    const cplx f = m.fine[idx & (WF_FINE - 1)];
    const int hi = idx >> 9;
    return hi ? cmul(m.coarse[hi], f) : f;
}

// one forward Stockham stage by one warp; x -> y, `batch` interleaved sequences
template <int R>
__device__ __forceinline__ void wstage(const cplx* __restrict__ x, cplx* __restrict__ y, int n, int s, int batch,
                                       const cplx* __restrict__ tw, int lane) {
    const int m = n / R;
    const int work = m * batch;
    for (int w = lane; w < work; w += 32) {
        const int b = w % batch, j = w / batch;
        const int q = j % s, base = j - q;
        cplx a[R], o[R];
#pragma unroll
        for (int t = 0; t < R; t++) a[t] = x[(j + t * m) * batch + b];
        PencilDft<R>::run(a, o);
        const int dst = (base * R + q) * batch + b;
        y[dst] = o[0];
        if (base == 0) {
#pragma unroll
            for (int u = 1; u < R; u++) y[dst + s * u * batch] = o[u];
        } else {
#pragma unroll
            for (int u = 1; u < R; u++) y[dst + s * u * batch] = cmul(o[u], tw[base * u]);
        }
    }
}

#define DP_WSTAGE_CASE(r) case r: wstage<r>(x, y, n, s, batch, tw, lane); break;
// forward FFT of `batch` interleaved sequences of length plan.n in the warp's slabs x/y;
// returns the slab holding the result.  Warp-synchronous: call with all 32 lanes.
__device__ __forceinline__ cplx* warp_fft(cplx* x, cplx* y, const FftPlan& plan, int batch,
                                          const cplx* __restrict__ tw, int lane) {
    int s = 1;
    const int n = plan.n;
    for (int st = 0; st < plan.nstage; st++) {
        const int r = plan.radix[st];
        switch (r) {
            DP_WSTAGE_CASE(2) DP_WSTAGE_CASE(3) DP_WSTAGE_CASE(4) DP_WSTAGE_CASE(5) DP_WSTAGE_CASE(7)
            DP_WSTAGE_CASE(8) DP_WSTAGE_CASE(11) DP_WSTAGE_CASE(13) DP_WSTAGE_CASE(17) DP_WSTAGE_CASE(19)
            DP_WSTAGE_CASE(23) DP_WSTAGE_CASE(29) DP_WSTAGE_CASE(31)
            default: break;
        }
        __syncwarp();
        cplx* t = x; x = y; y = t;
        s *= r;
    }
    return x;
}
#undef DP_WSTAGE_CASE

// Column pass: tiles of c1 columns distributed over the warps.  load(idx) -> element j1*n2 + c,
// store(k1, c, val).  CONJ: inverse transform (conjugate on load and on store).
// TW_BEFORE: inter-step twiddle applied to the input (inverse order) instead of the output.
template <bool CONJ, bool TW_BEFORE, class Load, class Store>
__device__ __forceinline__ void wpass_columns(const WSplit& sp, const WSmem& m, Load load, Store store, int warp,
                                              int nwarps, int lane) {
    const int n1 = sp.n1, n2 = sp.n2, c1 = sp.c1;
    cplx* x = m.slabs + (size_t)warp * 2 * m.slab;
    cplx* y = x + m.slab;
    const int ntiles = (n2 + c1 - 1) / c1;
    for (int tile = warp; tile < ntiles; tile += nwarps) {
        const int c0 = tile * c1;
        const int nc = imin(c1, n2 - c0);
        for (int e = lane; e < n1 * nc; e += 32) {
            const int j1 = e / nc, c = e % nc;
            cplx v = load(j1 * n2 + c0 + c);
            if (CONJ) v.y = -v.y;
            if (TW_BEFORE) v = cmul(v, wtwiddle(m, j1 * (c0 + c)));
            x[e] = v;
        }
        __syncwarp();
        cplx* r = warp_fft(x, y, sp.p1, nc, m.tw1, lane);
        for (int e = lane; e < n1 * nc; e += 32) {
            const int k1 = e / nc, c = e % nc;
            cplx v = r[e];
            if (!TW_BEFORE) v = cmul(v, wtwiddle(m, k1 * (c0 + c)));
            if (CONJ) v.y = -v.y;
            store(k1, c0 + c, v);
        }
        __syncwarp();
    }
}

// Row pass: tiles of r2 rows per warp.  load(row, j2), store(row, k2, val).
// POWER_INVERSE: after the forward transform replace X by |X|^2 and transform back
// (IFFT of a real sequence = conj(FFT)), i.e. the circular autocorrelation of the row.
// CONJ: inverse transform (conjugate on load and on store).
template <bool POWER_INVERSE, bool CONJ, class Load, class Store>
__device__ __forceinline__ void wpass_rows(const WSplit& sp, const WSmem& m, Load load, Store store, int warp,
                                           int nwarps, int lane) {
    const int n1 = sp.n1, n2 = sp.n2, r2 = sp.r2;
    cplx* x = m.slabs + (size_t)warp * 2 * m.slab;
    cplx* y = x + m.slab;
    const int ntiles = (n1 + r2 - 1) / r2;
    for (int tile = warp; tile < ntiles; tile += nwarps) {
        const int r0 = tile * r2;
        const int nr = imin(r2, n1 - r0);
        for (int e = lane; e < n2 * nr; e += 32) {
            const int b = e / n2, j2 = e % n2;
            cplx v = load(r0 + b, j2);
            if (CONJ) v.y = -v.y;
            x[j2 * nr + b] = v;
        }
        __syncwarp();
        cplx* r = warp_fft(x, y, sp.p2, nr, m.tw2, lane);
        if (POWER_INVERSE) {
            cplx* other = (r == x) ? y : x;
            for (int e = lane; e < n2 * nr; e += 32) {
                const cplx v = r[e];
                r[e] = make_double2(v.x * v.x + v.y * v.y, 0.0);
            }
            __syncwarp();
            r = warp_fft(r, other, sp.p2, nr, m.tw2, lane);
        }
        for (int e = lane; e < n2 * nr; e += 32) {
            const int b = e / n2, k2 = e % n2;
            cplx v = r[k2 * nr + b];
            if (POWER_INVERSE || CONJ) v.y = -v.y;
            store(r0 + b, k2, v);
        }
        __syncwarp();
    }
}

}  // namespace dp
