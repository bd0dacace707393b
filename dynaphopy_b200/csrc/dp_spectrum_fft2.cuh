// A7, fast path ("v2") of the FFT power spectrum: one persistent CTA per (series, window) item,
// four passes over an L2-resident scratch slab, every sub-transform done in registers by the
// warp-private engine of dp_fft_reg.cuh.  Included by dp_spectrum_fft.cu only.
//
//   P = n1*n2 = power of two >= L + L/2 (linear autocorrelation of an L-sample window), n2 = 256 (64)
//   pass 1  x (HBM, read once)  -> column FFTs (length n1, stride n2) * w_P^(col k1)      -> A[k1][col]
//   pass 2  rows of A: FFT_n2 -> |X|^2 -> FFT_n2 (DIF, so no reordering) * w_P^(row k)    -> A in place
//   pass 3  column FFTs of A -> conj -> r[kappa], only the lags 0..L/2 (r is Hermitian)   -> A[kappa]
//   pass 4  spectrum bins k of FFT_L(a), a[m] = r[m - L/2] / L, only the bins the frequency grid
//           needs:  FFT_L(a)[k] = sum_j2 w_L^(j2 k) C_j2[k mod n1'],  C_j2 = DFT_n1'(column j2 of a)
//           (L = n1' * n2', n1' from a menu of register-codelet sizes, n2' arbitrary), evaluated by
//           Horner's rule in z = w_L^k over the columns                                    -> spec
// The slab (P complex per CTA, <= 1 MB; pass 3 writes r over the rows it has consumed) is written and
// re-read by the same SM within microseconds and stays in the 126 MB L2; HBM sees the window once
// (16 B per sample; the next item's window is prefetched into L2 while passes 3-4 run) and the kept
// bins (8 B each).
#pragma once

#include "dp_fft_reg.cuh"

namespace dp {

#ifndef DP_S2_THREADS
#define DP_S2_THREADS 256
#endif
constexpr int S2_THREADS = DP_S2_THREADS;          // 8 warps x <= 255 registers
constexpr int S2_WARPS = S2_THREADS / 32;
constexpr int S2_BINSLOTS = 8;                    // bins per residue thread in pass 4
constexpr int S2_TILE_COLS = S2_THREADS / 16;     // pass-4 column tile: 16 threads per column
constexpr int S2_NP_MAX = 192;                    // largest n1' of the menu
constexpr int S2_CPITCH = S2_TILE_COLS + 1;       // odd pitch (in 16-byte elements)

typedef SubCfg<16, 1, 1> C16;
typedef SubCfg<8, 4, 4> C32;
typedef SubCfg<8, 8, 4> C64;
typedef SubCfg<16, 8, 8> C128;
typedef SubCfg<16, 16, 16> C256;

// mixed-radix menu for n1' (pass 4): all with 16 lanes per column
#define DP_S2_MIXED_MENU(X) \
    X(16, 12) X(16, 10) X(15, 10) X(16, 9) X(16, 8) X(12, 10) X(10, 10) X(16, 6) X(16, 5) \
    X(8, 8) X(10, 5) X(8, 5) X(8, 4) X(5, 5) X(4, 4)

struct S2Plan {
    int L, half, P, n1, n2;        // n1, n2: powers of two
    int m1, m2;                    // n1' = m1 * m2 (menu entry), n2L = L / n1'
    int n2L;
};

struct S2Params {
    const cplx* series;
    int64_t stride_s, stride_t;    // element strides of the series index and of time
    int tb_len;                    // 0: plain time axis, else time is blocked ...
    int tb_shift;                  // log2(tb_len) when it is a power of two (no integer division), else -1
    int64_t tb_stride;             // ... block b = t / tb_len starts at b * tb_stride
    int S, npieces;
    int w_first, w_count;          // windows [w_first, w_first + w_count) are transformed by this launch
    const int* starts;
    S2Plan pl;
    const cplx* g_tables;          // 6 x 256: T1, T2, T2f, Ts, fine, coarse (s2_tables_kernel)
    const cplx* g_twL;             // w_L^k, k < L
    cplx* scratch;                 // gridDim.x slabs of slab_elems
    size_t slab_elems;
    double* spec;                  // [S][npieces][nkeep]
    int lo, nkeep;
    double dt, acf_scale;
    int skip_mask;                 // development aid (DP_S2_SKIP): bit i set = skip pass i+1
    int pair;                      // 1: process the windows of a series two at a time (see spectrum2_kernel)
};

struct S2Smem {
    cplx* T1;       // [k1*R2 + j2] = w_n1^(j2 k1)      stage twiddles of the n1 sub-transform (DIT)
    cplx* T2;       // same for n2
    cplx* T2f;      // [c*R1 + k1]  = w_n2^(k1 c)       DIF order
    cplx* Ts;       // [k1*R2 + j2] = w_n1'^(j2 k1)     pass 4
    cplx* fine;     // [256] w_P^j
    cplx* coarse;   // [256] w_P^(256 j)
    cplx* E;        // S2_WARPS * SUBFFT_EWARP_MAX      exchange slabs (pass 4: the column tile)
    cplx* PF;       // S2_WARPS * 512                   per-warp landing zone of the NEXT tile (cp.async, passes 1-3)
};
// Measured (B200, cfg 5-A): 138.4 ms with the cp.async pipeline vs 129.7 ms without -- the passes are limited by
// LSU / L2-port throughput (LSU wavefronts 50 %), not by exposed load latency, and the extra shared-memory round trip
// costs more than the overlap buys.  Kept as a build option (-DDP_S2_PREFETCH=1) for other window lengths.
#ifndef DP_S2_PREFETCH
#define DP_S2_PREFETCH 0
#endif
// Register software pipeline of passes 1-3 (see s2_pass1): next tile's loads issued between the two butterfly stages.
#ifndef DP_S2_SWP
#define DP_S2_SWP 0
#endif
#ifdef DP_EMUL
#define DP_LAMBDA_INLINE
#else
#define DP_LAMBDA_INLINE __attribute__((always_inline))
#endif
constexpr size_t S2_PF_ELEMS = DP_S2_PREFETCH ? (size_t)S2_WARPS * 512 : 0;
constexpr size_t S2_E_ELEMS = S2_WARPS * SUBFFT_EWARP_MAX > 2 * S2_NP_MAX * S2_CPITCH ? S2_WARPS * SUBFFT_EWARP_MAX
                                                                                  : 2 * S2_NP_MAX * S2_CPITCH;
constexpr size_t S2_SMEM_BYTES = (size_t)(6 * 256 + S2_E_ELEMS + S2_PF_ELEMS) * sizeof(cplx);

// carve the CTA's dynamic shared memory; done inside every pass (from the __shared__ symbol itself) so
// that the compiler sees shared-space pointers and emits LDS/STS instead of generic loads
__device__ __forceinline__ S2Smem s2_carve(cplx* smem) {
    S2Smem sm;
    sm.T1 = smem; sm.T2 = smem + 256; sm.T2f = smem + 512; sm.Ts = smem + 768;
    sm.fine = smem + 1024; sm.coarse = smem + 1280; sm.E = smem + 1536;
    sm.PF = smem + 1536 + S2_E_ELEMS;
    return sm;
}
#define S2_SMEM_HERE() DP_DYNSMEM(cplx, s2_smem_base_); const S2Smem sm = s2_carve(s2_smem_base_)

__device__ __forceinline__ int64_t s2_time_offset(const S2Params& p, int t) {
    if (p.tb_len == 0) return (int64_t)t * p.stride_t;
    const int b = p.tb_shift >= 0 ? (t >> p.tb_shift) : t / p.tb_len;
    return (int64_t)b * p.tb_stride + (int64_t)(t - b * p.tb_len) * p.stride_t;
}

__device__ __forceinline__ void s2_prefetch_l2(const void* ptr) {
#ifndef DP_EMUL
    asm volatile("prefetch.global.L2 [%0];\n" ::"l"(ptr));
#endif
}

// Slab accesses carry an L2 evict_last policy (the slab must survive in L2 between the passes), the window
// samples are read with the streaming (evict_first) hint: they are used once.
__device__ __forceinline__ unsigned long long s2_slab_policy() {
#ifdef DP_EMUL
    return 0;
#else
    unsigned long long pol;
#ifndef DP_S2_EVICT_FRAC
#define DP_S2_EVICT_FRAC 1.0
#endif
#define DP_S2_STR2(x) #x
#define DP_S2_STR(x) DP_S2_STR2(x)
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, " DP_S2_STR(DP_S2_EVICT_FRAC) ";\n" : "=l"(pol));
    return pol;
#endif
}
#ifndef DP_S2_LDMODE
#define DP_S2_LDMODE 2
#endif
__device__ __forceinline__ cplx s2_ld_slab(const cplx* ptr, unsigned long long pol) {
#if defined(DP_EMUL) || DP_S2_LDMODE == 0
    return *ptr;
#elif DP_S2_LDMODE == 1
    return __ldcg(ptr);
#else
    cplx v;
    asm volatile("ld.global.cg.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;\n" : "=d"(v.x), "=d"(v.y) : "l"(ptr), "l"(pol) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void s2_st_slab(cplx* ptr, cplx v, unsigned long long pol) {
#if defined(DP_EMUL) || DP_S2_LDMODE == 0
    *ptr = v;
#elif DP_S2_LDMODE == 1
    __stcg(ptr, v);
#else
    asm volatile("st.global.cg.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;\n" ::"l"(ptr), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
#endif
}


// ---- software pipeline of passes 1-3: the NEXT tile of a warp travels global -> shared memory with cp.async while
// the current one is transformed; a lane copies exactly the 16 values it will load into registers (slot r*32 + lane:
// conflict-free), so no cross-lane synchronisation is involved.  Zero fill for samples beyond the window.
__device__ __forceinline__ unsigned long long s2_stream_policy() {
#ifdef DP_EMUL
    return 0;
#else
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
#endif
}
__device__ __forceinline__ void s2_cp16(cplx* dst_smem, const cplx* src, bool valid, unsigned long long pol) {
#ifdef DP_EMUL
    *dst_smem = valid ? *src : make_double2(0.0, 0.0);
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;\n" ::"r"(d), "l"(src), "r"(sz), "l"(pol) : "memory");
#endif
}
__device__ __forceinline__ void s2_cp_commit() {
#ifndef DP_EMUL
    asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
}
__device__ __forceinline__ void s2_cp_wait() {
#ifndef DP_EMUL
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#endif
}

// ---- pass 1: window -> column FFTs -> A -----------------------------------------------------------
// LAYOUT 0: time axis contiguous (stride_t == 1, no blocks) -- plain pointer arithmetic in the unrolled loads;
// 1: contiguous power-of-two time blocks (the all-to-all's receive layout); 2: anything else.
template <class C, int LAYOUT>
__device__ __noinline__ void s2_pass1(const S2Params& p, const cplx* __restrict__ x, int t0, cplx* __restrict__ A,
                                      int warp, int lane) {
    S2_SMEM_HERE();
    const unsigned long long pol = s2_slab_policy();
    const int n2 = p.pl.n2, L = p.pl.L;
    const int tb_shift = p.tb_shift;
    const int64_t tb_stride = p.tb_stride;
    const int seq = lane % C::NS, g = lane / C::NS;
    cplx* E = sm.E + (size_t)warp * SUBFFT_EWARP_MAX + seq * C::ESEQ;
    const int ntiles = n2 / C::NS;
    auto load_tile = [&](int tile, cplx (&w)[16]) DP_LAMBDA_INLINE {
        const int col = tile * C::NS + seq;
#pragma unroll
        for (int i = 0; i < C::NB1; i++)
#pragma unroll
            for (int a = 0; a < C::R1; a++) {
                const int idx = (a * C::R2 + g + C::TG * i) * n2 + col;
                cplx val = make_double2(0.0, 0.0);
                if (idx < L) {
                    if (LAYOUT == 0) val = __ldcs(x + t0 + idx);
                    else if (LAYOUT == 1) {
                        const int t = t0 + idx, b = t >> tb_shift;
                        val = __ldcs(x + (int64_t)b * tb_stride + (t - (b << tb_shift)));
                    } else val = __ldcs(x + s2_time_offset(p, t0 + idx));
                }
                w[i * C::R1 + a] = val;
            }
    };
#if DP_S2_PREFETCH
    cplx* PB = sm.PF + (size_t)warp * 512 + lane;
    const unsigned long long spol = s2_stream_policy();
    auto prefetch_tile = [&](int tile) {
        const int col = tile * C::NS + seq;
#pragma unroll
        for (int i = 0; i < C::NB1; i++)
#pragma unroll
            for (int a = 0; a < C::R1; a++) {
                const int idx = (a * C::R2 + g + C::TG * i) * n2 + col;
                const bool ok = idx < L;
                const int t = t0 + (ok ? idx : 0);
                const cplx* src;
                if (LAYOUT == 0) src = x + t;
                else if (LAYOUT == 1) { const int b = t >> tb_shift; src = x + (int64_t)b * tb_stride + (t - (b << tb_shift)); }
                else src = x + s2_time_offset(p, t);
                s2_cp16(PB + (i * C::R1 + a) * 32, src, ok, spol);
            }
        s2_cp_commit();
    };
    if (warp < ntiles) prefetch_tile(warp);
#endif
#if DP_S2_SWP
    cplx v[16], nx[16];
    if (warp < ntiles) load_tile(warp, v);
#endif
    for (int tile = warp; tile < ntiles; tile += S2_WARPS) {
        const int col = tile * C::NS + seq;
#if DP_S2_SWP
        // software pipeline: the next tile's samples are requested as soon as this tile's first butterflies are
        // parked in the exchange slab, and travel while the second butterflies, twiddles and stores run
        sub_dit_a<C>(v, E, sm.T1, g);
        const bool more = tile + S2_WARPS < ntiles;
        if (more) load_tile(tile + S2_WARPS, nx);
        sub_dit_b<C>(v, E, g);
#else
        cplx v[16];
#if DP_S2_PREFETCH
        s2_cp_wait();
#pragma unroll
        for (int r = 0; r < 16; r++) v[r] = PB[r * 32];
        if (tile + S2_WARPS < ntiles) prefetch_tile(tile + S2_WARPS);
#else
        load_tile(tile, v);
#endif
        sub_dit<C>(v, E, sm.T1, g);
#endif
        // four-step twiddle w_P^(col*kk), kk = k1 + R1*k2: geometric in k2
        const cplx ratio = tw2level(sm.fine, sm.coarse, col * C::R1);
#pragma unroll
        for (int i = 0; i < C::NB2; i++) {
            const int k1 = g + C::TG * i;
            cplx tw = tw2level(sm.fine, sm.coarse, col * k1);
#pragma unroll
            for (int k2 = 0; k2 < C::R2; k2++) {
                s2_st_slab(A + (size_t)(k1 + C::R1 * k2) * n2 + col, cmul(v[i * C::R2 + k2], tw), pol);
                if (k2 + 1 < C::R2) tw = cmul(tw, ratio);
            }
        }
        __syncwarp();
#if DP_S2_SWP
        if (more) {
#pragma unroll
            for (int r = 0; r < 16; r++) v[r] = nx[r];
        }
#endif
    }
}

// ---- pass 2: rows: FFT -> |.|^2 -> FFT (DIF) -> twiddle, in place ------------------------------------
// MODE 0: single window.  Paired windows (two windows of one series share the second half of the pipeline,
// see spectrum2_kernel): MODE 1 = first window, rows FFT -> |X|^2 parked as REAL values in Sr (digit order);
// MODE 2 = second window, rows FFT -> |X|^2, z = Sr + i |X|^2, FFT (DIF) -> twiddle -> A.
template <class C, int MODE>
__device__ __noinline__ void s2_pass2(const S2Params& p, cplx* __restrict__ A, double* __restrict__ Sr, int warp,
                                      int lane) {
    S2_SMEM_HERE();
    const unsigned long long pol = s2_slab_policy();
    const int n1 = p.pl.n1, n2 = p.pl.n2;
    const int seq = lane % C::NS, g = lane / C::NS;
    cplx* E = sm.E + (size_t)warp * SUBFFT_EWARP_MAX + seq * C::ESEQ;
    const int ntiles = n1 / C::NS;
    auto load_tile = [&](int tile, cplx (&w)[16]) DP_LAMBDA_INLINE {
        const cplx* Ar = A + (size_t)(tile * C::NS + seq) * n2;
#pragma unroll
        for (int i = 0; i < C::NB1; i++)
#pragma unroll
            for (int a = 0; a < C::R1; a++) w[i * C::R1 + a] = s2_ld_slab(Ar + a * C::R2 + g + C::TG * i, pol);
    };
#if DP_S2_PREFETCH
    cplx* PB = sm.PF + (size_t)warp * 512 + lane;
    auto prefetch_tile = [&](int tile) {
        const cplx* Ar = A + (size_t)(tile * C::NS + seq) * n2;
#pragma unroll
        for (int i = 0; i < C::NB1; i++)
#pragma unroll
            for (int a = 0; a < C::R1; a++) s2_cp16(PB + (i * C::R1 + a) * 32, Ar + a * C::R2 + g + C::TG * i, true, pol);
        s2_cp_commit();
    };
    if (warp < ntiles) prefetch_tile(warp);
#endif
#if DP_S2_SWP
    cplx v[16], nx[16];
    if (warp < ntiles) load_tile(warp, v);
#endif
    for (int tile = warp; tile < ntiles; tile += S2_WARPS) {
        const int row = tile * C::NS + seq;
        cplx* __restrict__ Ar = A + (size_t)row * n2;
#if !DP_S2_SWP
        cplx v[16];
#if DP_S2_PREFETCH
        s2_cp_wait();
#pragma unroll
        for (int r = 0; r < 16; r++) v[r] = PB[r * 32];
        if (tile + S2_WARPS < ntiles) prefetch_tile(tile + S2_WARPS);
#else
        load_tile(tile, v);
#endif
#endif
        double sa[16];
        double* __restrict__ Srow = Sr + (size_t)row * n2;
        if (MODE == 2) {                       // |X_a|^2 of the partner window, requested before the butterflies
#pragma unroll
            for (int i = 0; i < C::NB2; i++)
#pragma unroll
                for (int k2 = 0; k2 < C::R2; k2++) sa[i * C::R2 + k2] = __ldcg(Srow + g + C::TG * i + C::R1 * k2);
        }
#if DP_S2_SWP
        sub_dit_a<C>(v, E, sm.T2, g);
        const bool more = tile + S2_WARPS < ntiles;
        if (more) load_tile(tile + S2_WARPS, nx);      // next rows travel underneath the rest of this tile
        sub_dit_b<C>(v, E, g);
#else
        sub_dit<C>(v, E, sm.T2, g);
#endif
        if constexpr (MODE == 1) {
#pragma unroll
            for (int i = 0; i < C::NB2; i++)
#pragma unroll
                for (int k2 = 0; k2 < C::R2; k2++) {
                    const cplx x = v[i * C::R2 + k2];
                    __stcg(Srow + g + C::TG * i + C::R1 * k2, x.x * x.x + x.y * x.y);
                }
        } else {
#pragma unroll
            for (int r = 0; r < C::NB2 * C::R2; r++) {
                const double sb = v[r].x * v[r].x + v[r].y * v[r].y;
                v[r] = MODE == 2 ? make_double2(sa[r], sb) : make_double2(sb, 0.0);
            }
            __syncwarp();
            sub_dif<C>(v, E, sm.T2f, g);
            // four-step twiddle w_P^(row*k), k = d*R2 + c: geometric in d
            const cplx ratio = tw2level(sm.fine, sm.coarse, row * C::R2);
#pragma unroll
            for (int i = 0; i < C::NB1; i++) {
                const int c = g + C::TG * i;
                cplx tw = tw2level(sm.fine, sm.coarse, row * c);
#pragma unroll
                for (int d = 0; d < C::R1; d++) {
                    s2_st_slab(Ar + d * C::R2 + c, cmul(v[i * C::R1 + d], tw), pol);
                    if (d + 1 < C::R1) tw = cmul(tw, ratio);
                }
            }
        }
        __syncwarp();
#if DP_S2_SWP
        if (more) {
#pragma unroll
            for (int r = 0; r < 16; r++) v[r] = nx[r];
        }
#endif
    }
}

// ---- pass 3: column FFTs -> conj -> lags 0..half of the autocorrelation, in place ------------------------
// r[kappa] lands at A[kappa]: a warp overwrites only rows of the columns it has just read.
// PAIRED: A holds Y = FFT_P(|X_a|^2 + i |X_b|^2); rho[kappa] = r_a[kappa] + i r_b[kappa] = Y[-kappa] / P, so the
// lags -half..half are Y at indices 0..half and P-half..P-1 (no conjugation, no Hermitian shortcut).
template <class C, bool PAIRED>
__device__ __noinline__ void s2_pass3(const S2Params& p, cplx* __restrict__ A, int warp, int lane) {
    S2_SMEM_HERE();
    const unsigned long long pol = s2_slab_policy();
    const int n2 = p.pl.n2, half = p.pl.half;
    const double sc = p.acf_scale;
    const int seq = lane % C::NS, g = lane / C::NS;
    cplx* E = sm.E + (size_t)warp * SUBFFT_EWARP_MAX + seq * C::ESEQ;
    const int ntiles = n2 / C::NS;
    auto load_tile = [&](int tile, cplx (&w)[16]) DP_LAMBDA_INLINE {
        const int col = tile * C::NS + seq;
#pragma unroll
        for (int i = 0; i < C::NB1; i++)
#pragma unroll
            for (int a = 0; a < C::R1; a++) w[i * C::R1 + a] = s2_ld_slab(A + (size_t)(a * C::R2 + g + C::TG * i) * n2 + col, pol);
    };
#if DP_S2_PREFETCH
    cplx* PB = sm.PF + (size_t)warp * 512 + lane;
    auto prefetch_tile = [&](int tile) {
        const int col = tile * C::NS + seq;
#pragma unroll
        for (int i = 0; i < C::NB1; i++)
#pragma unroll
            for (int a = 0; a < C::R1; a++)
                s2_cp16(PB + (i * C::R1 + a) * 32, A + (size_t)(a * C::R2 + g + C::TG * i) * n2 + col, true, pol);
        s2_cp_commit();
    };
    if (warp < ntiles) prefetch_tile(warp);
#endif
#if DP_S2_SWP
    cplx v[16], nx[16];
    if (warp < ntiles) load_tile(warp, v);
#endif
    for (int tile = warp; tile < ntiles; tile += S2_WARPS) {
        const int col = tile * C::NS + seq;
#if DP_S2_SWP
        sub_dit_a<C>(v, E, sm.T1, g);
        const bool more = tile + S2_WARPS < ntiles;
        if (more) load_tile(tile + S2_WARPS, nx);      // other columns: independent of this tile's in-place stores
        sub_dit_b<C>(v, E, g);
#else
        cplx v[16];
#if DP_S2_PREFETCH
        s2_cp_wait();
#pragma unroll
        for (int r = 0; r < 16; r++) v[r] = PB[r * 32];
        if (tile + S2_WARPS < ntiles) prefetch_tile(tile + S2_WARPS);
#else
        load_tile(tile, v);
#endif
        sub_dit<C>(v, E, sm.T1, g);
#endif
        __syncwarp();          // every lane of the warp has its column in registers before rows are overwritten
#pragma unroll
        for (int i = 0; i < C::NB2; i++)
#pragma unroll
            for (int k2 = 0; k2 < C::R2; k2++) {
                const int kap = (g + C::TG * i + C::R1 * k2) * n2 + col;
                if (PAIRED) {
                    if (kap <= half || kap >= p.pl.P - half)
                        s2_st_slab(A + kap, make_double2(v[i * C::R2 + k2].x * sc, v[i * C::R2 + k2].y * sc), pol);
                } else if (kap <= half) {
                    s2_st_slab(A + kap, make_double2(v[i * C::R2 + k2].x * sc, -v[i * C::R2 + k2].y * sc), pol);
                }
            }
#if DP_S2_SWP
        if (more) {
#pragma unroll
            for (int r = 0; r < 16; r++) v[r] = nx[r];
        }
#endif
    }
}

// ---- pass 4: the kept bins of FFT_L(a) ------------------------------------------------------------------
// C: SubCfg<m1, m2, 16>, n1' = m1*m2 <= 240.  Column phase (block-wide): a tile of S2_TILE_COLS columns of a;
// thread (c, g) runs the radix-m1 butterfly g < m2 of column c, the tile is exchanged through shared
// memory (column index fastest, so the lags are fetched in contiguous runs), thread (c, k1 < m1) runs
// the radix-m2 butterfly.  Bin phase: thread r < n1' owns the kept bins k = k_lo + r + n1'*b (all of
// residue (k_lo + r) mod n1', so one tile value feeds up to S2_BINSLOTS bins) and advances Horner's rule
//   X_k <- X_k * z_k + C_col[k mod n1'],  z_k = w_L^k,  over the columns in descending order.
// PAIRED: a[m] = rho[m - half] = Y[(half - m) mod P]; the Horner sums give FFT_L(a_a) + i FFT_L(a_b), which are
// separated at the end (each is real up to the phase w_L^(half k) and, for even L, the unpaired lag -L/2).
template <class C, bool PAIRED>
__device__ __noinline__ void s2_pass4(const S2Params& p, const cplx* __restrict__ R, double* __restrict__ spec,
                                      int tid) {
    S2_SMEM_HERE();
    const unsigned long long pol = s2_slab_policy();
    const int L = p.pl.L, half = p.pl.half, n2L = p.pl.n2L;
    constexpr int NP = C::N, R1 = C::R1, R2 = C::R2;
    cplx* Ca = sm.E;                                    // [NP][S2_CPITCH] after the first butterfly
    cplx* Cb = sm.E + NP * S2_CPITCH;                   // [NP][S2_CPITCH] column DFTs
    const int c = tid % S2_TILE_COLS, g = tid / S2_TILE_COLS;      // g < 16
    auto load_tile = [&](int c0, cplx (&v)[16]) {
        const int col = c0 + c;
        if (g < R2 && col < n2L) {
#pragma unroll
            for (int a = 0; a < R1; a++) {
                const int kap = (a * R2 + g) * n2L + col - half;
                if (PAIRED) {
                    v[a] = s2_ld_slab(R + (kap > 0 ? p.pl.P - kap : -kap), pol);      // Y[-kappa]
                } else {
                    v[a] = s2_ld_slab(R + (kap >= 0 ? kap : -kap), pol);
                    if (kap < 0) v[a].y = -v[a].y;                 // negative lag: conj(r[-kappa])
                }
            }
        } else {
#pragma unroll
            for (int a = 0; a < R1; a++) v[a] = make_double2(0.0, 0.0);
        }
    };
    // bins owned by this thread (residue thread r = tid < NP)
    int k_lo = p.lo - half;
    if (k_lo < 0) k_lo += L;
    const int nbins = tid < NP ? (p.nkeep - tid + NP - 1) / NP : 0;   // kept indices tid, tid+NP, ... < nkeep
    cplx acc[S2_BINSLOTS], z[S2_BINSLOTS];
#pragma unroll
    for (int b = 0; b < S2_BINSLOTS; b++) {
        int k = k_lo + tid + NP * b;
        k %= L;
        z[b] = b < nbins ? __ldg(p.g_twL + k) : make_double2(0.0, 0.0);
        acc[b] = make_double2(0.0, 0.0);
    }
    const cplx* crow = Cb + ((k_lo + tid) % NP) * S2_CPITCH;
    const int ntile = (n2L + S2_TILE_COLS - 1) / S2_TILE_COLS;
    cplx v[16];
    load_tile((ntile - 1) * S2_TILE_COLS, v);
    for (int t = ntile - 1; t >= 0; t--) {
        const int c0 = t * S2_TILE_COLS;
        // ---- column DFT_n1' of a[j1*n2L + col] = r[j1*n2L + col - half] ----
        if (g < R2) {
            dft_inplace<R1>(v);
#pragma unroll
            for (int k1 = 0; k1 < R1; k1++)
                Ca[(k1 * R2 + g) * S2_CPITCH + c] = k1 ? cmul(v[k1], sm.Ts[k1 * R2 + g]) : v[0];
        }
        __syncthreads();
        if (g < R1) {                                   // here g plays k1
#pragma unroll
            for (int b = 0; b < R2; b++) v[b] = Ca[(g * R2 + b) * S2_CPITCH + c];
            dft_inplace<R2>(v);
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) Cb[(g + R1 * k2) * S2_CPITCH + c] = v[k2];
        }
        if (t > 0) load_tile(c0 - S2_TILE_COLS, v);    // next tile's lags travel while the bins are updated
        __syncthreads();
        // ---- Horner step over the columns of this tile, last column first ----
        if (nbins > 0) {
            const int ncols = imin(S2_TILE_COLS, n2L - c0);
            for (int cc = ncols - 1; cc >= 0; cc--) {
                const cplx cv = crow[cc];
#pragma unroll
                for (int b = 0; b < S2_BINSLOTS; b++) {
                    if (b < nbins) {
                        const cplx a = acc[b];
                        acc[b].x = fma(a.x, z[b].x, fma(-a.y, z[b].y, cv.x));
                        acc[b].y = fma(a.x, z[b].y, fma(a.y, z[b].x, cv.y));
                    }
                }
            }
        }
    }
    if (!PAIRED) {
#pragma unroll
        for (int b = 0; b < S2_BINSLOTS; b++)
            if (b < nbins) spec[tid + NP * b] = __dmul_rn(hypot(acc[b].x, acc[b].y), p.dt);
    } else {
        // Im r_a[-L/2], Im r_b[-L/2] from rho[-half] = Y[half], rho[+half] = Y[P - half] (even L only)
        double Ia = 0.0, Ib = 0.0;
        if ((L & 1) == 0) {
            const cplx rm = s2_ld_slab(R + half, pol), rp = s2_ld_slab(R + p.pl.P - half, pol);
            Ia = 0.5 * (rm.y - rp.y);
            Ib = 0.5 * (rp.x - rm.x);
        }
        double* __restrict__ spec_b = spec + p.nkeep;             // the pair's second window
#pragma unroll
        for (int b = 0; b < S2_BINSLOTS; b++) {
            if (b < nbins) {
                const int k = (k_lo + tid + NP * b) % L;
                const cplx ph = __ldg(p.g_twL + (int)(((int64_t)half * k) % L));      // w_L^(half k)
                const cplx wt = cmulc(acc[b], ph);                                     // remove the phase
                const double sgn = (k & 1) ? -1.0 : 1.0;
                spec[tid + NP * b] = __dmul_rn(hypot(wt.x + sgn * Ib, Ia), p.dt);
                spec_b[tid + NP * b] = __dmul_rn(hypot(wt.y - sgn * Ia, Ib), p.dt);
            }
        }
    }
}

#define DP_S2_PASS1_L(xptr, t0v, LAY)                                               \
    switch (p.pl.n1) {                                                              \
        case 16: s2_pass1<C16, LAY>(p, xptr, t0v, A, warp, lane); break;            \
        case 32: s2_pass1<C32, LAY>(p, xptr, t0v, A, warp, lane); break;            \
        case 64: s2_pass1<C64, LAY>(p, xptr, t0v, A, warp, lane); break;            \
        case 128: s2_pass1<C128, LAY>(p, xptr, t0v, A, warp, lane); break;          \
        default: s2_pass1<C256, LAY>(p, xptr, t0v, A, warp, lane); break;           \
    }
#define DP_S2_PASS1(xptr, t0v)                                                      \
    if (layout == 0) { DP_S2_PASS1_L(xptr, t0v, 0) }                                \
    else if (layout == 1) { DP_S2_PASS1_L(xptr, t0v, 1) }                           \
    else { DP_S2_PASS1_L(xptr, t0v, 2) }
#define DP_S2_PASS2(MODE)                                                           \
    if (p.pl.n2 == 64) s2_pass2<C64, MODE>(p, A, Sr, warp, lane);                   \
    else s2_pass2<C256, MODE>(p, A, Sr, warp, lane);
#define DP_S2_PASS3(PAIRED)                                                         \
    switch (p.pl.n1) {                                                              \
        case 16: s2_pass3<C16, PAIRED>(p, A, warp, lane); break;                    \
        case 32: s2_pass3<C32, PAIRED>(p, A, warp, lane); break;                    \
        case 64: s2_pass3<C64, PAIRED>(p, A, warp, lane); break;                    \
        case 128: s2_pass3<C128, PAIRED>(p, A, warp, lane); break;                  \
        default: s2_pass3<C256, PAIRED>(p, A, warp, lane); break;                   \
    }

// Items: (series, pair of consecutive windows).  Two windows of one series share everything after the
// forward transforms: |X_a|^2 + i |X_b|^2 goes through ONE second transform, one column pass and one bin
// evaluation (the two autocorrelations are Hermitian, so their transforms separate exactly) -- 3/4 of the
// row work and 1/2 of passes 3-4.  An odd last window runs the single-window passes.
__global__ void __launch_bounds__(S2_THREADS, 1)
spectrum2_kernel(const __grid_constant__ S2Params p) {
    S2_SMEM_HERE();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = tid; k < 6 * 256; k += S2_THREADS) sm.T1[k] = p.g_tables[k];
    __syncthreads();
    cplx* A = p.scratch + (size_t)blockIdx.x * p.slab_elems;
    double* Sr = reinterpret_cast<double*>(A + p.pl.P);            // [P] real: |X_a|^2 of a pair's first window
    const int layout = p.stride_t != 1 ? 2 : (p.tb_len == 0 ? 0 : (p.tb_shift >= 0 ? 1 : 2));
    const int npairs = p.pair ? (p.w_count + 1) / 2 : p.w_count;
    const int64_t nitems = (int64_t)p.S * npairs;
    for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int pj = (int)(item / p.S), s = (int)(item % p.S);
        const int w0 = p.w_first + (p.pair ? 2 * pj : pj);
        const bool paired = p.pair && (w0 + 1 < p.w_first + p.w_count);
        const cplx* x = p.series + (int64_t)s * p.stride_s;
        double* spec = p.spec + ((size_t)s * p.npieces + w0) * p.nkeep;
        // the next item's first window starts its way from HBM to L2 (time-contiguous layouts only)
        auto prefetch_next = [&]() {
            if (item + gridDim.x < nitems && p.stride_t == 1) {
                const int64_t nxt = item + gridDim.x;
                const cplx* xn = p.series + (int64_t)(nxt % p.S) * p.stride_s;
                const int tn = p.starts[p.w_first + (p.pair ? 2 * (int)(nxt / p.S) : (int)(nxt / p.S))];
                for (int i = tid * 8; i < p.pl.L; i += S2_THREADS * 8) s2_prefetch_l2(xn + s2_time_offset(p, tn + i));
            }
        };
        if (!paired) {
            if (!(p.skip_mask & 1)) { DP_S2_PASS1(x, p.starts[w0]) }
            __syncthreads();
            if (!(p.skip_mask & 2)) { DP_S2_PASS2(0) }
            __syncthreads();
            if (!(p.skip_mask & 4)) { DP_S2_PASS3(false) }
            __syncthreads();
            prefetch_next();
            if (!(p.skip_mask & 8)) switch (p.pl.m1 * 32 + p.pl.m2) {
#define DP_S2_CASE(a, b) case (a) * 32 + (b): s2_pass4<SubCfg<a, b, 16>, false>(p, A, spec, tid); break;
                DP_S2_MIXED_MENU(DP_S2_CASE)
#undef DP_S2_CASE
                default: break;
            }
        } else {
            if (!(p.skip_mask & 1)) { DP_S2_PASS1(x, p.starts[w0]) }
            __syncthreads();
            // the pair's second window is needed right after this pass: bring it to L2 now
            if (p.stride_t == 1)
                for (int i = tid * 8; i < p.pl.L; i += S2_THREADS * 8) s2_prefetch_l2(x + s2_time_offset(p, p.starts[w0 + 1] + i));
            if (!(p.skip_mask & 2)) { DP_S2_PASS2(1) }
            __syncthreads();
            if (!(p.skip_mask & 1)) { DP_S2_PASS1(x, p.starts[w0 + 1]) }
            __syncthreads();
            if (!(p.skip_mask & 2)) { DP_S2_PASS2(2) }
            __syncthreads();
            if (!(p.skip_mask & 4)) { DP_S2_PASS3(true) }
            __syncthreads();
            prefetch_next();
            if (!(p.skip_mask & 8)) switch (p.pl.m1 * 32 + p.pl.m2) {
#define DP_S2_CASE(a, b) case (a) * 32 + (b): s2_pass4<SubCfg<a, b, 16>, true>(p, A, spec, tid); break;
                DP_S2_MIXED_MENU(DP_S2_CASE)
#undef DP_S2_CASE
                default: break;
            }
        }
        __syncthreads();
    }
}
#undef DP_S2_PASS1
#undef DP_S2_PASS1_L
#undef DP_S2_PASS2
#undef DP_S2_PASS3

// ---- host planning ---------------------------------------------------------------------------------------
static inline void s2_radices(int n, int* r1, int* r2) {       // the SubCfg used for a power-of-two length
    switch (n) {
        case 16: *r1 = 16; *r2 = 1; break;
        case 32: *r1 = 8; *r2 = 4; break;
        case 64: *r1 = 8; *r2 = 8; break;
        case 128: *r1 = 16; *r2 = 8; break;
        default: *r1 = 16; *r2 = 16; break;
    }
}

// returns false when this (L, kept-bin range) is better served by the general engine
static bool make_s2_plan(int L, int nkeep, S2Plan* pl) {
    if (L < 48) return false;
    const int half = L / 2;
    int P = 1024;
    while (P < L + half) P *= 2;
    if (P > 65536) return false;
    pl->L = L; pl->half = half; pl->P = P;
    pl->n2 = P >= 4096 ? 256 : 64;
    pl->n1 = P / pl->n2;
    static const int menu[][2] = {
#define DP_S2_ROW(a, b) {a, b},
        DP_S2_MIXED_MENU(DP_S2_ROW)
#undef DP_S2_ROW
    };
    int best = -1, best_n = 0;
    for (size_t i = 0; i < sizeof(menu) / sizeof(menu[0]); i++) {
        const int n = menu[i][0] * menu[i][1];
        if (L % n == 0 && n > best_n) { best = (int)i; best_n = n; }
    }
    if (best < 0) return false;
    pl->m1 = menu[best][0]; pl->m2 = menu[best][1];
    pl->n2L = L / best_n;
    if (nkeep > S2_BINSLOTS * best_n) return false;             // bins per residue thread
    // direct evaluation of the kept bins costs nkeep * n2L complex MACs per item: keep it below the
    // cost of the transforms themselves
    if ((double)nkeep * pl->n2L > 12.0 * (double)L) return false;
    return true;
}

// tables (256 entries each): T1, T2, T2f, Ts, fine, coarse -- see S2Smem
static __global__ void __launch_bounds__(256) s2_tables_kernel(cplx* __restrict__ tab, int a1, int a2, int b1, int b2,
                                                               int m1, int m2, int P) {
    const int k = threadIdx.x;
    auto root = [](long long num, long long den) {
        double s, c;
        sincospi(2.0 * (double)(num % den) / (double)den, &s, &c);
        return make_double2(c, -s);
    };
    tab[k] = root((long long)(k / a2) * (k % a2), a1 * a2);                 // T1[k1*R2 + j2]
    tab[256 + k] = root((long long)(k / b2) * (k % b2), b1 * b2);           // T2
    tab[512 + k] = root((long long)(k / b1) * (k % b1), b1 * b2);           // T2f[c*R1 + k1]
    tab[768 + k] = root((long long)(k / m2) * (k % m2), m1 * m2);           // Ts
    tab[1024 + k] = root(k, P);
    tab[1280 + k] = root(256LL * k, P);
}

}  // namespace dp
