// A7, long-window path ("v3") of the FFT power spectrum: windows whose padded length P = 2^p >= L + L/2 exceeds
// what one CTA's L2-resident slab can hold in the v2 engine (P > 65536, e.g. cfg 5-B: ONE window of 2^17 samples,
// P = 2^18), or whose kept-bin range is too wide for the v2 bin pass.  Reference arithmetic:
// dynaphopy/power_spectrum/__init__.py:226-245 (np.correlate 'same' / L, |FFT_L| * dt).
//
// P = n1 * n2 * n3 (each 16..256, register codelets of dp_fft_reg.cuh).  A BATCH of series goes through seven
// grid-wide passes over a slab of `batch` x P complex values that is sized to stay in the 126 MB L2 (HBM sees the
// window once and the kept bins once); every warp owns whole sub-transforms, there is no block barrier in passes 1-5:
//   1  x[j1 M + c]            -> n1-point FFT over j1 (stride M = n2 n3), * w_P^(c k1)         -> A[k1][c]
//   2  A[k1][j2 n3 + j3]      -> n2-point FFT over j2 (stride n3),       * w_M^(j3 k2)         -> A[k1][k2][j3]
//   3  rows A[k1][k2][.]      -> n3-point FFT -> |X|^2 -> n3-point FFT (DIF: digit order in, natural order out)
//                                                                       * w_M^(k2 g3)         -> A[k1][k2][g3]
//   4  A[k1][k2][g3]          -> n2-point FFT over k2,                   * w_P^(k1 (g2 n3 + g3)) -> A[k1][g2][g3]
//   5  A[k1][g]               -> n1-point FFT over k1 -> Y[kappa1 M + g]; r[kappa] = conj(Y[kappa]) / (P L) kept for
//                                the lags 0..L/2 only (the autocorrelation is Hermitian)       -> A[kappa]
//   6a a[m] = r[m - L/2], L = n1' n2L: column DFTs C_j2[k1] = sum_j1 a[j1 n2L + j2] w_n1'^(j1 k1) -> C[j2][k1]
//   6b the kept bins only: FFT_L(a)[k] = sum_j2 w_L^(j2 k) C_j2[k mod n1'] by Horner's rule in z = w_L^k   -> spec
// The forward transform leaves X in digit-scrambled order; |X|^2 is position independent and the second transform
// consumes that order, so both transforms need five passes together instead of six.
#pragma once

#include "dp_fft_reg.cuh"

namespace dp {

constexpr int S3_THREADS = 256;
constexpr int S3_WARPS = S3_THREADS / 32;
constexpr int S3_TILE_COLS = S3_THREADS / 16;      // pass 6a: 16 threads per column
constexpr int S3_CPITCH = S3_TILE_COLS + 1;        // odd pitch (16-byte elements)
constexpr int S3_HGROUPS = S3_THREADS / 32;        // pass 6b: warp h owns the bins j = h (mod 8) of its 32 residues
constexpr int S3_BCOLS = 32;                       // pass 6b: columns per shared-memory stage (16 KB)
constexpr int S3_BSTAGES = 4;                      // ring depth
constexpr size_t S3_BINS_B_SMEM = (size_t)S3_BSTAGES * S3_BCOLS * 32 * sizeof(double2);

typedef SubCfg<16, 1, 1> D16;
typedef SubCfg<8, 4, 4> D32;
typedef SubCfg<8, 8, 4> D64;
typedef SubCfg<8, 8, 8> R64;         // row pass: 8 points per lane (the next tile is prefetched), 8 lanes = one 128-byte run
typedef SubCfg<16, 8, 8> D128;
typedef SubCfg<16, 16, 16> D256;

// n1' menu of pass 6a (16 threads per column): the v2 menu plus 16 x 16
#define DP_S3_MENU(X) \
    X(16, 16) X(16, 12) X(16, 10) X(15, 10) X(16, 9) X(16, 8) X(12, 10) X(10, 10) X(16, 6) X(16, 5) \
    X(8, 8) X(10, 5) X(8, 5) X(8, 4) X(5, 5) X(4, 4)

struct S3Plan {
    int L, half, P;
    int n1, n2, n3, M;        // P = n1 * n2 * n3, M = n2 * n3
    int m1, m2, np, n2L;      // L = np * n2L, np = m1 * m2 (menu entry)
    int r_lo;                 // rows kappa1 < r_lo of Y hold the lags 0..half
};

struct S3Params {
    const cplx* series;
    int64_t stride_s, stride_t;
    int tb_len, tb_shift;
    int64_t tb_stride;
    int s0, bs;               // this launch works on the series [s0, s0 + bs)
    int t0;                   // first sample of the window
    S3Plan pl;
    const cplx* g_tab;        // 14 x 256, see s3_tables_kernel
    const cplx* g_twL;        // w_L^k, k < L
    cplx* slab;               // bs x P
    cplx* cbuf;               // bs x L
    double* spec;             // kept bins of series s0 for this window; next series spec_stride doubles further
    int64_t spec_stride;
    int lo, nkeep;
    double dt, acf_scale;     // acf_scale = 1 / (P * L)
};

// table blocks of 256 entries
constexpr int S3_TAB_DIT = 0;       // + log2(n) - 4: [k1*R2 + j2] = w_n^(j2 k1)
constexpr int S3_TAB_DIF = 5;       // + log2(n) - 4: [c*R1 + k1]  = w_n^(k1 c)
constexpr int S3_TAB_FINE = 10;     // w_P^j, w_P^(256 j), w_P^(65536 j)
constexpr int S3_TAB_TS = 13;       // [k1*m2 + j2] = w_n1'^(j2 k1)
constexpr int S3_TAB_BLOCKS = 14;

__host__ __device__ constexpr int s3_log2(int n) { return n <= 1 ? 0 : 1 + s3_log2(n / 2); }

// w_P^idx, idx < P <= 2^24, from three exactly rounded 256-entry tables (at most two complex products)
__device__ __forceinline__ cplx tw3level(const cplx* __restrict__ fine, int idx) {
    cplx r = fine[idx & 255];
    const int mi = (idx >> 8) & 255, ci = idx >> 16;
    if (mi) r = cmul(fine[256 + mi], r);
    if (ci) r = cmul(fine[512 + ci], r);
    return r;
}

__device__ __forceinline__ int64_t s3_time_offset(const S3Params& p, int t) {
    if (p.tb_len == 0) return (int64_t)t * p.stride_t;
    const int b = p.tb_shift >= 0 ? (t >> p.tb_shift) : t / p.tb_len;
    return (int64_t)b * p.tb_stride + (int64_t)(t - b * p.tb_len) * p.stride_t;
}

// ---- passes 1, 2, 4, 5: strided sub-transforms, one tile of C::NS adjacent columns per warp and step --------
template <class C, int PASS>
__global__ void __launch_bounds__(S3_THREADS, 2) s3_col_kernel(const __grid_constant__ S3Params p) {
    DP_DYNSMEM(cplx, smem);
    cplx* T = smem;                    // stage twiddles of the C::N-point sub-transform
    cplx* fine = smem + 256;           // fine / mid / coarse
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = tid; k < 256; k += S3_THREADS) T[k] = p.g_tab[(S3_TAB_DIT + s3_log2(C::N) - 4) * 256 + k];
    for (int k = tid; k < 768; k += S3_THREADS) fine[k] = p.g_tab[S3_TAB_FINE * 256 + k];
    __syncthreads();
    pdl_wait();                        // the tables are older than the previous kernel; the slab is not
    const int seq = lane % C::NS, g = lane / C::NS;
    cplx* E = smem + 1024 + (size_t)warp * SUBFFT_EWARP_MAX + seq * C::ESEQ;
    const int P = p.pl.P, M = p.pl.M, n1 = p.pl.n1, n3 = p.pl.n3, L = p.pl.L;
    constexpr bool OUTER = (PASS == 1 || PASS == 5);       // n1-point transforms at stride M
    const int tps = OUTER ? M / C::NS : n1 * (n3 / C::NS); // tiles per series
    const int64_t total = (int64_t)p.bs * tps;
    const double sc = p.acf_scale;
    // tile -> slab address of the warp's sub-transforms
    auto decode = [&](int64_t tile, int& b, int& col, int& k1, cplx*& base, int& stride) {
        b = (int)(tile / tps);
        const int tt = (int)(tile % tps);
        cplx* A = p.slab + (size_t)b * P;
        k1 = 0;
        if (OUTER) { col = tt * C::NS + seq; base = A + col; stride = M; }
        else {
            const int tpb = n3 / C::NS;
            k1 = tt / tpb;
            col = (tt - k1 * tpb) * C::NS + seq;
            base = A + (size_t)k1 * M + col;
            stride = n3;
        }
    };
    const bool plain = p.tb_len == 0 && p.stride_t == 1;
    auto load = [&](cplx* dst, int b, int col, const cplx* base, int stride) {
        if (PASS == 1) {
            const cplx* __restrict__ x = p.series + (int64_t)(p.s0 + b) * p.stride_s;
#pragma unroll
            for (int i = 0; i < C::NB1; i++)
#pragma unroll
                for (int a = 0; a < C::R1; a++) {
                    const int idx = (a * C::R2 + g + C::TG * i) * M + col;
                    cplx val = make_double2(0.0, 0.0);
                    if (idx < L) val = plain ? __ldcs(x + p.t0 + idx) : __ldcs(x + s3_time_offset(p, p.t0 + idx));
                    dst[i * C::R1 + a] = val;
                }
        } else {
#pragma unroll
            for (int i = 0; i < C::NB1; i++)
#pragma unroll
                for (int a = 0; a < C::R1; a++) dst[i * C::R1 + a] = __ldcg(base + (size_t)(a * C::R2 + g + C::TG * i) * stride);
        }
    };
    // sub-transforms of <= 8 points per lane: the samples of the next tile are fetched while this one is transformed
    constexpr int NV = C::NB1 * C::R1;
    constexpr bool PRE = NV <= 8;
    cplx nx[PRE ? NV : 1];
    const int64_t tile0 = (int64_t)blockIdx.x * S3_WARPS + warp, tstep = (int64_t)gridDim.x * S3_WARPS;
    if (PRE && tile0 < total) {
        int b, col, k1, stride; cplx* base;
        decode(tile0, b, col, k1, base, stride);
        load(nx, b, col, base, stride);
    }
    for (int64_t tile = tile0; tile < total; tile += tstep) {
        int b, col, k1, stride;
        cplx* base;
        decode(tile, b, col, k1, base, stride);
        cplx v[16];
        if (PRE) {
#pragma unroll
            for (int i = 0; i < NV; i++) v[i] = nx[i];
            if (tile + tstep < total) {
                int b2, col2, k12, stride2; cplx* base2;
                decode(tile + tstep, b2, col2, k12, base2, stride2);
                load(nx, b2, col2, base2, stride2);
            }
        } else {
            load(v, b, col, base, stride);
        }
        __syncwarp();                  // the exchange slab of the previous tile has been read by every lane
        sub_dit<C>(v, E, T, g);
        // lane g, register i*R2 + k2 now holds output kk = (g + TG i) + R1 k2
        if (PASS == 5) {
            const int r_lo = p.pl.r_lo;
#pragma unroll
            for (int i = 0; i < C::NB2; i++)
#pragma unroll
                for (int k2 = 0; k2 < C::R2; k2++) {
                    const int kk = g + C::TG * i + C::R1 * k2;
                    if (kk < r_lo) __stcg(base + (size_t)kk * stride, make_double2(v[i * C::R2 + k2].x * sc, -v[i * C::R2 + k2].y * sc));
                }
        } else {
            // twiddle w_P^(off + step kk): geometric in k2
            const int step = PASS == 1 ? col : PASS == 2 ? n1 * col : k1 * n3;
            const int off = PASS == 4 ? k1 * col : 0;
            const cplx ratio = tw3level(fine, step * C::R1);
#pragma unroll
            for (int i = 0; i < C::NB2; i++) {
                const int kk0 = g + C::TG * i;
                cplx tw = tw3level(fine, off + step * kk0);
#pragma unroll
                for (int k2 = 0; k2 < C::R2; k2++) {
                    __stcg(base + (size_t)(kk0 + C::R1 * k2) * stride, cmul(v[i * C::R2 + k2], tw));
                    if (k2 + 1 < C::R2) tw = cmul(tw, ratio);
                }
            }
        }
    }
}

// ---- pass 3: rows: FFT -> |.|^2 -> FFT (DIF) -> twiddle, in place ------------------------------------------
template <class C>
__global__ void __launch_bounds__(S3_THREADS, 2) s3_row_kernel(const __grid_constant__ S3Params p) {
    DP_DYNSMEM(cplx, smem);
    cplx* T = smem;                    // DIT stage twiddles, DIF stage twiddles, fine / mid / coarse
    cplx* Tf = smem + 256;
    cplx* fine = smem + 512;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = tid; k < 256; k += S3_THREADS) {
        T[k] = p.g_tab[(S3_TAB_DIT + s3_log2(C::N) - 4) * 256 + k];
        Tf[k] = p.g_tab[(S3_TAB_DIF + s3_log2(C::N) - 4) * 256 + k];
    }
    for (int k = tid; k < 768; k += S3_THREADS) fine[k] = p.g_tab[S3_TAB_FINE * 256 + k];
    __syncthreads();
    pdl_wait();
    // rows are contiguous along the element index: group lane fastest, so a quarter-warp reads one 128-byte run
    // (the column passes map the sequence index fastest for the same reason)
    const int seq = lane / C::TG, g = lane % C::TG;
    cplx* E = smem + 1280 + (size_t)warp * SUBFFT_EWARP_MAX + seq * C::ESEQ;
    const int P = p.pl.P, n1 = p.pl.n1, n2 = p.pl.n2, n3 = p.pl.n3;
    const int tps = n1 * n2 / C::NS;
    const int64_t total = (int64_t)p.bs * tps;
    auto row_of = [&](int64_t tile, int& row) -> cplx* {
        const int b = (int)(tile / tps), tt = (int)(tile % tps);
        row = tt * C::NS + seq;
        return p.slab + (size_t)b * P + (size_t)row * n3;
    };
    auto load = [&](cplx* dst, const cplx* Ar) {
#pragma unroll
        for (int i = 0; i < C::NB1; i++)
#pragma unroll
            for (int a = 0; a < C::R1; a++) dst[i * C::R1 + a] = __ldcg(Ar + a * C::R2 + g + C::TG * i);
    };
    constexpr int NV = C::NB1 * C::R1;
    constexpr bool PRE = NV <= 8;                  // see s3_col_kernel
    cplx nx[PRE ? NV : 1];
    const int64_t tile0 = (int64_t)blockIdx.x * S3_WARPS + warp, tstep = (int64_t)gridDim.x * S3_WARPS;
    if (PRE && tile0 < total) { int row; load(nx, row_of(tile0, row)); }
    for (int64_t tile = tile0; tile < total; tile += tstep) {
        int row;
        cplx* __restrict__ Ar = row_of(tile, row);
        cplx v[16];
        if (PRE) {
#pragma unroll
            for (int i = 0; i < NV; i++) v[i] = nx[i];
            if (tile + tstep < total) { int row2; load(nx, row_of(tile + tstep, row2)); }
        } else {
            load(v, Ar);
        }
        __syncwarp();
        sub_dit<C>(v, E, T, g);
#pragma unroll
        for (int r = 0; r < C::NB2 * C::R2; r++) v[r] = make_double2(v[r].x * v[r].x + v[r].y * v[r].y, 0.0);
        __syncwarp();
        sub_dif<C>(v, E, Tf, g);
        // natural order: register i*R1 + d holds element d*R2 + c, c = g + TG i; twiddle w_M^(k2 elem) = w_P^(n1 k2 elem)
        const int step = n1 * (row % n2);
        const cplx ratio = tw3level(fine, step * C::R2);
#pragma unroll
        for (int i = 0; i < C::NB1; i++) {
            const int c = g + C::TG * i;
            cplx tw = tw3level(fine, step * c);
#pragma unroll
            for (int d = 0; d < C::R1; d++) {
                __stcg(Ar + d * C::R2 + c, cmul(v[i * C::R1 + d], tw));
                if (d + 1 < C::R1) tw = cmul(tw, ratio);
            }
        }
    }
}

// ---- pass 6a: column DFTs of a[m] = r[m - half] -> C[j2][k1] --------------------------------------------------
// C: SubCfg<m1, m2, 16>.  Thread (c, g): column c of the tile, radix-m1 butterfly g < m2, exchange through shared
// memory (column index fastest: the lags are fetched in contiguous runs), then radix-m2 butterfly g < m1.
template <class C>
__global__ void __launch_bounds__(S3_THREADS) s3_bins_a_kernel(const __grid_constant__ S3Params p) {
    DP_DYNSMEM(cplx, smem);
    cplx* Ts = smem;                   // [256]
    cplx* Ca = smem + 256;             // [NP][S3_CPITCH]
    constexpr int NP = C::N, R1 = C::R1, R2 = C::R2;
    const int tid = threadIdx.x;
    for (int k = tid; k < 256; k += S3_THREADS) Ts[k] = p.g_tab[S3_TAB_TS * 256 + k];
    __syncthreads();
    pdl_wait();
    const int L = p.pl.L, half = p.pl.half, n2L = p.pl.n2L, P = p.pl.P;
    const int c = tid % S3_TILE_COLS, g = tid / S3_TILE_COLS;
    const int ntile = (n2L + S3_TILE_COLS - 1) / S3_TILE_COLS;
    const int64_t total = (int64_t)p.bs * ntile;
    for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
        const int b = (int)(item / ntile), c0 = (int)(item % ntile) * S3_TILE_COLS;
        const cplx* __restrict__ R = p.slab + (size_t)b * P;
        cplx* __restrict__ Cout = p.cbuf + (size_t)b * L;
        const int col = c0 + c;
        cplx v[16];
        if (g < R2 && col < n2L) {
#pragma unroll
            for (int a = 0; a < R1; a++) {
                const int kap = (a * R2 + g) * n2L + col - half;
                v[a] = __ldcg(R + (kap >= 0 ? kap : -kap));
                if (kap < 0) v[a].y = -v[a].y;                 // negative lag: conj(r[-kappa])
            }
            dft_inplace<R1>(v);
#pragma unroll
            for (int k1 = 0; k1 < R1; k1++) Ca[(k1 * R2 + g) * S3_CPITCH + c] = k1 ? cmul(v[k1], Ts[k1 * R2 + g]) : v[0];
        } else if (g < R2) {
#pragma unroll
            for (int k1 = 0; k1 < R1; k1++) Ca[(k1 * R2 + g) * S3_CPITCH + c] = make_double2(0.0, 0.0);
        }
        __syncthreads();
        if (g < R1) {                                   // here g plays k1
#pragma unroll
            for (int bb = 0; bb < R2; bb++) v[bb] = Ca[(g * R2 + bb) * S3_CPITCH + c];
            dft_inplace<R2>(v);
        }
        __syncthreads();                                // every row group has been read: results go back in place
        if (g < R1) {
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) Ca[(g + R1 * k2) * S3_CPITCH + c] = v[k2];
        }
        __syncthreads();
        const int ncols = imin(S3_TILE_COLS, n2L - c0);
        for (int e = tid; e < NP * ncols; e += S3_THREADS) {
            const int cc = e / NP, k1 = e - cc * NP;
            __stcg(Cout + (size_t)(c0 + cc) * NP + k1, Ca[k1 * S3_CPITCH + cc]);
        }
        __syncthreads();
    }
}

// ---- pass 6b: the kept bins by Horner's rule -------------------------------------------------------------------
// Kept bin e (sorted-frequency index lo + e) is FFT bin k = (k_lo + e) mod L of residue (k_lo + e) mod n1'.  A CTA owns
// 32 consecutive residues of one series: lane = residue, warp h owns the bins e = r + n1' j with j = h (mod 8).  The
// columns C[j2][32 residues] (one 512-byte run each) stream from the highest j2 down through a ring of shared-memory
// stages filled with cp.async, S3_BSTAGES - 1 stages ahead of the arithmetic: every value is fetched once per CTA and
// feeds SLOTS bins of each of the eight warps.
__device__ __forceinline__ void s3_cp16(cplx* dst_smem, const cplx* src) {
#ifdef DP_EMUL
    *dst_smem = *src;
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void s3_cp_commit() {
#ifndef DP_EMUL
    asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
}
template <int N> __device__ __forceinline__ void s3_cp_wait() {
#ifndef DP_EMUL
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
#endif
}

template <int SLOTS>
__global__ void __launch_bounds__(S3_THREADS) s3_bins_b_kernel(const __grid_constant__ S3Params p) {
    DP_DYNSMEM(cplx, ring);            // [S3_BSTAGES][S3_BCOLS][32]
    const int L = p.pl.L, half = p.pl.half, NP = p.pl.np, n2L = p.pl.n2L;
    const int lane = threadIdx.x & 31, h = threadIdx.x >> 5;
    const int nrb = (NP + 31) / 32;
    const int b = blockIdx.x / nrb, r = (blockIdx.x % nrb) * 32 + lane;
    pdl_wait();
    if (b >= p.bs) return;
    int k_lo = p.lo - half;
    if (k_lo < 0) k_lo += L;
    const cplx* __restrict__ Cin = p.cbuf + (size_t)b * L + (k_lo + (r < NP ? r : 0)) % NP;
    double* __restrict__ spec = p.spec + (int64_t)b * p.spec_stride;
    const int nres = r < NP ? (p.nkeep - r + NP - 1) / NP : 0;        // kept bins of this residue: e = r + NP j, j < nres
    const int nres_max = (p.nkeep + NP - 1) / NP;
    const int nchunks = (n2L + S3_BCOLS - 1) / S3_BCOLS;
    auto issue = [&](int c) {          // chunk c = the columns n2L-1 - c*S3_BCOLS - i, i < S3_BCOLS; row i of its stage
        if (c < nchunks) {
            const int hi = n2L - 1 - c * S3_BCOLS;
            cplx* dst = ring + (size_t)(c % S3_BSTAGES) * (S3_BCOLS * 32);
#pragma unroll
            for (int u = 0; u < S3_BCOLS / S3_HGROUPS; u++) {
                const int i = h + S3_HGROUPS * u;
                if (hi - i >= 0) s3_cp16(dst + i * 32 + lane, Cin + (size_t)(hi - i) * NP);
            }
        }
        s3_cp_commit();                // one group per call, empty or not: the wait below counts groups
    };
    for (int j0 = 0; j0 < nres_max; j0 += S3_HGROUPS * SLOTS) {        // rounds (uniform over the CTA)
        cplx acc[SLOTS], z[SLOTS];
#pragma unroll
        for (int s = 0; s < SLOTS; s++) {
            const int j = j0 + h + S3_HGROUPS * s;
            acc[s] = make_double2(0.0, 0.0);
            z[s] = make_double2(0.0, 0.0);
            if (j < nres) z[s] = __ldg(p.g_twL + (k_lo + r + NP * j) % L);
        }
        const bool active = j0 + h < nres_max;                         // warp-uniform
        __syncthreads();               // the previous round has left the ring
        for (int c = 0; c < S3_BSTAGES - 1; c++) issue(c);
        for (int c = 0; c < nchunks; c++) {
            s3_cp_wait<S3_BSTAGES - 2>();                              // this thread's part of chunk c has landed
            __syncthreads();           // ... everybody's has, and chunk c-1 has been consumed by every warp
            issue(c + S3_BSTAGES - 1); // into the stage of chunk c-1
            if (active) {
                const cplx* st = ring + (size_t)(c % S3_BSTAGES) * (S3_BCOLS * 32) + lane;
                const int cnt = imin(S3_BCOLS, n2L - c * S3_BCOLS);
                int i = 0;
                for (; i + 4 <= cnt; i += 4) {
                    cplx cv[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) cv[u] = st[(i + u) * 32];
#pragma unroll
                    for (int u = 0; u < 4; u++)
#pragma unroll
                        for (int s = 0; s < SLOTS; s++) {
                            const cplx a = acc[s];
                            acc[s].x = fma(a.x, z[s].x, fma(-a.y, z[s].y, cv[u].x));
                            acc[s].y = fma(a.x, z[s].y, fma(a.y, z[s].x, cv[u].y));
                        }
                }
                for (; i < cnt; i++) {
                    const cplx cv = st[i * 32];
#pragma unroll
                    for (int s = 0; s < SLOTS; s++) {
                        const cplx a = acc[s];
                        acc[s].x = fma(a.x, z[s].x, fma(-a.y, z[s].y, cv.x));
                        acc[s].y = fma(a.x, z[s].y, fma(a.y, z[s].x, cv.y));
                    }
                }
            }
        }
        s3_cp_wait<0>();
#pragma unroll
        for (int s = 0; s < SLOTS; s++) {
            const int j = j0 + h + S3_HGROUPS * s;
            if (j < nres) spec[r + NP * j] = __dmul_rn(hypot(acc[s].x, acc[s].y), p.dt);
        }
    }
}

// tables, blocks of 256 entries: DIT and DIF stage twiddles of the five sub-transform sizes, fine / mid / coarse
// roots of P, stage twiddles of n1' = m1 m2
static __global__ void __launch_bounds__(256) s3_tables_kernel(cplx* __restrict__ tab, int m1, int m2, int P) {
    const int k = threadIdx.x;
    auto root = [](long long num, long long den) {
        double s, c;
        sincospi(2.0 * (double)(num % den) / (double)den, &s, &c);
        return make_double2(c, -s);
    };
    const int r1s[5] = {16, 8, 8, 16, 16}, r2s[5] = {1, 4, 8, 8, 16};
    for (int i = 0; i < 5; i++) {
        const int r1 = r1s[i], r2 = r2s[i];
        tab[(S3_TAB_DIT + i) * 256 + k] = root((long long)(k / r2) * (k % r2), r1 * r2);      // [k1*R2 + j2]
        tab[(S3_TAB_DIF + i) * 256 + k] = root((long long)(k / r1) * (k % r1), r1 * r2);      // [c*R1 + k1]
    }
    tab[S3_TAB_FINE * 256 + k] = root(k, P);
    tab[(S3_TAB_FINE + 1) * 256 + k] = root(256LL * k, P);
    tab[(S3_TAB_FINE + 2) * 256 + k] = root(65536LL * k, P);
    tab[S3_TAB_TS * 256 + k] = root((long long)(k / m2) * (k % m2), m1 * m2);
}

// ---- host planning -----------------------------------------------------------------------------------------------
// `wide` = the v2 engine declined because of its bin-pass limits (P may be small); otherwise P > 65536 brought us here
static bool make_s3_plan(int L, int nkeep, S3Plan* pl) {
    if (L < 48) return false;
    const int half = L / 2;
    int p2 = 13;
    while ((1LL << p2) < (long long)L + half) p2++;
    if (p2 > 24) return false;
    int e3 = (p2 + 2) / 3;
    if (e3 > 8) e3 = 8;
    const int e1 = (p2 - e3) / 2, e2 = p2 - e3 - e1;
    if (e1 < 4 || e2 < 4 || e2 > 8) return false;
    pl->L = L; pl->half = half; pl->P = 1 << p2;
    pl->n1 = 1 << e1; pl->n2 = 1 << e2; pl->n3 = 1 << e3;
    pl->M = pl->n2 * pl->n3;
    const int ns2 = pl->n2 == 16 ? 32 : pl->n2 <= 64 ? 8 : pl->n2 == 128 ? 4 : 2;      // columns per warp tile of pass 2 / 4
    if (pl->n3 % ns2) return false;
    static const int menu[][2] = {
#define DP_S3_ROW(a, b) {a, b},
        DP_S3_MENU(DP_S3_ROW)
#undef DP_S3_ROW
    };
    int best = -1, best_n = 0;
    for (size_t i = 0; i < sizeof(menu) / sizeof(menu[0]); i++) {
        const int n = menu[i][0] * menu[i][1];
        if (L % n == 0 && n > best_n) { best = (int)i; best_n = n; }
    }
    if (best < 0) return false;
    pl->m1 = menu[best][0]; pl->m2 = menu[best][1];
    pl->np = best_n;
    pl->n2L = L / best_n;
    // direct evaluation of the kept bins costs nkeep * n2L complex MACs per window: keep it comparable to the transforms
    if ((double)nkeep * pl->n2L > 32.0 * (double)L) return false;
    pl->r_lo = imin(pl->n1, half / pl->M + 1);
    return true;
}

constexpr size_t S3_COL_SMEM = (size_t)(1024 + S3_WARPS * SUBFFT_EWARP_MAX) * sizeof(cplx);
constexpr size_t S3_ROW_SMEM = (size_t)(1280 + S3_WARPS * SUBFFT_EWARP_MAX) * sizeof(cplx);
static inline size_t s3_bins_a_smem(int np) { return (size_t)(256 + np * S3_CPITCH) * sizeof(cplx); }

template <int PASS> static int s3_launch_col(int n, int grid, cudaStream_t st, const S3Params& p) {
#define DP_S3_COL(CFG)                                                          \
    {                                                                           \
        auto k = s3_col_kernel<CFG, PASS>;                                      \
        DP_SET_SMEM_ONCE(k, S3_COL_SMEM);                                       \
        DP_CUDA_OK(DP_LAUNCH_PDL(k, dim3(grid), dim3(S3_THREADS), S3_COL_SMEM, st, p)); \
    }
    switch (n) {
        case 16: DP_S3_COL(D16) break;
        case 32: DP_S3_COL(D32) break;
#ifdef DP_S3_COL64_PRE
        case 64: DP_S3_COL(R64) break;
#else
        case 64: DP_S3_COL(D64) break;
#endif
        case 128: DP_S3_COL(D128) break;
        default: DP_S3_COL(D256) break;
    }
#undef DP_S3_COL
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

static int s3_launch_row(int n, int grid, cudaStream_t st, const S3Params& p) {
#define DP_S3_ROWK(CFG)                                                         \
    {                                                                           \
        auto k = s3_row_kernel<CFG>;                                            \
        DP_SET_SMEM_ONCE(k, S3_ROW_SMEM);                                       \
        DP_CUDA_OK(DP_LAUNCH_PDL(k, dim3(grid), dim3(S3_THREADS), S3_ROW_SMEM, st, p)); \
    }
    switch (n) {
        case 16: DP_S3_ROWK(D16) break;
        case 32: DP_S3_ROWK(D32) break;
        case 64: DP_S3_ROWK(R64) break;
        case 128: DP_S3_ROWK(D128) break;
        default: DP_S3_ROWK(D256) break;
    }
#undef DP_S3_ROWK
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

static int s3_launch_bins_a(int grid, cudaStream_t st, const S3Params& p) {
    const size_t smem = s3_bins_a_smem(p.pl.np);
    switch (p.pl.m1 * 32 + p.pl.m2) {
#define DP_S3_CASE(a, b)                                                        \
    case (a) * 32 + (b): {                                                      \
        auto k = s3_bins_a_kernel<SubCfg<a, b, 16>>;                            \
        DP_SET_SMEM_ONCE(k, smem);                                              \
        DP_CUDA_OK(DP_LAUNCH_PDL(k, dim3(grid), dim3(S3_THREADS), smem, st, p)); \
    } break;
        DP_S3_MENU(DP_S3_CASE)
#undef DP_S3_CASE
        default: set_error("power_fft: internal: no pass-6a kernel for %d x %d", p.pl.m1, p.pl.m2); return DP_ERR_UNSUPPORTED;
    }
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

// all seven passes of one window for one batch of series
static int s3_run_window(const S3Params& p, int max_ctas, cudaStream_t st) {
    const S3Plan& pl = p.pl;
    const int cap = max_ctas > 0 ? max_ctas : 2 * imax(sm_count(), 1);
    auto grid_for = [&](int64_t warp_tiles) { return (int)imin<int64_t>((warp_tiles + S3_WARPS - 1) / S3_WARPS, cap); };
#ifdef DP_S3_COL64_PRE
    auto ns_of = [](int n) { return n == 16 ? 32 : n == 32 ? 8 : (n == 64 || n == 128) ? 4 : 2; };
#else
    auto ns_of = [](int n) { return n == 16 ? 32 : (n == 32 || n == 64) ? 8 : n == 128 ? 4 : 2; };      // column passes
#endif
    auto ns_row = [](int n) { return n == 16 ? 32 : n == 32 ? 8 : (n == 64 || n == 128) ? 4 : 2; };       // row pass
    const int64_t bs = p.bs;
    int rc;
    if ((rc = s3_launch_col<1>(pl.n1, grid_for(bs * (pl.M / ns_of(pl.n1))), st, p))) return rc;
    if ((rc = s3_launch_col<2>(pl.n2, grid_for(bs * pl.n1 * (pl.n3 / ns_of(pl.n2))), st, p))) return rc;
    if ((rc = s3_launch_row(pl.n3, grid_for(bs * (pl.n1 * pl.n2 / ns_row(pl.n3))), st, p))) return rc;
    if ((rc = s3_launch_col<4>(pl.n2, grid_for(bs * pl.n1 * (pl.n3 / ns_of(pl.n2))), st, p))) return rc;
    if ((rc = s3_launch_col<5>(pl.n1, grid_for(bs * (pl.M / ns_of(pl.n1))), st, p))) return rc;
    const int64_t items_a = bs * ((pl.n2L + S3_TILE_COLS - 1) / S3_TILE_COLS);
    if ((rc = s3_launch_bins_a((int)imin<int64_t>(items_a, max_ctas > 0 ? max_ctas : 8 * imax(sm_count(), 1)), st, p))) return rc;
    // bins per thread and round: the fewest slot-rounds, then the fewest passes over the columns
    const int nres_max = (p.nkeep + pl.np - 1) / pl.np;
    int slots = 4, best = 1 << 30;
    for (int sl = 4; sl >= 2; sl--) {
        const int cost = ((nres_max + S3_HGROUPS * sl - 1) / (S3_HGROUPS * sl)) * sl;
        if (cost < best) { best = cost; slots = sl; }
    }
    const dim3 gb((unsigned)(bs * ((pl.np + 31) / 32)));
#define DP_S3_BINS_B(SL)                                                                        \
    {                                                                                           \
        auto kb = s3_bins_b_kernel<SL>;                                                         \
        DP_SET_SMEM_ONCE(kb, S3_BINS_B_SMEM);                                                   \
        DP_CUDA_OK(DP_LAUNCH_PDL(kb, gb, dim3(S3_THREADS), S3_BINS_B_SMEM, st, p));             \
    }
    if (slots == 4) DP_S3_BINS_B(4) else if (slots == 3) DP_S3_BINS_B(3) else DP_S3_BINS_B(2)
#undef DP_S3_BINS_B
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

}  // namespace dp
