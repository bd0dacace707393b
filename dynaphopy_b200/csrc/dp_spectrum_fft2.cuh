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
// (16 B per sample) and the kept bins (8 B each).
#pragma once

#include "dp_fft_reg.cuh"
#include "dp_tma.cuh"
#include "dp_tmem.cuh"

namespace dp {

#ifndef DP_S2_THREADS
#define DP_S2_THREADS 384
#endif
constexpr int S2_THREADS = DP_S2_THREADS;          // 12 warps x <= 168 registers (measured best of 8 / 12 / 16 warps)
constexpr int S2_WARPS = S2_THREADS / 32;
#ifndef DP_S2_PREEMPT_ITEMS
#define DP_S2_PREEMPT_ITEMS 1
#endif
constexpr int S2_PREEMPT_ITEMS = DP_S2_PREEMPT_ITEMS;   // items per CTA in the preemptible form (dp_power_fft_windows_preemptible)
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
    // bulk stores of passes 1-3: all slabs form ONE 2-D tensor of doubles (gridDim.x * slab_rows rows of 2*n2);
    // boxes of one warp tile (2*NS doubles wide): n1 rows (pass 1), the r_lo first / n1 - r_hi last rows (pass 3)
    int slab_rows, r_lo, r_hi;
    TensorMap2D tm_col, tm_lo, tm_hi;
    int park, park_cols;           // |X_a|^2 parked in tensor memory (park_cols columns allocated) instead of the slab
    // Preemptible form (dp_power_fft_windows_preemptible): S2_PREEMPT_ITEMS (series, window pair) items per CTA (one),
    // grid = items / that, so the block scheduler can place CTAs of higher-priority streams whenever a CTA ends (~0.15 ms)
    // instead of waiting for a persistent launch to finish.  A CTA then borrows one of `n_slabs` slabs for its lifetime: slab_busy
    // [dev, n_slabs ints, zero when the launch starts] is the free list (one CTA per SM is resident at a time -- the
    // kernel's shared memory fills an SM -- so n_slabs = number of SMs always leaves a free one).  slab_busy == nullptr:
    // persistent form, CTA b owns slab b.
    int* slab_busy;
    int n_slabs;
};

struct S2Smem {
    cplx* T1;       // [k1*R2 + j2] = w_n1^(j2 k1)      stage twiddles of the n1 sub-transform (DIT)
    cplx* T2;       // same for n2
    cplx* T2f;      // [c*R1 + k1]  = w_n2^(k1 c)       DIF order
    cplx* Ts;       // [k1*R2 + j2] = w_n1'^(j2 k1)     pass 4
    cplx* fine;     // [256] w_P^j
    cplx* coarse;   // [256] w_P^(256 j)
    cplx* E;        // S2_WARPS * SUBFFT_EWARP_MAX      per-warp exchange slab, reused as the warp's staging tile of
                    //                                  the bulk stores (pass 4: the column tile)
};
// Slab traffic of passes 1-3 (history): per-lane ld/st.global with L2 eviction hints, plain .cg accesses and a
// cp.async landing zone all measured the same or slower (B200, cfg 5-A: 129-138 ms) -- the passes were limited by
// the in-order load/store unit: every warp's 16 stores per tile (8 wavefronts of 64 bytes each in the column
// passes) drain at the SM's L2 write rate (~26 B/clk) and the next tile's loads queue up behind them.  Stores now
// leave through the bulk-copy engine (dp_tma.cuh): a tile is written to the warp's staging area in shared memory
// (conflict-free STS) and ONE elected lane issues a tensor-box / bulk store; the load/store unit sees loads only.
constexpr size_t S2_E_ELEMS = S2_WARPS * SUBFFT_EWARP_MAX > 2 * S2_NP_MAX * S2_CPITCH ? S2_WARPS * SUBFFT_EWARP_MAX
                                                                                  : 2 * S2_NP_MAX * S2_CPITCH;
constexpr size_t S2_SMEM_BYTES = (size_t)(6 * 256 + S2_E_ELEMS) * sizeof(cplx);

// carve the CTA's dynamic shared memory; done inside every pass (from the __shared__ symbol itself) so
// that the compiler sees shared-space pointers and emits LDS/STS instead of generic loads
__device__ __forceinline__ S2Smem s2_carve(cplx* smem) {
    S2Smem sm;
    sm.T1 = smem; sm.T2 = smem + 256; sm.T2f = smem + 512; sm.Ts = smem + 768;
    sm.fine = smem + 1024; sm.coarse = smem + 1280; sm.E = smem + 1536;
    return sm;
}
#define S2_SMEM_HERE() DP_DYNSMEM(cplx, s2_smem_base_); const S2Smem sm = s2_carve(s2_smem_base_)

__device__ __forceinline__ int64_t s2_time_offset(const S2Params& p, int t) {
    if (p.tb_len == 0) return (int64_t)t * p.stride_t;
    const int b = p.tb_shift >= 0 ? (t >> p.tb_shift) : t / p.tb_len;
    return (int64_t)b * p.tb_stride + (int64_t)(t - b * p.tb_len) * p.stride_t;
}

// streaming load of a window sample (evict-first); DP_S2_LD128 (experiment): with the 128-byte L2 prefetch size, so that
// the 64-byte piece of the neighbouring column tile (same 128-byte line, fetched by another warp a moment later) is an
// L2 hit -- measured neutral (4.83-4.85 ms per 1480 series either way), left off
__device__ __forceinline__ cplx s2_ld_window(const cplx* ptr) {
#if defined(DP_S2_LD128) && !defined(DP_EMUL)
    cplx v;
    asm volatile("ld.global.cs.L2::128B.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(ptr));
    return v;
#else
    return __ldcs(ptr);
#endif
}

__device__ __forceinline__ void s2_prefetch_l2(const void* ptr) {
#ifndef DP_EMUL
    asm volatile("prefetch.global.L2 [%0];\n" ::"l"(ptr));
#endif
}

// slab accesses of the load/store unit: L2 only (the slab is shared with the bulk-copy engine, never cached in L1)
__device__ __forceinline__ cplx s2_ld_slab(const cplx* ptr) {
#ifdef DP_EMUL
    return *ptr;
#else
    return __ldcg(ptr);
#endif
}
__device__ __forceinline__ void s2_st_slab(cplx* ptr, cplx v) {
#ifdef DP_EMUL
    *ptr = v;
#else
    __stcg(ptr, v);
#endif
}

// ---- tile scheduling of passes 1-3: a CTA-wide counter (monotonic over the whole kernel) hands out the tiles of a pass,
// so that any number of warps stays balanced (64 tiles over 12 warps).  The claim for the NEXT tile is issued at the top
// of a tile and read at its end: the shared-memory atomic's latency is never exposed.
struct S2Tiles {
    int* counter;
    int base;           // counter value at which the current pass starts (uniform over the CTA)
};
__device__ __forceinline__ int s2_claim_issue(const S2Tiles& ts, int lane) {
    return lane == 0 ? atomicAdd(ts.counter, 1) - ts.base : 0;
}
__device__ __forceinline__ int s2_claim_get(int raw) { return __shfl_sync(0xffffffffu, raw, 0); }
__device__ __forceinline__ void s2_tiles_done(S2Tiles& ts, int ntiles) { ts.base += ntiles + S2_WARPS; }   // one claim beyond the end per warp

// ---- staging-tile protocol of one warp -------------------------------------------------------------
// The staging tile shares the warp's exchange slab E: (1) before the first write to E of a tile, the bulk stores of the
// previous tile must have finished reading it: s2_stage_acquire; (2) after the last exchange read, a warp barrier, then
// the results are written to E in the tile layout; (3) s2_stage_publish orders those writes before the asynchronous
// proxy; the elected lanes then issue their stores and commit.  s2_stage_drain at the end of a pass: the global writes
// are complete (and ordered before the ordinary loads of the next pass, after the CTA barrier that follows).
__device__ __forceinline__ void s2_stage_acquire() {
    bulk_wait_read<0>();          // lanes that issued nothing return at once
    __syncwarp();
}
__device__ __forceinline__ void s2_stage_publish() {
    fence_proxy_async_smem();
    __syncwarp();
}
__device__ __forceinline__ void s2_stage_drain() {
    bulk_wait<0>();
    fence_proxy_async();
}
// row pitch (in elements of 16 / 8 bytes) of a staged ROW tile of NS rows: spreads the rows of a quarter-/half-warp
// over the shared-memory bank groups (same rule as SubCfg::ESEQ)
template <class C> struct S2RowPitch {
    static constexpr int NSQ = C::NS < 8 ? C::NS : 8;
    __device__ static __forceinline__ int cplx_pitch(int n2) { return n2 + 8 / NSQ; }
    __device__ static __forceinline__ int real_pitch(int n2) { return n2 + (16 / C::NS > 2 ? 16 / C::NS : 2); }
};

// ---- pass 1: window -> column FFTs -> A -----------------------------------------------------------
// LAYOUT 0: time axis contiguous (stride_t == 1, no blocks) -- plain pointer arithmetic in the unrolled loads;
// 1: contiguous power-of-two time blocks (the all-to-all's receive layout); 2: anything else.
template <class C, int LAYOUT>
__device__ __noinline__ void s2_pass1(const S2Params& p, const cplx* __restrict__ x, int t0, cplx* __restrict__ A,
                                      S2Tiles& ts, int warp, int lane) {
    S2_SMEM_HERE();
    const int n2 = p.pl.n2, L = p.pl.L;
    const int tb_shift = p.tb_shift;
    const int64_t tb_stride = p.tb_stride;
    const int seq = lane % C::NS, g = lane / C::NS;
    cplx* Ew = sm.E + (size_t)warp * SUBFFT_EWARP_MAX;       // the warp's exchange slab / staging tile
    cplx* E = Ew + seq * C::ESEQ;
    const int ntiles = n2 / C::NS;
    for (int tile = s2_claim_get(s2_claim_issue(ts, lane)); tile < ntiles;) {
        const int next_raw = s2_claim_issue(ts, lane);
        const int col = tile * C::NS + seq;
        cplx v[16];
#pragma unroll
        for (int i = 0; i < C::NB1; i++)
#pragma unroll
            for (int a = 0; a < C::R1; a++) {
                const int idx = (a * C::R2 + g + C::TG * i) * n2 + col;
                cplx val = make_double2(0.0, 0.0);
                if (idx < L) {
                    if (LAYOUT == 0) val = s2_ld_window(x + t0 + idx);
                    else if (LAYOUT == 1) {
                        const int t = t0 + idx, b = t >> tb_shift;
                        val = s2_ld_window(x + (int64_t)b * tb_stride + (t - (b << tb_shift)));
                    } else val = s2_ld_window(x + s2_time_offset(p, t0 + idx));
                }
                v[i * C::R1 + a] = val;
            }
        // (Requesting the strip of the warp's NEXT tile into L2 here -- 78 prefetch.global.L2 per tile, one tile of
        // look-ahead -- was measured slower as well: 4.92 vs 4.86 ms per 1480 series.)
        s2_stage_acquire();
        sub_dit<C>(v, E, sm.T1, g);
        __syncwarp();                 // every lane has read its exchange values: E becomes the staging tile [n1][NS]
        // four-step twiddle w_P^(col*kk), kk = k1 + R1*k2: geometric in k2
        const cplx ratio = tw2level(sm.fine, sm.coarse, col * C::R1);
#pragma unroll
        for (int i = 0; i < C::NB2; i++) {
            const int k1 = g + C::TG * i;
            cplx tw = tw2level(sm.fine, sm.coarse, col * k1);
#pragma unroll
            for (int k2 = 0; k2 < C::R2; k2++) {
                const cplx val = cmul(v[i * C::R2 + k2], tw);
                Ew[(k1 + C::R1 * k2) * C::NS + seq] = val;
                if (k2 + 1 < C::R2) tw = cmul(tw, ratio);
            }
        }
        s2_stage_publish();
        if (lane == 0) {
            tma_store_2d(&p.tm_col, 2 * tile * C::NS, (int)((A - p.scratch) / n2), Ew);        // first row of this slab
            bulk_commit();
        }
        tile = s2_claim_get(next_raw);
    }
    s2_stage_drain();
    s2_tiles_done(ts, ntiles);
}

// ---- pass 2: rows: FFT -> |.|^2 -> FFT (DIF) -> twiddle, in place ------------------------------------
// MODE 0: single window.  Paired windows (two windows of one series share the second half of the pipeline,
// see spectrum2_kernel): MODE 1 = first window, rows FFT -> |X|^2 parked as REAL values in Sr (digit order);
// MODE 2 = second window, rows FFT -> |X|^2, z = Sr + i |X|^2, FFT (DIF) -> twiddle -> A.
// PARK: |X_a|^2 waits in tensor memory (dp_tmem.cuh) instead of Sr: the thread that produced the 16 values of a tile in
// MODE 1 is the one that consumes them in MODE 2 (static tile -> warp assignment in those two modes; warp % 4 is the
// TMEM lane quarter, so S2_WARPS must be a multiple of 4; tile t uses columns 32 * (t / 4) of quarter t % 4).  The slab
// shrinks from 1.5 P to P elements -- 148 x 512 KB at cfg 5-A, which the 126 MB L2 holds.
template <class C, int MODE, bool PARK>
__device__ __noinline__ void s2_pass2(const S2Params& p, cplx* __restrict__ A, double* __restrict__ Sr, S2Tiles& ts,
                                      uint32_t tmem, int warp, int lane) {
    static_assert(!PARK || S2_WARPS % 4 == 0, "TMEM parking maps tile % 4 to the warp's lane quarter");
    S2_SMEM_HERE();
    const int n1 = p.pl.n1, n2 = p.pl.n2;
    // rows are contiguous along the element index: group lane fastest, a quarter-warp reads one 128-byte run of its row
    // (the column passes map the sequence index fastest for the same reason)
#ifdef DP_S2_ROW_SEQFAST
    const int seq = lane % C::NS, g = lane / C::NS;
#else
    const int seq = lane / C::TG, g = lane % C::TG;
#endif
    cplx* Ew = sm.E + (size_t)warp * SUBFFT_EWARP_MAX;
    cplx* E = Ew + seq * C::ESEQ;
    const int ntiles = n1 / C::NS;
    const int rp = S2RowPitch<C>::cplx_pitch(n2), rpd = S2RowPitch<C>::real_pitch(n2);
    constexpr bool STATIC = PARK && MODE != 0;
    for (int tile = STATIC ? warp : s2_claim_get(s2_claim_issue(ts, lane)); tile < ntiles;) {
        const int next_raw = STATIC ? 0 : s2_claim_issue(ts, lane);
        const int row = tile * C::NS + seq;
        cplx* __restrict__ Ar = A + (size_t)row * n2;
        cplx v[16];
#pragma unroll
        for (int i = 0; i < C::NB1; i++)
#pragma unroll
            for (int a = 0; a < C::R1; a++) v[i * C::R1 + a] = s2_ld_slab(Ar + a * C::R2 + g + C::TG * i);
        double sa[16];
        uint32_t parked[32];
        double* __restrict__ Srow = Sr + (size_t)row * n2;
        if (MODE == 2) {                       // |X_a|^2 of the partner window, requested before the butterflies
            if (PARK) tmem_fetch16(tmem, warp, lane, 32 * (tile >> 2), parked);
            else {
#pragma unroll
                for (int i = 0; i < C::NB2; i++)
#pragma unroll
                    for (int k2 = 0; k2 < C::R2; k2++) sa[i * C::R2 + k2] = __ldcg(Srow + g + C::TG * i + C::R1 * k2);
            }
        }
        s2_stage_acquire();
        sub_dit<C>(v, E, sm.T2, g);
        if constexpr (MODE == 1 && PARK) {
            double pw[16];
#pragma unroll
            for (int r = 0; r < 16; r++) pw[r] = v[r].x * v[r].x + v[r].y * v[r].y;
            tmem_park16(tmem, warp, lane, 32 * (tile >> 2), pw);
        } else if constexpr (MODE == 1) {
            __syncwarp();             // exchange reads done: E becomes the staging tile, NS rows of n2 doubles
            double* Sw = reinterpret_cast<double*>(Ew) + (size_t)seq * rpd;
#pragma unroll
            for (int i = 0; i < C::NB2; i++)
#pragma unroll
                for (int k2 = 0; k2 < C::R2; k2++) {
                    const cplx x = v[i * C::R2 + k2];
                    Sw[g + C::TG * i + C::R1 * k2] = x.x * x.x + x.y * x.y;
                }
            s2_stage_publish();
            if (lane < C::NS) {       // lane s ships row s of the tile: n2 contiguous doubles
                bulk_s2g(Sr + (size_t)(tile * C::NS + lane) * n2, reinterpret_cast<double*>(Ew) + (size_t)lane * rpd,
                         (unsigned)(n2 * sizeof(double)));
                bulk_commit();
            }
        } else {
            if (MODE == 2 && PARK) {
                tmem_fetch_wait();
#pragma unroll
                for (int r = 0; r < 16; r++) sa[r] = tmem_value(parked, r);
            }
#pragma unroll
            for (int r = 0; r < C::NB2 * C::R2; r++) {
                const double sb = v[r].x * v[r].x + v[r].y * v[r].y;
                v[r] = MODE == 2 ? make_double2(sa[r], sb) : make_double2(sb, 0.0);
            }
            __syncwarp();
            sub_dif<C>(v, E, sm.T2f, g);
            __syncwarp();             // exchange reads done: E becomes the staging tile, NS rows of n2 complex values
            cplx* Ow = Ew + (size_t)seq * rp;
            // four-step twiddle w_P^(row*k), k = d*R2 + c: geometric in d
            const cplx ratio = tw2level(sm.fine, sm.coarse, row * C::R2);
#pragma unroll
            for (int i = 0; i < C::NB1; i++) {
                const int c = g + C::TG * i;
                cplx tw = tw2level(sm.fine, sm.coarse, row * c);
#pragma unroll
                for (int d = 0; d < C::R1; d++) {
                    const cplx val = cmul(v[i * C::R1 + d], tw);
                    Ow[d * C::R2 + c] = val;
                    if (d + 1 < C::R1) tw = cmul(tw, ratio);
                }
            }
            s2_stage_publish();
            if (lane < C::NS) {
                bulk_s2g(A + (size_t)(tile * C::NS + lane) * n2, Ew + (size_t)lane * rp, (unsigned)(n2 * sizeof(cplx)));
                bulk_commit();
            }
        }
        tile = STATIC ? tile + S2_WARPS : s2_claim_get(next_raw);
    }
    if (MODE == 1 && PARK) tmem_park_wait();
    s2_stage_drain();
    if (!STATIC) s2_tiles_done(ts, ntiles);
}

// ---- pass 3: column FFTs -> conj -> lags 0..half of the autocorrelation, in place ------------------------
// r[kappa] lands at A[kappa]: a warp overwrites only rows of the columns it has just read.
// PAIRED: A holds Y = FFT_P(|X_a|^2 + i |X_b|^2); rho[kappa] = r_a[kappa] + i r_b[kappa] = Y[-kappa] / P, so the
// lags -half..half are Y at indices 0..half and P-half..P-1 (no conjugation, no Hermitian shortcut).
// Only the rows that hold needed lags are written: rows [0, r_lo) and, PAIRED, rows [r_hi, n1) (whole rows of the
// warp's own columns: the few extra elements of the boundary rows are never read).
template <class C, bool PAIRED>
__device__ __noinline__ void s2_pass3(const S2Params& p, cplx* __restrict__ A, S2Tiles& ts, int warp, int lane) {
    S2_SMEM_HERE();
    const int n2 = p.pl.n2, half = p.pl.half;
    const double sc = p.acf_scale;
    const int seq = lane % C::NS, g = lane / C::NS;
    cplx* Ew = sm.E + (size_t)warp * SUBFFT_EWARP_MAX;
    cplx* E = Ew + seq * C::ESEQ;
    const int ntiles = n2 / C::NS;
    const int r_lo = p.r_lo, r_hi = p.r_hi;
    for (int tile = s2_claim_get(s2_claim_issue(ts, lane)); tile < ntiles;) {
        const int next_raw = s2_claim_issue(ts, lane);
        const int col = tile * C::NS + seq;
        cplx v[16];
#pragma unroll
        for (int i = 0; i < C::NB1; i++)
#pragma unroll
            for (int a = 0; a < C::R1; a++) v[i * C::R1 + a] = s2_ld_slab(A + (size_t)(a * C::R2 + g + C::TG * i) * n2 + col);
        s2_stage_acquire();
        sub_dit<C>(v, E, sm.T1, g);
        __syncwarp();          // every lane of the warp has its column in registers before rows are overwritten
#pragma unroll
        for (int i = 0; i < C::NB2; i++)
#pragma unroll
            for (int k2 = 0; k2 < C::R2; k2++) {
                const int row = g + C::TG * i + C::R1 * k2;
                if (row < r_lo || (PAIRED && row >= r_hi))
                    Ew[row * C::NS + seq] = make_double2(v[i * C::R2 + k2].x * sc, PAIRED ? v[i * C::R2 + k2].y * sc : -v[i * C::R2 + k2].y * sc);
            }
        s2_stage_publish();
        if (lane == 0) {
            tma_store_2d(&p.tm_lo, 2 * tile * C::NS, (int)((A - p.scratch) / n2), Ew);
            if (PAIRED) tma_store_2d(&p.tm_hi, 2 * tile * C::NS, (int)((A - p.scratch) / n2) + r_hi, Ew + (size_t)r_hi * C::NS);
            bulk_commit();
        }
        tile = s2_claim_get(next_raw);
    }
    s2_stage_drain();
    s2_tiles_done(ts, ntiles);
    (void)half;
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
                    v[a] = s2_ld_slab(R + (kap > 0 ? p.pl.P - kap : -kap));      // Y[-kappa]
                } else {
                    v[a] = s2_ld_slab(R + (kap >= 0 ? kap : -kap));
                    if (kap < 0) v[a].y = -v[a].y;                 // negative lag: conj(r[-kappa])
                }
            }
        } else {
#pragma unroll
            for (int a = 0; a < R1; a++) v[a] = make_double2(0.0, 0.0);
        }
    };
    // bins owned by this thread: residue r = tid % NP, and of that residue's kept bins e = r + NP*j every NSPLIT-th one
    // (j = h + NSPLIT*b, h = tid / NP) -- all threads of the CTA take part in the Horner phase, not only the first NP
    constexpr int NSPLIT = S2_THREADS / NP < 1 ? 1 : (S2_THREADS / NP > S2_BINSLOTS ? S2_BINSLOTS : S2_THREADS / NP);
    constexpr int SLOTS = (S2_BINSLOTS + NSPLIT - 1) / NSPLIT;
    const int r = tid % NP, h = tid / NP;
    int k_lo = p.lo - half;
    if (k_lo < 0) k_lo += L;
    const int nres = h < NSPLIT ? (p.nkeep - r + NP - 1) / NP : 0;       // kept bins of residue r: e = r + NP*j, j < nres
    const int nbins = nres > h ? (nres - h + NSPLIT - 1) / NSPLIT : 0;    // ... of which this thread owns j = h + NSPLIT*b
    cplx acc[SLOTS], z[SLOTS];
#pragma unroll
    for (int b = 0; b < SLOTS; b++) {
        int k = k_lo + r + NP * (h + NSPLIT * b);
        k %= L;
        z[b] = b < nbins ? __ldg(p.g_twL + k) : make_double2(0.0, 0.0);
        acc[b] = make_double2(0.0, 0.0);
    }
    const cplx* crow = Cb + ((k_lo + r) % NP) * S2_CPITCH;
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
                for (int b = 0; b < SLOTS; b++) {
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
        for (int b = 0; b < SLOTS; b++)
            if (b < nbins) spec[r + NP * (h + NSPLIT * b)] = __dmul_rn(hypot(acc[b].x, acc[b].y), p.dt);
    } else {
        // Im r_a[-L/2], Im r_b[-L/2] from rho[-half] = Y[half], rho[+half] = Y[P - half] (even L only)
        double Ia = 0.0, Ib = 0.0;
        if ((L & 1) == 0) {
            const cplx rm = s2_ld_slab(R + half), rp = s2_ld_slab(R + p.pl.P - half);
            Ia = 0.5 * (rm.y - rp.y);
            Ib = 0.5 * (rp.x - rm.x);
        }
        double* __restrict__ spec_b = spec + p.nkeep;             // the pair's second window
#pragma unroll
        for (int b = 0; b < SLOTS; b++) {
            if (b < nbins) {
                const int e = r + NP * (h + NSPLIT * b);
                const int k = (k_lo + e) % L;
                const cplx ph = __ldg(p.g_twL + (int)(((int64_t)half * k) % L));      // w_L^(half k)
                const cplx wt = cmulc(acc[b], ph);                                     // remove the phase
                const double sgn = (k & 1) ? -1.0 : 1.0;
                spec[e] = __dmul_rn(hypot(wt.x + sgn * Ib, Ia), p.dt);
                spec_b[e] = __dmul_rn(hypot(wt.y - sgn * Ia, Ib), p.dt);
            }
        }
    }
}

#define DP_S2_PASS1_L(xptr, t0v, LAY)                                               \
    switch (p.pl.n1) {                                                              \
        case 16: s2_pass1<C16, LAY>(p, xptr, t0v, A, ts, warp, lane); break;            \
        case 32: s2_pass1<C32, LAY>(p, xptr, t0v, A, ts, warp, lane); break;            \
        case 64: s2_pass1<C64, LAY>(p, xptr, t0v, A, ts, warp, lane); break;            \
        case 128: s2_pass1<C128, LAY>(p, xptr, t0v, A, ts, warp, lane); break;          \
        default: s2_pass1<C256, LAY>(p, xptr, t0v, A, ts, warp, lane); break;           \
    }
#define DP_S2_PASS1(xptr, t0v)                                                      \
    if (layout == 0) { DP_S2_PASS1_L(xptr, t0v, 0) }                                \
    else if (layout == 1) { DP_S2_PASS1_L(xptr, t0v, 1) }                           \
    else { DP_S2_PASS1_L(xptr, t0v, 2) }
#define DP_S2_PASS2(MODE)                                                                   \
    if (p.park) {                                                                           \
        if (p.pl.n2 == 64) s2_pass2<C64, MODE, true>(p, A, Sr, ts, tmem_base, warp, lane);  \
        else s2_pass2<C256, MODE, true>(p, A, Sr, ts, tmem_base, warp, lane);               \
    } else {                                                                                \
        if (p.pl.n2 == 64) s2_pass2<C64, MODE, false>(p, A, Sr, ts, 0u, warp, lane);        \
        else s2_pass2<C256, MODE, false>(p, A, Sr, ts, 0u, warp, lane);                     \
    }
#define DP_S2_PASS3(PAIRED)                                                         \
    switch (p.pl.n1) {                                                              \
        case 16: s2_pass3<C16, PAIRED>(p, A, ts, warp, lane); break;                    \
        case 32: s2_pass3<C32, PAIRED>(p, A, ts, warp, lane); break;                    \
        case 64: s2_pass3<C64, PAIRED>(p, A, ts, warp, lane); break;                    \
        case 128: s2_pass3<C128, PAIRED>(p, A, ts, warp, lane); break;                  \
        default: s2_pass3<C256, PAIRED>(p, A, ts, warp, lane); break;                   \
    }

// Items: (series, pair of consecutive windows).  Two windows of one series share everything after the
// forward transforms: |X_a|^2 + i |X_b|^2 goes through ONE second transform, one column pass and one bin
// evaluation (the two autocorrelations are Hermitian, so their transforms separate exactly) -- 3/4 of the
// row work and 1/2 of passes 3-4.  An odd last window runs the single-window passes.
__global__ void __launch_bounds__(S2_THREADS, 1)
spectrum2_kernel(const __grid_constant__ S2Params p) {
    S2_SMEM_HERE();
    __shared__ int s2_tile_counter;
    __shared__ uint32_t s2_tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = tid; k < 6 * 256; k += S2_THREADS) sm.T1[k] = p.g_tables[k];
    if (tid == 0) s2_tile_counter = 0;
    if (p.park && warp == 0) tmem_alloc(&s2_tmem_slot, (uint32_t)p.park_cols);
    tmem_alloc_done();             // includes the CTA barrier
    const uint32_t tmem_base = p.park ? s2_tmem_slot : 0u;
    S2Tiles ts;
    ts.counter = &s2_tile_counter;
    ts.base = 0;
    __shared__ int s2_slab;
    if (p.slab_busy) {
        if (tid == 0) {
            // start the search at the SM's own number: an SM keeps re-using the same slab (and so the same L2 lines),
            // as a persistent CTA does; with n_slabs = number of SMs the first try succeeds
            unsigned smid = blockIdx.x;
#ifndef DP_EMUL
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
#endif
            int slot = (int)(smid % (unsigned)p.n_slabs);
            while (atomicCAS(p.slab_busy + slot, 0, 1) != 0) slot = slot + 1 == p.n_slabs ? 0 : slot + 1;
            s2_slab = slot;
        }
        __syncthreads();
    }
    const int slab = p.slab_busy ? s2_slab : (int)blockIdx.x;
    cplx* A = p.scratch + (size_t)slab * p.slab_elems;
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
        // (No explicit L2 prefetch of the coming windows.  Per-line prefetch.global.L2 and bulk prefetches, with normal
        // and with evict-first priority, one or two passes ahead, were all measured to HURT -- 5.66 -> 5.39 ms per 1480
        // series without them at 114 MB of slabs, 4.83 -> 5.11 ms with them at 76 MB: the slabs, dirty all the time and
        // partly mirrored in the far L2 partition, fill the L2 and every prefetched window evicts lines of some CTA.)
        if (!paired) {
            if (!(p.skip_mask & 1)) { DP_S2_PASS1(x, p.starts[w0]) }
            __syncthreads();
            if (!(p.skip_mask & 2)) { DP_S2_PASS2(0) }
            __syncthreads();
            if (!(p.skip_mask & 4)) { DP_S2_PASS3(false) }
            __syncthreads();
            if (!(p.skip_mask & 8)) switch (p.pl.m1 * 32 + p.pl.m2) {
#define DP_S2_CASE(a, b) case (a) * 32 + (b): s2_pass4<SubCfg<a, b, 16>, false>(p, A, spec, tid); break;
                DP_S2_MIXED_MENU(DP_S2_CASE)
#undef DP_S2_CASE
                default: break;
            }
        } else {
            if (!(p.skip_mask & 1)) { DP_S2_PASS1(x, p.starts[w0]) }
            __syncthreads();
            if (!(p.skip_mask & 2)) { DP_S2_PASS2(1) }
            __syncthreads();
            if (!(p.skip_mask & 1)) { DP_S2_PASS1(x, p.starts[w0 + 1]) }
            __syncthreads();
            if (!(p.skip_mask & 2)) { DP_S2_PASS2(2) }
            __syncthreads();
            if (!(p.skip_mask & 4)) { DP_S2_PASS3(true) }
            __syncthreads();
            if (!(p.skip_mask & 8)) switch (p.pl.m1 * 32 + p.pl.m2) {
#define DP_S2_CASE(a, b) case (a) * 32 + (b): s2_pass4<SubCfg<a, b, 16>, true>(p, A, spec, tid); break;
                DP_S2_MIXED_MENU(DP_S2_CASE)
#undef DP_S2_CASE
                default: break;
            }
        }
        __syncthreads();
    }
    if (p.park && warp == 0) tmem_free(tmem_base, (uint32_t)p.park_cols);      // every TMEM access precedes the barrier above
    // the slab goes back to the free list: every access of this CTA to it (the bulk stores were drained at the end of
    // their passes) precedes the barrier that closes the item
    if (p.slab_busy && tid == 0) { __threadfence(); atomicExch(p.slab_busy + slab, 0); }
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

// |X_a|^2 of a window pair (P doubles) fits the SM's tensor memory (128 lanes x 512 columns x 4 bytes)
static inline bool s2_parks_in_tmem(const S2Plan& pl) {
#ifdef DP_S2_NO_PARK
    (void)pl;
    return false;
#else
    return S2_WARPS % 4 == 0 && (size_t)pl.P * sizeof(double) <= 256 * 1024;
#endif
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
