// A2, commensurate form: wave-vector projection of the whole q list as a batched 3-D DFT.
// Reference semantics: dynaphopy/projection.py:4-40 with the supercell atom order of
// dynaphopy/atoms.py:139-156 (i = u*Nc + x + Nx*(y + Ny*z)).
//
// For q commensurate with the (diagonal) supercell, exp(-i q.R_cell) = exp(-2 pi i m.n/dims), so
//   vc[t,q,p,k] = sum_{u in p} twiddle[q,u] * DFT3_{cells}( v[t, u, cell, k] )[ m(q) ]
// and one frame is read from HBM exactly once for all Q wave vectors (24 B per atom-frame in,
// 16*3*n_prim*Q B per frame out) instead of once per q.
//
// Work unit = (frame, job).  A job packs TWO real sequences into one complex 3-D transform
// c = a + i b (e.g. the x and y components of one unit-cell atom, or the z components of two
// atoms) and separates them in the epilogue with A[m] = (C[m] + conj C[-m]) / 2,
// B[m] = (C[m] - conj C[-m]) / 2i.  The transform runs in shared memory (Nc*16 B, row pitch
// padded to an odd number of elements): three axis passes, each thread owning whole pencils in
// registers and using compile-time unrolled mixed-radix codelets (radix 4/2/3/5, generic odd
// primes) whose twiddles are immediate constants.  Persistent grid, one CTA per work unit,
// sized as a multiple of the SM count.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "dp_common.cuh"
#include "dp_fft.cuh"

#include "dp_pencil.cuh"

namespace dp {

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// pencil pc: base = (pc % A)*sa + (pc / A)*sb, elements at base + i*st
// (measured: the in-place 16-point codelet of dp_fft_reg.cuh in this pass removes the spills but is slower: 24.7 vs 24.1 ms)
// The last decimation-in-time stage is done here rather than inside PencilDft<N>::run: the R sub-transforms read
// their inputs straight from shared memory and the final butterflies store straight back, so only the N intermediate
// values are live (2N doubles instead of the 4N-6N of the out-of-place codelet, which spilled for N >= 16 under the
// 96-register budget of two CTAs per SM).
template <int N>
__device__ __noinline__ void axis_pass_n(int npencil, int A, int sa, int sb, int st, int tid, int nthreads) {
    DP_DYNSMEM(cplx, W);             // the transform is the CTA's dynamic shared memory: keep the address space visible
    constexpr int R = PencilDft<N>::R, M = PencilDft<N>::M;
    for (int pc = tid; pc < npencil; pc += nthreads) {
        const int base = (pc % A) * sa + (pc / A) * sb;
        if constexpr (M == 1) {
            cplx t[N];
#pragma unroll
            for (int i = 0; i < N; i++) t[i] = W[base + i * st];
            pencil_bfly<N>(t);
#pragma unroll
            for (int i = 0; i < N; i++) W[base + i * st] = t[i];
        } else {
            cplx sub[R][M];
#pragma unroll
            for (int r = 0; r < R; r++) {
                cplx g[M];
#pragma unroll
                for (int mm = 0; mm < M; mm++) g[mm] = W[base + (mm * R + r) * st];
                PencilDft<M>::run(g, sub[r]);
            }
#pragma unroll
            for (int k = 0; k < M; k++) {
                cplx t[R];
#pragma unroll
                for (int r = 0; r < R; r++) t[r] = (r == 0 || k == 0) ? sub[r][k] : cmul(sub[r][k], root_const(N, r * k));
                pencil_bfly<R>(t);
#pragma unroll
                for (int u = 0; u < R; u++) W[base + (k + M * u) * st] = t[u];
            }
        }
    }
}

#define DP_AXIS_CASE(n) case n: axis_pass_n<n>(npencil, A, sa, sb, st, tid, nthreads); break;
__device__ __forceinline__ void axis_pass(int n, int npencil, int A, int sa, int sb, int st, int tid,
                                          int nthreads) {
    switch (n) {
        DP_AXIS_CASE(2) DP_AXIS_CASE(3) DP_AXIS_CASE(4) DP_AXIS_CASE(5) DP_AXIS_CASE(6) DP_AXIS_CASE(7)
        DP_AXIS_CASE(8) DP_AXIS_CASE(9) DP_AXIS_CASE(10) DP_AXIS_CASE(11) DP_AXIS_CASE(12) DP_AXIS_CASE(13)
        DP_AXIS_CASE(14) DP_AXIS_CASE(15) DP_AXIS_CASE(16) DP_AXIS_CASE(17) DP_AXIS_CASE(18) DP_AXIS_CASE(19)
        DP_AXIS_CASE(20) DP_AXIS_CASE(21) DP_AXIS_CASE(22) DP_AXIS_CASE(23) DP_AXIS_CASE(24) DP_AXIS_CASE(25)
        DP_AXIS_CASE(26) DP_AXIS_CASE(27) DP_AXIS_CASE(28) DP_AXIS_CASE(29) DP_AXIS_CASE(30) DP_AXIS_CASE(31)
        DP_AXIS_CASE(32)
        default: break;                      // n == 1: nothing to do
    }
}
#undef DP_AXIS_CASE

// ---------------- kernel ------------------------------------------------------------------------
constexpr int PF_MAX_BLOCKS = 16;
struct PfJob { int pA, kA, pB, kB, term0, nterms; };      // pB < 0: single sequence
struct PfTerm { int uA, uB; };                              // uB < 0: none

struct PfParams {
    const double* v;
    int Nx, Ny, Nz, Px, Nc, N, n_unit, n_prim, Q, njobs, nterms, npad;
    int64_t T;
    const PfJob* jobs;
    const PfTerm* terms;
    const int2* qoff;        // (offset of bin m(q), offset of bin -m(q)) in the padded layout
    const cplx* tw;          // [Q][n_unit], or nullptr: no twiddle (raw lattice DFT of every unit atom)
    cplx* out;               // [T][Q][n_prim][3]
    int pair_major;          // output layout: 0 = out[t][q][3*n_prim], 1 = out[t][slot pair][q][2]
    // pair-major output split into blocks of q_block wave vectors with their own base pointer (one per destination rank
    // of the multi-GPU exchange): blocks[q / q_block][t][slot pair][q % q_block][2]; n_blocks == 0: one array (out)
    int q_block, n_blocks;
    cplx* blocks[PF_MAX_BLOCKS];
    int skip;                // development aid (DP_PF_SKIP): 1 = no loads, 2 = no transforms, 4 = no epilogue
};

__global__ void __launch_bounds__(256)
qoffset_kernel(const int* __restrict__ qbin, int Q, int Nx, int Ny, int Nz, int Px, int2* __restrict__ qoff,
               int* __restrict__ /*unused*/) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < Q; q += gridDim.x * blockDim.x) {
        int mx = qbin[q * 3], my = qbin[q * 3 + 1], mz = qbin[q * 3 + 2];
        const bool ok = !(mx < 0 || mx >= Nx || my < 0 || my >= Ny || mz < 0 || mz >= Nz);
        if (!ok) mx = my = mz = 0;
        int nx = (Nx - mx) % Nx, ny = (Ny - my) % Ny, nz = (Nz - mz) % Nz;
        // a bin outside the grid: both offsets point at the poison element W[npad] (NaN), see project_fft_kernel
        qoff[q] = ok ? make_int2((mz * Ny + my) * Px + mx, (nz * Ny + ny) * Px + nx) : make_int2(-1, -1);
    }
}

// ---- phases of one (frame, transform) unit; each is its own function so that the register allocation of
// the wide pencil codelets does not leak into the streaming phases (and vice versa) -----------------------
#ifndef DP_PF_LD
#define DP_PF_LD 4
#endif
constexpr int PF_LD = DP_PF_LD;

struct PfUnit {
    const double* srcA; const double* srcB;      // real sequences (stride 3 doubles), or nullptr
    cplx* outA; cplx* outB;                      // first output element of slot A / B for q = 0
    const cplx* twA; const cplx* twB;            // tw + unit atom
    int64_t qstride;
    bool hasA, hasB;
};

// (the phases take their operands as scalars: a struct passed by reference to a __noinline__ function lives on the
// local-memory stack and every field costs a local load at the top of the phase)
__device__ __noinline__ void pf_load(const double* __restrict__ srcA, const double* __restrict__ srcB, int Nc, int Nx,
                                     int Px, int tid, int nthreads) {
    DP_DYNSMEM(cplx, W);
    // two real sequences -> one complex array.  PF_LD cells (2 * PF_LD independent 8-byte loads) are in flight per
    // thread (measured on B200, loads alone, T = 4096: 4 cells 0.376 ms, 10 cells 0.418 ms, 20 cells 0.501 ms)
    for (int c0 = tid; c0 < Nc; c0 += PF_LD * nthreads) {
        double a[PF_LD], b[PF_LD];
#pragma unroll
        for (int i = 0; i < PF_LD; i++) {
            const int c = c0 + i * nthreads;
            a[i] = (srcA && c < Nc) ? __ldg(srcA + (int64_t)c * 3) : 0.0;
            b[i] = (srcB && c < Nc) ? __ldg(srcB + (int64_t)c * 3) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < PF_LD; i++) {
            const int c = c0 + i * nthreads;
            if (c < Nc) W[(c / Nx) * Px + (c % Nx)] = make_double2(a[i], b[i]);
        }
    }
}

// epilogue: separate the two real sequences (A = (C + conj D)/2, B = (C - conj D)/(2i) with C, D the bins
// m(q), -m(q)), twiddle, store.  ACC: add to what earlier unit atoms of the same type already stored.
template <bool ACC, bool TW>
__device__ __noinline__ void pf_epilogue(cplx* __restrict__ outA, cplx* __restrict__ outB, const cplx* __restrict__ twA,
                                         const cplx* __restrict__ twB, int qstride_i, int flags, int npad, int Q,
                                         int n_unit, int tid, int nthreads, cplx* const* blk, int q_block,
                                         int64_t row_off) {
    DP_DYNSMEM(cplx, W);
    const unsigned* __restrict__ qo = reinterpret_cast<const unsigned*>(W + npad + 1);   // packed bin offsets (m, -m)
    const bool hasA = flags & 1, hasB = flags & 2;
    const int64_t qstride = qstride_i;
    // blk != nullptr (pair-major only): row_off is the offset of this (frame, slot pair) row inside a block, in
    // elements; wave vector q lives in block q / q_block at column q % q_block (outA / outB are not used for stores)
    if (!ACC && hasA && hasB && qstride == 2 && outB == outA + 1 && (nthreads & 31) == 0) {
        // Pair-major layout: (A, B) of 32 consecutive q are 1 KB contiguous.  A lane computes ONE of the two
        // sequences (even lanes A, odd lanes B) of wave vector 16h + lane/2 of its group, so consecutive lanes hold
        // consecutive 16-byte values: every warp store is one fully coalesced 512-byte run, no shuffles, no
        // divergence; the pair of lanes that shares a wave vector reads the same two bins (shared-memory broadcast).
        // Two groups (four outputs) are in flight per thread.
        const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
        const bool par = lane & 1;
        const int sub = lane >> 1;
        const int ngroups = (Q + 31) >> 5;
        const cplx* __restrict__ tw = par ? twB : twA;
        for (int g0 = warp; g0 < ngroups; g0 += 2 * nwarps) {
            int q[4];
            unsigned o[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                q[j] = (g0 + (j >> 1) * nwarps) * 32 + (j & 1) * 16 + sub;
                o[j] = qo[imin(q[j], Q - 1)];
            }
            cplx C[4], D[4], w[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                C[j] = W[o[j] & 0xffffu];
                D[j] = W[o[j] >> 16];
                if (TW) w[j] = __ldg(tw + (int64_t)imin(q[j], Q - 1) * n_unit);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                // A = (0.5 (C.x + D.x), 0.5 (C.y - D.y)),  B = (0.5 (C.y + D.y), 0.5 (D.x - C.x))
                const double a1 = par ? C[j].y : C[j].x, b1 = par ? D[j].y : D[j].x;
                const double a2 = par ? D[j].x : C[j].y, b2 = par ? C[j].x : D[j].y;
                cplx r = make_double2(0.5 * (a1 + b1), 0.5 * (a2 - b2));
                if (TW) r = cmul(r, w[j]);
                if (q[j] < Q) {
                    if (blk) {
                        const int b = q[j] / q_block;
                        __stcs(blk[b] + row_off + (int64_t)(q[j] - b * q_block) * 2 + (lane & 1), r);
                    } else __stcs(outA + (int64_t)(q[j] - sub) * 2 + lane, r);
                }
            }
        }
        return;
    }
    for (int q0 = tid; q0 < Q; q0 += 4 * nthreads) {
        cplx wa[4], wb[4];
        if (TW) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int q = imin(q0 + i * nthreads, Q - 1);
                wa[i] = __ldg(twA + (int64_t)q * n_unit);
                wb[i] = __ldg(twB + (int64_t)q * n_unit);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int q = q0 + i * nthreads;
            if (q < Q) {
                const unsigned o = qo[q];
                const cplx C = W[o & 0xffffu], D = W[o >> 16];
                cplx ra = make_double2(0.5 * (C.x + D.x), 0.5 * (C.y - D.y));
                cplx rb = make_double2(0.5 * (C.y + D.y), 0.5 * (D.x - C.x));
                if (TW) { ra = cmul(ra, wa[i]); rb = cmul(rb, wb[i]); }
                if (ACC) {
                    if (hasA) ra = cadd(outA[q * qstride], ra);
                    if (hasB) rb = cadd(outB[q * qstride], rb);
                }
                // paired slots are adjacent: the two 16-byte stores of a thread fill one 32-byte sector
                if (blk) {
                    const int b = q / q_block;
                    cplx* d = blk[b] + row_off + (int64_t)(q - b * q_block) * 2;
                    if (hasA) __stcs(d, ra);
                    if (hasB) __stcs(d + 1, rb);
                } else {
                    if (hasA) __stcs(outA + q * qstride, ra);
                    if (hasB) __stcs(outB + q * qstride, rb);
                }
            }
        }
    }
}

// MAXT / MINB: launch bounds (register budget): <=320 threads with 2 CTAs per SM, or <=512 with 1
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
project_fft_kernel(const __grid_constant__ PfParams p) {
    DP_DYNSMEM(cplx, W);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int Nx = p.Nx, Ny = p.Ny, Nz = p.Nz, Px = p.Px, Nc = p.Nc, Q = p.Q, n_unit = p.n_unit;
    // shared memory: transform [npad] | poison element (NaN: the "bin" of a q outside the grid) | packed bin
    // offsets [Q] | jobs | terms  (staged once per CTA)
    unsigned* qo = reinterpret_cast<unsigned*>(W + p.npad + 1);
    if (tid == 0) W[p.npad] = make_double2(nan(""), nan(""));
    PfJob* sjobs = reinterpret_cast<PfJob*>(qo + ((Q + 3) & ~3));
    PfTerm* sterms = reinterpret_cast<PfTerm*>(sjobs + p.njobs);
    cplx** sblk = reinterpret_cast<cplx**>(sterms + ((p.nterms + 1) & ~1));      // 16-byte aligned
    if (tid < p.n_blocks) sblk[tid] = p.blocks[tid];
    for (int q = tid; q < Q; q += nthreads) {
        const int2 o = p.qoff[q];
        qo[q] = o.x < 0 ? ((unsigned)p.npad | ((unsigned)p.npad << 16)) : ((unsigned)o.x | ((unsigned)o.y << 16));
    }
    for (int i = tid; i < p.njobs; i += nthreads) sjobs[i] = p.jobs[i];
    for (int i = tid; i < p.nterms; i += nthreads) sterms[i] = p.terms[i];
    __syncthreads();
    const int64_t nunits = p.T * p.njobs;
    for (int64_t unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
        const int64_t t = unit / p.njobs;
        const PfJob job = sjobs[(int)(unit % p.njobs)];
        const double* __restrict__ vt = p.v + t * (int64_t)p.N * 3;
        PfUnit u;
        int64_t row_off = 0;
        if (p.pair_major) {       // jobs are exactly the slot pairs (2j, 2j+1): [t][j][q][2], contiguous per job
            const int j = (job.pA * 3 + job.kA) >> 1;
            u.outA = p.n_blocks > 0 ? p.blocks[0] : p.out + ((t * (3 * p.n_prim / 2) + j) * Q) * 2;
            row_off = ((t * (3 * p.n_prim / 2) + j) * p.q_block) * 2;
            u.outB = job.pB >= 0 ? u.outA + 1 : nullptr;
            u.qstride = 2;
        } else {
            u.outA = p.out + ((t * Q) * p.n_prim + job.pA) * 3 + job.kA;
            u.outB = job.pB >= 0 ? p.out + ((t * Q) * p.n_prim + job.pB) * 3 + job.kB : nullptr;
            u.qstride = (int64_t)p.n_prim * 3;
        }
        for (int term = 0; term < job.nterms; term++) {
            const PfTerm tm = sterms[job.term0 + term];
            u.srcA = tm.uA >= 0 ? vt + (int64_t)tm.uA * Nc * 3 + job.kA : nullptr;
            u.srcB = (tm.uB >= 0 && job.pB >= 0) ? vt + (int64_t)tm.uB * Nc * 3 + job.kB : nullptr;
            u.hasA = tm.uA >= 0;
            u.hasB = u.srcB != nullptr;
            u.twA = p.tw + (u.hasA ? tm.uA : 0);
            u.twB = p.tw + (u.hasB ? tm.uB : 0);
            __syncthreads();
            if (!(p.skip & 1)) pf_load(u.srcA, u.srcB, Nc, Nx, Px, tid, nthreads);
            __syncthreads();
            if (!(p.skip & 2)) {
                axis_pass(Nx, Ny * Nz, Ny * Nz, Px, 0, 1, tid, nthreads);
                __syncthreads();
                axis_pass(Ny, Nx * Nz, Nx, 1, Ny * Px, Px, tid, nthreads);
                __syncthreads();
                axis_pass(Nz, Nx * Ny, Nx, 1, Px, Ny * Px, tid, nthreads);
                __syncthreads();
            }
            if (!(p.skip & 4)) {
                const int fl = (u.hasA ? 1 : 0) | (u.hasB ? 2 : 0), qs = (int)u.qstride;
                cplx* const* blk = p.n_blocks > 0 ? sblk : nullptr;
                if (p.tw) {
                    if (term) pf_epilogue<true, true>(u.outA, u.outB, u.twA, u.twB, qs, fl, p.npad, Q, n_unit, tid, nthreads, blk, p.q_block, row_off);
                    else pf_epilogue<false, true>(u.outA, u.outB, u.twA, u.twB, qs, fl, p.npad, Q, n_unit, tid, nthreads, blk, p.q_block, row_off);
                } else {
                    if (term) pf_epilogue<true, false>(u.outA, u.outB, u.twA, u.twB, qs, fl, p.npad, Q, n_unit, tid, nthreads, blk, p.q_block, row_off);
                    else pf_epilogue<false, false>(u.outA, u.outB, u.twA, u.twB, qs, fl, p.npad, Q, n_unit, tid, nthreads, blk, p.q_block, row_off);
                }
            }
        }
    }
}

struct PfLayout {
    size_t off_jobs, off_terms, off_qoff, off_bad, total;
};

static PfLayout pf_layout(int n_unit, int n_prim, int Q) {
    PfLayout l;
    size_t o = 0;
    l.off_jobs = o;  o += align_up((size_t)(2 * n_prim + 2) * sizeof(PfJob), 256);
    l.off_terms = o; o += align_up((size_t)(3 * n_unit + 4) * sizeof(PfTerm), 256);
    l.off_qoff = o;  o += align_up((size_t)Q * sizeof(int2), 256);
    l.off_bad = o;   o += 256;
    l.total = o;
    return l;
}

static int pick_threads(int Nx, int Ny, int Nz, int Q, int max_threads) {
    // every pass should keep at least 8 warps busy and waste few thread slots in its last round
    const char* env = dev_env("DP_PF_THREADS");
    if (env && atoi(env) >= 32 && atoi(env) <= max_threads) return atoi(env) / 32 * 32;
    int best = 256;
    double best_cost = -1;
    for (int th = 256; th <= max_threads; th += 32) {
        auto rounds = [&](int n) { return (double)((n + th - 1) / th); };
        // time ~ rounds (latency bound); pencil passes weigh more than the element-wise phases
        double cost = 4.0 * (rounds(Ny * Nz) + rounds(Nx * Nz) + rounds(Nx * Ny)) + rounds(Nx * Ny * Nz) / 4.0 +
                      rounds(Q) / 4.0;
        cost *= 1.0 + 0.1 * th / 512.0;       // mild preference for fewer threads (register budget)
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = th; }
    }
    return best;
}

}  // namespace dp

extern "C" {

size_t dp_project_commensurate_workspace_bytes(const int* dims, int n_unit, int n_prim, int Q) {
    if (!dims || n_unit <= 0 || n_prim <= 0 || Q <= 0) return 0;
    return dp::pf_layout(n_unit, n_prim, Q).total;
}

int dp_project_commensurate(const double* v, const int* dims, int n_unit, const int* unit_type, int n_prim,
                            int64_t T, const int* qbin, const double* twiddle, int Q, int project_on_atom,
                            double* out_vc, void* workspace, size_t workspace_bytes, void* stream) {
    return dp_project_commensurate_ex(v, dims, n_unit, unit_type, n_prim, T, qbin, twiddle, Q, project_on_atom,
                                      DP_VC_ROWS, out_vc, workspace, workspace_bytes, stream);
}

static int project_commensurate_impl(const double* v, const int* dims, int n_unit, const int* unit_type, int n_prim,
                                     int64_t T, const int* qbin, const double* twiddle, int Q, int project_on_atom,
                                     int vc_layout, double* out_vc, void* workspace, size_t workspace_bytes,
                                     void* stream, int q_block, double* const* block_out, int n_blocks);

int dp_project_commensurate_ex(const double* v, const int* dims, int n_unit, const int* unit_type, int n_prim,
                               int64_t T, const int* qbin, const double* twiddle, int Q, int project_on_atom,
                               int vc_layout, double* out_vc, void* workspace, size_t workspace_bytes,
                               void* stream) {
    return project_commensurate_impl(v, dims, n_unit, unit_type, n_prim, T, qbin, twiddle, Q, project_on_atom, vc_layout,
                                     out_vc, workspace, workspace_bytes, stream, Q, nullptr, 0);
}

int dp_project_commensurate_blocks(const double* v, const int* dims, int n_unit, const int* unit_type, int n_prim,
                                   int64_t T, const int* qbin, int Q, int q_block, double* const* block_out, int n_blocks,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    DP_REQUIRE(block_out && q_block > 0 && Q % q_block == 0 && n_blocks == Q / q_block && n_blocks <= dp::PF_MAX_BLOCKS,
               DP_ERR_INVALID, "dp_project_commensurate_blocks: %d block pointers for Q=%d in blocks of %d (at most %d blocks)",
               n_blocks, Q, q_block, dp::PF_MAX_BLOCKS);
    for (int g = 0; g < n_blocks; g++)
        DP_REQUIRE(block_out[g] != nullptr, DP_ERR_INVALID, "dp_project_commensurate_blocks: null pointer for block %d", g);
    return project_commensurate_impl(v, dims, n_unit, unit_type, n_prim, T, qbin, nullptr, Q, -1, DP_VC_PAIRS,
                                     block_out[0], workspace, workspace_bytes, stream, q_block, block_out, n_blocks);
}

static int project_commensurate_impl(const double* v, const int* dims, int n_unit, const int* unit_type, int n_prim,
                                     int64_t T, const int* qbin, const double* twiddle, int Q, int project_on_atom,
                                     int vc_layout, double* out_vc, void* workspace, size_t workspace_bytes,
                                     void* stream, int q_block, double* const* block_out, int n_blocks) {
    DP_REQUIRE(vc_layout == DP_VC_ROWS || vc_layout == DP_VC_PAIRS, DP_ERR_INVALID, "dp_project_commensurate: bad vc_layout %d", vc_layout);
    DP_REQUIRE(vc_layout == DP_VC_ROWS || (n_unit == n_prim && project_on_atom < 0 && (3 * n_prim) % 2 == 0), DP_ERR_INVALID,
               "dp_project_commensurate: the pair-major layout needs one unit-cell atom per type, no atom filter and an even 3*n_prim");
    DP_REQUIRE((T == 0 || (v && out_vc)) && dims && unit_type && qbin && workspace, DP_ERR_INVALID,
               "dp_project_commensurate: null pointer");
    DP_REQUIRE(n_unit > 0 && n_prim > 0 && T >= 0 && Q > 0 && project_on_atom < n_prim, DP_ERR_INVALID,
               "dp_project_commensurate: bad sizes");
    DP_REQUIRE(twiddle || n_unit == n_prim, DP_ERR_INVALID,
               "dp_project_commensurate: twiddle == NULL (raw lattice DFT) needs one unit-cell atom per primitive type");
    const int Nx = dims[0], Ny = dims[1], Nz = dims[2];
    DP_REQUIRE(Nx >= 1 && Ny >= 1 && Nz >= 1, DP_ERR_INVALID, "dp_project_commensurate: bad dims");
    DP_REQUIRE(Nx <= DP_ROOTS_NMAX && Ny <= DP_ROOTS_NMAX && Nz <= DP_ROOTS_NMAX, DP_ERR_UNSUPPORTED,
               "dp_project_commensurate: supercell multiplicity > %d (use dp_project_wave_vector)", DP_ROOTS_NMAX);
    const int Px = (Nx % 2 == 0) ? Nx + 1 : Nx;
    const int64_t Nc64 = (int64_t)Nx * Ny * Nz;
    const size_t smem_w = (size_t)Nz * Ny * Px * sizeof(dp::cplx);
    const size_t smem = smem_w + sizeof(dp::cplx) + dp::align_up((size_t)Q * sizeof(unsigned), 16) +
                        (size_t)(2 * n_prim + 2) * sizeof(dp::PfJob) + (size_t)(3 * n_unit + 4) * sizeof(dp::PfTerm) +
                        (size_t)dp::PF_MAX_BLOCKS * sizeof(void*);
    DP_REQUIRE((int64_t)Nz * Ny * Px <= 65535, DP_ERR_UNSUPPORTED,
               "dp_project_commensurate: %lld cells exceed the 16-bit bin offsets (use dp_project_wave_vector)", (long long)Nc64);
    DP_REQUIRE(smem <= 220 * 1024, DP_ERR_UNSUPPORTED,
               "dp_project_commensurate: %lld cells need %zu B of shared memory (use dp_project_wave_vector)",
               (long long)Nc64, smem);
    dp::PfLayout lay = dp::pf_layout(n_unit, n_prim, Q);
    DP_REQUIRE(workspace_bytes >= lay.total, DP_ERR_WORKSPACE, "dp_project_commensurate: workspace too small");
    if (T == 0) return DP_OK;
    const int Nc = (int)Nc64;
    cudaStream_t st = (cudaStream_t)stream;

    std::vector<std::vector<int>> units(n_prim);
    for (int u = 0; u < n_unit; u++) {
        DP_REQUIRE(unit_type[u] >= 0 && unit_type[u] < n_prim, DP_ERR_INVALID, "dp_project_commensurate: unit_type[%d]=%d", u, unit_type[u]);
        units[unit_type[u]].push_back(u);
    }
    std::vector<dp::PfJob> jobs;
    std::vector<dp::PfTerm> terms;
    std::vector<int> types;
    for (int p = 0; p < n_prim; p++)
        if (project_on_atom < 0 || project_on_atom == p) types.push_back(p);
    auto add_job = [&](int pA, int kA, int pB, int kB) {
        dp::PfJob j{pA, kA, pB, kB, (int)terms.size(), 0};
        size_t na = units[pA].size(), nb = pB >= 0 ? units[pB].size() : 0;
        size_t n = std::max(na, nb);
        for (size_t i = 0; i < n; i++) terms.push_back(dp::PfTerm{i < na ? units[pA][i] : -1, i < nb ? units[pB][i] : -1});
        j.nterms = (int)n;
        if (n) jobs.push_back(j);
    };
    // output slots (p, k) = 3p + k of the selected types, two per transform; adjacent slots (2j, 2j+1) are
    // paired so that a thread's two stores fill 32 contiguous bytes (a full sector) per wave vector
    std::vector<int> slots;
    for (int p : types)
        for (int k = 0; k < 3; k++) slots.push_back(3 * p + k);
    for (size_t i = 0; i < slots.size();) {
        const int sa = slots[i];
        if (sa % 2 == 0 && i + 1 < slots.size() && slots[i + 1] == sa + 1) {
            add_job(sa / 3, sa % 3, (sa + 1) / 3, (sa + 1) % 3);
            i += 2;
        } else if (i + 1 < slots.size()) {              // unaligned pair: still one transform
            add_job(sa / 3, sa % 3, slots[i + 1] / 3, slots[i + 1] % 3);
            i += 2;
        } else {
            add_job(sa / 3, sa % 3, -1, 0);
            i += 1;
        }
    }
    // a slot whose type has fewer unit atoms than its partner still needs its first term to STORE:
    // terms are ordered so that index 0 always carries both sequences when both exist.
    bool need_zero = project_on_atom >= 0;
    for (int p : types) if (units[p].empty()) need_zero = true;
    for (const auto& j : jobs)
        if (j.pB >= 0 && units[j.pA].size() != units[j.pB].size()) need_zero = true;
    DP_REQUIRE(!(need_zero && block_out), DP_ERR_INVALID, "dp_project_commensurate_blocks: needs one unit-cell atom per type");
    if (need_zero) DP_CUDA_OK(cudaMemsetAsync(out_vc, 0, (size_t)T * Q * n_prim * 3 * sizeof(dp::cplx), st));
    if (jobs.empty()) return DP_OK;

    char* ws = (char*)workspace;
    dp::PfJob* d_jobs = (dp::PfJob*)(ws + lay.off_jobs);
    dp::PfTerm* d_terms = (dp::PfTerm*)(ws + lay.off_terms);
    int2* d_qoff = (int2*)(ws + lay.off_qoff);
    int* d_bad = (int*)(ws + lay.off_bad);
    // job tables travel as kernel parameters and the q-bin check stays on the device: this entry point never blocks
    // the host.  A bin outside the supercell grid poisons the output of that q with NaN (see the header).
    int rc = dp::upload_table(d_jobs, jobs.data(), jobs.size() * sizeof(dp::PfJob), st);
    if (rc) return rc;
    rc = dp::upload_table(d_terms, terms.data(), terms.size() * sizeof(dp::PfTerm), st);
    if (rc) return rc;
    {
        auto k = dp::qoffset_kernel;
        DP_LAUNCH(k, dim3((unsigned)dp::imin((Q + 255) / 256, 1024)), dim3(256), 0, st, qbin, Q, Nx, Ny, Nz, Px, d_qoff, d_bad);
        DP_CUDA_OK(cudaGetLastError());
    }

    dp::PfParams p;
    p.v = v; p.Nx = Nx; p.Ny = Ny; p.Nz = Nz; p.Px = Px; p.Nc = Nc; p.N = n_unit * Nc;
    p.n_unit = n_unit; p.n_prim = n_prim; p.Q = Q; p.njobs = (int)jobs.size(); p.nterms = (int)terms.size();
    p.npad = Nz * Ny * Px; p.T = T; p.pair_major = (vc_layout == DP_VC_PAIRS);
    p.q_block = q_block; p.n_blocks = block_out ? n_blocks : 0;
    for (int g = 0; g < dp::PF_MAX_BLOCKS; g++) p.blocks[g] = (block_out && g < n_blocks) ? (dp::cplx*)block_out[g] : nullptr;
    p.skip = dp::dev_env("DP_PF_SKIP") ? atoi(dp::dev_env("DP_PF_SKIP")) : 0;
    p.jobs = d_jobs; p.terms = d_terms; p.qoff = d_qoff; p.tw = (const dp::cplx*)twiddle; p.out = (dp::cplx*)out_vc;
    int per_sm = (int)dp::imin<size_t>(2, dp::imax<size_t>(1, (size_t)(226 * 1024) / (smem + 1024)));
    int threads = dp::pick_threads(Nx, Ny, Nz, Q, per_sm >= 2 ? 320 : 512);
    if (threads > 320) per_sm = 1;
    int64_t nunits = T * (int64_t)jobs.size();
    // DP_RESERVE_SMS: leave some SMs to a concurrently running collective (the chunked all-to-all of the
    // multi-GPU pipeline); a persistent grid that fills every SM would make the NCCL kernel wait or be waited for
    int sms = dp::imax(dp::sm_count(), 1);
    if (dp::dev_env("DP_RESERVE_SMS")) sms = dp::imax(sms - atoi(dp::dev_env("DP_RESERVE_SMS")), 1);
    int grid = (int)dp::imin<int64_t>(nunits, (int64_t)sms * per_sm);
    if (per_sm >= 2 && threads <= 256) {
        auto k = dp::project_fft_kernel<256, 2>;
        DP_CUDA_OK(DP_SET_SMEM(k, smem));
        DP_LAUNCH(k, dim3(grid), dim3(threads), smem, st, p);
    } else if (per_sm >= 2) {
        auto k = dp::project_fft_kernel<320, 2>;
        DP_CUDA_OK(DP_SET_SMEM(k, smem));
        DP_LAUNCH(k, dim3(grid), dim3(threads), smem, st, p);
    } else {
        auto k = dp::project_fft_kernel<512, 1>;
        DP_CUDA_OK(DP_SET_SMEM(k, smem));
        DP_LAUNCH(k, dim3(grid), dim3(threads), smem, st, p);
    }
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

}  // extern "C"
