// A2 (dense form) and A3: wave-vector projection for arbitrary q lists and the eigenvector
// contraction.  Reference: dynaphopy/projection.py:4-40 and :43-54.
//
// Dense projection = real(3T x N) x complex(N x Q) contraction per primitive type; the
// phases exp(-i q.r_i) * sqrt(m_i) / sqrt(N/n_p) are tabulated once per call (Q x N), the
// main kernel is a tensor-core GEMM (mma.sync m8n8k4 f64 = DMMA, see project_dense_kernel).
// Bound: fp64 (12 flop per atom-frame-q).
// The commensurate fast path (3-D DFT, HBM-bound) lives in dp_projection_fft.cu.
#include <algorithm>
#include <vector>

#include "dp_common.cuh"

namespace dp {

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ph[q][a] for atoms in type-sorted order a -> atom_list[a]
__global__ void __launch_bounds__(256)
phase_table_kernel(const double* __restrict__ q_cart, const double* __restrict__ pos0,
                   const int* __restrict__ atom_list, const double* __restrict__ sqrt_mass, int N, int Q,
                   double norm, cplx* __restrict__ ph) {
    int64_t total = (int64_t)Q * N;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        int q = (int)(idx / N), a = (int)(idx % N);
        int i = atom_list[a];
        double x = q_cart[q * 3 + 0] * pos0[i * 3 + 0] + q_cart[q * 3 + 1] * pos0[i * 3 + 1] +
                   q_cart[q * 3 + 2] * pos0[i * 3 + 2];
        double s, c;
        sincos(x, &s, &c);
        double w = norm * (sqrt_mass ? sqrt_mass[i] : 1.0);
        ph[idx] = make_double2(c * w, -s * w);
    }
}

// ---- dense projection as a tensor-core GEMM (DMMA) ------------------------------------------------------------
// Per primitive type p:  C[(k, t), (q, re|im)] = sum_a V[(k, t), a] * Ph[a, (q, re|im)]  -- a real (3T x N_p) times
// real (N_p x 2Q) product in fp64, computed with mma.sync.aligned.m8n8k4.f64 (SASS: DMMA.8x8x4).  Complex velocities
// double the contraction length: A' = [Re V | Im V], B' = [[Pr, Pi], [-Pi, Pr]].
// CTA = 16 warps, tile PJ_F frames x 3 components (96 rows, row = k * PJ_F + frame) by PJ_Q wave vectors (128 columns,
// column = 2 q + re|im); warp tile 24 x 32 = 3 x 4 fragments of 8 x 8 (24 fp64 accumulators per thread).  The atoms are
// consumed in chunks of PJ_KA atoms staged in shared memory -- A as [row][step], B TRANSPOSED as [column][step], both
// with a row pitch of 4 (mod 16) doubles, so that every fragment load and every staging store of a warp is conflict
// free and both global reads are coalesced (consecutive threads: consecutive components / atoms of one frame,
// consecutive atoms of one wave vector).  The next chunk travels global -> registers while the current one is multiplied.
constexpr int PJ_F = 32;     // frames per CTA
constexpr int PJ_Q = 64;     // wave vectors per CTA
constexpr int PJ_KA = 32;    // atoms per chunk
constexpr int PJ_THREADS = 512;
constexpr int PJ_ROWS = 3 * PJ_F, PJ_COLS = 2 * PJ_Q;
template <int NC> struct PjCfg {
    static constexpr int KC = PJ_KA * NC;                 // contraction steps (doubles per row) per chunk
    static constexpr int PITCH = KC + 4;                  // shared-memory pitch of As and Bt rows (doubles)
    static constexpr int NA = PJ_F * PJ_KA * 3 * NC / PJ_THREADS;       // staged doubles of A per thread
    static constexpr int NB = PJ_Q * PJ_KA / PJ_THREADS;                // staged phases (complex) per thread
    static constexpr size_t SMEM = (size_t)(PJ_ROWS + PJ_COLS) * PITCH * sizeof(double) + PJ_KA * sizeof(int);
};

__device__ __forceinline__ void dmma_884(double (&c)[2], double a, double b) {
#ifdef DP_EMUL
    (void)c; (void)a; (void)b;      // the emulated kernel below does not use fragments
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
#endif
}

template <int NC>   // 1: real velocities, 2: complex128 velocities
__global__ void __launch_bounds__(PJ_THREADS, 1)
project_dense_kernel(const double* __restrict__ v, const cplx* __restrict__ ph,
                     const int* __restrict__ atom_list, const int* __restrict__ type_off, int N,
                     int n_prim, int64_t T, int Q, int p_first, int p_last, cplx* __restrict__ out) {
    typedef PjCfg<NC> Cfg;
    constexpr int KC = Cfg::KC, PITCH = Cfg::PITCH;
    DP_DYNSMEM(double, pj_smem);                       // As[PJ_ROWS][PITCH] | Bt[PJ_COLS][PITCH] | atoms[PJ_KA]
    double* As = pj_smem;
    double* Bt = pj_smem + PJ_ROWS * PITCH;
    int* atoms = reinterpret_cast<int*>(pj_smem + (PJ_ROWS + PJ_COLS) * PITCH);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * PJ_F;
    const int q0 = blockIdx.y * PJ_Q;
    const int wr = (warp >> 2) * 24, wc = (warp & 3) * 32;       // warp tile origin
    for (int p = p_first; p <= p_last; p++) {
        const int a_begin = type_off[p], a_end = type_off[p + 1];
        const int natoms = a_end - a_begin;
        double acc[3][4][2];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        double ra[Cfg::NA];
        cplx rb[Cfg::NB];
        // A: element e -> frame f = e / (KA*3*NC), then (atom, component k, re|im) fastest: the three (six) doubles of
        // an atom are adjacent in global memory.  B: element e -> wave vector e / KA, atom e % KA.
        auto fetch = [&](int a0) {
#pragma unroll
            for (int i = 0; i < Cfg::NA; i++) {
                const int e = tid + i * PJ_THREADS, f = e / (PJ_KA * 3 * NC), r = e % (PJ_KA * 3 * NC);
                const int a = r / (3 * NC), kc = r % (3 * NC);
                double val = 0.0;
                if (a0 + a < natoms && t0 + f < T)
                    val = __ldg(v + ((t0 + f) * N + __ldg(atom_list + a_begin + a0 + a)) * (3 * NC) + kc);
                ra[i] = val;
            }
#pragma unroll
            for (int i = 0; i < Cfg::NB; i++) {
                const int e = tid + i * PJ_THREADS, q = q0 + e / PJ_KA, a = e % PJ_KA;
                rb[i] = (a0 + a < natoms && q < Q) ? __ldg(ph + (int64_t)q * N + a_begin + a0 + a) : make_double2(0.0, 0.0);
            }
        };
        auto stash = [&]() {
#pragma unroll
            for (int i = 0; i < Cfg::NA; i++) {
                const int e = tid + i * PJ_THREADS, f = e / (PJ_KA * 3 * NC), r = e % (PJ_KA * 3 * NC);
                const int a = r / (3 * NC), kc = r % (3 * NC), k = kc / NC, c = kc % NC;
                As[(k * PJ_F + f) * PITCH + a * NC + c] = ra[i];
            }
#pragma unroll
            for (int i = 0; i < Cfg::NB; i++) {
                const int e = tid + i * PJ_THREADS, ql = e / PJ_KA, a = e % PJ_KA;
                double* b0 = Bt + (2 * ql) * PITCH + a * NC;       // column (q, re), then (q, im) one pitch further
                if (NC == 1) { b0[0] = rb[i].x; b0[PITCH] = rb[i].y; }
                else { b0[0] = rb[i].x; b0[1] = -rb[i].y; b0[PITCH] = rb[i].y; b0[PITCH + 1] = rb[i].x; }
            }
        };
        fetch(0);
        for (int a0 = 0; a0 < natoms; a0 += PJ_KA) {
            __syncthreads();                       // the previous chunk has been consumed
            stash();
            __syncthreads();
            if (a0 + PJ_KA < natoms) fetch(a0 + PJ_KA);       // next chunk in flight underneath the multiplications
#ifdef DP_EMUL
            // host emulation (tests only): the same tile product without tensor-core fragments; thread (lane, warp)
            // owns the accumulator entries the DMMA layout gives it
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    for (int kk = 0; kk < KC; kk++) {
                        const double a = As[(wr + i * 8 + lane / 4) * PITCH + kk];
                        acc[i][j][0] += a * Bt[(wc + j * 8 + 2 * (lane % 4)) * PITCH + kk];
                        acc[i][j][1] += a * Bt[(wc + j * 8 + 2 * (lane % 4) + 1) * PITCH + kk];
                    }
#else
#pragma unroll
            for (int ks = 0; ks < KC / 4; ks++) {
                double fa[3], fb[4];
#pragma unroll
                for (int i = 0; i < 3; i++) fa[i] = As[(wr + i * 8 + lane / 4) * PITCH + ks * 4 + lane % 4];
#pragma unroll
                for (int j = 0; j < 4; j++) fb[j] = Bt[(wc + j * 8 + lane / 4) * PITCH + ks * 4 + lane % 4];
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma_884(acc[i][j], fa[i], fb[j]);
            }
#endif
        }
        // accumulator (row, col) = (wr + 8 i + lane / 4, wc + 8 j + 2 (lane % 4) + {0, 1}): one complex value per register pair
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int row = wr + i * 8 + lane / 4, k = row / PJ_F;
            const int64_t t = t0 + row % PJ_F;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int q = q0 + ((wc + j * 8) >> 1) + lane % 4;
                if (t < T && q < Q) out[((t * Q + q) * n_prim + p) * 3 + k] = make_double2(acc[i][j][0], acc[i][j][1]);
            }
        }
    }
    (void)atoms;
}

// vq[t,q,s] = sum_p ( sum_k vc[t,q,p,k] conj(e[q,s,p,k]) ): CTA = one q, PH_TF frames.
constexpr int PH_TF = 64;

__global__ void __launch_bounds__(256)
project_phonon_kernel(const cplx* vc, const cplx* __restrict__ eig, int64_t T, int Q, int M, int q_block,
                      cplx* vq) {
    DP_DYNSMEM(cplx, sm);
    cplx* es = sm;                      // [M][M]   e[q][s][j]
    cplx* vs = sm + (size_t)M * M;      // [PH_TF][M]
    const int q = blockIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * PH_TF;
    const int nthreads = blockDim.x;
    for (int e = threadIdx.x; e < M * M; e += nthreads) es[e] = eig[(int64_t)q * M * M + e];
    for (int e = threadIdx.x; e < PH_TF * M; e += nthreads) {
        int f = e / M, j = e % M;
        cplx val = make_double2(0.0, 0.0);
        if (t0 + f < T) val = vc[((t0 + f) * Q + q) * M + j];
        vs[e] = val;
    }
    __syncthreads();
    const int np = M / 3;
    for (int e = threadIdx.x; e < PH_TF * M; e += nthreads) {
        int f = e / M, s = e % M;
        if (t0 + f >= T) continue;
        cplx acc = make_double2(0.0, 0.0);
        for (int p = 0; p < np; p++) {
            cplx part = make_double2(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                cplx a = vs[f * M + p * 3 + k], b = es[s * M + p * 3 + k];
                part = cadd(part, cmulc(a, b));
            }
            acc = cadd(acc, part);
        }
        // [q / q_block][t][q % q_block][s]; q_block == Q gives the plain [t][q][s] layout
        vq[(((int64_t)(q / q_block) * T + t0 + f) * q_block + (q % q_block)) * M + s] = acc;
    }
}


// Same contraction, series-major output for the spectrum kernels:
//   out[((q / q_block) * q_block * M + (q % q_block) * M + s) * T_total + t_offset + t]
// i.e. one contiguous time series per (q, s) (HBM-friendly for the time FFTs) and, for q_block < Q, one
// contiguous block per destination rank of the all-to-all.  CTA = PS_TT frames x qt wave vectors:
// the tile is read as PS_TT runs of qt*M contiguous values, contracted from shared memory (lane = frame,
// so the eigenvector reads are broadcasts) and written as runs of PS_TT consecutive time samples.
constexpr int PS_TT = 32;
constexpr int PS_QT_MAX = 16;       // wave vectors per CTA (fewer for wide M: shared-memory budget)
constexpr int PS_MAX_BLOCKS = 16;   // destination blocks with their own base pointer (ranks of one NVSwitch domain)

// Per-destination output bases: block g = q / q_block is written at p[g] instead of out + g*q_block*M*T_total.
// The pointers may be PEER device memory mapped over NVLink: the contraction then IS the exchange step (its
// 512-byte runs of consecutive time samples go straight into the destination rank's receive buffer).
struct PsBlocks { cplx* p[PS_MAX_BLOCKS]; int n; };

__device__ __forceinline__ void ps_cp16(cplx* dst_smem, const cplx* src) {
#ifdef DP_EMUL
    *dst_smem = *src;
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void ps_cp_wait_all() {
#ifndef DP_EMUL
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
#endif
}

// MIRROR (velocities are real, so the lattice DFT is Hermitian: vc[t, -q, :] = conj(vc[t, q, :])): the intermediate holds
// only one wave vector h of every pair (q, -q); the tile contracts it twice -- with the eigenvectors of qmap[h].x as it
// is, and with the CONJUGATED eigenvectors of the partner qmap[h].y (second half of eig[h]) followed by a conjugation of
// the result: vq[-q, s] = sum conj(e[-q,s,j]) conj(vc[q,j]) = conj(sum conj(conj e[-q,s,j]) vc[q,j]), bit for bit what
// the unpaired contraction of the explicitly stored vc[-q] gives.  Q is then the number of stored wave vectors.
template <int MT, bool MIRROR>   // MT > 0: compile-time M (inputs kept in registers); 0: any M
__global__ void __launch_bounds__(256)
project_phonon_series_kernel(const cplx* __restrict__ vc, const cplx* __restrict__ eig, int64_t T, int Q, int Mdyn,
                             int q_block, int qt, int vc_pairs, int64_t t_offset, int64_t T_total,
                             cplx* __restrict__ out, const __grid_constant__ PsBlocks blocks,
                             const int2* __restrict__ qmap) {
    DP_DYNSMEM(cplx, sm);
    const int M = MT > 0 ? MT : Mdyn;
    const int pitch = qt * M + 1;                   // odd: conflict-free column reads
    constexpr int NE = MIRROR ? 2 : 1;                  // eigenvector sets per stored wave vector
    cplx* in = sm;                                      // [PS_TT][pitch]
    cplx* es = sm + (size_t)PS_TT * pitch;              // [qt][NE][M*M]
    const int64_t t0 = (int64_t)blockIdx.x * PS_TT;
    const int q0 = blockIdx.y * qt;
    const int nq = imin(qt, Q - q0);
    const int nthreads = blockDim.x, tid = threadIdx.x;
    if (!vc_pairs) {
        for (int e = tid; e < PS_TT * nq * M; e += nthreads) {
            const int f = e / (nq * M), j = e % (nq * M);
            cplx val = make_double2(0.0, 0.0);
            if (t0 + f < T) val = __ldcs(vc + ((t0 + f) * Q + q0) * M + j);
            in[f * pitch + j] = val;
        }
    } else if (MT > 0 && nq == PS_QT_MAX && t0 + PS_TT <= T) {
        // full tile, compile-time M: shifts instead of divisions, and the 16-byte elements travel global -> shared
        // memory with cp.async (no register staging, 12 copies in flight per thread)
        constexpr int np2 = MT > 0 ? MT / 2 : 1, run = PS_QT_MAX * 2;       // (MT == 0 never takes this branch)
#pragma unroll
        for (int e = tid; e < PS_TT * np2 * run; e += 256) {
            const int f = e / (np2 * run), r = e % (np2 * run), pr = r / run, w = r % run;
            ps_cp16(in + f * pitch + (w >> 1) * MT + pr * 2 + (w & 1), vc + (((t0 + f) * np2 + pr) * Q + q0) * 2 + w);
        }
        ps_cp_wait_all();
    } else {                       // vc[t][slot pair][q][2]: runs of nq*2 contiguous values per (frame, pair)
        const int np2 = M / 2, run = nq * 2;
        for (int e = tid; e < PS_TT * np2 * run; e += nthreads) {
            const int f = e / (np2 * run), r = e % (np2 * run), pr = r / run, w = r % run;
            cplx val = make_double2(0.0, 0.0);
            if (t0 + f < T) val = __ldcs(vc + (((t0 + f) * np2 + pr) * Q + q0) * 2 + w);
            in[f * pitch + (w >> 1) * M + pr * 2 + (w & 1)] = val;
        }
    }
    for (int e = tid; e < nq * NE * M * M; e += nthreads) es[e] = __ldg(eig + (int64_t)q0 * NE * M * M + e);
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const int64_t t = t0 + lane;
    const int np = M / 3;
    for (int qq = warp; qq < nq; qq += nwarps) {
        const cplx* row = in + lane * pitch + qq * M;
        cplx a[MT > 0 ? MT : 1];
        if (MT > 0) {
#pragma unroll
            for (int j = 0; j < MT; j++) a[j] = row[j];
        }
#pragma unroll
        for (int side = 0; side < NE; side++) {
            int q = q0 + qq;
            if (MIRROR) {
                const int2 qm = __ldg(qmap + q);
                q = side ? qm.y : qm.x;
                if (q < 0) continue;                  // self-conjugate wave vector (or a partner outside the list)
            }
            const cplx* e = es + (qq * NE + side) * M * M;
            cplx* dst = (blocks.n > 0 ? blocks.p[q / q_block] : out + (int64_t)(q / q_block) * q_block * M * T_total) +
                        (int64_t)(q % q_block) * M * T_total + t_offset + t;
            if (MT > 0) {
#pragma unroll
                for (int s = 0; s < MT; s++) {
                    cplx acc = make_double2(0.0, 0.0);
#pragma unroll
                    for (int p = 0; p < MT / 3; p++) {
                        cplx part = make_double2(0.0, 0.0);
#pragma unroll
                        for (int k = 0; k < 3; k++) part = cadd(part, cmulc(a[p * 3 + k], e[s * MT + p * 3 + k]));
                        acc = cadd(acc, part);
                    }
                    if (MIRROR && side) acc.y = -acc.y;
                    if (t < T) __stcs(dst + (int64_t)s * T_total, acc);
                }
            } else {
                for (int s = 0; s < M; s++) {
                    cplx acc = make_double2(0.0, 0.0);
                    for (int p = 0; p < np; p++) {
                        cplx part = make_double2(0.0, 0.0);
                        for (int k = 0; k < 3; k++) part = cadd(part, cmulc(row[p * 3 + k], e[s * M + p * 3 + k]));
                        acc = cadd(acc, part);
                    }
                    if (MIRROR && side) acc.y = -acc.y;
                    if (t < T) __stcs(dst + (int64_t)s * T_total, acc);
                }
            }
        }
    }
}

}  // namespace dp

extern "C" {

size_t dp_project_wave_vector_workspace_bytes(int N, int Q) {
    if (N <= 0 || Q <= 0) return 0;
    return dp::align_up((size_t)Q * 3 * sizeof(double), 256) + dp::align_up((size_t)N * sizeof(int), 256) +
           dp::align_up(1024 * sizeof(int), 256) + dp::align_up((size_t)Q * N * sizeof(dp::cplx), 256);
}

int dp_project_wave_vector(const double* v, int v_is_complex, const double* sqrt_mass,
                           const double* pos0, const int* type_idx, int N, int n_prim, int64_t T,
                           const double* q_cart, int Q, int project_on_atom, double* out_vc,
                           void* workspace, size_t workspace_bytes, void* stream) {
    DP_REQUIRE((T == 0 || (v && out_vc)) && pos0 && type_idx && q_cart && workspace, DP_ERR_INVALID,
               "dp_project_wave_vector: null pointer");
    DP_REQUIRE(N > 0 && n_prim > 0 && n_prim < 1024 && T >= 0 && Q > 0, DP_ERR_INVALID,
               "dp_project_wave_vector: bad sizes N=%d n_prim=%d T=%lld Q=%d", N, n_prim, (long long)T, Q);
    DP_REQUIRE(project_on_atom < n_prim, DP_ERR_INVALID, "dp_project_wave_vector: project_on_atom=%d >= n_prim=%d",
               project_on_atom, n_prim);
    DP_REQUIRE(workspace_bytes >= dp_project_wave_vector_workspace_bytes(N, Q), DP_ERR_WORKSPACE,
               "dp_project_wave_vector: workspace too small");
    if (T == 0) return DP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // type-sorted atom list (stable: reference order inside a type is preserved)
    std::vector<int> list(N), off(n_prim + 1, 0);
    for (int i = 0; i < N; i++) {
        DP_REQUIRE(type_idx[i] >= 0 && type_idx[i] < n_prim, DP_ERR_INVALID,
                   "dp_project_wave_vector: type_idx[%d]=%d out of range", i, type_idx[i]);
        off[type_idx[i] + 1]++;
    }
    for (int p = 0; p < n_prim; p++) off[p + 1] += off[p];
    {
        std::vector<int> cur(off.begin(), off.end() - 1);
        for (int i = 0; i < N; i++) list[cur[type_idx[i]]++] = i;
    }
    char* ws = (char*)workspace;
    double* d_q = (double*)ws;            ws += dp::align_up((size_t)Q * 3 * sizeof(double), 256);
    int* d_list = (int*)ws;               ws += dp::align_up((size_t)N * sizeof(int), 256);
    int* d_off = (int*)ws;                ws += dp::align_up(1024 * sizeof(int), 256);
    dp::cplx* d_ph = (dp::cplx*)ws;
    // small tables travel as kernel parameters (no host synchronisation); tables above 64 KB use one pageable
    // cudaMemcpyAsync, which is staged before it returns
    int rc;
    if ((rc = dp::upload_table(d_q, q_cart, (size_t)Q * 3 * sizeof(double), st))) return rc;
    if ((rc = dp::upload_table(d_list, list.data(), (size_t)N * sizeof(int), st))) return rc;
    if ((rc = dp::upload_table(d_off, off.data(), (size_t)(n_prim + 1) * sizeof(int), st))) return rc;
    double norm = 1.0 / sqrt((double)N / (double)n_prim);
    {
        int64_t total = (int64_t)Q * N;
        int grid = (int)dp::imin<int64_t>((total + 255) / 256, (int64_t)dp::imax(dp::sm_count(), 1) * 16);
        auto k = dp::phase_table_kernel;
        DP_LAUNCH(k, dim3(grid), dim3(256), 0, st, d_q, pos0, d_list, sqrt_mass, N, Q, norm, d_ph);
        DP_CUDA_OK(cudaGetLastError());
    }
    int p_first = 0, p_last = n_prim - 1;
    if (project_on_atom > -1) {
        DP_CUDA_OK(cudaMemsetAsync(out_vc, 0, (size_t)T * Q * n_prim * 3 * sizeof(dp::cplx), st));
        p_first = p_last = project_on_atom;
    }
    dim3 grid((unsigned)((T + dp::PJ_F - 1) / dp::PJ_F), (unsigned)((Q + dp::PJ_Q - 1) / dp::PJ_Q));
    dim3 block(dp::PJ_THREADS);
    DP_REQUIRE(grid.y <= 65535, DP_ERR_UNSUPPORTED, "dp_project_wave_vector: Q=%d too large for one launch", Q);
    if (v_is_complex) {
        auto k = dp::project_dense_kernel<2>;
        DP_CUDA_OK(DP_SET_SMEM(k, dp::PjCfg<2>::SMEM));
        DP_LAUNCH(k, grid, block, dp::PjCfg<2>::SMEM, st, v, d_ph, d_list, d_off, N, n_prim, T, Q, p_first, p_last, (dp::cplx*)out_vc);
    } else {
        auto k = dp::project_dense_kernel<1>;
        DP_CUDA_OK(DP_SET_SMEM(k, dp::PjCfg<1>::SMEM));
        DP_LAUNCH(k, grid, block, dp::PjCfg<1>::SMEM, st, v, d_ph, d_list, d_off, N, n_prim, T, Q, p_first, p_last, (dp::cplx*)out_vc);
    }
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

int dp_project_phonon(const double* vc, const double* eig, int64_t T, int Q, int M, double* vq, void* stream) {
    return dp_project_phonon_blocked(vc, eig, T, Q, M, Q, vq, stream);
}

int dp_project_phonon_blocked(const double* vc, const double* eig, int64_t T, int Q, int M, int q_block,
                              double* vq, void* stream) {
    DP_REQUIRE((T == 0 || (vc && vq)) && eig, DP_ERR_INVALID, "dp_project_phonon: null pointer");
    DP_REQUIRE(q_block > 0 && Q % q_block == 0, DP_ERR_INVALID, "dp_project_phonon: q_block=%d must divide Q=%d", q_block, Q);
    DP_REQUIRE(vc != vq || q_block == Q, DP_ERR_INVALID, "dp_project_phonon: in-place needs q_block == Q");
    DP_REQUIRE(T >= 0 && Q > 0 && M > 0 && M % 3 == 0, DP_ERR_INVALID, "dp_project_phonon: bad sizes (M must be 3*n_prim)");
    DP_REQUIRE(Q <= 65535, DP_ERR_UNSUPPORTED, "dp_project_phonon: Q=%d too large for one launch", Q);
    if (T == 0) return DP_OK;
    size_t smem = ((size_t)M * M + (size_t)dp::PH_TF * M) * sizeof(dp::cplx);
    DP_REQUIRE(smem <= 200 * 1024, DP_ERR_UNSUPPORTED, "dp_project_phonon: M=%d too large", M);
    auto k = dp::project_phonon_kernel;
    DP_CUDA_OK(DP_SET_SMEM(k, smem));
    dim3 grid((unsigned)((T + dp::PH_TF - 1) / dp::PH_TF), (unsigned)Q);
    DP_LAUNCH(k, grid, dim3(256), smem, (cudaStream_t)stream, (const dp::cplx*)vc, (const dp::cplx*)eig, T, Q, M,
              q_block, (dp::cplx*)vq);
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

int dp_project_phonon_series(const double* vc, const double* eig, int64_t T, int Q, int M, int q_block,
                             int64_t t_offset, int64_t T_total, double* out, void* stream) {
    return dp_project_phonon_series_ex(vc, DP_VC_ROWS, eig, T, Q, M, q_block, t_offset, T_total, out, stream);
}

// qmap != nullptr: mirror form -- vc / eig hold Q stored wave vectors, the outputs are addressed on the full mesh of
// Q_full wave vectors (see project_phonon_series_kernel)
static int project_phonon_series_impl(const double* vc, int vc_layout, const double* eig, int64_t T, int Q, int M,
                                      int q_block, int64_t t_offset, int64_t T_total, double* out,
                                      double* const* block_out, int n_blocks, void* stream,
                                      const int* qmap = nullptr, int Q_full = 0) {
    if (!qmap) Q_full = Q;
    DP_REQUIRE(!qmap || vc_layout == DP_VC_PAIRS, DP_ERR_INVALID, "dp_project_phonon_series_mirror: needs the pair-major layout");
    DP_REQUIRE(Q_full >= Q, DP_ERR_INVALID, "dp_project_phonon_series_mirror: Q_full=%d < stored Q=%d", Q_full, Q);
    DP_REQUIRE(vc_layout == DP_VC_ROWS || (vc_layout == DP_VC_PAIRS && M % 2 == 0), DP_ERR_INVALID,
               "dp_project_phonon_series: bad vc_layout %d (the pair-major layout needs an even M)", vc_layout);
    const int vc_pairs = vc_layout == DP_VC_PAIRS;
    DP_REQUIRE((T == 0 || (vc && (out || block_out))) && eig, DP_ERR_INVALID, "dp_project_phonon_series: null pointer");
    DP_REQUIRE(vc != out, DP_ERR_INVALID, "dp_project_phonon_series: cannot run in place (the output is transposed)");
    DP_REQUIRE(q_block > 0 && Q_full % q_block == 0, DP_ERR_INVALID, "dp_project_phonon_series: q_block=%d must divide Q=%d", q_block, Q_full);
    DP_REQUIRE(T >= 0 && Q > 0 && M > 0 && M % 3 == 0, DP_ERR_INVALID, "dp_project_phonon_series: bad sizes (M must be 3*n_prim)");
    DP_REQUIRE(t_offset >= 0 && t_offset + T <= T_total, DP_ERR_INVALID, "dp_project_phonon_series: frames [%lld, %lld) outside the series length %lld",
               (long long)t_offset, (long long)(t_offset + T), (long long)T_total);
    dp::PsBlocks blocks;
    blocks.n = 0;
    for (int g = 0; g < dp::PS_MAX_BLOCKS; g++) blocks.p[g] = nullptr;
    if (block_out) {
        DP_REQUIRE(n_blocks == Q_full / q_block && n_blocks <= dp::PS_MAX_BLOCKS, DP_ERR_INVALID,
                   "dp_project_phonon_series_peers: %d block pointers for Q/q_block = %d blocks (at most %d)", n_blocks,
                   Q_full / q_block, dp::PS_MAX_BLOCKS);
        for (int g = 0; g < n_blocks; g++) {
            DP_REQUIRE(block_out[g] != nullptr && (const double*)block_out[g] != vc, DP_ERR_INVALID,
                       "dp_project_phonon_series_peers: bad pointer for block %d", g);
            blocks.p[g] = (dp::cplx*)block_out[g];
        }
        blocks.n = n_blocks;
    }
    if (T == 0) return DP_OK;
    const int ne = qmap ? 2 : 1;
    const size_t per_q = ((size_t)dp::PS_TT * M + (size_t)ne * M * M) * sizeof(dp::cplx);
    const int qt = (int)dp::imax<size_t>(1, dp::imin<size_t>(dp::PS_QT_MAX, (size_t)(72 * 1024) / per_q));
    const unsigned gy = (unsigned)((Q + qt - 1) / qt);
    DP_REQUIRE(gy <= 65535, DP_ERR_UNSUPPORTED, "dp_project_phonon_series: Q=%d too large for one launch", Q);
    size_t smem = ((size_t)dp::PS_TT * (qt * M + 1) + (size_t)qt * ne * M * M) * sizeof(dp::cplx);
    DP_REQUIRE(smem <= 200 * 1024, DP_ERR_UNSUPPORTED, "dp_project_phonon_series: M=%d too large", M);
    dim3 grid((unsigned)((T + dp::PS_TT - 1) / dp::PS_TT), gy);
    cudaStream_t st = (cudaStream_t)stream;
#define DP_PS_LAUNCH(MT, MIRROR)                                                                                  \
    do {                                                                                                          \
        auto k = dp::project_phonon_series_kernel<MT, MIRROR>;                                                    \
        DP_CUDA_OK(DP_SET_SMEM(k, smem));                                                                         \
        DP_LAUNCH(k, grid, dim3(256), smem, st, (const dp::cplx*)vc, (const dp::cplx*)eig, T, Q, M, q_block,     \
                  qt, vc_pairs, t_offset, T_total, (dp::cplx*)out, blocks, (const int2*)qmap);                    \
    } while (0)
    if (qmap) {
        if (M == 6) DP_PS_LAUNCH(6, true);
        else if (M == 12) DP_PS_LAUNCH(12, true);
        else DP_PS_LAUNCH(0, true);
    } else {
        if (M == 6) DP_PS_LAUNCH(6, false);
        else if (M == 12) DP_PS_LAUNCH(12, false);
        else DP_PS_LAUNCH(0, false);
    }
#undef DP_PS_LAUNCH
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

int dp_project_phonon_series_ex(const double* vc, int vc_layout, const double* eig, int64_t T, int Q, int M, int q_block,
                                int64_t t_offset, int64_t T_total, double* out, void* stream) {
    DP_REQUIRE(T == 0 || out, DP_ERR_INVALID, "dp_project_phonon_series: null pointer");
    return project_phonon_series_impl(vc, vc_layout, eig, T, Q, M, q_block, t_offset, T_total, out, nullptr, 0, stream);
}

int dp_project_phonon_series_peers(const double* vc, int vc_layout, const double* eig, int64_t T, int Q, int M,
                                   int q_block, int64_t t_offset, int64_t T_total, double* const* block_out,
                                   int n_blocks, void* stream) {
    DP_REQUIRE(block_out, DP_ERR_INVALID, "dp_project_phonon_series_peers: null pointer");
    return project_phonon_series_impl(vc, vc_layout, eig, T, Q, M, q_block, t_offset, T_total, nullptr, block_out,
                                      n_blocks, stream);
}
int dp_project_phonon_series_mirror(const double* vc, const double* eig2, const int* qmap, int64_t T, int Q_stored,
                                    int Q, int M, int q_block, int64_t t_offset, int64_t T_total, double* out,
                                    double* const* block_out, int n_blocks, void* stream) {
    DP_REQUIRE(qmap && (out || block_out), DP_ERR_INVALID, "dp_project_phonon_series_mirror: null pointer");
    return project_phonon_series_impl(vc, DP_VC_PAIRS, eig2, T, Q_stored, M, q_block, t_offset, T_total, out,
                                      block_out, block_out ? n_blocks : 0, stream, qmap, Q);
}

}  // extern "C"
