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
template <int N>
__device__ __noinline__ void axis_pass_n(cplx* W, int npencil, int A, int sa, int sb, int st, int tid, int nthreads) {
    for (int pc = tid; pc < npencil; pc += nthreads) {
        const int base = (pc % A) * sa + (pc / A) * sb;
        cplx in[N], out[N];
#pragma unroll
        for (int i = 0; i < N; i++) in[i] = W[base + i * st];
        PencilDft<N>::run(in, out);
#pragma unroll
        for (int i = 0; i < N; i++) W[base + i * st] = out[i];
    }
}

#define DP_AXIS_CASE(n) case n: axis_pass_n<n>(W, npencil, A, sa, sb, st, tid, nthreads); break;
__device__ __forceinline__ void axis_pass(int n, cplx* W, int npencil, int A, int sa, int sb, int st, int tid,
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
struct PfJob { int pA, kA, pB, kB, term0, nterms; };      // pB < 0: single sequence
struct PfTerm { int uA, uB; };                              // uB < 0: none

struct PfParams {
    const double* v;
    int Nx, Ny, Nz, Px, Nc, N, n_unit, n_prim, Q, njobs;
    int64_t T;
    const PfJob* jobs;
    const PfTerm* terms;
    const int2* qoff;        // (offset of bin m(q), offset of bin -m(q)) in the padded layout
    const cplx* tw;          // [Q][n_unit]
    cplx* out;               // [T][Q][n_prim][3]
};

__global__ void __launch_bounds__(256)
qoffset_kernel(const int* __restrict__ qbin, int Q, int Nx, int Ny, int Nz, int Px, int2* __restrict__ qoff,
               int* __restrict__ bad) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < Q; q += gridDim.x * blockDim.x) {
        int mx = qbin[q * 3], my = qbin[q * 3 + 1], mz = qbin[q * 3 + 2];
        if (mx < 0 || mx >= Nx || my < 0 || my >= Ny || mz < 0 || mz >= Nz) { atomicAdd(bad, 1); mx = my = mz = 0; }
        int nx = (Nx - mx) % Nx, ny = (Ny - my) % Ny, nz = (Nz - mz) % Nz;
        qoff[q] = make_int2((mz * Ny + my) * Px + mx, (nz * Ny + ny) * Px + nx);
    }
}

// MAXT / MINB: launch bounds (register budget): <=320 threads with 2 CTAs per SM, or <=512 with 1
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
project_fft_kernel(PfParams p) {
    DP_DYNSMEM(cplx, W);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int Nx = p.Nx, Ny = p.Ny, Nz = p.Nz, Px = p.Px, Nc = p.Nc, Q = p.Q, n_unit = p.n_unit;
    const int2* __restrict__ qoff = p.qoff;
    const cplx* __restrict__ tw = p.tw;
    const int64_t nunits = p.T * p.njobs;
    for (int64_t unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
        const int64_t t = unit / p.njobs;
        const PfJob job = p.jobs[(int)(unit % p.njobs)];
        const double* __restrict__ vt = p.v + t * (int64_t)p.N * 3;
        cplx* __restrict__ outA = p.out + ((t * Q) * p.n_prim + job.pA) * 3 + job.kA;
        cplx* __restrict__ outB = job.pB >= 0 ? p.out + ((t * Q) * p.n_prim + job.pB) * 3 + job.kB : nullptr;
        const int64_t qstride = (int64_t)p.n_prim * 3;
        for (int term = 0; term < job.nterms; term++) {
            const PfTerm tm = p.terms[job.term0 + term];
            const double* __restrict__ srcA = tm.uA >= 0 ? vt + (int64_t)tm.uA * Nc * 3 + job.kA : nullptr;
            const double* __restrict__ srcB = (tm.uB >= 0 && job.pB >= 0) ? vt + (int64_t)tm.uB * Nc * 3 + job.kB : nullptr;
            __syncthreads();
            // ---- load: two real sequences -> one complex array (4 independent loads in flight) ----
            for (int c0 = tid; c0 < Nc; c0 += 4 * nthreads) {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int c = c0 + i * nthreads;
                    a[i] = (srcA && c < Nc) ? __ldg(srcA + (int64_t)c * 3) : 0.0;
                    b[i] = (srcB && c < Nc) ? __ldg(srcB + (int64_t)c * 3) : 0.0;
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int c = c0 + i * nthreads;
                    if (c < Nc) W[(c / Nx) * Px + (c % Nx)] = make_double2(a[i], b[i]);
                }
            }
            __syncthreads();
            axis_pass(Nx, W, Ny * Nz, Ny * Nz, Px, 0, 1, tid, nthreads);
            __syncthreads();
            axis_pass(Ny, W, Nx * Nz, Nx, 1, Ny * Px, Px, tid, nthreads);
            __syncthreads();
            axis_pass(Nz, W, Nx * Ny, Nx, 1, Px, Ny * Px, tid, nthreads);
            __syncthreads();
            // ---- epilogue: separate the two sequences, twiddle, store (4 q in flight per thread) ----
            const bool hasA = tm.uA >= 0, hasB = srcB != nullptr;
            const cplx* __restrict__ twA = tw + (hasA ? tm.uA : 0);
            const cplx* __restrict__ twB = tw + (hasB ? tm.uB : 0);
            for (int q0 = tid; q0 < Q; q0 += 4 * nthreads) {
                int2 o[4];
                cplx wa[4], wb[4], olda[4], oldb[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int q = q0 + i * nthreads;
                    if (q < Q) {
                        o[i] = __ldg(qoff + q);
                        if (hasA) wa[i] = __ldg(twA + (int64_t)q * n_unit);
                        if (hasB) wb[i] = __ldg(twB + (int64_t)q * n_unit);
                        if (term) {
                            if (hasA) olda[i] = outA[q * qstride];
                            if (hasB) oldb[i] = outB[q * qstride];
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int q = q0 + i * nthreads;
                    if (q < Q) {
                        const cplx C = W[o[i].x], D = W[o[i].y];
                        // A = (C + conj D)/2,  B = (C - conj D)/(2i)
                        if (hasA) {
                            cplx r = cmul(make_double2(0.5 * (C.x + D.x), 0.5 * (C.y - D.y)), wa[i]);
                            if (term) r = cadd(olda[i], r);
                            outA[q * qstride] = r;
                        }
                        if (hasB) {
                            cplx r = cmul(make_double2(0.5 * (C.y + D.y), 0.5 * (D.x - C.x)), wb[i]);
                            if (term) r = cadd(oldb[i], r);
                            outB[q * qstride] = r;
                        }
                    }
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
    const char* env = getenv("DP_PF_THREADS");
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
    DP_REQUIRE((T == 0 || (v && out_vc)) && dims && unit_type && qbin && twiddle && workspace, DP_ERR_INVALID,
               "dp_project_commensurate: null pointer");
    DP_REQUIRE(n_unit > 0 && n_prim > 0 && T >= 0 && Q > 0 && project_on_atom < n_prim, DP_ERR_INVALID,
               "dp_project_commensurate: bad sizes");
    const int Nx = dims[0], Ny = dims[1], Nz = dims[2];
    DP_REQUIRE(Nx >= 1 && Ny >= 1 && Nz >= 1, DP_ERR_INVALID, "dp_project_commensurate: bad dims");
    DP_REQUIRE(Nx <= DP_ROOTS_NMAX && Ny <= DP_ROOTS_NMAX && Nz <= DP_ROOTS_NMAX, DP_ERR_UNSUPPORTED,
               "dp_project_commensurate: supercell multiplicity > %d (use dp_project_wave_vector)", DP_ROOTS_NMAX);
    const int Px = (Nx % 2 == 0) ? Nx + 1 : Nx;
    const int64_t Nc64 = (int64_t)Nx * Ny * Nz;
    const size_t smem = (size_t)Nz * Ny * Px * sizeof(dp::cplx);
    DP_REQUIRE(smem <= 220 * 1024, DP_ERR_UNSUPPORTED,
               "dp_project_commensurate: %lld cells need %zu B of shared memory (use dp_project_wave_vector)",
               (long long)Nc64, smem);
    dp::PfLayout lay = dp::pf_layout(n_unit, n_prim, Q);
    DP_REQUIRE(workspace_bytes >= lay.total, DP_ERR_WORKSPACE, "dp_project_commensurate: workspace too small");
    if (T == 0) return DP_OK;
    const int Nc = (int)Nc64;
    cudaStream_t st = (cudaStream_t)stream;

    // jobs: (p,x)+(p,y) per type; z components of two types share a transform
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
    for (int p : types) add_job(p, 0, p, 1);
    for (size_t i = 0; i < types.size(); i += 2) add_job(types[i], 2, i + 1 < types.size() ? types[i + 1] : -1, 2);
    // a slot whose type has fewer unit atoms than its partner still needs its first term to STORE:
    // terms are ordered so that index 0 always carries both sequences when both exist.
    bool need_zero = project_on_atom >= 0;
    for (int p : types) if (units[p].empty()) need_zero = true;
    for (const auto& j : jobs)
        if (j.pB >= 0 && units[j.pA].size() != units[j.pB].size()) need_zero = true;
    if (need_zero) DP_CUDA_OK(cudaMemsetAsync(out_vc, 0, (size_t)T * Q * n_prim * 3 * sizeof(dp::cplx), st));
    if (jobs.empty()) return DP_OK;

    char* ws = (char*)workspace;
    dp::PfJob* d_jobs = (dp::PfJob*)(ws + lay.off_jobs);
    dp::PfTerm* d_terms = (dp::PfTerm*)(ws + lay.off_terms);
    int2* d_qoff = (int2*)(ws + lay.off_qoff);
    int* d_bad = (int*)(ws + lay.off_bad);
    DP_CUDA_OK(cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(dp::PfJob), cudaMemcpyHostToDevice, st));
    DP_CUDA_OK(cudaMemcpyAsync(d_terms, terms.data(), terms.size() * sizeof(dp::PfTerm), cudaMemcpyHostToDevice, st));
    DP_CUDA_OK(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    {
        auto k = dp::qoffset_kernel;
        DP_LAUNCH(k, dim3((unsigned)dp::imin((Q + 255) / 256, 1024)), dim3(256), 0, st, qbin, Q, Nx, Ny, Nz, Px, d_qoff, d_bad);
        DP_CUDA_OK(cudaGetLastError());
    }
    int bad = 0;
    DP_CUDA_OK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    DP_CUDA_OK(cudaStreamSynchronize(st));      // also keeps the host job vectors alive long enough
    DP_REQUIRE(bad == 0, DP_ERR_INVALID, "dp_project_commensurate: %d q bins outside the supercell grid", bad);

    dp::PfParams p;
    p.v = v; p.Nx = Nx; p.Ny = Ny; p.Nz = Nz; p.Px = Px; p.Nc = Nc; p.N = n_unit * Nc;
    p.n_unit = n_unit; p.n_prim = n_prim; p.Q = Q; p.njobs = (int)jobs.size(); p.T = T;
    p.jobs = d_jobs; p.terms = d_terms; p.qoff = d_qoff; p.tw = (const dp::cplx*)twiddle; p.out = (dp::cplx*)out_vc;
    int per_sm = (int)dp::imin<size_t>(2, dp::imax<size_t>(1, (size_t)(226 * 1024) / (smem + 1024)));
    int threads = dp::pick_threads(Nx, Ny, Nz, Q, per_sm >= 2 ? 320 : 512);
    if (threads > 320) per_sm = 1;
    int64_t nunits = T * (int64_t)jobs.size();
    int grid = (int)dp::imin<int64_t>(nunits, (int64_t)dp::imax(dp::sm_count(), 1) * per_sm);
    if (per_sm >= 2) {
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
