// A8 / A9 / A1: minimum-image displacements, np.gradient velocities, sqrt(mass) weighting.
//
// Reference: c/displacements.c:158-183 (per-frame arithmetic), :222-265 (cell transpose and
// cofactor inverse), dynaphopy/dynamics.py:313-329 (np.gradient), :143-156 (mass weighting).
// All three are HBM-bound element-wise maps; the fused kernel reads each position once
// (24 B per atom-frame) and writes the mass-weighted velocity once (24 B).
// Arithmetic uses explicit round-to-nearest mul/add (no FMA contraction) in the reference's
// operation order, so the displacement is bit-identical to the compiled reference.
#include <cmath>

#include "dp_common.cuh"

namespace dp {

struct CellMats {
    double c[9];    // C = cell^T  (c/displacements.c:236-238)
    double ci[9];   // C^-1 by cofactors (c/displacements.c:250-265)
};

// Host replica of the reference's inverse, same operation order (first-row expansion).
static double det2(double a, double b, double c, double d) { return a * d - c * b; }

static void make_cell_mats(const double* cell, CellMats* m) {
    double a[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) a[i][j] = cell[j * 3 + i];
    double det = 0.0;
    for (int j1 = 0; j1 < 3; j1++) {
        double mm[2][2];
        for (int i = 1; i < 3; i++) {
            int j2 = 0;
            for (int j = 0; j < 3; j++) {
                if (j == j1) continue;
                mm[i - 1][j2++] = a[i][j];
            }
        }
        det += std::pow(-1.0, j1 + 2.0) * a[0][j1] * det2(mm[0][0], mm[0][1], mm[1][0], mm[1][1]);
    }
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) {
            double c[2][2];
            int i1 = 0;
            for (int ii = 0; ii < 3; ii++) {
                if (ii == i) continue;
                int j1 = 0;
                for (int jj = 0; jj < 3; jj++) {
                    if (jj == j) continue;
                    c[i1][j1++] = a[ii][jj];
                }
                i1++;
            }
            double cof = std::pow(-1.0, i + j + 2.0) * det2(c[0][0], c[0][1], c[1][0], c[1][1]);
            m->ci[j * 3 + i] = cof / det;      // b[j][i] = cof[i][j] / det
        }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) m->c[i * 3 + j] = a[i][j];
}

__device__ __forceinline__ void displacement3(const double x[3], const double p[3], const CellMats& m,
                                              double u[3]) {
    double d[3], f[3];
#pragma unroll
    for (int k = 0; k < 3; k++) d[k] = __dsub_rn(x[k], p[k]);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double s = __dmul_rn(m.ci[i * 3 + 0], d[0]);
        s = __dadd_rn(s, __dmul_rn(m.ci[i * 3 + 1], d[1]));
        s = __dadd_rn(s, __dmul_rn(m.ci[i * 3 + 2], d[2]));
        f[i] = round(s);                       // C round(): half away from zero (:170)
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double s = __dmul_rn(m.c[i * 3 + 0], f[0]);
        s = __dadd_rn(s, __dmul_rn(m.c[i * 3 + 1], f[1]));
        s = __dadd_rn(s, __dmul_rn(m.c[i * 3 + 2], f[2]));
        u[i] = __dsub_rn(__dsub_rn(x[i], s), p[i]);
    }
}

template <int NC>   // NC = 1 real input, 2 complex128 input (real part used)
__device__ __forceinline__ void load_pos(const double* __restrict__ base, int64_t atom_frame, double x[3]) {
#pragma unroll
    for (int k = 0; k < 3; k++) x[k] = __ldg(base + (atom_frame * 3 + k) * NC);
}

template <int NC>
__global__ void __launch_bounds__(256)
displacements_kernel(const double* __restrict__ traj, const double* __restrict__ pos0, CellMats m,
                     int64_t total, int N, double* __restrict__ out) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        int atom = (int)(idx % N);
        double x[3], p[3], u[3];
        load_pos<NC>(traj, idx, x);
#pragma unroll
        for (int k = 0; k < 3; k++) p[k] = __ldg(pos0 + atom * 3 + k);
        displacement3(x, p, m, u);
#pragma unroll
        for (int k = 0; k < 3; k++) out[idx * 3 + k] = u[k];
    }
}

// One thread walks `run` consecutive frames of one atom with a 3-frame sliding window, so every
// position is read once (plus one halo frame per run) and no displacement array is re-read.
template <int NC>
__global__ void __launch_bounds__(256)
velocity_kernel(const double* __restrict__ traj, const double* __restrict__ pos0, CellMats m,
                const double* __restrict__ sqrt_mass, int64_t Tl, int N, double dt, int run,
                const double* __restrict__ prev_frame, const double* __restrict__ next_frame,
                double* __restrict__ out_u, double* __restrict__ out_v) {
    int64_t nruns = (Tl + run - 1) / run;
    int64_t total = nruns * N;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // The reference differentiates complex128 arrays, and numpy divides complex by real as
    // (a + b*0) * (1.0 / (c + 0*0)): a multiplication by the rounded reciprocal, reproduced here.
    const double inv_two_dt = __ddiv_rn(1.0, __dmul_rn(2.0, dt));
    const double inv_dt = __ddiv_rn(1.0, dt);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        int atom = (int)(idx % N);
        int64_t t0 = (idx / N) * run;
        int64_t t1 = imin<int64_t>(t0 + run, Tl);
        double p[3], x[3], up[3], uc[3], un[3];
#pragma unroll
        for (int k = 0; k < 3; k++) p[k] = __ldg(pos0 + atom * 3 + k);
        double sm = sqrt_mass ? __ldg(sqrt_mass + atom) : 1.0;
        bool has_prev = true;
        if (t0 > 0) {
            load_pos<NC>(traj, (t0 - 1) * N + atom, x);
            displacement3(x, p, m, up);
        } else if (prev_frame) {
            load_pos<NC>(prev_frame, atom, x);
            displacement3(x, p, m, up);
        } else {
            has_prev = false;
            up[0] = up[1] = up[2] = 0.0;
        }
        load_pos<NC>(traj, t0 * N + atom, x);
        displacement3(x, p, m, uc);
        for (int64_t t = t0; t < t1; t++) {
            bool has_next = true;
            if (t + 1 < Tl) {
                load_pos<NC>(traj, (t + 1) * N + atom, x);
                displacement3(x, p, m, un);
            } else if (next_frame) {
                load_pos<NC>(next_frame, atom, x);
                displacement3(x, p, m, un);
            } else {
                has_next = false;
                un[0] = un[1] = un[2] = 0.0;
            }
            int64_t o = (t * N + atom) * 3;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                double v;
                if (has_prev && has_next) v = __dmul_rn(__dsub_rn(un[k], up[k]), inv_two_dt);
                else if (has_next) v = __dmul_rn(__dsub_rn(un[k], uc[k]), inv_dt);
                else v = __dmul_rn(__dsub_rn(uc[k], up[k]), inv_dt);
                if (out_u) out_u[o + k] = uc[k];
                out_v[o + k] = sqrt_mass ? __dmul_rn(v, sm) : v;
                up[k] = uc[k];
                uc[k] = un[k];
            }
            has_prev = true;
        }
    }
}

__global__ void __launch_bounds__(256)
mass_weight_kernel(const double* __restrict__ v, const double* __restrict__ sqrt_mass, int64_t total,
                   int N, int per_atom, double* __restrict__ out) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        int atom = (int)((idx / per_atom) % N);
        out[idx] = __dmul_rn(__ldg(v + idx), __ldg(sqrt_mass + atom));
    }
}

// (f)3: per-atom moments of the displacement, fused with the minimum-image map: every position is read once
// (24 B per atom*frame) and nothing of size T is written.  Thread = (atom, time split); lanes run over atoms so a
// warp reads 768 contiguous bytes per frame.  part[split][atom][9] = (sum u_k ; sum u_a u_b for a <= b).
template <int NC>
__global__ void __launch_bounds__(256)
displacement_moments_kernel(const double* __restrict__ traj, const double* __restrict__ pos0, CellMats m,
                            int64_t T, int N, int splits, double* __restrict__ part) {
    const int atom = blockIdx.x * blockDim.x + threadIdx.x;
    const int split = blockIdx.y;
    if (atom >= N) return;
    const int64_t per = (T + splits - 1) / splits;
    const int64_t t0 = split * per, t1 = imin<int64_t>(t0 + per, T);
    double p[3], x[3], u[3];
#pragma unroll
    for (int k = 0; k < 3; k++) p[k] = __ldg(pos0 + atom * 3 + k);
    double s[9];
#pragma unroll
    for (int k = 0; k < 9; k++) s[k] = 0.0;
    for (int64_t t = t0; t < t1; t++) {
        load_pos<NC>(traj, t * N + atom, x);
        displacement3(x, p, m, u);
        s[0] += u[0]; s[1] += u[1]; s[2] += u[2];
        s[3] += u[0] * u[0]; s[4] += u[0] * u[1]; s[5] += u[0] * u[2];
        s[6] += u[1] * u[1]; s[7] += u[1] * u[2]; s[8] += u[2] * u[2];
    }
    double* o = part + ((size_t)split * N + atom) * 9;
#pragma unroll
    for (int k = 0; k < 9; k++) o[k] = s[k];
}

// fixed-order sum over the time splits (run-to-run reproducible, no atomics)
__global__ void __launch_bounds__(256)
moments_reduce_kernel(const double* __restrict__ part, int64_t n9, int splits, double* __restrict__ sums) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n9; idx += stride) {
        double a = 0.0;
        for (int sp = 0; sp < splits; sp++) a += part[(size_t)sp * n9 + idx];
        sums[idx] = a;
    }
}

static int moment_splits(int64_t T, int N) {
    // enough (atom, split) threads to fill the machine, runs of at least 16 frames
    int64_t want = ((int64_t)imax(sm_count(), 1) * 2048 + N - 1) / N;
    int64_t cap = imax<int64_t>(1, T / 16);
    return (int)imax<int64_t>(1, imin(imin(want, cap), (int64_t)65535));
}

static int grid_for(int64_t work, int block) {
    int64_t g = (work + block - 1) / block;
    int64_t cap = (int64_t)imax(sm_count(), 1) * 16;
    return (int)imax<int64_t>(1, imin(g, cap));
}

}  // namespace dp

extern "C" {

int dp_atomic_displacements(const double* traj, int traj_is_complex, const double* pos0,
                            const double* cell, int64_t T, int N, double* out, void* stream) {
    DP_REQUIRE(T >= 0 && N > 0, DP_ERR_INVALID, "dp_atomic_displacements: bad sizes T=%lld N=%d", (long long)T, N);
    if (T == 0) return DP_OK;      // empty trajectory: nothing to do (pointers may be null)
    DP_REQUIRE(traj && pos0 && cell && out, DP_ERR_INVALID, "dp_atomic_displacements: null pointer");
    dp::CellMats m;
    dp::make_cell_mats(cell, &m);
    int64_t total = T * N;
    int grid = dp::grid_for(total, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (traj_is_complex) {
        auto k = dp::displacements_kernel<2>;
        DP_LAUNCH(k, dim3(grid), dim3(256), 0, st, traj, pos0, m, total, N, out);
    } else {
        auto k = dp::displacements_kernel<1>;
        DP_LAUNCH(k, dim3(grid), dim3(256), 0, st, traj, pos0, m, total, N, out);
    }
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

int dp_velocity_from_positions(const double* traj, int traj_is_complex, const double* pos0,
                               const double* cell, const double* sqrt_mass, int64_t Tl, int N,
                               double dt, const double* prev_frame, const double* next_frame,
                               double* out_u, double* out_v, void* stream) {
    DP_REQUIRE(traj && pos0 && cell && out_v, DP_ERR_INVALID, "dp_velocity_from_positions: null pointer");
    DP_REQUIRE(Tl > 0 && N > 0 && dt > 0, DP_ERR_INVALID, "dp_velocity_from_positions: bad sizes");
    int64_t frames = Tl + (prev_frame ? 1 : 0) + (next_frame ? 1 : 0);
    DP_REQUIRE(frames >= 2, DP_ERR_INVALID,
               "dp_velocity_from_positions: np.gradient needs at least 2 frames (got %lld)", (long long)frames);
    dp::CellMats m;
    dp::make_cell_mats(cell, &m);
    // runs long enough to amortise the halo frame, short enough to fill the machine
    int run = 32;
    while (run > 4 && ((Tl + run - 1) / run) * N < (int64_t)dp::imax(dp::sm_count(), 1) * 2048) run /= 2;
    int64_t total = ((Tl + run - 1) / run) * N;
    int grid = dp::grid_for(total, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (traj_is_complex) {
        auto k = dp::velocity_kernel<2>;
        DP_LAUNCH(k, dim3(grid), dim3(256), 0, st, traj, pos0, m, sqrt_mass, Tl, N, dt, run, prev_frame,
                  next_frame, out_u, out_v);
    } else {
        auto k = dp::velocity_kernel<1>;
        DP_LAUNCH(k, dim3(grid), dim3(256), 0, st, traj, pos0, m, sqrt_mass, Tl, N, dt, run, prev_frame,
                  next_frame, out_u, out_v);
    }
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

size_t dp_displacement_moments_workspace_bytes(int64_t T, int N) {
    if (T <= 0 || N <= 0) return 0;
    return (size_t)dp::moment_splits(T, N) * (size_t)N * 9 * sizeof(double);
}

int dp_displacement_moments(const double* traj, int traj_is_complex, const double* pos0, const double* cell,
                            int64_t T, int N, double* sums, void* workspace, size_t workspace_bytes, void* stream) {
    DP_REQUIRE(T > 0 && N > 0, DP_ERR_INVALID, "dp_displacement_moments: bad sizes T=%lld N=%d", (long long)T, N);
    DP_REQUIRE(traj && pos0 && cell && sums && workspace, DP_ERR_INVALID, "dp_displacement_moments: null pointer");
    const int splits = dp::moment_splits(T, N);
    const size_t need = (size_t)splits * (size_t)N * 9 * sizeof(double);
    DP_REQUIRE(workspace_bytes >= need, DP_ERR_WORKSPACE, "dp_displacement_moments: workspace %zu < %zu", workspace_bytes, need);
    dp::CellMats m;
    dp::make_cell_mats(cell, &m);
    cudaStream_t st = (cudaStream_t)stream;
    double* part = (double*)workspace;
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)splits);
    if (traj_is_complex) {
        auto k = dp::displacement_moments_kernel<2>;
        DP_LAUNCH(k, grid, dim3(256), 0, st, traj, pos0, m, T, N, splits, part);
    } else {
        auto k = dp::displacement_moments_kernel<1>;
        DP_LAUNCH(k, grid, dim3(256), 0, st, traj, pos0, m, T, N, splits, part);
    }
    DP_CUDA_OK(cudaGetLastError());
    auto kr = dp::moments_reduce_kernel;
    const int64_t n9 = (int64_t)N * 9;
    DP_LAUNCH(kr, dim3(dp::grid_for(n9, 256)), dim3(256), 0, st, (const double*)part, n9, splits, sums);
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

int dp_mass_weight(const double* v, int ncomp, const double* sqrt_mass, int64_t T, int N, double* out,
                   void* stream) {
    DP_REQUIRE((ncomp == 1 || ncomp == 2) && T >= 0 && N > 0, DP_ERR_INVALID, "dp_mass_weight: bad sizes");
    if (T == 0) return DP_OK;
    DP_REQUIRE(v && sqrt_mass && out, DP_ERR_INVALID, "dp_mass_weight: null pointer");
    int per_atom = 3 * ncomp;
    int64_t total = T * N * per_atom;
    auto k = dp::mass_weight_kernel;
    DP_LAUNCH(k, dim3(dp::grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, v, sqrt_mass, total, N,
              per_atom, out);
    DP_CUDA_OK(cudaGetLastError());
    return DP_OK;
}

}  // extern "C"
