/* TEST INFRASTRUCTURE -- CPU oracle (plain-C restatement), never the product.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the dynaphopy_b200 package never does.
 *
 * Each function restates, with its own loop structure and 0-based indexing, the
 * arithmetic of one reference routine (citations are into /root/reference):
 *
 *   op_burg_part        c/mem.c:199-252   GetCoefficients (canonical Data[N] := 0)
 *   op_mem_spectrum     c/mem.c:102-196   MaximumEntropyMethod + FrequencyEvaluation
 *   op_correlation      c/correlation.c:101-178 correlation_par / EvaluateCorrelation
 *   op_displacements    c/displacements.c:95-196,222-265 atomic_displacements
 *   op_gradient         dynaphopy/dynamics.py:313-329 (np.gradient, edge_order=1)
 *   op_project_q        dynaphopy/projection.py:4-40 project_onto_wave_vector
 *
 * The oracle is pinned against the compiled reference itself (oracle/_ref, built by
 * oracle/Makefile from the unmodified sources) in tests/test_oracle_pinning.py.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---- Burg linear-prediction coefficients ---------------------------------------
 * x[0..n): one real series.  The reference walks Data[1..n] on a 0-based buffer, so
 * sample x[0] is dropped and a virtual sample x[n] = 0 is appended (c/mem.c:210-219).
 * d[0..m): coefficient d_k stored at d[k-1].  Returns the residual power xms, or 0.0
 * when the |denominator| < 1e-6 early-out fires (c/mem.c:229).                      */
double op_burg_part(const double *x, int n, int m, double *d)
{
    int len = n - 1;                       /* number of (forward, backward) pairs   */
    double *f = (double *)malloc(sizeof(double) * (size_t)(len > 0 ? len : 1));
    double *b = (double *)malloc(sizeof(double) * (size_t)(len > 0 ? len : 1));
    double *prev = (double *)calloc((size_t)(m > 0 ? m : 1), sizeof(double));
    double p = 0.0;
    for (int j = 1; j <= n; j++) {
        double s = (j < n) ? x[j] : 0.0;
        p += pow(s, 2);
    }
    double xms = p / n;
    for (int j = 0; j < len; j++) {
        f[j] = x[j + 1];
        b[j] = (j + 2 < n) ? x[j + 2] : 0.0;
    }
    for (int k = 1; k <= m; k++) {
        double num = 0.0, den = 0.0;
        int cnt = n - k;                   /* pairs j = 0 .. n-k-1                  */
        for (int j = 0; j < cnt; j++) {
            num += f[j] * b[j];
            den += pow(f[j], 2) + pow(b[j], 2);
        }
        if (fabs(den) < 1.0e-6) { xms = 0.0; break; }
        double c = 2.0 * num / den;
        d[k - 1] = c;
        xms *= (1.0 - pow(c, 2));
        for (int i = 1; i <= k - 1; i++) d[i - 1] = prev[i - 1] - c * prev[k - i - 1];
        if (k == m) break;
        for (int i = 0; i < k; i++) prev[i] = d[i];
        for (int j = 0; j < cnt - 1; j++) {
            f[j] -= c * b[j];
            b[j] = b[j + 1] - c * f[j + 1];    /* f[j+1] not yet updated            */
        }
    }
    free(f); free(b); free(prev);
    return xms;
}

/* P(f) = dt * sum_parts xms / |1 - sum_k d_k z^k|^2, z = exp(i 2 pi f dt)
 * (c/mem.c:155-168,184-196).  vel: n complex samples interleaved (re,im).          */
void op_mem_spectrum(const double *freq, int nfreq, const double *vel, int n,
                     double dt, int m, double *out)
{
    double *part = (double *)malloc(sizeof(double) * (size_t)n);
    double *d[2];
    double xms[2];
    for (int c = 0; c < 2; c++) {
        d[c] = (double *)calloc((size_t)m + 1, sizeof(double));
        for (int j = 0; j < n; j++) part[j] = vel[2 * j + c];
        xms[c] = op_burg_part(part, n, m, d[c]);
    }
#pragma omp parallel for
    for (int i = 0; i < nfreq; i++) {
        double delta = freq[i] * 2.0 * M_PI * dt;
        double _Complex z = cexp(delta * 1.0j);
        double acc = 0.0;
        for (int c = 0; c < 2; c++) {
            if (xms[c] == 0.0) continue;
            double _Complex a = 1.0;
            for (int k = 1; k <= m; k++) {
                double _Complex prod = d[c][k - 1] * cpow(z, (double)k);
                a = (creal(a) - creal(prod)) + (cimag(a) - cimag(prod)) * 1.0j;
            }
            acc += xms[c] / creal(a * conj(a));
        }
        out[i] = acc * dt;
    }
    free(part); free(d[0]); free(d[1]);
}

/* Direct correlation spectrum (c/correlation.c:153-178).
 * weight_mode 0: as shipped, weight(tau) = exp(w*tau + 1i)   (real exponential!)
 * weight_mode 1: intended,   weight(tau) = exp(i*w*tau)      (comments :166-172)
 * method 1 (rectangular) adds each lag product twice; method 0 adds three terms, one
 * of them with an argument missing the time step (c/correlation.c:162-164).
 * The accumulation order (lag-major, origin-minor, one running complex sum) is kept. */
static double _Complex op_weight(double arg, int weight_mode)
{
    return weight_mode ? cexp(arg * 1.0j) : cexp(arg + 1.0j);
}

void op_correlation(const double *freq, int nfreq, const double *vel, int n, double dt,
                    int step, int method, int weight_mode, double *out)
{
    const double _Complex *v = (const double _Complex *)vel;
#pragma omp parallel for
    for (int fi = 0; fi < nfreq; fi++) {
        double w = freq[fi] * 2.0 * M_PI;
        double _Complex acc = 0.0;
        for (int lag = 0; lag < n; lag += step) {
            double _Complex w0 = op_weight(w * (lag * dt), weight_mode);
            double _Complex w1 = op_weight(w * ((lag + step) * dt), weight_mode);
            double _Complex w2 = op_weight(w * (double)(lag + step), weight_mode);
            for (int j = 0; j < n - lag - step; j++) {
                double _Complex t = (conj(v[j]) * v[j + lag]) * w0;
                acc = (creal(acc) + creal(t)) + (cimag(acc) + cimag(t)) * 1.0j;
                if (method == 0) {
                    double _Complex ti = (conj(v[j]) * v[j + lag + step]) * w1;
                    double _Complex tj = (conj(v[j]) * v[j + lag]) * w2;
                    acc = (creal(acc) + creal(ti) + creal(tj))
                        + (cimag(acc) + cimag(ti) + cimag(tj)) * 1.0j;
                } else {
                    acc = (creal(acc) + creal(t)) + (cimag(acc) + cimag(t)) * 1.0j;
                }
            }
        }
        out[fi] = creal(acc) * dt / (n / step);      /* integer division, :177      */
    }
}

/* ---- minimum-image displacements (c/displacements.c:158-183) -------------------
 * traj: T x 3 real positions of ONE atom; pos0: its lattice site; cell: 3x3 rows =
 * supercell vectors.  C = cell^T (c/displacements.c:222-248); C^-1 by cofactors with
 * first-row determinant expansion (:250-335).  out: T x 3 real.                      */
static double det2(double a, double b, double c, double d) { return a * d - c * b; }

static double det3_first_row(const double m[3][3])
{
    /* expansion by minors along row 0, j1 = 0,1,2, sign (-1)^(j1+2) (:291)          */
    double det = 0.0;
    for (int j1 = 0; j1 < 3; j1++) {
        double mm[2][2];
        for (int i = 1; i < 3; i++) {
            int j2 = 0;
            for (int j = 0; j < 3; j++) {
                if (j == j1) continue;
                mm[i - 1][j2++] = m[i][j];
            }
        }
        det += pow(-1.0, j1 + 2.0) * m[0][j1] * det2(mm[0][0], mm[0][1], mm[1][0], mm[1][1]);
    }
    return det;
}

void op_cell_inverse(const double *cell, double *cmat, double *cinv)
{
    double a[3][3], cof[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) a[i][j] = cell[j * 3 + i];       /* transpose   */
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
            cof[i][j] = pow(-1.0, i + j + 2.0) * det2(c[0][0], c[0][1], c[1][0], c[1][1]);
        }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            cmat[i * 3 + j] = a[i][j];
            cinv[i * 3 + j] = cof[j][i] / det3_first_row(a);
        }
}

void op_displacements(const double *traj, int nt, const double *pos0, const double *cell,
                      double *out)
{
    double cm[9], ci[9];
    op_cell_inverse(cell, cm, ci);
    for (int t = 0; t < nt; t++) {
        double diff[3], frac[3], shift[3];
        for (int k = 0; k < 3; k++) diff[k] = traj[t * 3 + k] - pos0[k];
        for (int i = 0; i < 3; i++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += ci[i * 3 + k] * diff[k];
            frac[i] = round(s);
        }
        for (int i = 0; i < 3; i++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += cm[i * 3 + k] * frac[k];
            shift[i] = s;
        }
        for (int k = 0; k < 3; k++) out[t * 3 + k] = traj[t * 3 + k] - shift[k] - pos0[k];
    }
}

/* np.gradient(u, dt) along time, 2nd-order interior, 1st-order ends (dynamics.py:324-327) */
void op_gradient(const double *u, int nt, int stride, double dt, double *out)
{
    if (nt < 2) return;
    out[0] = (u[stride] - u[0]) / dt;
    for (int t = 1; t < nt - 1; t++)
        out[(size_t)t * stride] = (u[(size_t)(t + 1) * stride] - u[(size_t)(t - 1) * stride]) / (2.0 * dt);
    out[(size_t)(nt - 1) * stride] = (u[(size_t)(nt - 1) * stride] - u[(size_t)(nt - 2) * stride]) / dt;
}

/* project_onto_wave_vector for one q (projection.py:24-39).  vm: T x N x 3 real
 * mass-weighted velocities; out: T x n_prim x 3 complex (interleaved re,im).         */
void op_project_q(const double *vm, int nt, int natoms, const double *pos0, const int *type_idx,
                  int nprim, const double *q, int project_on_atom, double *out)
{
    memset(out, 0, sizeof(double) * 2 * (size_t)nt * nprim * 3);
    for (int i = 0; i < natoms; i++) {
        if (project_on_atom > -1 && type_idx[i] != project_on_atom) continue;
        double ph = q[0] * pos0[i * 3] + q[1] * pos0[i * 3 + 1] + q[2] * pos0[i * 3 + 2];
        double _Complex e = cexp(-ph * 1.0j);
        for (int t = 0; t < nt; t++)
            for (int k = 0; k < 3; k++) {
                double _Complex c = vm[((size_t)t * natoms + i) * 3 + k] * e;
                size_t o = (((size_t)t * nprim + type_idx[i]) * 3 + k) * 2;
                out[o] += creal(c);
                out[o + 1] += cimag(c);
            }
    }
    double norm = sqrt((double)natoms / nprim);
    for (size_t o = 0; o < 2 * (size_t)nt * nprim * 3; o++) out[o] /= norm;
}
