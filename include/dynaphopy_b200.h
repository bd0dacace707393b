/*
 * dynaphopy_b200 -- C ABI of the B200-native normal-mode-decomposition hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / numpy types.  Each
 * entry point names the interface of the reference (abelcarreras/DynaPhoPy v1.19.0,
 * paths relative to the reference tree) it replaces.  The reference's native interface
 * is three CPython extension modules (c/correlation.c, c/mem.c, c/displacements.c) plus
 * numpy code in dynaphopy/projection.py and dynaphopy/power_spectrum/__init__.py.
 *
 * Conventions
 *   - Pointers marked [dev] are CUDA device pointers, [host] are host pointers.
 *   - Complex numbers are interleaved (re, im) doubles (numpy complex128 layout).
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*); calls are
 *     asynchronous with respect to the host, re-entrant, and keep no global state except the
 *     per-thread last-error string.  No entry point calls cudaStreamSynchronize: small [host]
 *     tables travel as kernel parameters; a [host] table above 64 KB (dp_project_wave_vector with
 *     Q > 2730 or N > 16384) is staged by one pageable cudaMemcpyAsync, which may wait for earlier
 *     work of the stream before it returns.  Development knobs (DP_* environment variables) are
 *     compiled in only with -DDP_DEVELOP.
 *   - No hidden device allocation: scratch memory is a caller-owned workspace whose
 *     size is returned by the matching *_workspace_bytes function.
 *   - Return value: 0 on success, negative DP_ERR_* on failure; dp_last_error() gives
 *     the text.  There is no CPU fallback anywhere: without a CUDA device every compute
 *     entry point fails with DP_ERR_CUDA.
 */
#ifndef DYNAPHOPY_B200_H
#define DYNAPHOPY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DP_OK 0
#define DP_ERR_INVALID (-1)     /* bad argument                                          */
#define DP_ERR_UNSUPPORTED (-2) /* valid request this build cannot serve (e.g. FFT length
                                   with a prime factor > 61)                              */
#define DP_ERR_WORKSPACE (-3)   /* workspace too small                                   */
#define DP_ERR_CUDA (-4)        /* CUDA runtime error (includes "no device")             */

/* vc_layout of the *_ex projection / contraction entry points */
#define DP_VC_ROWS 0   /* vc[t][q][3*n_prim]        (the reference's (T, n_p, 3) per q; default)            */
#define DP_VC_PAIRS 1  /* vc[t][j][q][2], j = slot pair (2j, 2j+1) of the 3*n_prim (p, k) slots: every       */
                       /* transform of the projection kernel stores one contiguous run -- an INTERMEDIATE    */
                       /* layout between dp_project_commensurate_ex and dp_project_phonon_series_ex          */

/* weight_mode of dp_power_correlation */
#define DP_CORR_WEIGHT_REFERENCE 0 /* as shipped: exp(w*tau + 1i)  (c/correlation.c:158) */
#define DP_CORR_WEIGHT_INTENDED 1  /* exp(i*w*tau)  (comments at c/correlation.c:166-172) */

const char* dp_version(void);
const char* dp_last_error(void);
/* Number of SMs of the current device (host query), or a negative error. */
int dp_sm_count(void);
/* Number of kernels this library has launched in this process so far (monotonic counter). */
long long dp_launch_count(void);

/* ------------------------------------------------------------------------------------
 * A8  minimum-image displacements.
 * Replaces: displacements.atomic_displacements(trajectory, positions, cell)
 *           (c/displacements.c:95-196), called once per atom by
 *           Dynamics.get_relative_trajectory (dynaphopy/dynamics.py:158-179).
 * Here all atoms are processed in one launch.
 *   traj  [dev]  T x N x 3 positions; real doubles, or complex128 when traj_is_complex
 *                (only the real part is used, as in the reference, :160,176)
 *   pos0  [dev]  N x 3 lattice sites (Structure.get_positions(supercell))
 *   cell  [host] 3 x 3, rows = supercell lattice vectors
 *   out   [dev]  T x N x 3 real displacements (the reference returns complex with
 *                imag == 0; the host shim widens when asked to)
 * ------------------------------------------------------------------------------------ */
int dp_atomic_displacements(const double* traj, int traj_is_complex, const double* pos0,
                            const double* cell, int64_t T, int N, double* out, void* stream);

/* ------------------------------------------------------------------------------------
 * A8+A9+A1 fused: positions -> displacement -> np.gradient velocity -> sqrt(mass) weight.
 * Replaces: Dynamics.get_relative_trajectory + Dynamics.velocity +
 *           Dynamics.get_velocity_mass_average (dynaphopy/dynamics.py:158-179, 313-329,
 *           143-156).
 *   traj       [dev] Tl x N x 3 local block of frames (real or complex128)
 *   prev_frame [dev] N x 3 positions of the frame before the block, or NULL when the
 *                    block starts the trajectory (one-sided difference, np.gradient)
 *   next_frame [dev] same for the frame after the block (time-sharded multi-GPU halo)
 *   sqrt_mass  [dev] N, or NULL to skip mass weighting
 *   out_u      [dev] Tl x N x 3 displacements, or NULL
 *   out_v      [dev] Tl x N x 3 velocities (mass weighted when sqrt_mass != NULL)
 * ------------------------------------------------------------------------------------ */
int dp_velocity_from_positions(const double* traj, int traj_is_complex, const double* pos0,
                               const double* cell, const double* sqrt_mass, int64_t Tl, int N,
                               double dt, const double* prev_frame, const double* next_frame,
                               double* out_u, double* out_v, void* stream);

/* ------------------------------------------------------------------------------------
 * Per-atom moments of the minimum-image displacement, fused with A8 (no T-sized output).
 * Replaces: the per-atom np.dot / np.average loops of Dynamics.get_mean_displacement_matrix
 *           (dynaphopy/dynamics.py:207-239) and Dynamics.average_positions (:241-295) over
 *           Dynamics.get_relative_trajectory; the host combines the N x 9 sums per atom type.
 *   traj, pos0, cell: as dp_atomic_displacements
 *   sums  [dev]  N x 9: sum_t u_x, u_y, u_z, u_x u_x, u_x u_y, u_x u_z, u_y u_y, u_y u_z, u_z u_z
 *   workspace [dev] dp_displacement_moments_workspace_bytes(T, N)
 * ------------------------------------------------------------------------------------ */
size_t dp_displacement_moments_workspace_bytes(int64_t T, int N);
int dp_displacement_moments(const double* traj, int traj_is_complex, const double* pos0,
                            const double* cell, int64_t T, int N, double* sums, void* workspace,
                            size_t workspace_bytes, void* stream);

/* A1 stand-alone: out = v * sqrt_mass[atom]  (dynaphopy/dynamics.py:143-156).
 * v/out: T x N x 3, real (ncomp = 1) or complex128 (ncomp = 2).                         */
int dp_mass_weight(const double* v, int ncomp, const double* sqrt_mass, int64_t T, int N,
                   double* out, void* stream);

/* ------------------------------------------------------------------------------------
 * A2  wave-vector projection, dense form, batched over q (arbitrary q, arbitrary sites).
 * Replaces: projection.project_onto_wave_vector(trajectory, q_vector, project_on_atom)
 *           (dynaphopy/projection.py:4-40), one launch for Q wave vectors.
 *   vc[t,q,p,k] = (N/n_prim)^-1/2 * sum_{i: type(i)=p} sqrt_mass[i] v[t,i,k] exp(-i q.r_i)
 *   v         [dev]  T x N x 3, real or complex128 (v_is_complex)
 *   sqrt_mass [dev]  N, or NULL when v is already mass weighted
 *   pos0      [dev]  N x 3 equilibrium positions
 *   type_idx  [host] N primitive-atom index of every supercell atom
 *   q_cart    [host] Q x 3 Cartesian wave vectors (Quasiparticle.get_q_vector)
 *   project_on_atom  -1, or a primitive atom index: other atoms are skipped while the
 *                    normalisation is unchanged (projection.py:26-28,38-39)
 *   out_vc    [dev]  T x Q x n_prim x 3 complex128
 *   workspace [dev]  dp_project_wave_vector_workspace_bytes(N, Q) bytes
 * ------------------------------------------------------------------------------------ */
size_t dp_project_wave_vector_workspace_bytes(int N, int Q);
int dp_project_wave_vector(const double* v, int v_is_complex, const double* sqrt_mass,
                           const double* pos0, const int* type_idx, int N, int n_prim, int64_t T,
                           const double* q_cart, int Q, int project_on_atom, double* out_vc,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * A2  wave-vector projection, commensurate form: for q commensurate with a diagonal
 * supercell the lattice sum is a 3-D DFT over unit-cell translations (atom order of
 * dynaphopy/atoms.py:139-156: i = u*Nc + x + Nx*(y + Ny*z)), so every frame is read from
 * HBM once for the whole q list.
 *   v         [dev]  T x N x 3 real velocities, N = n_unit * dims[0]*dims[1]*dims[2]
 *   dims      [host] 3 supercell multiplicities (Dynamics.get_supercell_matrix)
 *   unit_type [host] n_unit primitive-atom index of every unit-cell atom
 *   qbin      [dev]  Q x 3 int32 DFT bin of every q: round(q_unitcell * dims) mod dims.  The bins live on the
 *                    device and are checked there (the call never waits for the device): a bin outside
 *                    [0, dims) makes every out_vc entry of that q NaN
 *   twiddle   [dev]  Q x n_unit complex128: (N/n_prim)^-1/2 sqrt(mass_u) exp(-i q.tau_u); or NULL (only when
 *                    n_unit == n_prim): out_vc[t,q,u,k] is the bare lattice DFT of unit atom u, and the caller
 *                    folds the twiddle into the eigenvectors of the contraction that follows
 *                    (e'[q,s,u,k] = e[q,s,u,k] * conj(twiddle[q,u]) gives the same vq)
 *   out_vc    [dev]  T x Q x n_prim x 3 complex128
 * ------------------------------------------------------------------------------------ */
size_t dp_project_commensurate_workspace_bytes(const int* dims, int n_unit, int n_prim, int Q);
int dp_project_commensurate(const double* v, const int* dims, int n_unit, const int* unit_type,
                            int n_prim, int64_t T, const int* qbin, const double* twiddle, int Q,
                            int project_on_atom, double* out_vc, void* workspace,
                            size_t workspace_bytes, void* stream);

int dp_project_commensurate_ex(const double* v, const int* dims, int n_unit, const int* unit_type,
                               int n_prim, int64_t T, const int* qbin, const double* twiddle, int Q,
                               int project_on_atom, int vc_layout, double* out_vc, void* workspace,
                               size_t workspace_bytes, void* stream);
/* The pair-major bare lattice DFT (twiddle == NULL, DP_VC_PAIRS) written as n_blocks = Q / q_block separate arrays, one
 * per block of q_block consecutive wave vectors of the list: block_out[g][((t * (3 n_prim / 2) + j) * q_block + qq) * 2 + c]
 * for q = g * q_block + qq.  block_out is a HOST array of device pointers (<= 16), e.g. the send slots / the receive
 * buffer of the multi-GPU exchange: each destination rank receives the intermediate of ITS wave vectors and contracts
 * it there (half the exchange volume of the contracted series when the list holds one q per pair (q, -q)). */
int dp_project_commensurate_blocks(const double* v, const int* dims, int n_unit, const int* unit_type, int n_prim,
                                   int64_t T, const int* qbin, int Q, int q_block, double* const* block_out, int n_blocks,
                                   void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * A3  eigenvector contraction, batched over q.
 * Replaces: projection.project_onto_phonon(vc, eigenvectors) (dynaphopy/projection.py:43-54)
 *   vq[t,q,s] = sum_{p,k} vc[t,q,p,k] * conj(e[q,s,p,k])
 *   vc  [dev] T x Q x M complex128 (M = 3*n_prim, (p,k) flattened)
 *   eig [dev] Q x M x M complex128, e[q][s][3p+k] (phonopy_link.py:189-192 arrangement)
 *   vq  [dev] T x Q x M complex128 (may alias vc only if M*16 bytes rows are private: no)
 * ------------------------------------------------------------------------------------ */
int dp_project_phonon(const double* vc, const double* eig, int64_t T, int Q, int M, double* vq,
                      void* stream);
/* Same contraction, output grouped by blocks of q_block wave vectors:
 *   vq[(q / q_block)][t][q % q_block][s]   (q_block must divide Q; q_block == Q is the layout above).
 * This is the send layout of the time-sharded -> series-sharded all-to-all: block g is the
 * contiguous message for rank g, and concatenating the received blocks in rank order yields a
 * time-major (T_total, q_block*M) series matrix on every rank.
 * vq may alias vc when q_block == Q. */
int dp_project_phonon_blocked(const double* vc, const double* eig, int64_t T, int Q, int M, int q_block,
                              double* vq, void* stream);

/* Same contraction, SERIES-MAJOR output: one contiguous time series per (q, s) -- the layout the spectrum
 * kernels stream at full HBM bandwidth -- grouped by blocks of q_block wave vectors (one block per
 * destination rank of the all-to-all, q_block == Q: a plain (Q*M, T_total) matrix):
 *   out[((q / q_block)*q_block*M + (q % q_block)*M + s) * T_total + t_offset + t],   t < T.
 * t_offset / T_total let a long trajectory be contracted chunk by chunk into one series matrix. */
int dp_project_phonon_series(const double* vc, const double* eig, int64_t T, int Q, int M, int q_block,
                             int64_t t_offset, int64_t T_total, double* out, void* stream);
int dp_project_phonon_series_ex(const double* vc, int vc_layout, const double* eig, int64_t T, int Q, int M,
                                int q_block, int64_t t_offset, int64_t T_total, double* out, void* stream);

/* Contraction FUSED with the exchange step of the multi-GPU path: block g = q / q_block is
 * written at block_out[g] + ((q % q_block)*M + s) * T_total + t_offset + t instead of into one array.  The pointers
 * may be peer device memory of the other ranks of an NVLink/NVSwitch domain (mapped with CUDA IPC / symmetric
 * memory): the kernel's 512-byte runs then go straight into the destination rank's receive buffer and no
 * all-to-all is needed.  block_out is a HOST array of n_blocks = Q / q_block (<= 16) device pointers.  The caller
 * orders the accesses across ranks (a barrier before reuse of the receive buffers and after the last chunk). */
int dp_project_phonon_series_peers(const double* vc, int vc_layout, const double* eig, int64_t T, int Q, int M,
                                   int q_block, int64_t t_offset, int64_t T_total, double* const* block_out,
                                   int n_blocks, void* stream);

/* Contraction of a HALF mesh (projection.py:43-54 on top of the Hermitian symmetry of projection.py:24-37 for real
 * velocities: without the per-atom twiddle, vc[t, -q, :] = conj(vc[t, q, :])).  vc is the pair-major intermediate
 * (DP_VC_PAIRS) of dp_project_commensurate_ex with twiddle == NULL for Q_stored wave vectors h, one of every pair
 * (q, -q) of the full list of Q wave vectors:
 *   qmap [dev] Q_stored x 2 ints: index (in the full list) of the stored wave vector and of its partner -q
 *              (< 0: none -- a self-conjugate q, or -q is not in the list)
 *   eig2 [dev] Q_stored x 2 x M x M complex128: the eigenvectors of qmap[h][0], then the CONJUGATED eigenvectors of
 *              qmap[h][1] (anything when there is no partner)
 * Output as dp_project_phonon_series_ex / _peers (out, or block_out != NULL with n_blocks = Q / q_block pointers), for
 * all Q wave vectors, bit-identical to contracting an explicitly projected full mesh: the intermediate and the
 * projection's stores are half the size. */
int dp_project_phonon_series_mirror(const double* vc, const double* eig2, const int* qmap, int64_t T, int Q_stored,
                                    int Q, int M, int q_block, int64_t t_offset, int64_t T_total, double* out,
                                    double* const* block_out, int n_blocks, void* stream);

/* ------------------------------------------------------------------------------------
 * A7  power spectrum, FFT method (selector keys 2 and 3).
 * Replaces: get_fft_numpy_spectra / _numpy_power / _division_of_data
 *           (dynaphopy/power_spectrum/__init__.py:226-270, 33-51), all columns at once.
 *   series [dev]  T x ld complex128, column s of row t at series[(t*ld + s)*2]
 *                 (the (T, M') array the reference passes; ld >= S)
 *   freq   [host] F requested frequencies (parameters.frequency_range), F >= 2
 *   scale         multiplied into the result (the reference's unit_conversion, :9)
 *   out    [dev]  F x S doubles, row-major (the reference's (F, M') return layout)
 * dp_power_fft_pieces reports the window list for inspection (host only).
 * ------------------------------------------------------------------------------------ */
size_t dp_power_fft_workspace_bytes(int64_t T, int S, double dt, const double* freq, int F);
int dp_power_fft(const double* series, int64_t T, int S, int64_t ld, double dt, const double* freq,
                 int F, double scale, double* out, void* workspace, size_t workspace_bytes,
                 void* stream);
/* Same spectrum for series stored with arbitrary strides (in complex128 elements): sample t of series s
 * is at  series[s*stride_s + (t / tb_len)*tb_stride + (t % tb_len)*stride_t]  when tb_len > 0 (the time axis
 * arrives in blocks, e.g. one block per source rank after the time-sharded -> series-sharded all-to-all,
 * SURVEY.md section 8e), or  series[s*stride_s + t*stride_t]  when tb_len == 0.  dp_power_fft is the
 * case stride_s = 1, stride_t = ld.  The series-major layout (stride_t = 1) is the one the kernels stream
 * at full HBM bandwidth. */
int dp_power_fft_strided(const double* series, int64_t T, int S, int64_t stride_s, int64_t stride_t,
                         int64_t tb_len, int64_t tb_stride, double dt, const double* freq, int F, double scale,
                         double* out, void* workspace, size_t workspace_bytes, void* stream);
/* The same spectrum in two steps, for callers that produce the series incrementally (e.g. a host-resident
 * trajectory streamed to the device): dp_power_fft_windows transforms the distinct windows
 * [first_window, first_window + n_windows) (n_windows < 0: all remaining; see dp_power_fft_pieces for the
 * list) into the workspace and reads no sample outside them, so it can run as soon as those frames exist;
 * dp_power_fft_finish averages the windows and interpolates onto the grid once all of them are done.  Every
 * call of one spectrum must use the same workspace and the same (T, S, dt, freq, F). */
int dp_power_fft_windows(const double* series, int64_t T, int S, int64_t stride_s, int64_t stride_t,
                         int64_t tb_len, int64_t tb_stride, double dt, const double* freq, int F, int first_window,
                         int n_windows, void* workspace, size_t workspace_bytes, void* stream);
/* max_ctas > 0: the persistent kernel occupies at most that many SMs (it needs a whole SM per CTA), so that kernels of
 * other streams -- the projection / contraction of the frames still to come, which feed the multi-GPU exchange -- keep
 * running beside it; 0: all SMs. */
int dp_power_fft_windows_ex(const double* series, int64_t T, int S, int64_t stride_s, int64_t stride_t,
                            int64_t tb_len, int64_t tb_stride, double dt, const double* freq, int F, int first_window,
                            int n_windows, int max_ctas, void* workspace, size_t workspace_bytes, void* stream);
/* As dp_power_fft_windows_ex, PREEMPTIBLE form for pipelines that run urgent kernels on higher-priority streams (the
 * multi-GPU pipeline: the projection / contraction kernels that feed and drain the exchange): the slab engine launches
 * one short-lived CTA per (series, window pair) item instead of one persistent CTA per SM, so the block scheduler places
 * the waiting high-priority CTAs whenever an item ends (~0.15 ms) and the spectrum CTAs take every SM back afterwards.
 * Each CTA borrows the slab of the SM it runs on (free list in the workspace).  Same results; 3 % slower than the
 * persistent form on an otherwise idle GPU (tools/preempt_overhead.py).  Other engines: as dp_power_fft_windows_ex. */
int dp_power_fft_windows_preemptible(const double* series, int64_t T, int S, int64_t stride_s, int64_t stride_t,
                                     int64_t tb_len, int64_t tb_stride, double dt, const double* freq, int F,
                                     int first_window, int n_windows, int max_ctas, void* workspace,
                                     size_t workspace_bytes, void* stream);
int dp_power_fft_finish(int64_t T, int S, double dt, const double* freq, int F, double scale, double* out,
                        void* workspace, size_t workspace_bytes, void* stream);
int dp_power_fft_pieces(int64_t T, double dt, const double* freq, int F, int* piece_len,
                        int* n_pieces, int* starts, int max_pieces);
/* Which transform engine serves this (T, dt, freq) (host query, for tests and benchmarks): 2 = one persistent CTA per
 * series and window pair over an L2-resident slab (padded window length <= 65536), 3 = batched three-factor passes
 * (longer windows, e.g. cfg 5-B's single 2^17-sample window, or wide kept-bin ranges), 1 = the general mixed-radix
 * engine (window lengths without a usable factorisation); < 0: the error code dp_power_fft would return. */
int dp_power_fft_engine(int64_t T, double dt, const double* freq, int F);

/* ------------------------------------------------------------------------------------
 * A6  power spectrum, maximum-entropy (Burg) method (selector key 1).
 * Replaces: get_mem_power_spectra -> mem.mem(frequency, velocity, time_step, coefficients)
 *           (dynaphopy/power_spectrum/__init__.py:81-103, c/mem.c:102-252) with the
 *           canonical semantics Data[N] := 0 for the reference's out-of-bounds read.
 *   nan_to_num != 0 applies np.nan_to_num (power_spectrum/__init__.py:101).
 * ------------------------------------------------------------------------------------ */
size_t dp_power_mem_workspace_bytes(int64_t T, int S, int F, int coefficients);
int dp_power_mem(const double* series, int64_t T, int S, int64_t ld, double dt, const double* freq,
                 int F, int coefficients, double scale, int nan_to_num, double* out,
                 void* workspace, size_t workspace_bytes, void* stream);
/* Same for series stored with arbitrary strides / a blocked time axis (see dp_power_fft_strided): the series-major
 * output of dp_project_phonon_series and the (source rank, chunk) time blocks of the multi-GPU exchange are read in
 * place.  dp_power_mem is the case stride_s = 1, stride_t = ld. */
int dp_power_mem_strided(const double* series, int64_t T, int S, int64_t stride_s, int64_t stride_t, int64_t tb_len,
                         int64_t tb_stride, double dt, const double* freq, int F, int coefficients, double scale,
                         int nan_to_num, double* out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * A5  power spectrum, direct correlation method (selector key 0).
 * Replaces: get_fourier_direct_power_spectra -> correlation.correlation_par(frequency,
 *           velocity, time_step, step, integration_method)
 *           (dynaphopy/power_spectrum/__init__.py:57-75, c/correlation.c:101-178).
 *   weight_mode DP_CORR_WEIGHT_REFERENCE reproduces the shipped real-exponential weights,
 *   DP_CORR_WEIGHT_INTENDED uses exp(i w tau).  integration_method outside {0,1} fails
 *   (the reference returns NULL without setting an exception, c/correlation.c:138-141).
 * ------------------------------------------------------------------------------------ */
size_t dp_power_correlation_workspace_bytes(int64_t T, int S, int F, int step);
int dp_power_correlation(const double* series, int64_t T, int S, int64_t ld, double dt,
                         const double* freq, int F, int step, int integration_method,
                         int weight_mode, double scale, double* out, void* workspace,
                         size_t workspace_bytes, void* stream);
/* Same for strided series / a blocked time axis (see dp_power_fft_strided); dp_power_correlation is the case
 * stride_s = 1, stride_t = ld. */
int dp_power_correlation_strided(const double* series, int64_t T, int S, int64_t stride_s, int64_t stride_t,
                                 int64_t tb_len, int64_t tb_stride, double dt, const double* freq, int F, int step,
                                 int integration_method, int weight_mode, double scale, double* out, void* workspace,
                                 size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Utility: batched complex FFT on device data with the library's own radix engine
 * (the transform used inside dp_power_fft; exposed for validation and benchmarking).
 *   in/out [dev] batch x n complex128 (contiguous rows), direction -1 forward, +1 inverse
 *   (unnormalised).  n must factor into primes <= 61.
 * ------------------------------------------------------------------------------------ */
size_t dp_fft_c2c_workspace_bytes(int n, int batch);
int dp_fft_c2c(const double* in, double* out, int n, int batch, int direction, void* workspace,
               size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Hand-off to the fitting stage: degenerate-set average + initial guess for every series.
 * Replaces: degenerate_sets / average_phonon and the np.max / np.argmax guess of
 *           phonon_fitting_analysis (dynaphopy/analysis/fitting/__init__.py:17-39, 60-61),
 *           run per mode inside the q loop of get_commensurate_points_data
 *           (dynaphopy/__init__.py:1130-1174).  Lorentzian fitting stays in scipy.
 *   spectra     [dev] F x S (the selector's output layout)
 *   set_ptr     [dev] n_sets + 1 offsets into set_members, or NULL (no averaging)
 *   set_members [dev] series indices grouped by degenerate set (reference order), or NULL
 *   series_set  [dev] S: set of each series, or NULL
 *   averaged    [dev] F x S averaged spectra, or NULL
 *   peak_bin    [dev] S   np.argmax of the (averaged) spectrum
 *   peak_height [dev] S   np.max
 * ------------------------------------------------------------------------------------ */
int dp_spectra_peaks(const double* spectra, int F, int S, const int* set_ptr, const int* set_members,
                     const int* series_set, int n_sets, double* averaged, int* peak_bin,
                     double* peak_height, void* stream);

/* ------------------------------------------------------------------------------------
 * Trajectory ingest (host only): LAMMPS text dump -> (T, N, ndim) doubles.
 * Replaces: the per-line Python parsing of read_lammps_trajectory
 *           (dynaphopy/interface/iofile/trajectory_parsers.py:135-352).  Same conventions: frames
 *           counted at every "TIMESTEP" marker, the first ndim columns of every atom line, cell
 *           and atom count from the first BOX BOUNDS / NUMBER OF ATOMS blocks.
 * dp_lammps_scan: n_frames (complete frames), n_atoms, bounds[3][3] row-major (lo, hi, tilt; only
 *   bound_cols = 2 or 3 columns are present), labels = the last "ITEM: ATOMS ..." line
 *   (position or velocity dump is decided from it: ' x y' / ' vx vy', :335-349).
 * dp_lammps_read: frames [first_frame, first_frame + n_frames) (0-based) into `out` [host,
 *   caller-owned, e.g. pinned], their TIMESTEP values into `timesteps` (or NULL); n_threads <= 0:
 *   all host threads.  Conversion with strtod (correctly rounded, as Python's float()).
 * ------------------------------------------------------------------------------------ */
int dp_lammps_scan(const char* path, int64_t* n_frames, int* n_atoms, double* bounds, int* bound_cols,
                   char* labels, int labels_len);
int dp_lammps_read(const char* path, int64_t first_frame, int64_t n_frames, int n_atoms, int ndim,
                   double* out, double* timesteps, int n_threads);

/* Same for VASP XDATCAR files (read_VASP_XDATCAR, trajectory_parsers.py:355-483): header cell (3 x 3, the scale
 * line is ignored like in the reference), atom count = sum of the counts line, frames at every
 * "Direct configuration= k" marker, fractional coordinates.  `steps` receives the configuration numbers of
 * frames [0, n_steps) -- the reference keeps the times of the frames it skips. */
int dp_xdatcar_scan(const char* path, int64_t* n_frames, int* n_atoms, double* cell);
int dp_xdatcar_read(const char* path, int64_t first_frame, int64_t n_frames, int n_atoms, double* out,
                    double* steps, int64_t n_steps, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* DYNAPHOPY_B200_H */
