/*
 * shenb200.h -- C ABI of libshenb200.so: sm_100a kernels for shenfun's
 * tensor-product transform-and-solve hot path.
 *
 * Conventions
 *  - every pointer named d_* is a DEVICE pointer owned by the caller
 *    (torch tensors on the Python side); the library owns only twiddle /
 *    coefficient tables that it caches per device.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - all functions return 0 on success, a negative shb_status otherwise;
 *    shb_last_error() gives a human readable message for the calling thread.
 *  - arrays are C-ordered, fp64 (`double`) or complex128 (`double[2]`).
 *    An N-D array transformed along `axis` is passed collapsed to 3-D as
 *    [outer][n][inner] (outer = prod(shape[:axis]), inner = prod(shape[axis+1:])).
 *
 * Each entry point names the reference interface it stands in for
 * (paths relative to the reference checkout, spectralDNS/shenfun v4.3.0).
 */
#ifndef SHENB200_H
#define SHENB200_H

#include <stdint.h>

#if defined(__GNUC__)
#define SHB_API __attribute__((visibility("default")))
#else
#define SHB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    SHB_OK = 0,
    SHB_ERR_ARG = -1,      /* invalid argument */
    SHB_ERR_CUDA = -2,     /* CUDA runtime error (see shb_last_error) */
    SHB_ERR_UNSUPPORTED = -3,
    SHB_ERR_NOMEM = -4
} shb_status;

SHB_API const char *shb_last_error(void);
SHB_API int shb_version(void);
/* 1 when a CUDA device of compute capability 10.x is usable, else 0 */
SHB_API int shb_device_ok(void);

/* ------------------------------------------------------------------ */
/* Fourier axis transforms (replace mpi4py_fft.fftw.{fftn,ifftn,rfftn,irfftn}
 * planned in shenfun/fourier/bases.py:218-279 and the pad/truncate/normalise
 * around them: FourierBase.forward :131-149, R2C._truncation_forward /
 * _padding_backward :409-435, C2C :570-598).                              */

/* Forward complex DFT of length n_fft along the middle axis, then 3/2-rule
 * truncation to N (N <= n_fft; N == n_fft means no padding) and scaling:
 *   out[o][k][c] = scale * trunc_N( sum_j in[o][j][c] exp(-2 pi i j k / n_fft) )
 * in : complex [outer][n_fft][inner];  out : complex [outer][N][inner].
 * in == out is allowed when N == n_fft.                                   */
SHB_API int shb_fft_c2c_forward(const void *d_in, void *d_out, int64_t outer, int64_t inner,
                        int n_fft, int N, double scale, void *stream);

/* Zero-pad N -> n_fft by the C2C rule (both +-N/2 entries halved when N is
 * even and N < n_fft), then un-normalised inverse DFT of length n_fft.
 * in : complex [outer][N][inner];  out : complex [outer][n_fft][inner].
 * flags bit 0: dealias_direct (2/3-rule zeroing, only with N == n_fft).   */
SHB_API int shb_fft_c2c_backward(const void *d_in, void *d_out, int64_t outer, int64_t inner,
                         int n_fft, int N, double scale, int flags, void *stream);

/* Complex-to-real backward transform of the three components of a vector field fused with the
 * cross product:  out = irfft_pad(in) x b  on [3][nlines][.] arrays (in: N//2+1 complex modes per
 * line, b / out: n_fft reals per line).  The velocity never reaches HBM in physical space.
 * Replaces `U = U_hat.backward(U); W[:] = np.cross(U, curl_, axis=0)` of demo/NavierStokes.py:39-40
 * (utilities.cross, optimization/cython/cross.pyx:25-40).  flags bit 0: dealias_direct; bit 1:
 * `d_b` is itself a SPECTRAL vector field with the layout of `d_in` and is transformed in the same
 * kernel (out = irfft_pad(in) x irfft_pad(b), `scale` applies to `in` only) -- then neither
 * physical field reaches HBM: the whole of NavierStokes.py:35-41 after the x/y passes.
 * SHB_ERR_UNSUPPORTED when no register-resident kernel covers n_fft (callers then run c2r and
 * shb_cross3d separately). */
SHB_API int shb_fft_c2r_cross(const void *d_in, const void *d_b, void *d_out, int64_t nlines, int n_fft,
                              int N, double scale, int flags, void *stream);
/* First pass of the Navier-Stokes nonlinear term (demo/NavierStokes.py:37-39: `curl_hat = cross(K, U_hat) * 1j`
 * followed by the padded backward of both fields): u_hat [3][N][inner] (inner = n1*n2 spectral columns) ->
 * out [6][n_fft][inner], n_fft = 3N/2: fields 0..2 the 3/2-rule padded backward C2C of u_hat along the
 * leading axis, fields 3..5 the same of i k x u_hat, formed on the fly from the wavenumber tables
 * k0 [N], k1 [n1], k2 [n2].  The vorticity spectrum never exists in memory.  SHB_ERR_UNSUPPORTED unless
 * N/2 is one of 64, 128, 256, 512. */
SHB_API int shb_fft_c2c_backward_curl(const void *d_u_hat, void *d_out, int64_t inner, int n2, int n_fft, int N,
                                      const double *d_k0, const double *d_k1, const double *d_k2, double scale,
                                      void *stream);

/* C2C transform fused with the local half of the slab all-to-all (replaces the pack /
 * unpack that mpi4py_fft.pencil.Transfer expresses as MPI sub-array datatypes; call sites
 * shenfun/tensorproductspace.py:89,183,252).  With chunk_rows = cr != 0 the rows of the
 * transform axis are addressed peer by peer: row r of outer index o is at
 *   (r / cr) * (outer*cr*inner) + (o*cr + r % cr) * inner + c
 * i.e. the send buffer (out_chunk_rows) / receive buffer (in_chunk_rows) of an all-to-all
 * with equal blocks.  Returns SHB_ERR_UNSUPPORTED when no register-resident kernel covers
 * the size (callers then pack / unpack separately).                                   */
SHB_API int shb_fft_c2c_chunked(int backward, const void *d_in, void *d_out, int64_t outer, int64_t inner,
                                int n_fft, int N, double scale, int flags, int in_chunk_rows,
                                int out_chunk_rows, void *stream);

/* Same, with the all-to-all done by the kernel itself: block p of the output is stored
 * straight into d_peer_out[p] -- the receive buffer of rank p (peer-mapped memory over
 * NVLink / NVSwitch), already offset to this rank's slot -- so the transfer overlaps the
 * transform tile by tile and no separate collective runs.  The caller synchronises the
 * ranks (barrier) before the consumers read.  d_peer_out is a HOST array of nranks pointers. */
SHB_API int shb_fft_c2c_peer(int backward, const void *d_in, void *const *d_peer_out, int nranks,
                             int64_t outer, int64_t inner, int n_fft, int N, double scale, int flags,
                             int in_chunk_rows, int out_chunk_rows, void *stream);

/* The same transform on an arbitrary family of lines, for exchanges that are cut into pieces and
 * overlapped with the neighbouring passes: line (a, b), a < na, b < ncols, has element r at
 *     a*as + r*rs + b + c0                                   (plain side)
 *     (r / cr)*chunk_stride + a*as + (r % cr)*rs + b + c0     (peer-chunked side, cr = chunk_rows)
 * with separate (as, rs) for input and output.  The [outer][n][inner] array of the entry points
 * above is na = outer, ncols = inner, c0 = 0, rs = inner, as = n*inner (chunked: cr*inner); a column
 * sub-range, or the lines of a flattened (j, c) inner index for one range of c, are other settings.
 * Only columns b < ncols_valid are transformed (ncols may be rounded up so that the column tiles
 * of the kernel line up with a padded row pitch: aligned 128 / 256 B stores into peer memory).
 * d_peer_out (may be NULL): nranks peer receive buffers as in shb_fft_c2c_peer; else d_out.
 * SHB_ERR_UNSUPPORTED for line lengths without a register-resident kernel.                        */
SHB_API int shb_fft_c2c_lines(int backward, const void *d_in, void *d_out, void *const *d_peer_out, int nranks,
                              int64_t na, int64_t ncols, int64_t ncols_valid, int64_t c0, int64_t in_as, int64_t in_rs, int64_t out_as,
                              int64_t out_rs, int n_fft, int N, double scale, int flags, int in_chunk_rows,
                              int64_t in_chunk_stride, int out_chunk_rows, int64_t out_chunk_stride, void *stream);

/* Real-to-complex forward DFT of length n_fft, truncation to N//2+1 modes
 * (Nyquist -> 2*Re when N even and N < n_fft), scaling.
 * in : real [outer][n_fft][inner];  out : complex [outer][N/2+1][inner].  */
SHB_API int shb_fft_r2c_forward(const void *d_in, void *d_out, int64_t outer, int64_t inner,
                        int n_fft, int N, double scale, void *stream);

/* Pad N//2+1 modes to n_fft//2+1 (Nyquist -> 0.5*Re when N even and
 * N < n_fft), un-normalised complex-to-real inverse DFT of length n_fft.
 * in : complex [outer][N/2+1][inner];  out : real [outer][n_fft][inner].  */
SHB_API int shb_fft_c2r_backward(const void *d_in, void *d_out, int64_t outer, int64_t inner,
                         int n_fft, int N, double scale, int flags, void *stream);

/* ------------------------------------------------------------------ */
/* N-D Fourier transform plans (single device): the C-level counterpart of
 * TensorProductSpace.__init__ / configure_backwards (shenfun/tensorproductspace.py:314-541) and of
 * ForwardTransform / BackwardTransform.__call__ (:205-259, :46-96) for Fourier bases.
 * Axis i has `shape[i]` modes (an R2C axis stores shape/2+1 of them) and `padded[i]` >= shape[i]
 * quadrature points (3/2-rule padding when larger).  `axes` is the backward transform order; the
 * forward transform runs it in reverse ("last item is transformed first", :274-276), so an R2C axis
 * must be the last entry.  Forward is normalised by the padded lengths, backward is not
 * (fourier/bases.py:138-142, 407).  Arrays are C-ordered, `ncomp` components stacked in front
 * (VectorSpace / TensorSpace batches, :1786-1795).  The plan owns no device memory; the caller passes
 * shb_plan_workspace_bytes(plan, ncomp) bytes of scratch.                                     */
#define SHB_PLAN_MAXDIM 4
enum shb_axis_kind { SHB_AXIS_C2C = 0, SHB_AXIS_R2C = 1 };
typedef struct shb_plan_desc {
    int ndim;
    int64_t shape[SHB_PLAN_MAXDIM];
    int64_t padded[SHB_PLAN_MAXDIM];
    int kind[SHB_PLAN_MAXDIM];
    int axes[SHB_PLAN_MAXDIM];
    int dealias_direct;
} shb_plan_desc;
typedef struct shb_plan shb_plan;

SHB_API int shb_plan_create(const shb_plan_desc *desc, shb_plan **plan);
SHB_API int shb_plan_destroy(shb_plan *plan);
SHB_API int64_t shb_plan_workspace_bytes(const shb_plan *plan, int ncomp);
/* elements of the physical (spectral = 0: real when there is an R2C axis, else complex) or spectral
 * (spectral = 1: complex) array of ONE component */
SHB_API int64_t shb_plan_array_elems(const shb_plan *plan, int spectral);
SHB_API int shb_plan_forward(const shb_plan *plan, const void *d_in, void *d_out, void *d_work, int ncomp,
                             void *stream);
SHB_API int shb_plan_backward(const shb_plan *plan, const void *d_in, void *d_out, void *d_work, int ncomp,
                              void *stream);

/* ------------------------------------------------------------------ */
/* Point-wise kernels of the Navier-Stokes path.                        */

/* c = a x b on [3][n] arrays (real fp64 or complex128).  Replaces
 * shenfun/optimization/cython/cross.pyx:25-40 (`cross3D`), reached through
 * shenfun/utilities/__init__.py:177-210 (`cross`).                        */
SHB_API int shb_cross3d(const void *d_a, const void *d_b, void *d_c, int64_t n, int is_complex, void *stream);

/* out = i k x u on [3][n0][n1][n2] complex with per-axis wavenumber vectors:
 * what `project(curl(U_hat), TV)` evaluates for Fourier^3
 * (demo/NavierStokes.py:37; shenfun/fourier/matrices.py:149-176).          */
/* Convective term in physical space, H_i = sum_j u_j * grad[3*i+j], on [3][n] / nine [n] real
 * arrays (d_grad: HOST array of nine device pointers, d u_i / d x_j at index 3*i+j).  Replaces the
 * products and sums of demo/ChannelFlow.py:213-215 (conv = 0). */
SHB_API int shb_dot_grad3d(const double *d_u, const double *const *d_grad, double *d_out, int64_t n,
                           void *stream);

SHB_API int shb_curl_hat(const void *d_u, void *d_out, const double *d_k0, const double *d_k1,
                         const double *d_k2, int n0, int n1, int n2, void *stream);

/* One Runge-Kutta stage of the triply periodic Navier-Stokes demo, fused:
 * pressure solve + gradient (demo/NavierStokes.py:41-43), viscous term and
 * the RK4 stage update (shenfun/utilities/integrators.py:590-598).
 * d_P may be NULL; d_curl (may be NULL) receives i k x u_hat of the UPDATED velocity, i.e. the
 * curl the next stage starts from.  `last` bit 0: final stage (u_hat <- U1); bit 1: do not write dU back. */
SHB_API int shb_ns_rk_stage(void *d_dU, void *d_u_hat, const void *d_U0, void *d_U1, void *d_P,
                            void *d_curl, const double *d_k0, const double *d_k1, const double *d_k2,
                            int n0, int n1, int n2, double nu, double a_dt, double b_dt, int last,
                            void *stream);

/* Generic RK4 stage on [ncomp][n] complex arrays with a real diagonal linear
 * operator L[n] (may be NULL): integrators.py:590-598.                      */
SHB_API int shb_rk4_stage(void *d_dU, void *d_u_hat, const void *d_U0, void *d_U1, const double *d_L,
                          int64_t n, int ncomp, double a_dt, double b_dt, int last, void *stream);
/* Same with a COMPLEX diagonal operator L[n] (complex128): odd-derivative / complex-coefficient
 * linear terms such as KdV's -u_xxx (integrators.py:569-585 uses the complex scale array as is). */
SHB_API int shb_rk4_stage_cplx(void *d_dU, void *d_u_hat, const void *d_U0, void *d_U1, const void *d_L,
                               int64_t n, int ncomp, double a_dt, double b_dt, int last, void *stream);

/* out = mask * sum_j coefs[j] * T_j on n real or complex entries (d_terms: HOST array of nterms <= 10
 * device pointers, coefs: HOST array; mask real [n_mask], applied per entry modulo n_mask, may be NULL);
 * up to two terms are also copied out unchanged (d_save_dst[k] = T_{save_idx[k]}).  One pass for the
 * right-hand side of an IMEX Runge-Kutta stage, rhs = (u0, v) + dt sum_j b_j N_j + dt sum_j a_j L_j with
 * the newest N and L kept for the later stages (shenfun/utilities/integrators.py:692-696, 803-812).
 * is_complex bit 1 (complex data, no save slots): the result is also multiplied by i -- with a wavenumber
 * mask this is a Fourier derivative of odd order.  */
SHB_API int shb_lincomb(void *d_out, int nterms, const void *const *d_terms, const double *coefs,
                        const double *d_mask, int64_t n_mask, int64_t n, int is_complex, int nsave,
                        const int *save_idx, void *const *d_save_dst, void *stream);

/* u *= mask on [ncomp][n] complex (Nyquist masking): replaces
 * shenfun/optimization/cython/applymask.pyx:15-169 (`apply_mask`).          */
SHB_API int shb_apply_mask(void *d_u, const double *d_mask, int64_t n, int ncomp, void *stream);
/* out = mask * in (times i when times_i != 0) on complex arrays of n entries, mask real [n_mask] repeated along the
 * leading index (n a multiple of n_mask, n / n_mask <= 65535): a derivative along periodic axes applied in spectral
 * space, what `Project(Dx(u, periodic axis, k), space of u)` reduces to (forms/project.py:172-272). */
SHB_API int shb_scale_mask(void *d_out, const void *d_in, const double *d_mask, int64_t n_mask, int64_t n, int times_i,
                           void *stream);
/* v = i (a f + b g), w = i (b f - a g) on complex arrays of n entries; a, b real tables of n12 entries
 * broadcast over the leading index (demo/ChannelFlow.py:235-239, `compute_vw`). */
SHB_API int shb_kmm_vw(const void *d_f, const void *d_g, const double *d_a, const double *d_b, void *d_v, void *d_w,
                       int64_t n, int64_t n12, void *stream);

/* ------------------------------------------------------------------ */
/* Batched banded line solvers along the middle axis of [outer][n][inner].
 * Diagonals are spaced `step` apart (2 for Shen bases: offsets ..,-2,0,2,..;
 * 1 otherwise), kl lower and ku upper.  Matrix storage on the device:
 * mat[kl+1+ku][n], row d holds offset (d-kl)*step, entry [d][i] = M[i, i+off].
 * Factor storage: fac[kl+1+ku][outer][n][inner] per line, or fac[kl+1+ku][n]
 * when `fac_shared`.                                                         */

/* Per-line LU (no pivoting) of A(line) = sum_m scale_m[line] * M_m, nmat <= 4;
 * d_mats / d_scales are HOST arrays of device pointers (scale NULL = 1).
 * `pin_line` >= 0 replaces row 0 of that line by the identity (gauge).
 * Replaces la.SolverGeneric1ND.assemble/perform_lu/get_data
 * (shenfun/la.py:1231-1319) and the X_LU kernels (optimization/cython/la.pyx:199-303). */
SHB_API int shb_banded_factor(int kl, int ku, int step, int nmat, const double *const *d_mats,
                              const double *const *d_scales, double *d_fac, int64_t outer, int64_t inner,
                              int n, int64_t pin_line, void *stream);

/* In-place forward/backward substitution on u[outer][n_u][inner] (first n rows).
 * fac_shared: 0 = per-line factors in the batch layout above; 1 = one table [d][n];
 * 2 = one table in the reference's scipy-DIA layout (data[d][j] holds the entry of COLUMN j,
 * exactly what TDMA_LU / PDMA_LU / ... leave in `data`, shenfun/la.py:346-353); 3 = per-line
 * DIA tables [line][d][n] (the `data` argument of SolverGeneric1ND_solve_data).
 * Replaces TDMA_Solve / PDMA_Solve / FDMA_Solve / TwoDMA / ThreeDMA / HeptaDMA /
 * DiagMA_Solve and SolverGeneric1ND_solve_data (la.pyx:371-611).             */
SHB_API int shb_banded_solve(int kl, int ku, int step, void *d_u, const double *d_fac, int64_t outer,
                             int64_t inner, int n, int n_u, int fac_shared, int is_complex, void *stream);

/* out (+)= sum_m scale_m[line] * M_m u along the axis; M_m given as mat[ndiag][n_rows]
 * with explicit offsets.  `accumulate` bit 0: add to out; bit 1: multiply the result by i
 * (imaginary Fourier scales, complex data).  Replaces SparseMatrix.matvec / Matvec.pyx:430-512. */
SHB_API int shb_banded_matvec(const void *d_u, void *d_out, int nmat, int ndiag, const int *offsets,
                              const double *const *d_mats, const double *const *d_scales, int64_t outer,
                              int64_t inner, int n_rows, int n_cols, int n_u, int n_out, int accumulate,
                              int is_complex, void *stream);

/* Separable upper tail of a derivative matrix between Chebyshev bases:
 *   out_k (+)= scale(line) * r[k] * sum_{j >= k + w, (j - k - w) even} t[j] * u_j,   k < n_rows,
 * as running sums per parity (O(n) per line; rows n_rows..n_out-1 are zeroed unless accumulating).
 * `accumulate` as in shb_banded_matvec.  Replaces the dense `SpectralMatrix.matvec` of the
 * upper-triangular matrices (shenfun/matrixbase.py:87-144) and the specialised CDN/CTSD/... kernels
 * of optimization/cython/Matvec.pyx:57-420. */
SHB_API int shb_tail_matvec(const void *d_u, void *d_out, const double *d_r, const double *d_t,
                            const double *d_scale, int w, int64_t outer, int64_t inner, int n_rows, int n_cols,
                            int n_u, int n_out, int accumulate, int is_complex, void *stream);

/* Chebyshev-Galerkin Helmholtz (alfa*A + beta*B), per-line alfa/beta[outer*inner]:
 * factor into the four arrays d0, d1, d2, L, each [outer][n_fac][inner] with n_fac >= n rows
 * along the axis (the reference allocates n + 2, chebyshev/la.py:270-274), and solve.
 * `neumann` (factor) / flags bit 0 (solve): the reference's Neumann variant -- the caller passes
 * the column-scaled diagonals (la.pyx:737-745), lines with |beta| < 1e-8 get d0[0] = 1, and the
 * solve divides u_k by k^2 and zeroes the pinned mode (la.pyx:942-947).  flags bit 1 (solve): the
 * four factor arrays are single lines [n_fac] shared by every line (Poisson_Solve_ADD).  Replaces
 * LU_Helmholtz / Solve_Helmholtz (chebyshev/la.py:131-351, la.pyx:708-947).   */
SHB_API int shb_helmholtz_factor(const double *d_A0, const double *d_A2, const double *d_A4,
                                 const double *d_Bm2, const double *d_B0, const double *d_B2,
                                 const double *d_alfa, const double *d_beta, double *d_d0, double *d_d1,
                                 double *d_d2, double *d_L, int64_t outer, int64_t inner, int n, int n_fac,
                                 int neumann, void *stream);
SHB_API int shb_helmholtz_solve(const void *d_b, void *d_u, const double *d_d0, const double *d_d1,
                                const double *d_d2, const double *d_L, int64_t outer, int64_t inner, int n,
                                int n_fac, int n_u, int is_complex, int flags, void *stream);
/* The same solver from ONE stored array per line: d_invd0[outer][n][inner] = 1/d0 (written by
 * shb_helmholtz_factor_lean); d1, d2 and L are regenerated from the O(N) diagonals and the per-line
 * (alfa, beta) inside the solve (d2 obeys the recurrence of the forward substitution and rides along
 * as a third right-hand side).  Every line is cut into chunks of rows solved by different threads
 * (chunk maps + scan), the line stays in registers between the sweeps: b is read once, u written
 * once (1.25x the algorithmic bytes with the 1/d0 read).  Dirichlet variant only.
 * SHB_ERR_UNSUPPORTED when n is too long for one CTA (callers use the four-array form).           */
SHB_API int shb_helmholtz_factor_lean(const double *d_A0, const double *d_A2, const double *d_A4,
                                      const double *d_Bm2, const double *d_B0, const double *d_B2,
                                      const double *d_alfa, const double *d_beta, double *d_invd0,
                                      int64_t outer, int64_t inner, int n, void *stream);
/* The solve takes the four diagonals it needs as two interleaved tables of
 * shb_helmholtz_lean_table_rows(n) rows, zero beyond row n:
 *   d_tabF[i] = { Bm2[i-2] (0 for i < 2), A4[i] },   d_tabB[i] = { A2[i], B2[i] }.               */
SHB_API int shb_helmholtz_lean_table_rows(int n);
SHB_API int shb_helmholtz_solve_lean(const void *d_b, void *d_u, const double *d_tabF, const double *d_tabB,
                                     const double *d_alfa, const double *d_beta, const double *d_invd0,
                                     int64_t outer, int64_t inner, int n, int n_u, int is_complex, void *stream);
/* c = (alfa*A + beta*B) v with running sums over the constant rows of A; replaces
 * Helmholtz_matvec (optimization/cython/Matvec.pyx:656-712).                    */
SHB_API int shb_helmholtz_matvec(const void *d_v, void *d_c, const double *d_A0, const double *d_A2,
                                 const double *d_A4, const double *d_Bm2, const double *d_B0,
                                 const double *d_B2, const double *d_alfa, const double *d_beta, int64_t outer,
                                 int64_t inner, int n, int n_u, int is_complex, void *stream);

/* Chebyshev-Galerkin biharmonic (a*S + b*A + c*B) on the ShenBiharmonic basis, per-line scales
 * a, b, c[outer*inner].  d_tab[13][n] holds the diagonals of S (0,+2,+4), the coefficients of its
 * rank-2 tail, A (-2,0,2) and B (-4..4); d_fac[7][outer][n][inner] receives L1, L2, D, U1, U2 and
 * the eliminated tail coefficients.  O(n) per line.  Replaces LU_Biharmonic,
 * Biharmonic_factor_pr, Biharmonic_Solve, Biharmonic_matvec (shenfun/chebyshev/la.py:354-558,
 * optimization/cython/la.pyx:975-2121, Matvec.pyx:820-922).                                  */
SHB_API int shb_biharmonic_factor(const double *d_tab, const double *d_a, const double *d_b,
                                  const double *d_c, double *d_fac, int64_t outer, int64_t inner, int n,
                                  void *stream);
SHB_API int shb_biharmonic_solve(const void *d_b, void *d_u, const double *d_fac, int64_t outer,
                                 int64_t inner, int n, int n_u, int is_complex, void *stream);
SHB_API int shb_biharmonic_matvec(const void *d_v, void *d_c, const double *d_tab, const double *d_a,
                                  const double *d_b, const double *d_cs, int64_t outer, int64_t inner,
                                  int n, int n_u, int is_complex, void *stream);

/* ------------------------------------------------------------------ */
/* Chebyshev (Gauss-Chebyshev points) axis transforms on [outer][n][inner],
 * real (is_real) or complex data (real and imaginary parts in ONE pass --
 * the reference's DCTWrap runs two, chebyshev/bases.py:100-115).
 * A composite (Shen) basis phi_k = sum_i sten[i][k] T_{k+soff[i]}, k < dim,
 * is given by its stencil diagonals (device arrays, NULL = ones); the
 * orthogonal basis is nsten = 1, soff = {0}, sten = NULL, dim = n_spec.       */

/* Weighted scalar product: out_k = sum_i sten_i[k] * S_{k+soff_i},
 * S = scale * DCT-II(in) (FFTW REDFT10 convention), rows dim..n_spec-1 zeroed.
 * Replaces chebyshev.Orthogonal._evaluate_scalar_product (chebyshev/bases.py:340-353),
 * CompositeBase._evaluate_scalar_product (spectralbase.py:1462-1474) and the
 * truncation (spectralbase.py:1363-1367).  scale = pi/(2*domain_factor*n_phys). */
SHB_API int shb_cheb_scalar_product(const void *d_in, void *d_out, int64_t outer, int64_t inner, int n_phys,
                                    int n_spec, double scale, int is_real, int dim, int nsten, const int *soff,
                                    const double *const *d_sten, void *stream);

/* Evaluation on the quadrature points: c = K^T u_hat (+ boundary lift from rows
 * dim..dim+nbc-1 through bc_lift[nbc][4]), zero-padded to n_phys, then
 * u = 0.5*DCT-III(c) + c_0/2.  Replaces spectralbase.py:1369-1390 (padding),
 * CompositeBase.to_ortho :1616-1634 and chebyshev/bases.py:315-338.             */
SHB_API int shb_cheb_backward(const void *d_in, void *d_out, int64_t outer, int64_t inner, int n_phys,
                              int n_spec, int is_real, int dim, int nsten, const int *soff,
                              const double *const *d_sten, int nbc, const double *d_bc_lift, void *stream);

/* ------------------------------------------------------------------ */
/* Dense contraction along the middle axis with a real operator matrix A[M][K]
 * (fp64 tensor cores):  out[o][i][c] = sum_j A[i][j] * u[o][j][c],
 * u : [outer][K][inner] real or complex, out : [outer][M][inner].  Legendre/Jacobi
 * Vandermonde transforms: replaces SpectralBase._evaluate_expansion_all /
 * _evaluate_scalar_product kinds 'vandermonde' and 'recursive'
 * (shenfun/spectralbase.py:448-491, 516-558; optimization/cython/transforms.pyx:175-323). */
SHB_API int shb_gemm_axis(const double *d_A, const void *d_u, void *d_out, int M, int K, int64_t outer,
                          int64_t inner, int is_complex, void *stream);
/* Same contraction for an operator with the parity structure of the symmetric families (Legendre, Jacobi(a,a),
 * Chebyshev: basis functions even / odd, quadrature points symmetric): half the tensor-core work.
 *   forward != 0: A[i][K-1-j] = (-1)^i A[i][j] (K even; values -> coefficients),
 *   forward == 0: A[M-1-i][k] = (-1)^k A[i][k] (M even; coefficients -> values).
 * d_A2 = the two half operators [2][Mh][Kp], Kp even and zero padded: forward: A2[p][i][j] = A[2i+p][j]
 * (Mh = ceil(M/2), j < K/2); backward: A2[p][i][k] = A[i][2k+p] (Mh = M/2, k < ceil(K/2)).
 * `forward` bits 0-1 = 2: checkerboard operator (A[i][j] = 0 for i + j odd, e.g. the Legendre <-> Chebyshev connection
 * matrices, shenfun/legendre/dlt.py:353-474): A2[p][i][k] = A[2i+p][2k+p] (Mh = ceil(M/2), k < ceil(K/2)); bit 2 / bit 3:
 * the halves are upper / lower triangular and their zero k-tiles are skipped.
 * Returns SHB_ERR_UNSUPPORTED for layouts it does not cover (inner*(complex?2:1) odd or 1, unaligned
 * pointers): call shb_gemm_axis then. */
SHB_API int shb_gemm_axis_parity(const double *d_A2, const void *d_u, void *d_out, int M, int K, int Kp, int forward,
                                 int64_t outer, int64_t inner, int is_complex, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SHENB200_H */
