// banded.cu -- batched banded line solvers and mat-vecs along the middle axis of
// [outer][n][inner], one line per (outer, inner) pair -- i.e. one line per Fourier
// wavenumber pair in the tensor-product solvers.
//
// Parallelisation: one thread per line with the lanes of a warp running along the
// contiguous `inner` index, so every global access of the sweep is coalesced along
// the batch; the recurrence itself runs sequentially down the (strided) solve axis.
//
//   * banded_factor_kernel : per-line LU without pivoting of  A(line) = sum_m s_m(line) M_m,
//       M_m shared banded matrices (diagonals spaced `step` apart: 2 for the Shen
//       bases, whose even/odd unknowns decouple), factors written in the batch layout
//   * banded_solve_kernel  : forward + backward substitution, in place, factors either
//       per line (SolverGeneric1ND) or shared (TDMA/PDMA/... of one matrix, inverse mass)
//   * banded_matvec_kernel : out (+)= sum_m s_m(line) M_m u
//   * helmholtz_factor/solve: Chebyshev-Galerkin (alfa*A + beta*B) with A upper
//       triangular with constant rows (LU with one lower and three upper "diagonals",
//       back substitution with running even/odd sums)
//
// Reference semantics restated (no code shared):
//   shenfun/la.py:291-820 (TDMA/PDMA/FDMA/TwoDMA/ThreeDMA/HeptaDMA LU + solve),
//   shenfun/la.py:1183-1434 (SolverGeneric1ND), optimization/cython/la.pyx:199-611,
//   chebyshev/la.py:131-351 + la.pyx:708-973 (Helmholtz), Matvec.pyx:430-512.
#include "common.cuh"

namespace shb {

int banded_tile_launch(int kl, int ku, int step, void *u, const double *fac, int64_t outer, int64_t inner, int n,
                       int n_u, int is_complex, cudaStream_t st, int per_line);

constexpr int BD_THREADS = 128;
constexpr int BD_MAXMAT = 4;

struct BandedDesc {
    int64_t outer, inner;
    int n;            // matrix size (rows solved)
    int n_u;          // rows of the data arrays along the axis (n_u >= n)
    int step;         // spacing of the diagonals (1 or 2)
    int nmat;
    const double *mats[BD_MAXMAT];    // each [KL+1+KU][n]: rows ordered by offset -KL*step .. +KU*step,
                                      // entry [d][i] = M[i, i + off_d] (0 outside the matrix)
    const double *scales[BD_MAXMAT];  // per line [outer*inner] or NULL (= 1)
    double *fac;                      // [KL+1+KU][outer][n][inner] (per line) or [KL+1+KU][n] (shared)
    int fac_shared;
    int64_t pin_line;                 // line whose row 0 is replaced by the identity (-1: none)
};

template <class T> struct Num;
template <> struct Num<double> {
    static __device__ __forceinline__ double zero() { return 0.0; }
    static __device__ __forceinline__ double fnma(double a, double x, double y) { return fma(-a, x, y); }  // y - a x
    static __device__ __forceinline__ double mul(double a, double x) { return a * x; }
    static __device__ __forceinline__ double div(double x, double a) { return x / a; }
    static __device__ __forceinline__ double add(double x, double y) { return x + y; }
};
template <> struct Num<cplx> {
    static __device__ __forceinline__ cplx zero() { return make_double2(0.0, 0.0); }
    static __device__ __forceinline__ cplx fnma(double a, cplx x, cplx y) {
        return make_double2(fma(-a, x.x, y.x), fma(-a, x.y, y.y));
    }
    static __device__ __forceinline__ cplx mul(double a, cplx x) { return make_double2(a * x.x, a * x.y); }
    static __device__ __forceinline__ cplx div(cplx x, double a) { return make_double2(x.x / a, x.y / a); }
    static __device__ __forceinline__ cplx add(cplx x, cplx y) { return make_double2(x.x + y.x, x.y + y.y); }
};

// ---- per-line LU ---------------------------------------------------------------------
// Row i of A: lo[p] = A[i, i-p*step] (p=1..KL), dg = A[i,i], up[q] = A[i, i+q*step] (q=1..KU).
// Factors stored per row i: L_p[i] (multiplier against row i-p*step), D[i], U_q[i].
// Layout of `fac`: diagonal index d = KL-p for L_p, KL for D, KL+q for U_q.
template <int KL, int KU>
__global__ void __launch_bounds__(BD_THREADS) banded_factor_kernel(const __grid_constant__ BandedDesc d) {
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nlines = d.outer * d.inner;
    if (line >= nlines) return;
    const int64_t o = line / d.inner, c = line % d.inner;
    const int n = d.n, step = d.step;
    constexpr int ND = KL + 1 + KU;
    double s[BD_MAXMAT];
#pragma unroll
    for (int m = 0; m < BD_MAXMAT; ++m)
        s[m] = (m < d.nmat) ? (d.scales[m] ? d.scales[m][line] : 1.0) : 0.0;
    const bool pin = (line == d.pin_line);
    const int64_t fstride_d = d.outer * (int64_t)n * d.inner;   // between diagonals
    double *fac = d.fac + (o * n) * d.inner + c;

    for (int par = 0; par < step; ++par) {
        // window of the previous KL pivot rows of this parity: D and U_1..U_KU
        double wD[KL > 0 ? KL : 1], wU[KL > 0 ? KL : 1][KU > 0 ? KU : 1];
#pragma unroll
        for (int p = 0; p < (KL > 0 ? KL : 1); ++p) {
            wD[p] = 1.0;
#pragma unroll
            for (int q = 0; q < (KU > 0 ? KU : 1); ++q) wU[p][q] = 0.0;
        }
        for (int i = par; i < n; i += step) {
            double a[ND];
#pragma unroll
            for (int dd = 0; dd < ND; ++dd) {
                double v = 0.0;
#pragma unroll
                for (int m = 0; m < BD_MAXMAT; ++m)
                    if (m < d.nmat) v = fma(s[m], __ldg(&d.mats[m][dd * n + i]), v);
                a[dd] = v;
            }
            if (pin && i == 0) {
#pragma unroll
                for (int dd = 0; dd < ND; ++dd) a[dd] = (dd == KL) ? 1.0 : 0.0;
            }
            // eliminate against rows i-KL*step ... i-step (farthest first)
            double l[KL > 0 ? KL : 1];
#pragma unroll
            for (int p = KL; p >= 1; --p) {
                // a[KL-p] is the entry in column i-p*step; pivot row is window slot p-1
                const double lp = (i - p * step >= 0) ? a[KL - p] / wD[p - 1] : 0.0;
                l[p - 1] = lp;
                // row i -= lp * (row i-p*step): that row has U_q at columns i-p*step+q*step
#pragma unroll
                for (int q = 1; q <= KU; ++q) {
                    const int col = KL - p + q;      // index into a[] of column i+(q-p)*step
                    if (col < ND) a[col] = fma(-lp, wU[p - 1][q - 1], a[col]);
                }
            }
            // store
#pragma unroll
            for (int p = 1; p <= KL; ++p) fac[(KL - p) * fstride_d + (int64_t)i * d.inner] = l[p - 1];
#pragma unroll
            for (int dd = KL; dd < ND; ++dd) fac[dd * fstride_d + (int64_t)i * d.inner] = a[dd];
            // slide the window
#pragma unroll
            for (int p = KL - 1; p >= 1; --p) {
                wD[p] = wD[p - 1];
#pragma unroll
                for (int q = 0; q < KU; ++q) wU[p][q] = wU[p - 1][q];
            }
            if (KL > 0) {
                wD[0] = a[KL];
#pragma unroll
                for (int q = 0; q < KU; ++q) wU[0][q] = a[KL + 1 + q];
            }
        }
    }
}

// ---- substitution -----------------------------------------------------------------------
template <int KL, int KU, class T>
__global__ void __launch_bounds__(BD_THREADS) banded_solve_kernel(T *__restrict__ u, const double *__restrict__ facp,
                                                                  int64_t outer, int64_t inner, int n, int n_u,
                                                                  int step, int fac_shared) {
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= outer * inner) return;
    const int64_t o = line / inner, c = line % inner;
    T *x = u + (o * n_u) * inner + c;
    // fac_shared: 0 = per line, batch layout [d][outer][n][inner];  1 = one table [d][n];
    //             2 = one table in scipy DIA layout (entry of column j stored at [d][j], as
    //                 the reference's X_LU leaves it);  3 = per line DIA, [line][d][n]
    //                 (the `data` of SolverGeneric1ND_solve_data)
    const int64_t fs_i = fac_shared ? 1 : inner;                                   // between rows
    const int64_t fs_d = fac_shared ? (int64_t)n : outer * (int64_t)n * inner;     // between diagonals
    const double *fac = fac_shared == 0 ? facp + (o * n) * inner + c
                      : fac_shared == 3 ? facp + line * (int64_t)(KL + 1 + KU) * n : facp;
    const int sh = fac_shared >= 2 ? step : 0;      // DIA: L_p of row i sits at column i - p*step, U_q at i + q*step
    // The recurrence is serial along the line, but the right-hand side and the factor entries of
    // later rows do not depend on it: rows are taken in batches of U whose loads are all issued
    // before the dependent arithmetic, so a thread keeps U*(1+KL) (resp. U*(2+KU)) loads in flight
    // instead of exposing the memory latency once per row.
    constexpr int U = (KL + KU <= 4) ? 8 : 4;
    for (int par = 0; par < step; ++par) {
        // forward: y_i = b_i - sum_p L_p[i] y_{i-p*step}
        T w[KL > 0 ? KL : 1];
#pragma unroll
        for (int p = 0; p < (KL > 0 ? KL : 1); ++p) w[p] = Num<T>::zero();
        if (KL > 0) {
            for (int i0 = par; i0 < n; i0 += step * U) {
                T bb[U];
                double ll[U][KL > 0 ? KL : 1];
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const int i = i0 + k * step;
                    if (i < n) {
                        bb[k] = x[(int64_t)i * inner];
#pragma unroll
                        for (int p = 1; p <= KL; ++p)
                            ll[k][p - 1] = (i - p * step >= 0) ? fac[(KL - p) * fs_d + i * fs_i - p * sh] : 0.0;
                    }
                }
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const int i = i0 + k * step;
                    if (i < n) {
                        T y = bb[k];
#pragma unroll
                        for (int p = 1; p <= KL; ++p) y = Num<T>::fnma(ll[k][p - 1], w[p - 1], y);
                        x[(int64_t)i * inner] = y;
#pragma unroll
                        for (int p = KL - 1; p >= 1; --p) w[p] = w[p - 1];
                        w[0] = y;
                    }
                }
            }
        }
        // backward: x_i = (y_i - sum_q U_q[i] x_{i+q*step}) / D[i]
        T z[KU > 0 ? KU : 1];
#pragma unroll
        for (int q = 0; q < (KU > 0 ? KU : 1); ++q) z[q] = Num<T>::zero();
        if (n - 1 - par < 0) continue;
        const int last = par + ((n - 1 - par) / step) * step;
        for (int i0 = last; i0 >= par; i0 -= step * U) {
            T bb[U];
            double dg[U];
            double uu[U][KU > 0 ? KU : 1];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int i = i0 - k * step;
                if (i >= par) {
                    bb[k] = x[(int64_t)i * inner];
                    dg[k] = fac[KL * fs_d + i * fs_i];
#pragma unroll
                    for (int q = 1; q <= KU; ++q)
                        uu[k][q - 1] = (i + q * step < n) ? fac[(KL + q) * fs_d + i * fs_i + q * sh] : 0.0;
                }
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int i = i0 - k * step;
                if (i >= par) {
                    T y = bb[k];
#pragma unroll
                    for (int q = 1; q <= KU; ++q) y = Num<T>::fnma(uu[k][q - 1], z[q - 1], y);
                    y = Num<T>::div(y, dg[k]);
                    x[(int64_t)i * inner] = y;
#pragma unroll
                    for (int q = KU - 1; q >= 1; --q) z[q] = z[q - 1];
                    if (KU > 0) z[0] = y;
                }
            }
        }
    }
}

// ---- mat-vec ------------------------------------------------------------------------------
// out[o][i][c] (+)= sum_m s_m(line) sum_d M_m[d][i] u[o][i+off_d][c],  offsets given explicitly
struct MatvecDesc {
    int64_t outer, inner;
    int n_rows, n_cols;     // matrix shape
    int n_u, n_out;         // rows of the arrays along the axis
    int nmat, ndiag;
    int offs[16];
    const double *mats[BD_MAXMAT];    // [ndiag][n_rows]
    const double *scales[BD_MAXMAT];
    int accumulate;
};

__device__ __forceinline__ double times_i(double a) { return a; }     // (real data never carries imaginary scales)
__device__ __forceinline__ cplx times_i(cplx a) { return make_double2(-a.y, a.x); }

template <class T>
__global__ void __launch_bounds__(256) banded_matvec_kernel(const T *__restrict__ u, T *__restrict__ out,
                                                            const __grid_constant__ MatvecDesc d) {
    const int64_t tot = d.outer * (int64_t)d.n_out * d.inner;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < tot;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx % d.inner;
        const int64_t r = idx / d.inner;
        const int i = (int)(r % d.n_out);
        const int64_t o = r / d.n_out;
        const int64_t line = o * d.inner + c;
        T acc = Num<T>::zero();
        if (i < d.n_rows) {
            const T *ul = u + (o * d.n_u) * d.inner + c;
            for (int m = 0; m < d.nmat; ++m) {
                const double s = d.scales[m] ? d.scales[m][line] : 1.0;
                T part = Num<T>::zero();
                for (int k = 0; k < d.ndiag; ++k) {
                    const int j = i + d.offs[k];
                    if (j >= 0 && j < d.n_cols) {
                        const double a = __ldg(&d.mats[m][k * d.n_rows + i]);
                        part = Num<T>::fnma(-a, ul[(int64_t)j * d.inner], part);
                    }
                }
                acc = Num<T>::fnma(-s, part, acc);
            }
        }
        if (d.accumulate & 2) acc = times_i(acc);          // scale arrays were the imaginary parts
        if (d.accumulate & 1) out[idx] = Num<T>::add(out[idx], acc);
        else out[idx] = acc;
    }
}

// Same product with the index arithmetic taken out of the inner loop: a thread owns one column
// (line) and R consecutive rows, the grid carries (column block, row block, outer index), so no
// 64-bit div/mod per element and the per-line scales are loaded once per R rows.  Used whenever
// the batch index is the contiguous one (inner >= 32).
template <class T, int R>
__global__ void __launch_bounds__(128) banded_matvec_col_kernel(const T *__restrict__ u, T *__restrict__ out,
                                                                const __grid_constant__ MatvecDesc d) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d.inner) return;
    const int64_t o = blockIdx.z;
    const int i0 = blockIdx.y * R;
    const int64_t line = o * d.inner + c;
    const T *ul = u + (o * d.n_u) * d.inner + c;
    T *ol = out + (o * d.n_out) * d.inner + c;
    double s[BD_MAXMAT];
#pragma unroll
    for (int m = 0; m < BD_MAXMAT; ++m) s[m] = (m < d.nmat && d.scales[m]) ? d.scales[m][line] : 1.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = i0 + r;
        if (i >= d.n_out) break;
        T acc = Num<T>::zero();
        if (i < d.n_rows) {
            for (int m = 0; m < d.nmat; ++m) {
                T part = Num<T>::zero();
                for (int k = 0; k < d.ndiag; ++k) {
                    const int j = i + d.offs[k];
                    if (j >= 0 && j < d.n_cols) {
                        const double a = __ldg(&d.mats[m][k * d.n_rows + i]);
                        part = Num<T>::fnma(-a, ul[(int64_t)j * d.inner], part);
                    }
                }
                acc = Num<T>::fnma(-s[m], part, acc);
            }
        }
        if (d.accumulate & 2) acc = times_i(acc);
        T *dst = ol + (int64_t)i * d.inner;
        if (d.accumulate & 1) *dst = Num<T>::add(*dst, acc);
        else *dst = acc;
    }
}


// Separable upper tail: out_k (+)= s(line) * r_k * sum_{j >= k + w, (j - k - w) even} t_j u_j.
// Derivative matrices between Chebyshev bases are upper triangular with entries that depend on the
// column only ((phi_j', T_k)_w = -2 pi for k < j, j + k odd) or on the row only (the Dirichlet
// stiffness matrix): the product is a running sum per parity carried down the line, O(n) instead of
// the O(n^2) dense contraction.  One thread per line, rows from the bottom up, the u values of eight
// rows loaded ahead of the sums.
template <class T>
__global__ void __launch_bounds__(BD_THREADS) tail_matvec_kernel(const T *__restrict__ u, T *__restrict__ out,
                                                                 const double *__restrict__ r,
                                                                 const double *__restrict__ t,
                                                                 const double *__restrict__ scale, int w,
                                                                 int64_t outer, int64_t inner, int n_rows, int n_cols,
                                                                 int n_u, int n_out, int accumulate) {
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= outer * inner) return;
    const int64_t o = line / inner, c = line % inner;
    const double s = scale ? scale[line] : 1.0;
    const T *ul = u + (o * n_u) * inner + c;
    T *ol = out + (o * n_out) * inner + c;
    if (!(accumulate & 1))
        for (int k = n_rows; k < n_out; ++k) ol[(int64_t)k * inner] = Num<T>::zero();
    constexpr int U = 8;
    T P[2] = {Num<T>::zero(), Num<T>::zero()};
    for (int k0 = n_rows - 1; k0 >= 0; k0 -= U) {
        T uu[U], old[U];
        double tt[U], rr[U];
#pragma unroll
        for (int b = 0; b < U; ++b) {
            const int k = k0 - b;
            if (k >= 0) {
                const int j = k + w;
                const bool in = (j >= 0 && j < n_cols);
                uu[b] = in ? ul[(int64_t)j * inner] : Num<T>::zero();
                tt[b] = in ? __ldg(&t[j]) : 0.0;
                rr[b] = __ldg(&r[k]);
                if (accumulate & 1) old[b] = ol[(int64_t)k * inner];
            }
        }
#pragma unroll
        for (int b = 0; b < U; ++b) {
            const int k = k0 - b;
            if (k >= 0) {
                const int par = (k + w) & 1;
                P[par] = Num<T>::fnma(-tt[b], uu[b], P[par]);
                T acc = Num<T>::mul(s * rr[b], P[par]);
                if (accumulate & 2) acc = times_i(acc);
                ol[(int64_t)k * inner] = (accumulate & 1) ? Num<T>::add(old[b], acc) : acc;
            }
        }
    }
}


// ---- Chebyshev Helmholtz ------------------------------------------------------------------
// (alfa*A + beta*B) u = b,  A: diagonals 0,2,4 with A[i, i+2q] = A4[i] for q >= 2 (constant rows),
// B: diagonals -2, 0, 2.  Per line: forward recurrence for L, d0, d1, d2 (la.pyx:744-756).
__global__ void __launch_bounds__(BD_THREADS) helmholtz_factor_kernel(
    const double *__restrict__ A0, const double *__restrict__ A2, const double *__restrict__ A4,
    const double *__restrict__ Bm2, const double *__restrict__ B0, const double *__restrict__ B2,
    const double *__restrict__ alfa, const double *__restrict__ beta, double *__restrict__ fd0,
    double *__restrict__ fd1, double *__restrict__ fd2, double *__restrict__ fL,
    int64_t outer, int64_t inner, int n, int n_fac, int neumann) {
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= outer * inner) return;
    const int64_t o = line / inner, c = line % inner;
    const double al = alfa[line], be = beta[line];
    const int64_t base = (o * n_fac) * inner + c;       // each factor array is [outer][n_fac][inner]
    fd0 += base; fd1 += base; fd2 += base; fL += base;
    // two interleaved chains (even / odd rows)
    double pd0[2], pd1[2], pd2[2];
    for (int i = 0; i < n; ++i) {
        const int par = i & 1;
        double d0 = al * __ldg(&A0[i]) + be * __ldg(&B0[i]);
        if (neumann && i == 0 && fabs(be) < 1e-8) d0 = 1.0;      // singular mode pinned (la.pyx:733-736)
        double d1 = (i < n - 2) ? al * __ldg(&A2[i]) + be * __ldg(&B2[i]) : 0.0;
        double d2 = (i < n - 4) ? al * __ldg(&A4[i]) : 0.0;
        if (i >= 2) {
            const double L = be * __ldg(&Bm2[i - 2]) / pd0[par];
            fL[(int64_t)(i - 2) * inner] = L;
            d0 -= L * pd1[par];
            if (i < n - 2) d1 -= L * pd2[par];
            if (i < n - 4) d2 -= L * pd2[par];
        }
        fd0[(int64_t)i * inner] = d0;
        fd1[(int64_t)i * inner] = d1;
        fd2[(int64_t)i * inner] = d2;
        pd0[par] = d0; pd1[par] = d1; pd2[par] = d2;
    }
    if (n >= 2) {
        fL[(int64_t)(n - 2) * inner] = 0.0;
        fL[(int64_t)(n - 1) * inner] = 0.0;
    } else if (n == 1) {
        fL[0] = 0.0;
    }
}

// Solve (la.pyx:871-947): y_i = b_i - L[i-2] y_{i-2};  u_i = (y_i - d1[i] u_{i+2} - d2[i] sum_{q>=2} u_{i+2q}) / d0[i]
template <class T>
__global__ void __launch_bounds__(BD_THREADS) helmholtz_solve_kernel(
    const T *b, T *u /* may alias b: in-place solve */, const double *__restrict__ fd0, const double *__restrict__ fd1,
    const double *__restrict__ fd2, const double *__restrict__ fL, int64_t outer, int64_t inner,
    int n, int n_fac, int n_u, int flags) {
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= outer * inner) return;
    const int64_t o = line / inner, c = line % inner;
    const bool shared = flags & 2, neumann = flags & 1;
    const int64_t fi = shared ? 1 : inner;                 // factor stride between rows
    if (!shared) {
        const int64_t base = (o * n_fac) * inner + c;
        fd0 += base; fd1 += base; fd2 += base; fL += base;
    }
    const T *bl = b + (o * n_u) * inner + c;
    T *x = u + (o * n_u) * inner + c;
    // rows in batches of U: all loads of a batch are issued before the serial arithmetic
    constexpr int U = 8;
    T y[2] = {Num<T>::zero(), Num<T>::zero()};
    for (int i0 = 0; i0 < n; i0 += U) {
        T bb[U];
        double ll[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int i = i0 + k;
            if (i < n) {
                bb[k] = bl[(int64_t)i * inner];
                ll[k] = (i >= 2) ? fL[(int64_t)(i - 2) * fi] : 0.0;
            }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int i = i0 + k;
            if (i < n) {
                const int par = k & 1;          // U is even: parity of i0 + k == parity of k
                T v = Num<T>::fnma(ll[k], y[par], bb[k]);
                x[(int64_t)i * inner] = v;
                y[par] = v;
            }
        }
    }
    T un[2] = {Num<T>::zero(), Num<T>::zero()};     // u_{i+2} per parity
    T sum[2] = {Num<T>::zero(), Num<T>::zero()};    // sum_{q>=2} u_{i+2q} per parity
    // walk down from the last row in batches whose first row is odd-aligned so that parity is static
    const int top = (n + 1) & ~1;                   // even, >= n
    for (int i0 = top - 1; i0 >= 0; i0 -= U) {
        T bb[U];
        double e0[U], e1[U], e2[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int i = i0 - k;
            if (i >= 0 && i < n) {
                bb[k] = x[(int64_t)i * inner];
                e0[k] = fd0[(int64_t)i * fi];
                e1[k] = fd1[(int64_t)i * fi];
                e2[k] = fd2[(int64_t)i * fi];
            }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int i = i0 - k;
            if (i >= 0 && i < n) {
                const int par = (k & 1) ^ 1;    // i0 is odd
                T v = bb[k];
                v = Num<T>::fnma(e1[k], un[par], v);
                v = Num<T>::fnma(e2[k], sum[par], v);
                v = Num<T>::div(v, e0[k]);
                sum[par] = Num<T>::add(sum[par], un[par]);
                un[par] = v;
                if (neumann) {
                    // unknowns were scaled by k^2; the pinned constant mode is zero (la.pyx:942-947)
                    if (i == 0) { if ((e0[k] - 1.0) * (e0[k] - 1.0) < 1e-16) v = Num<T>::zero(); }
                    else v = Num<T>::div(v, (double)i * (double)i);
                }
                x[(int64_t)i * inner] = v;
            }
        }
    }
}


// c = (alfa*A + beta*B) v with the same structure (Matvec.pyx:656-712 `Helmholtz_matvec`):
// (A v)_k = A0[k] v_k + A2[k] v_{k+2} + A4[k] sum_{q>=2} v_{k+2q}
template <class T>
__global__ void __launch_bounds__(BD_THREADS) helmholtz_matvec_kernel(
    const T *__restrict__ v, T *__restrict__ c, const double *__restrict__ A0, const double *__restrict__ A2,
    const double *__restrict__ A4, const double *__restrict__ Bm2, const double *__restrict__ B0,
    const double *__restrict__ B2, const double *__restrict__ alfa, const double *__restrict__ beta,
    int64_t outer, int64_t inner, int n, int n_u) {
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= outer * inner) return;
    const int64_t o = line / inner, cc = line % inner;
    const double al = alfa[line], be = beta[line];
    const T *x = v + (o * n_u) * inner + cc;
    T *y = c + (o * n_u) * inner + cc;
    T un[2] = {Num<T>::zero(), Num<T>::zero()};
    T sum[2] = {Num<T>::zero(), Num<T>::zero()};
    for (int i = n_u - 1; i >= n; --i) y[(int64_t)i * inner] = Num<T>::zero();
    for (int i = n - 1; i >= 0; --i) {
        const int par = i & 1;
        const T xi = x[(int64_t)i * inner];
        T acc = Num<T>::mul(al * __ldg(&A0[i]) + be * __ldg(&B0[i]), xi);
        if (i < n - 2) acc = Num<T>::fnma(-(al * __ldg(&A2[i]) + be * __ldg(&B2[i])), un[par], acc);
        if (i < n - 4) acc = Num<T>::fnma(-(al * __ldg(&A4[i])), sum[par], acc);
        if (i >= 2) acc = Num<T>::fnma(-(be * __ldg(&Bm2[i - 2])), x[(int64_t)(i - 2) * inner], acc);
        y[(int64_t)i * inner] = acc;
        sum[par] = Num<T>::add(sum[par], un[par]);
        un[par] = xi;
    }
}

// ---- Chebyshev biharmonic ------------------------------------------------------------------
// H = a*S + b*A + c*B on the Chebyshev ShenBiharmonic basis (chebyshev/la.py:354-558,
// la.pyx:975-1308): S = (d^4 phi_j / dx^4, phi_k)_w is upper triangular with diagonals 0, +2, +4
// and a rank-2 tail  S[k, j] = (alpha_k + beta_k (j+2)^2) / (j+3)  for j >= k+6 (same parity);
// A has offsets -2, 0, 2 and B offsets -4 .. 4.  LU without pivoting keeps that shape: two lower
// diagonals, three upper diagonals and tail coefficients (FA, FB) that follow the row recurrence,
// so the factorisation is O(n) per line (the reference eliminates full rows, O(n^2)).
// Table layout tab[13][n]: S0 S2 S4 alpha beta | A-2 A0 A2 | B-4 B-2 B0 B2 B4  (entry [d][i] = M[i, i+off]).
// Factor layout fac[7][outer][n][inner]: L1 (vs row i-2), L2 (vs row i-4), D, U1, U2, FA, FB.
__device__ __forceinline__ double bih_far(double fa, double fb, int j) {
    const double jp2 = (double)(j + 2);
    return (fa + fb * jp2 * jp2) / (double)(j + 3);
}

__global__ void __launch_bounds__(BD_THREADS) biharmonic_factor_kernel(
    const double *__restrict__ tab, const double *__restrict__ sa, const double *__restrict__ sb,
    const double *__restrict__ sc, double *__restrict__ fac, int64_t outer, int64_t inner, int n) {
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= outer * inner) return;
    const int64_t o = line / inner, c = line % inner;
    const double a = sa[line], b = sb[line], cc = sc[line];
    const int64_t fs = outer * (int64_t)n * inner;
    double *f = fac + (o * n) * inner + c;
    auto T = [&](int d, int i) { return __ldg(&tab[d * n + i]); };
    for (int par = 0; par < 2; ++par) {
        // previous two pivot rows of this parity: p1 = row i-2, p2 = row i-4
        double D1 = 1, U11 = 0, U21 = 0, FA1 = 0, FB1 = 0;
        double D2 = 1, U12 = 0, U22 = 0, FA2 = 0, FB2 = 0;
        for (int i = par; i < n; i += 2) {
            double lo2 = cc * T(8, i);
            double lo1 = b * T(5, i) + cc * T(9, i);
            double dg = a * T(0, i) + b * T(6, i) + cc * T(10, i);
            double up1 = a * T(1, i) + b * T(7, i) + cc * T(11, i);
            double up2 = a * T(2, i) + cc * T(12, i);
            double fa = a * T(3, i), fb = a * T(4, i);
            double l2 = 0.0, l1 = 0.0;
            if (i - 4 >= 0) {
                l2 = lo2 / D2;
                lo1 -= l2 * U12;
                dg -= l2 * U22;
                up1 -= l2 * bih_far(FA2, FB2, i + 2);
                up2 -= l2 * bih_far(FA2, FB2, i + 4);
                fa -= l2 * FA2;
                fb -= l2 * FB2;
            }
            if (i - 2 >= 0) {
                l1 = lo1 / D1;
                dg -= l1 * U11;
                up1 -= l1 * U21;
                up2 -= l1 * bih_far(FA1, FB1, i + 4);
                fa -= l1 * FA1;
                fb -= l1 * FB1;
            }
            const int64_t r = (int64_t)i * inner;
            f[0 * fs + r] = l1; f[1 * fs + r] = l2; f[2 * fs + r] = dg; f[3 * fs + r] = up1; f[4 * fs + r] = up2;
            f[5 * fs + r] = fa; f[6 * fs + r] = fb;
            D2 = D1; U12 = U11; U22 = U21; FA2 = FA1; FB2 = FB1;
            D1 = dg; U11 = up1; U21 = up2; FA1 = fa; FB1 = fb;
        }
    }
}

template <class T>
__global__ void __launch_bounds__(BD_THREADS) biharmonic_solve_kernel(
    const T *b, T *u /* may alias b: in-place solve */, const double *__restrict__ fac, int64_t outer, int64_t inner,
    int n, int n_u) {
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= outer * inner) return;
    const int64_t o = line / inner, c = line % inner;
    const int64_t fs = outer * (int64_t)n * inner;
    const double *f = fac + (o * n) * inner + c;
    const T *bl = b + (o * n_u) * inner + c;
    T *x = u + (o * n_u) * inner + c;
    // rows in batches of U: all loads of a batch are issued before the serial arithmetic
    constexpr int U = 4;
    for (int par = 0; par < 2; ++par) {
        T y1 = Num<T>::zero(), y2 = Num<T>::zero();
        for (int i0 = par; i0 < n; i0 += 2 * U) {
            T bb[U];
            double l1[U], l2[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int i = i0 + 2 * k;
                if (i < n) {
                    const int64_t r = (int64_t)i * inner;
                    bb[k] = bl[r];
                    l1[k] = f[0 * fs + r];
                    l2[k] = f[1 * fs + r];
                }
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int i = i0 + 2 * k;
                if (i < n) {
                    T v = bb[k];
                    v = Num<T>::fnma(l1[k], y1, v);
                    v = Num<T>::fnma(l2[k], y2, v);
                    x[(int64_t)i * inner] = v;
                    y2 = y1; y1 = v;
                }
            }
        }
        if (n - 1 - par < 0) continue;
        const int last = par + ((n - 1 - par) / 2) * 2;
        T x1 = Num<T>::zero(), x2 = Num<T>::zero(), x3 = Num<T>::zero();   // x_{i+2}, x_{i+4}, x_{i+6}
        T s1 = Num<T>::zero(), s2 = Num<T>::zero();
        for (int i0 = last; i0 >= par; i0 -= 2 * U) {
            T bb[U];
            double g[U][5];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int i = i0 - 2 * k;
                if (i >= par) {
                    const int64_t r = (int64_t)i * inner;
                    bb[k] = x[r];
#pragma unroll
                    for (int t = 0; t < 5; ++t) g[k][t] = f[(2 + t) * fs + r];
                }
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int i = i0 - 2 * k;
                if (i >= par) {
                    // tail sums over j >= i+6: add x_{i+6} (j = i+6)
                    if (i + 6 <= last) {
                        const int j = i + 6;
                        const double w = 1.0 / (double)(j + 3);
                        const double jp2 = (double)(j + 2);
                        s1 = Num<T>::fnma(-w, x3, s1);
                        s2 = Num<T>::fnma(-(w * jp2 * jp2), x3, s2);
                    }
                    T v = bb[k];
                    v = Num<T>::fnma(g[k][1], x1, v);
                    v = Num<T>::fnma(g[k][2], x2, v);
                    v = Num<T>::fnma(g[k][3], s1, v);
                    v = Num<T>::fnma(g[k][4], s2, v);
                    v = Num<T>::div(v, g[k][0]);
                    x[(int64_t)i * inner] = v;
                    x3 = x2; x2 = x1; x1 = v;
                }
            }
        }
    }
}

// c = (a*S + b*A + c*B) v with the same structure
template <class T>
__global__ void __launch_bounds__(BD_THREADS) biharmonic_matvec_kernel(
    const T *__restrict__ v, T *__restrict__ out, const double *__restrict__ tab, const double *__restrict__ sa,
    const double *__restrict__ sb, const double *__restrict__ sc, int64_t outer, int64_t inner, int n, int n_u) {
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= outer * inner) return;
    const int64_t o = line / inner, c = line % inner;
    const double a = sa[line], b = sb[line], cc = sc[line];
    const T *x = v + (o * n_u) * inner + c;
    T *y = out + (o * n_u) * inner + c;
    auto Tb = [&](int d, int i) { return __ldg(&tab[d * n + i]); };
    for (int i = n; i < n_u; ++i) y[(int64_t)i * inner] = Num<T>::zero();
    for (int par = 0; par < 2; ++par) {
        if (n - 1 - par < 0) continue;
        const int last = par + ((n - 1 - par) / 2) * 2;
        T x1 = Num<T>::zero(), x2 = Num<T>::zero(), x3 = Num<T>::zero();
        T s1 = Num<T>::zero(), s2 = Num<T>::zero();
        for (int i = last; i >= par; i -= 2) {
            if (i + 6 <= last) {
                const int j = i + 6;
                const double w = 1.0 / (double)(j + 3);
                const double jp2 = (double)(j + 2);
                s1 = Num<T>::fnma(-w, x3, s1);
                s2 = Num<T>::fnma(-(w * jp2 * jp2), x3, s2);
            }
            const T xi = x[(int64_t)i * inner];
            T acc = Num<T>::mul(a * Tb(0, i) + b * Tb(6, i) + cc * Tb(10, i), xi);
            acc = Num<T>::fnma(-(a * Tb(1, i) + b * Tb(7, i) + cc * Tb(11, i)), x1, acc);
            acc = Num<T>::fnma(-(a * Tb(2, i) + cc * Tb(12, i)), x2, acc);
            acc = Num<T>::fnma(-(a * Tb(3, i)), s1, acc);
            acc = Num<T>::fnma(-(a * Tb(4, i)), s2, acc);
            if (i - 2 >= 0) acc = Num<T>::fnma(-(b * Tb(5, i) + cc * Tb(9, i)), x[(int64_t)(i - 2) * inner], acc);
            if (i - 4 >= 0) acc = Num<T>::fnma(-(cc * Tb(8, i)), x[(int64_t)(i - 4) * inner], acc);
            y[(int64_t)i * inner] = acc;
            x3 = x2; x2 = x1; x1 = xi;
        }
    }
}

static inline unsigned bd_blocks(int64_t nlines) { return (unsigned)((nlines + BD_THREADS - 1) / BD_THREADS); }

template <int KL, int KU>
static int launch_factor(const BandedDesc &d, cudaStream_t st) {
    banded_factor_kernel<KL, KU><<<bd_blocks(d.outer * d.inner), BD_THREADS, 0, st>>>(d);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

template <int KL, int KU>
static int launch_solve(void *u, const double *fac, int64_t outer, int64_t inner, int n, int n_u, int step,
                        int fac_shared, int is_complex, cudaStream_t st) {
    if (is_complex)
        banded_solve_kernel<KL, KU, cplx><<<bd_blocks(outer * inner), BD_THREADS, 0, st>>>(
            (cplx *)u, fac, outer, inner, n, n_u, step, fac_shared);
    else
        banded_solve_kernel<KL, KU, double><<<bd_blocks(outer * inner), BD_THREADS, 0, st>>>(
            (double *)u, fac, outer, inner, n, n_u, step, fac_shared);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

#define SHB_BAND_DISPATCH(KLv, KUv, CALL)                                   \
    if (kl == 0 && ku == 0) { constexpr int KL = 0, KU = 0; CALL; }         \
    else if (kl == 0 && ku == 1) { constexpr int KL = 0, KU = 1; CALL; }    \
    else if (kl == 0 && ku == 2) { constexpr int KL = 0, KU = 2; CALL; }    \
    else if (kl == 1 && ku == 1) { constexpr int KL = 1, KU = 1; CALL; }    \
    else if (kl == 1 && ku == 2) { constexpr int KL = 1, KU = 2; CALL; }    \
    else if (kl == 2 && ku == 2) { constexpr int KL = 2, KU = 2; CALL; }    \
    else if (kl == 2 && ku == 4) { constexpr int KL = 2, KU = 4; CALL; }    \
    else if (kl == 1 && ku == 0) { constexpr int KL = 1, KU = 0; CALL; }    \
    else if (kl == 2 && ku == 0) { constexpr int KL = 2, KU = 0; CALL; }    \
    else if (kl == 2 && ku == 1) { constexpr int KL = 2, KU = 1; CALL; }    \
    else { set_error("banded: unsupported bandwidth kl=%d ku=%d", kl, ku); return SHB_ERR_UNSUPPORTED; }

}  // namespace shb

using namespace shb;

extern "C" {

int shb_banded_factor(int kl, int ku, int step, int nmat, const double *const *d_mats,
                      const double *const *d_scales, double *d_fac, int64_t outer, int64_t inner, int n,
                      int64_t pin_line, void *stream) {
    SHB_REQUIRE(nmat >= 1 && nmat <= BD_MAXMAT && d_mats && d_fac && n >= 1 && (step == 1 || step == 2),
                "bad arguments");
    if (outer * inner == 0) return SHB_OK;
    BandedDesc d;
    d.outer = outer; d.inner = inner; d.n = n; d.n_u = n; d.step = step; d.nmat = nmat;
    for (int m = 0; m < BD_MAXMAT; ++m) {
        d.mats[m] = m < nmat ? d_mats[m] : nullptr;
        d.scales[m] = (m < nmat && d_scales) ? d_scales[m] : nullptr;
    }
    d.fac = d_fac; d.fac_shared = 0; d.pin_line = pin_line;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SHB_BAND_DISPATCH(kl, ku, return (launch_factor<KL, KU>(d, st)));
}

int shb_banded_solve(int kl, int ku, int step, void *d_u, const double *d_fac, int64_t outer, int64_t inner,
                     int n, int n_u, int fac_shared, int is_complex, void *stream) {
    SHB_REQUIRE(d_u && d_fac && n >= 1 && n_u >= n && (step == 1 || step == 2), "bad arguments");
    if (outer * inner == 0) return SHB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (fac_shared == 1 || fac_shared == 0) {
        // chunk-parallel, tile-resident substitution (linesolve.cu): one factored matrix for every line
        // (shared table) or per-line factors in the batch layout
        const int rc = banded_tile_launch(kl, ku, step, d_u, d_fac, outer, inner, n, n_u, is_complex, st,
                                          fac_shared == 0);
        if (rc != SHB_ERR_UNSUPPORTED) return rc;
    }
    SHB_BAND_DISPATCH(kl, ku, return (launch_solve<KL, KU>(d_u, d_fac, outer, inner, n, n_u, step, fac_shared,
                                                           is_complex, st)));
}

int shb_banded_matvec(const void *d_u, void *d_out, int nmat, int ndiag, const int *offsets,
                      const double *const *d_mats, const double *const *d_scales, int64_t outer, int64_t inner,
                      int n_rows, int n_cols, int n_u, int n_out, int accumulate, int is_complex, void *stream) {
    SHB_REQUIRE(d_u && d_out && nmat >= 1 && nmat <= BD_MAXMAT && ndiag >= 1 && ndiag <= 16 && offsets && d_mats,
                "bad arguments");
    MatvecDesc d;
    d.outer = outer; d.inner = inner; d.n_rows = n_rows; d.n_cols = n_cols; d.n_u = n_u; d.n_out = n_out;
    d.nmat = nmat; d.ndiag = ndiag; d.accumulate = accumulate;
    for (int k = 0; k < ndiag; ++k) d.offs[k] = offsets[k];
    for (int m = 0; m < BD_MAXMAT; ++m) {
        d.mats[m] = m < nmat ? d_mats[m] : nullptr;
        d.scales[m] = (m < nmat && d_scales) ? d_scales[m] : nullptr;
    }
    const int64_t tot = outer * (int64_t)n_out * inner;
    if (tot == 0) return SHB_OK;
    if (inner >= 32 && outer <= 65535) {
        // rows per thread: a thread re-reads ndiag - 1 neighbour rows per block of R (SHB_MATVEC_R = 4 | 16; MKM step 98.2 -> 97.6 ms with 16)
        static const int rsel = getenv("SHB_MATVEC_R") ? atoi(getenv("SHB_MATVEC_R")) : 16;
        cudaStream_t st2 = reinterpret_cast<cudaStream_t>(stream);
        if (rsel == 16) {
            constexpr int R = 16;
            dim3 grid((unsigned)((inner + 127) / 128), (unsigned)((n_out + R - 1) / R), (unsigned)outer);
            if (is_complex) banded_matvec_col_kernel<cplx, R><<<grid, 128, 0, st2>>>((const cplx *)d_u, (cplx *)d_out, d);
            else banded_matvec_col_kernel<double, R><<<grid, 128, 0, st2>>>((const double *)d_u, (double *)d_out, d);
        } else {
            constexpr int R = 4;
            dim3 grid((unsigned)((inner + 127) / 128), (unsigned)((n_out + R - 1) / R), (unsigned)outer);
            if (is_complex) banded_matvec_col_kernel<cplx, R><<<grid, 128, 0, st2>>>((const cplx *)d_u, (cplx *)d_out, d);
            else banded_matvec_col_kernel<double, R><<<grid, 128, 0, st2>>>((const double *)d_u, (double *)d_out, d);
        }
        SHB_CUDA(cudaGetLastError());
        return SHB_OK;
    }
    int64_t blocks = (tot + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (is_complex) banded_matvec_kernel<cplx><<<(unsigned)blocks, 256, 0, st>>>((const cplx *)d_u, (cplx *)d_out, d);
    else banded_matvec_kernel<double><<<(unsigned)blocks, 256, 0, st>>>((const double *)d_u, (double *)d_out, d);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

int shb_tail_matvec(const void *d_u, void *d_out, const double *d_r, const double *d_t, const double *d_scale, int w,
                    int64_t outer, int64_t inner, int n_rows, int n_cols, int n_u, int n_out, int accumulate,
                    int is_complex, void *stream) {
    SHB_REQUIRE(d_u && d_out && d_u != d_out && d_r && d_t && n_rows >= 1 && n_cols >= 1 && n_u >= n_cols &&
                n_out >= n_rows, "bad arguments");
    if (outer * inner == 0) return SHB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (is_complex)
        tail_matvec_kernel<cplx><<<bd_blocks(outer * inner), BD_THREADS, 0, st>>>(
            (const cplx *)d_u, (cplx *)d_out, d_r, d_t, d_scale, w, outer, inner, n_rows, n_cols, n_u, n_out, accumulate);
    else
        tail_matvec_kernel<double><<<bd_blocks(outer * inner), BD_THREADS, 0, st>>>(
            (const double *)d_u, (double *)d_out, d_r, d_t, d_scale, w, outer, inner, n_rows, n_cols, n_u, n_out,
            accumulate);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

int shb_helmholtz_factor(const double *d_A0, const double *d_A2, const double *d_A4, const double *d_Bm2,
                         const double *d_B0, const double *d_B2, const double *d_alfa, const double *d_beta,
                         double *d_d0, double *d_d1, double *d_d2, double *d_L, int64_t outer, int64_t inner,
                         int n, int n_fac, int neumann, void *stream) {
    SHB_REQUIRE(d_A0 && d_A2 && d_A4 && d_Bm2 && d_B0 && d_B2 && d_alfa && d_beta && d_d0 && d_d1 && d_d2 && d_L &&
                n >= 1 && n_fac >= n, "bad arguments");
    if (outer * inner == 0) return SHB_OK;
    helmholtz_factor_kernel<<<bd_blocks(outer * inner), BD_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_A0, d_A2, d_A4, d_Bm2, d_B0, d_B2, d_alfa, d_beta, d_d0, d_d1, d_d2, d_L, outer, inner, n, n_fac, neumann);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

int shb_helmholtz_solve(const void *d_b, void *d_u, const double *d_d0, const double *d_d1, const double *d_d2,
                        const double *d_L, int64_t outer, int64_t inner, int n, int n_fac, int n_u, int is_complex,
                        int flags, void *stream) {
    SHB_REQUIRE(d_b && d_u && d_d0 && d_d1 && d_d2 && d_L && n >= 1 && n_u >= n && n_fac >= n, "bad arguments");
    if (outer * inner == 0) return SHB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (is_complex)
        helmholtz_solve_kernel<cplx><<<bd_blocks(outer * inner), BD_THREADS, 0, st>>>(
            (const cplx *)d_b, (cplx *)d_u, d_d0, d_d1, d_d2, d_L, outer, inner, n, n_fac, n_u, flags);
    else
        helmholtz_solve_kernel<double><<<bd_blocks(outer * inner), BD_THREADS, 0, st>>>(
            (const double *)d_b, (double *)d_u, d_d0, d_d1, d_d2, d_L, outer, inner, n, n_fac, n_u, flags);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

int shb_helmholtz_matvec(const void *d_v, void *d_c, const double *d_A0, const double *d_A2, const double *d_A4,
                         const double *d_Bm2, const double *d_B0, const double *d_B2, const double *d_alfa,
                         const double *d_beta, int64_t outer, int64_t inner, int n, int n_u, int is_complex,
                         void *stream) {
    SHB_REQUIRE(d_v && d_c && d_v != d_c && d_A0 && d_A2 && d_A4 && d_Bm2 && d_B0 && d_B2 && d_alfa && d_beta &&
                n >= 1 && n_u >= n, "bad arguments");
    if (outer * inner == 0) return SHB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (is_complex)
        helmholtz_matvec_kernel<cplx><<<bd_blocks(outer * inner), BD_THREADS, 0, st>>>(
            (const cplx *)d_v, (cplx *)d_c, d_A0, d_A2, d_A4, d_Bm2, d_B0, d_B2, d_alfa, d_beta, outer, inner, n, n_u);
    else
        helmholtz_matvec_kernel<double><<<bd_blocks(outer * inner), BD_THREADS, 0, st>>>(
            (const double *)d_v, (double *)d_c, d_A0, d_A2, d_A4, d_Bm2, d_B0, d_B2, d_alfa, d_beta, outer, inner, n, n_u);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

int shb_biharmonic_factor(const double *d_tab, const double *d_a, const double *d_b, const double *d_c,
                          double *d_fac, int64_t outer, int64_t inner, int n, void *stream) {
    SHB_REQUIRE(d_tab && d_a && d_b && d_c && d_fac && n >= 1, "bad arguments");
    if (outer * inner == 0) return SHB_OK;
    biharmonic_factor_kernel<<<bd_blocks(outer * inner), BD_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_tab, d_a, d_b, d_c, d_fac, outer, inner, n);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

int shb_biharmonic_solve(const void *d_b, void *d_u, const double *d_fac, int64_t outer, int64_t inner, int n,
                         int n_u, int is_complex, void *stream) {
    SHB_REQUIRE(d_b && d_u && d_fac && n >= 1 && n_u >= n, "bad arguments");
    if (outer * inner == 0) return SHB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (is_complex)
        biharmonic_solve_kernel<cplx><<<bd_blocks(outer * inner), BD_THREADS, 0, st>>>(
            (const cplx *)d_b, (cplx *)d_u, d_fac, outer, inner, n, n_u);
    else
        biharmonic_solve_kernel<double><<<bd_blocks(outer * inner), BD_THREADS, 0, st>>>(
            (const double *)d_b, (double *)d_u, d_fac, outer, inner, n, n_u);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

int shb_biharmonic_matvec(const void *d_v, void *d_c, const double *d_tab, const double *d_a, const double *d_b,
                          const double *d_cs, int64_t outer, int64_t inner, int n, int n_u, int is_complex,
                          void *stream) {
    SHB_REQUIRE(d_v && d_c && d_v != d_c && d_tab && d_a && d_b && d_cs && n >= 1 && n_u >= n, "bad arguments");
    if (outer * inner == 0) return SHB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (is_complex)
        biharmonic_matvec_kernel<cplx><<<bd_blocks(outer * inner), BD_THREADS, 0, st>>>(
            (const cplx *)d_v, (cplx *)d_c, d_tab, d_a, d_b, d_cs, outer, inner, n, n_u);
    else
        biharmonic_matvec_kernel<double><<<bd_blocks(outer * inner), BD_THREADS, 0, st>>>(
            (const double *)d_v, (double *)d_c, d_tab, d_a, d_b, d_cs, outer, inner, n, n_u);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

}  // extern "C"
