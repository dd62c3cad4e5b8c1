// linesolve.cu -- chunk-parallel line solvers: every line of [outer][n][inner] is solved by
// n / (2*RP) threads, the whole line resident in registers between the two sweeps.
//
//   * helm_chunked_kernel   : Chebyshev-Galerkin Helmholtz (alfa*A + beta*B) u = b from ONE stored
//       array per line (1/d0); the other three factor arrays of the reference
//       (d1, d2, L: chebyshev/la.py:270-274, la.pyx:744-756) are regenerated on the fly from the
//       O(N) diagonals and the per-line (alfa, beta):  L_i = beta*Bm2[i-2]/d0[i-2] needs no
//       recurrence, and d2_i = alfa*A4_i - L_i d2_{i-2} is the SAME first-order recurrence as the
//       forward substitution y_i = b_i - L_i y_{i-2}, so it rides along as a third right-hand side.
//       HBM traffic: b read once, u written once, 1/d0 read once = 1.25x the algorithmic bytes
//       (the thread-per-line kernel in banded.cu moves 3x: four factor arrays and y out and back).
//   * banded_chunked_kernel : forward / back substitution with the LU factors of banded.cu
//       (TDMA / PDMA / FDMA / ..., shared or per-line factors), same skeleton.
//
// Thread layout of a CTA: LW adjacent lines (lanes along the contiguous `inner` index: LW*16 B
// segments) x C chunks of RT = step*RP consecutive rows.  Sweeps:
//   pass 1  chunk map with unknown entering state  ->  shared memory
//   scan    exclusive scan of the chunk maps per (line, parity)   (linesolve.cuh: amap_scan)
//   pass 2  the recurrence over the thread's rows from the entering state
// once upwards (forward substitution) and once downwards (back substitution).
//
// Reference semantics restated: shenfun/chebyshev/la.py:131-351, optimization/cython/la.pyx:708-947
// (Helmholtz), shenfun/la.py:291-820 + la.pyx:199-611 (banded).  The elimination order within a
// line is the reference's; only the association of the recurrence changes (chunk maps), which
// moves results by O(eps).
#include "linesolve.cuh"
#include <stdlib.h>

namespace shb {

struct HelmDesc {
    const void *b;
    void *u;
    const double2 *tabF;                  // [n_pad] {Bm2[i-2] (0 for i < 2), A4[i]}, zero for i >= n
    const double2 *tabB;                  // [n_pad] {A2[i], B2[i]}, zero for i >= n
    const double *alfa, *beta;            // [outer*inner]
    const double *invd0;                  // [outer][n_fac][inner]
    int64_t outer, inner;
    int n, n_u, n_fac;
    int C;                                // chunks per line
    int ldc;                              // padded chunk stride of the shared-memory tables
    int64_t tiles_per_outer;
};

template <class T, int RP, int LW, int CL, int MAXT>
__global__ void __launch_bounds__(MAXT) helm_chunked_kernel(const __grid_constant__ HelmDesc d) {
    using V = Val<T>;
    constexpr int NC = V::NC;
    constexpr int RT = 2 * RP;            // rows per thread (both parities)
    constexpr int NP = 2 * LW;            // independent recurrences per CTA
    constexpr int NBF = NC + 1;           // forward right-hand sides: y (re, im) and d2
    extern __shared__ double sm[];
    const int C = d.C, ldc = d.ldc;
    double *maps = sm;                                      // max(E_fwd, E_bwd) * NP * ldc
    double *states = sm + (size_t)(4 + 2 * NC) * NP * ldc;  // backward map has 4 + 2*NC entries

    const int l = threadIdx.x % LW, c = threadIdx.x / LW;
    const int64_t o = blockIdx.x / d.tiles_per_outer;
    const int64_t cc = (blockIdx.x % d.tiles_per_outer) * LW + l;
    const bool lane_ok = (cc < d.inner) && (c < C);
    const int n = d.n;
    const int r0 = c * RT;
    const int nrow = lane_ok ? min(max(n - r0, 0), RT) : 0;     // valid rows of this thread
    const int64_t line = o * d.inner + cc;
    const double al = lane_ok ? d.alfa[line] : 0.0, be = lane_ok ? d.beta[line] : 0.0;
    const int64_t rs = d.inner;                                  // row stride (elements)

    const T *bp = reinterpret_cast<const T *>(d.b) + ((o * d.n_u + r0) * rs + cc);
    T *up = reinterpret_cast<T *>(d.u) + ((o * d.n_u + r0) * rs + cc);
    const double *ivq = d.invd0 + ((o * d.n_fac + r0) * rs + cc);

    // ---- loads: the thread's RT rows of the right-hand side and of 1/d0, plus the two 1/d0 above
    T x[RT];
    double iv[RT], ivp[2];
#pragma unroll
    for (int k = 0; k < RT; ++k) {
        const bool ok = k < nrow;
        x[k] = ok ? bp[k * rs] : V::zero();
        iv[k] = ok ? ivq[k * rs] : 1.0;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) ivp[j] = (nrow > 0 && r0 >= 2) ? ivq[(j - 2) * rs] : 0.0;
    // multipliers L_i = beta * Bm2[i-2] / d0[i-2] and a4_i = alfa * A4[i] (tables are zero beyond n)
    double Lk[RT], a4[RT];
    const double2 *tF = d.tabF + r0;
#pragma unroll
    for (int k = 0; k < RT; ++k) {
        const double2 t = (c < C) ? __ldg(&tF[k]) : make_double2(0.0, 0.0);
        const double ivm2 = (k >= 2) ? iv[k >= 2 ? k - 2 : 0] : ivp[k & 1];
        Lk[k] = be * t.x * ivm2;
        a4[k] = al * t.y;
    }

    // ---- forward substitution: y_i = b_i - L_i y_{i-2};  d2raw_i = a4_i - L_i d2raw_{i-2}
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        AMap<1, NBF> m;
        T yh = V::zero();
        double Dh = 0.0, M = 1.0;
#pragma unroll
        for (int q = 0; q < RP; ++q) {
            const int k = p + 2 * q;
            yh = V::fnma(Lk[k], yh, x[k]);
            Dh = fma(-Lk[k], Dh, a4[k]);
            M *= -Lk[k];
        }
        m.A[0] = M;
#pragma unroll
        for (int v = 0; v < NC; ++v) m.b[v] = V::get(yh, v);
        m.b[NC] = Dh;
        if (c < C) amap_store(maps, NP, ldc, p * LW + l, c, m);
    }
    __syncthreads();
    amap_scan<1, NBF, CL>(maps, states, NP, C, ldc, false);
    __syncthreads();
    double Dp[RT];                         // d2 of row i-2 (what d1_i and d2_i are eliminated with)
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        T y = V::zero();
        double Dc = 0.0;
        if (c < C) {
#pragma unroll
            for (int v = 0; v < NC; ++v) V::set(y, v, amap_state(states, NP, ldc, p * LW + l, c, v));
            Dc = amap_state(states, NP, ldc, p * LW + l, c, NC);
        }
#pragma unroll
        for (int q = 0; q < RP; ++q) {
            const int k = p + 2 * q;
            Dp[k] = Dc;
            y = V::fnma(Lk[k], y, x[k]);
            x[k] = y;
            Dc = fma(-Lk[k], Dc, a4[k]);
        }
    }
    __syncthreads();                        // maps / states are reused by the backward sweep

    // ---- back substitution: u_i = (y_i - d1_i u_{i+2} - d2_i sum_{q>=2} u_{i+2q}) / d0_i
    // d1_i = alfa*A2_i + beta*B2_i - L_i d2_{i-2} (i < n-2), d2_i = a4_i - L_i d2_{i-2} (i < n-4); rows at or
    // beyond n have x = 0 and (d1, d2) = 0, so they leave the state (u_{i+2}, sum) at zero
    double d1[RT], d2[RT];
    const double2 *tB = d.tabB + r0;
#pragma unroll
    for (int k = 0; k < RT; ++k) {
        const double2 t = (c < C) ? __ldg(&tB[k]) : make_double2(0.0, 0.0);
        const int i = r0 + k;
        d1[k] = (k < nrow && i < n - 2) ? fma(-Lk[k], Dp[k], fma(al, t.x, be * t.y)) : 0.0;
        d2[k] = (k < nrow && i < n - 4) ? fma(-Lk[k], Dp[k], a4[k]) : 0.0;
    }
    constexpr int NBB = NC;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        AMap<2, NBB> m;
        T un = V::zero(), sg = V::zero();
        double Hu[2] = {1.0, 0.0}, Hs[2] = {0.0, 1.0};
#pragma unroll
        for (int q = RP - 1; q >= 0; --q) {
            const int k = p + 2 * q;
            if (k < nrow) {
                T v = V::fnma(d1[k], un, x[k]);
                v = V::fnma(d2[k], sg, v);
                v = V::mul(iv[k], v);
                sg = V::add(sg, un);
                un = v;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const double hv = -(fma(d1[k], Hu[j], d2[k] * Hs[j])) * iv[k];
                    Hs[j] += Hu[j];
                    Hu[j] = hv;
                }
            }
        }
        m.A[0] = Hu[0]; m.A[1] = Hu[1]; m.A[2] = Hs[0]; m.A[3] = Hs[1];
#pragma unroll
        for (int v = 0; v < NC; ++v) { m.b[v * 2] = V::get(un, v); m.b[v * 2 + 1] = V::get(sg, v); }
        if (c < C) amap_store(maps, NP, ldc, p * LW + l, c, m);
    }
    __syncthreads();
    amap_scan<2, NBB, CL>(maps, states, NP, C, ldc, true);
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        T un = V::zero(), sg = V::zero();
        if (c < C) {
#pragma unroll
            for (int v = 0; v < NC; ++v) {
                V::set(un, v, amap_state(states, NP, ldc, p * LW + l, c, v * 2));
                V::set(sg, v, amap_state(states, NP, ldc, p * LW + l, c, v * 2 + 1));
            }
        }
#pragma unroll
        for (int q = RP - 1; q >= 0; --q) {
            const int k = p + 2 * q;
            if (k < nrow) {
                T v = V::fnma(d1[k], un, x[k]);
                v = V::fnma(d2[k], sg, v);
                v = V::mul(iv[k], v);
                sg = V::add(sg, un);
                un = v;
                x[k] = v;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < RT; ++k)
        if (k < nrow) up[k * rs] = x[k];
}

// 1/d0 per line: the d0 recurrence of la.pyx:744-756 (the only one with a division), one thread per
// line; d1 and d2 are carried in registers only.  Run once per solver.
__global__ void __launch_bounds__(128) helm_factor_inv_kernel(
    const double *__restrict__ A0, const double *__restrict__ A2, const double *__restrict__ A4,
    const double *__restrict__ Bm2, const double *__restrict__ B0, const double *__restrict__ B2,
    const double *__restrict__ alfa, const double *__restrict__ beta, double *__restrict__ inv,
    int64_t outer, int64_t inner, int n) {
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= outer * inner) return;
    const int64_t o = line / inner, c = line % inner;
    const double al = alfa[line], be = beta[line];
    inv += (o * n) * inner + c;
    double pd0[2] = {1.0, 1.0}, pd1[2] = {0.0, 0.0}, pd2[2] = {0.0, 0.0};
    for (int i = 0; i < n; ++i) {
        const int par = i & 1;
        double d0 = al * __ldg(&A0[i]) + be * __ldg(&B0[i]);
        double d1 = (i < n - 2) ? al * __ldg(&A2[i]) + be * __ldg(&B2[i]) : 0.0;
        double d2 = (i < n - 4) ? al * __ldg(&A4[i]) : 0.0;
        if (i >= 2) {
            const double L = be * __ldg(&Bm2[i - 2]) / pd0[par];
            d0 -= L * pd1[par];
            if (i < n - 2) d1 -= L * pd2[par];
            if (i < n - 4) d2 -= L * pd2[par];
        }
        inv[(int64_t)i * inner] = 1.0 / d0;
        pd0[par] = d0; pd1[par] = d1; pd2[par] = d2;
    }
}

template <class T, int RP, int LW, int CL, int MAXT>
static int launch_helm_cl(HelmDesc d, int threads, size_t smem, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        SHB_CUDA(cudaFuncSetAttribute(helm_chunked_kernel<T, RP, LW, CL, MAXT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set = true;
    }
    const int64_t grid = d.outer * d.tiles_per_outer;
    helm_chunked_kernel<T, RP, LW, CL, MAXT><<<(unsigned)grid, threads, smem, st>>>(d);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

template <class T, int RP, int LW, int MAXT>
static int launch_helm(HelmDesc d, cudaStream_t st) {
    constexpr int RT = 2 * RP;
    constexpr int NC = Val<T>::NC;
    d.C = (d.n + RT - 1) / RT;
    int ldc = d.C;
    while (ldc % 16 != 16 / LW) ++ldc;
    d.ldc = ldc;
    d.tiles_per_outer = (d.inner + LW - 1) / LW;
    const int threads = ((d.C * LW + 31) / 32) * 32;
    if (threads > MAXT || d.C > 128) return SHB_ERR_UNSUPPORTED;
    const size_t smem = (size_t)((4 + 2 * NC) + 2 * NC) * (2 * LW) * ldc * sizeof(double);
    if (smem > 160 * 1024) return SHB_ERR_UNSUPPORTED;
    if (d.C <= 32) return launch_helm_cl<T, RP, LW, 1, MAXT>(d, threads, smem, st);
    if (d.C <= 64) return launch_helm_cl<T, RP, LW, 2, MAXT>(d, threads, smem, st);
    return launch_helm_cl<T, RP, LW, 4, MAXT>(d, threads, smem, st);
}

int helm_rows_pad(int n) { return ((n + 15) / 16) * 16 + 16; }

}  // namespace shb

using namespace shb;

extern "C" {

int shb_helmholtz_factor_lean(const double *d_A0, const double *d_A2, const double *d_A4, const double *d_Bm2,
                              const double *d_B0, const double *d_B2, const double *d_alfa, const double *d_beta,
                              double *d_invd0, int64_t outer, int64_t inner, int n, void *stream) {
    SHB_REQUIRE(d_A0 && d_A2 && d_A4 && d_Bm2 && d_B0 && d_B2 && d_alfa && d_beta && d_invd0 && n >= 1,
                "bad arguments");
    if (outer * inner == 0) return SHB_OK;
    helm_factor_inv_kernel<<<(unsigned)((outer * inner + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_A0, d_A2, d_A4, d_Bm2, d_B0, d_B2, d_alfa, d_beta, d_invd0, outer, inner, n);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

int shb_helmholtz_lean_table_rows(int n) { return helm_rows_pad(n); }

int shb_helmholtz_solve_lean(const void *d_b, void *d_u, const double *d_tabF, const double *d_tabB,
                             const double *d_alfa, const double *d_beta, const double *d_invd0, int64_t outer,
                             int64_t inner, int n, int n_u, int is_complex, void *stream) {
    SHB_REQUIRE(d_b && d_u && d_tabF && d_tabB && d_alfa && d_beta && d_invd0 && n >= 1 && n_u >= n,
                "bad arguments");
    if (outer * inner == 0) return SHB_OK;
    HelmDesc d;
    d.b = d_b; d.u = d_u; d.tabF = reinterpret_cast<const double2 *>(d_tabF);
    d.tabB = reinterpret_cast<const double2 *>(d_tabB); d.alfa = d_alfa; d.beta = d_beta;
    d.invd0 = d_invd0; d.outer = outer; d.inner = inner; d.n = n; d.n_u = n_u; d.n_fac = n;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    static const int variant = getenv("SHB_HELM_VARIANT") ? atoi(getenv("SHB_HELM_VARIANT")) : 0;
    int rc;
    if (is_complex) {
        if (variant == 1 && n <= 512) rc = launch_helm<cplx, 4, 8, 512>(d, st);
        else rc = launch_helm<cplx, 4, 4, 512>(d, st);
    } else {
        rc = (n <= 512) ? launch_helm<double, 4, 8, 512>(d, st) : launch_helm<double, 4, 4, 512>(d, st);
    }
    if (rc == SHB_ERR_UNSUPPORTED) set_error("helmholtz_solve_lean: line length n=%d not covered", n);
    return rc;
}

}  // extern "C"
