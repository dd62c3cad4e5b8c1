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

template <class T> __device__ __forceinline__ void cp_async_elem(T *dst, const T *src, bool ok);
template <> __device__ __forceinline__ void cp_async_elem<cplx>(cplx *dst, const cplx *src, bool ok) { cp_async_16(dst, src, ok); }
template <> __device__ __forceinline__ void cp_async_elem<double>(double *dst, const double *src, bool ok) { cp_async_8(dst, src, ok); }

// ---- Helmholtz, tile-resident variant ------------------------------------------------------------
// A CTA stages LW adjacent lines x all rows in shared memory (cp.async, 16 B per copy, LW*16 B
// contiguous per row) together with their 1/d0, and a thread owns RP rows of ONE parity of ONE
// line.  The NCH chunks of a (line, parity) recurrence sit in NCH adjacent lanes, so the chunk maps are
// scanned with shuffles inside a lane segment (amap_seg_scan): no shared-memory maps, and only two
// CTA barriers in the kernel (tile loaded / tile ready to be written out).  The rows of a thread are
// 2*RP rows apart from its neighbour's, so the line index is rotated by the chunk index in shared
// memory ((l + c) mod LW): a quarter-warp reads 8 distinct 16 B slots.
template <class T, int LW, int RP, int NCH, int MINB, bool BULK>
__global__ void __launch_bounds__(2 * LW * NCH, MINB) helm_tile_kernel(const __grid_constant__ HelmDesc d) {
    using V = Val<T>;
    constexpr int NC = V::NC;
    constexpr int NT = 2 * LW * NCH;          // threads
    constexpr int RC = 2 * RP;                // rows of a chunk (both parities)
    constexpr int NR = RC * NCH;              // rows of the tile (>= n)
    constexpr int NBF = NC + 1;
    // BULK (complex data): rows are moved by the bulk-copy engine (cp.async.bulk, one LW*16 B row per
    // copy, both directions), so the tile traffic costs no LSU wavefronts.  A row must then stay
    // contiguous in shared memory: instead of rotating the line index, the rows of chunk c start c
    // entries (16 B) later -- the same bank spread for the lanes of a segment.
    constexpr int XSZ = NR * LW + (BULK ? NCH - 1 : 0), ISZ = NR * LW + (BULK ? 2 * (NCH - 1) : 0);
    extern __shared__ __align__(16) unsigned char smraw[];
    T *xs = reinterpret_cast<T *>(smraw);                      // [NR][LW]
    double *ivs = reinterpret_cast<double *>(xs + XSZ);        // [NR][LW]
    // the O(N) diagonals, skewed by one entry per chunk: the lanes of a segment read rows RC apart
    // (lanes run along the chunks), which from global memory costs one wavefront per lane
    double2 *tFs = reinterpret_cast<double2 *>(ivs + ISZ);      // [NR + NCH]
    double2 *tBs = tFs + (NR + NCH);
    unsigned long long &mbar = *reinterpret_cast<unsigned long long *>(tBs + (NR + NCH));   // BULK only
    const int n = d.n;
    const int64_t o = blockIdx.x / d.tiles_per_outer;
    const int64_t cc0 = (blockIdx.x % d.tiles_per_outer) * LW;
    const int64_t rs = d.inner;
    const T *bb = reinterpret_cast<const T *>(d.b) + (o * d.n_u) * rs + cc0;
    T *ub = reinterpret_cast<T *>(d.u) + (o * d.n_u) * rs + cc0;
    const double *ivb = d.invd0 + (o * d.n_fac) * rs + cc0;
    const int w = (d.inner - cc0 < LW) ? (int)(d.inner - cc0) : LW;     // lines of this tile inside the array
    const bool iv_bulk = BULK && (d.inner % 2 == 0);                    // 8-byte rows: 16 B alignment needs an even pitch

    // ---- stage the tile
    if constexpr (BULK) {
        const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
        if (threadIdx.x == 0) {
            const unsigned tx = (unsigned)n * (unsigned)w * (unsigned)(sizeof(T) + (iv_bulk ? sizeof(double) : 0));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(tx) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        // entries the copies do not write: rows >= n, lines outside the array
        for (int idx = threadIdx.x; idx < NR * LW; idx += NT) {
            const int i = idx / LW, l = idx % LW;
            if (i >= n || l >= w) {
                xs[i * LW + i / RC + l] = V::zero();
                ivs[i * LW + 2 * (i / RC) + l] = 0.0;
            } else if (!iv_bulk) {
                cp_async_8(&ivs[i * LW + 2 * (i / RC) + l], ivb + (int64_t)i * rs + l, true);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += NT) {
            const unsigned dx = (unsigned)__cvta_generic_to_shared(&xs[i * LW + i / RC]);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dx), "l"(bb + (int64_t)i * rs), "r"((unsigned)(w * sizeof(T))), "r"(mb) : "memory");
            if (iv_bulk) {
                const unsigned di = (unsigned)__cvta_generic_to_shared(&ivs[i * LW + 2 * (i / RC)]);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(di), "l"(ivb + (int64_t)i * rs), "r"((unsigned)(w * sizeof(double))), "r"(mb) : "memory");
            }
        }
    } else {
        for (int idx = threadIdx.x; idx < NR * LW; idx += NT) {
            const int i = idx / LW, l = idx % LW;
            const bool ok = (i < n) && (cc0 + l < d.inner);
            const int slot = i * LW + ((l + i / RC) % LW);
            cp_async_elem<T>(&xs[slot], ok ? bb + (int64_t)i * rs + l : bb, ok);
            cp_async_8(&ivs[slot], ok ? (const void *)(ivb + (int64_t)i * rs + l) : (const void *)ivb, ok);
        }
    }
    for (int i = threadIdx.x; i < NR; i += NT) {
        cp_async_16(&tFs[i + i / RC], &d.tabF[i], true);
        cp_async_16(&tBs[i + i / RC], &d.tabB[i], true);
    }
    const int c = threadIdx.x % NCH;
    const int pr = threadIdx.x / NCH;         // (line, parity)
    const int p = pr & 1, l = pr >> 1;
    const bool line_ok = cc0 + l < d.inner;
    const int64_t line = o * d.inner + cc0 + l;
    const double al = line_ok ? d.alfa[line] : 0.0, be = line_ok ? d.beta[line] : 0.0;
    const int i0 = c * RC + p;                // first row of this thread; rows i0 + 2q
    // this chunk's slot of line l, and the slot in the chunk below (row i0 - 2 when q == 0)
    const int sw = BULK ? c + l : (l + c) % LW;
    const int swi = BULK ? 2 * c + l : sw;
    const int swim = BULK ? 2 * (c - 1) + l : (l + c + LW - 1) % LW;
    cp_async_wait_all();
    if constexpr (BULK) {
        const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(mb) : "memory");
        }
    }
    __syncthreads();

    // ---- forward substitution
    double Lk[RP], a4[RP], Dp[RP];
    {
        AMap<1, NBF> m;
        T yh = V::zero();
        double Dh = 0.0, M = 1.0;
#pragma unroll
        for (int q = 0; q < RP; ++q) {
            const int i = i0 + 2 * q;
            const double2 t = tFs[i + c];
            const double ivm2 = (q == 0) ? ((i >= 2) ? ivs[(i - 2) * LW + swim] : 0.0) : ivs[(i - 2) * LW + swi];
            Lk[q] = be * t.x * ivm2;
            a4[q] = al * t.y;
            yh = V::fnma(Lk[q], yh, xs[i * LW + sw]);
            Dh = fma(-Lk[q], Dh, a4[q]);
            M *= -Lk[q];
        }
        m.A[0] = M;
#pragma unroll
        for (int v = 0; v < NC; ++v) m.b[v] = V::get(yh, v);
        m.b[NC] = Dh;
        double s[NBF];
        amap_seg_scan<1, NBF, NCH, false>(m, s, c);
        T y = V::zero();
#pragma unroll
        for (int v = 0; v < NC; ++v) V::set(y, v, s[v]);
        double Dc = s[NC];
#pragma unroll
        for (int q = 0; q < RP; ++q) {
            const int i = i0 + 2 * q;
            Dp[q] = Dc;
            y = V::fnma(Lk[q], y, xs[i * LW + sw]);
            xs[i * LW + sw] = y;
            Dc = fma(-Lk[q], Dc, a4[q]);
        }
    }
    // ---- back substitution; (Lk, Dp, a4) are recycled as (d1, d2, 1/d0)
    {
#pragma unroll
        for (int q = 0; q < RP; ++q) {
            const int i = i0 + 2 * q;
            const double2 t = tBs[i + c];
            const double e1 = (i < n - 2) ? fma(-Lk[q], Dp[q], fma(al, t.x, be * t.y)) : 0.0;
            const double e2 = (i < n - 4) ? fma(-Lk[q], Dp[q], a4[q]) : 0.0;
            Lk[q] = e1;
            Dp[q] = e2;
            a4[q] = ivs[i * LW + swi];
        }
        AMap<2, NC> m;
        T un = V::zero(), sg = V::zero();
        double Hu[2] = {1.0, 0.0}, Hs[2] = {0.0, 1.0};
#pragma unroll
        for (int q = RP - 1; q >= 0; --q) {
            const int i = i0 + 2 * q;
            T v = V::fnma(Lk[q], un, xs[i * LW + sw]);
            v = V::fnma(Dp[q], sg, v);
            v = V::mul(a4[q], v);
            sg = V::add(sg, un);
            un = v;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const double hv = -(fma(Lk[q], Hu[j], Dp[q] * Hs[j])) * a4[q];
                Hs[j] += Hu[j];
                Hu[j] = hv;
            }
        }
        m.A[0] = Hu[0]; m.A[1] = Hu[1]; m.A[2] = Hs[0]; m.A[3] = Hs[1];
#pragma unroll
        for (int v = 0; v < NC; ++v) { m.b[v * 2] = V::get(un, v); m.b[v * 2 + 1] = V::get(sg, v); }
        double s[2 * NC];
        amap_seg_scan<2, NC, NCH, true>(m, s, c);
#pragma unroll
        for (int v = 0; v < NC; ++v) { V::set(un, v, s[v * 2]); V::set(sg, v, s[v * 2 + 1]); }
#pragma unroll
        for (int q = RP - 1; q >= 0; --q) {
            const int i = i0 + 2 * q;
            T v = V::fnma(Lk[q], un, xs[i * LW + sw]);
            v = V::fnma(Dp[q], sg, v);
            v = V::mul(a4[q], v);
            sg = V::add(sg, un);
            un = v;
            xs[i * LW + sw] = v;
        }
    }
    if constexpr (BULK) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // ---- write the tile out (rows < n, lines inside the array)
    if constexpr (BULK) {
        for (int i = threadIdx.x; i < n; i += NT) {
            const unsigned sx = (unsigned)__cvta_generic_to_shared(&xs[i * LW + i / RC]);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(ub + (int64_t)i * rs), "r"(sx), "r"((unsigned)(w * sizeof(T))) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    } else {
        for (int idx = threadIdx.x; idx < NR * LW; idx += NT) {
            const int i = idx / LW, ll = idx % LW;
            if (i < n && cc0 + ll < d.inner) ub[(int64_t)i * rs + ll] = xs[i * LW + ((ll + i / RC) % LW)];
        }
    }
}

template <class T, int LW, int RP, int NCH, int MINB, bool BULK>
static int launch_helm_tile_b(HelmDesc d, cudaStream_t st) {
    constexpr int NR = 2 * RP * NCH;
    constexpr size_t smem = (size_t)(NR * LW + (BULK ? NCH - 1 : 0)) * sizeof(T) +
                            (size_t)(NR * LW + (BULK ? 2 * (NCH - 1) : 0)) * sizeof(double) +
                            2 * (size_t)(NR + NCH) * sizeof(double2) + (BULK ? 16 : 0);
    static bool attr_set = false;
    if (!attr_set) {
        SHB_CUDA(cudaFuncSetAttribute(helm_tile_kernel<T, LW, RP, NCH, MINB, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        attr_set = true;
    }
    d.tiles_per_outer = (d.inner + LW - 1) / LW;
    const int64_t grid = d.outer * d.tiles_per_outer;
    helm_tile_kernel<T, LW, RP, NCH, MINB, BULK><<<(unsigned)grid, 2 * LW * NCH, smem, st>>>(d);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

template <class T, int LW, int RP, int NCH, int MINB = 1>
static int launch_helm_tile(const HelmDesc &d, cudaStream_t st) {
    // SHB_HELM_BULK=1: complex tiles travel through the bulk-copy engine.  Measured slower on B200 (512^3:
    // 1.005 vs 0.759 ms; 1024 x 192 x 512: 1.49 vs 0.92 ms): one 128-byte row per copy is too fine a grain for
    // the engine, which pays off for multi-KB lines (fft_c2r_cross3_kernel).  Read per call so tests can switch.
    const char *e = getenv("SHB_HELM_BULK");
    const bool bulk = e && atoi(e) == 1;
    if constexpr (sizeof(T) == 16) {
        const bool aligned = ((reinterpret_cast<uintptr_t>(d.b) | reinterpret_cast<uintptr_t>(d.u) |
                               reinterpret_cast<uintptr_t>(d.invd0)) & 15) == 0;
        if (bulk && aligned) return launch_helm_tile_b<T, LW, RP, NCH, MINB, true>(d, st);
    }
    return launch_helm_tile_b<T, LW, RP, NCH, MINB, false>(d, st);
}

template <class T>
static int launch_helm_tile_any(const HelmDesc &d, cudaStream_t st) {
    const int n = d.n;
    if (n <= 128) return launch_helm_tile<T, 8, 8, 8, 2>(d, st);
    if (n <= 256) return launch_helm_tile<T, 8, 8, 16, 2>(d, st);
    static const int tv = getenv("SHB_HELM_TILE") ? atoi(getenv("SHB_HELM_TILE")) : 0;
    if (n <= 512) {
        // measured at 512^3 (B200): <8,16,16> at 164 registers (one CTA per SM) 0.95 ms; capped to 128
        // registers (two CTAs per SM) 0.76 ms; <4,8,32> 0.97 ms; <8,8,32> 0.95 ms
        if (tv == 1) return launch_helm_tile<T, 8, 16, 16, 1>(d, st);
        if (tv == 2) return launch_helm_tile<T, 4, 8, 32, 2>(d, st);
        return launch_helm_tile<T, 8, 16, 16, 2>(d, st);
    }
    if (n <= 1024) return launch_helm_tile<T, 4, 16, 32, 2>(d, st);
    return SHB_ERR_UNSUPPORTED;
}

// ---- banded LU substitution, tile-resident (shared factor table) -----------------------------------
// Same engine as helm_tile_kernel for the solvers that apply ONE factored banded matrix to every
// line (the inverse mass matrix of every Chebyshev / Legendre forward transform, la.TDMA / PDMA / ...
// along an axis): forward  y_i = b_i - sum_p L_p[i] y_{i-p*STEP},  backward
// x_i = (y_i - sum_q U_q[i] x_{i+q*STEP}) / D[i], factors in the table layout of banded.cu
// (fac[KL-p][i] = L_p, fac[KL][i] = D, fac[KL+q][i] = U_q).  The table goes to shared memory once
// per CTA (1/D formed there), skewed by one entry per chunk like the Helmholtz diagonals.
struct BandTileDesc {
    void *u;
    const double *fac;
    int64_t outer, inner;
    int n, n_u;
    int64_t tiles_per_outer;
};


// PL: per-line factors in the batch layout fac[KL+1+KU][outer][n][inner] (SolverGeneric1ND): staged
// into shared memory next to the data tile instead of the shared table.
template <class T, int KL, int KU, int STEP, int LW, int RP, int NCH, int MINB, bool PL>
__global__ void __launch_bounds__(STEP *LW *NCH, MINB) banded_tile_kernel(const __grid_constant__ BandTileDesc d) {
    using V = Val<T>;
    constexpr int NC = V::NC;
    constexpr int NT = STEP * LW * NCH;
    constexpr int RC = STEP * RP;
    constexpr int NR = RC * NCH;
    constexpr int NS = NR + NCH;              // skewed table length
    constexpr int KLm = KL > 0 ? KL : 1, KUm = KU > 0 ? KU : 1;
    extern __shared__ __align__(16) unsigned char smraw[];
    T *xs = reinterpret_cast<T *>(smraw);                      // [NR][LW]
    double *tl = reinterpret_cast<double *>(xs + NR * LW);     // [KLm][NS]
    double *tu = tl + KLm * NS;                                 // [KUm][NS]
    double *td = tu + KUm * NS;                                 // [NS]  1/D
    double *ft = reinterpret_cast<double *>(xs + NR * LW);     // PL: [KL+1+KU][NR][LW], same rotation as xs
    const int n = d.n;
    const int64_t o = blockIdx.x / d.tiles_per_outer;
    const int64_t cc0 = (blockIdx.x % d.tiles_per_outer) * LW;
    const int64_t rs = d.inner;
    T *ub = reinterpret_cast<T *>(d.u) + (o * d.n_u) * rs + cc0;

    for (int idx = threadIdx.x; idx < NR * LW; idx += NT) {
        const int i = idx / LW, l = idx % LW;
        const bool ok = (i < n) && (cc0 + l < d.inner);
        const int slot = i * LW + ((l + i / RC) % LW);
        cp_async_elem<T>(&xs[slot], ok ? ub + (int64_t)i * rs + l : ub, ok);
        if (PL) {
            const double *fb = d.fac + (o * n) * rs + cc0;
            const int64_t fsd = d.outer * (int64_t)n * rs;
#pragma unroll
            for (int dd = 0; dd < KL + 1 + KU; ++dd)
                cp_async_8(&ft[dd * NR * LW + slot], ok ? fb + dd * fsd + (int64_t)i * rs + l : fb, ok);
        }
    }
    if (!PL) {
        for (int i = threadIdx.x; i < NR; i += NT) {
            const int ii = i + i / RC;
#pragma unroll
            for (int p = 1; p <= KL; ++p)
                tl[(p - 1) * NS + ii] = (i < n && i - p * STEP >= 0) ? __ldg(&d.fac[(KL - p) * n + i]) : 0.0;
#pragma unroll
            for (int q = 1; q <= KU; ++q)
                tu[(q - 1) * NS + ii] = (i < n && i + q * STEP < n) ? __ldg(&d.fac[(KL + q) * n + i]) : 0.0;
            td[ii] = (i < n) ? 1.0 / __ldg(&d.fac[KL * n + i]) : 0.0;
        }
    }
    const int c = threadIdx.x % NCH;
    const int pr = threadIdx.x / NCH;
    const int p = pr % STEP, l = pr / STEP;
    const int i0 = c * RC + p;
    const int sw = (l + c) % LW;
    cp_async_wait_all();
    __syncthreads();

    if (KL > 0) {
        // ---- forward substitution: state s[k] = y_{i-(k+1)*STEP}
        double lc[RP][KLm];
        T sv[KLm];
        double H[KLm][KLm];
#pragma unroll
        for (int a = 0; a < KLm; ++a) {
            sv[a] = V::zero();
#pragma unroll
            for (int b = 0; b < KLm; ++b) H[a][b] = (a == b) ? 1.0 : 0.0;
        }
#pragma unroll
        for (int q = 0; q < RP; ++q) {
            const int i = i0 + STEP * q;
#pragma unroll
            for (int a = 0; a < KL; ++a)
                lc[q][a] = PL ? ((i < n && i - (a + 1) * STEP >= 0) ? ft[(KL - 1 - a) * NR * LW + i * LW + sw] : 0.0)
                              : tl[a * NS + i + c];
            T y = xs[i * LW + sw];
            double hn[KLm];
#pragma unroll
            for (int b = 0; b < KLm; ++b) hn[b] = 0.0;
#pragma unroll
            for (int a = 0; a < KL; ++a) {
                y = V::fnma(lc[q][a], sv[a], y);
#pragma unroll
                for (int b = 0; b < KL; ++b) hn[b] = fma(-lc[q][a], H[a][b], hn[b]);
            }
#pragma unroll
            for (int a = KL - 1; a >= 1; --a) {
                sv[a] = sv[a - 1];
#pragma unroll
                for (int b = 0; b < KL; ++b) H[a][b] = H[a - 1][b];
            }
            sv[0] = y;
#pragma unroll
            for (int b = 0; b < KL; ++b) H[0][b] = hn[b];
        }
        AMap<KLm, NC> m;
#pragma unroll
        for (int a = 0; a < KLm; ++a) {
#pragma unroll
            for (int b = 0; b < KLm; ++b) m.A[a * KLm + b] = H[a][b];
#pragma unroll
            for (int v = 0; v < NC; ++v) m.b[v * KLm + a] = V::get(sv[a], v);
        }
        double s[NC * KLm];
        amap_seg_scan<KLm, NC, NCH, false>(m, s, c);
#pragma unroll
        for (int a = 0; a < KLm; ++a)
#pragma unroll
            for (int v = 0; v < NC; ++v) V::set(sv[a], v, s[v * KLm + a]);
#pragma unroll
        for (int q = 0; q < RP; ++q) {
            const int i = i0 + STEP * q;
            T y = xs[i * LW + sw];
#pragma unroll
            for (int a = 0; a < KL; ++a) y = V::fnma(lc[q][a], sv[a], y);
#pragma unroll
            for (int a = KL - 1; a >= 1; --a) sv[a] = sv[a - 1];
            sv[0] = y;
            xs[i * LW + sw] = y;
        }
    }
    {
        // ---- back substitution: state s[k] = x_{i+(k+1)*STEP}
        double uc[RP][KUm], dc[RP];
        T sv[KUm];
        double H[KUm][KUm];
#pragma unroll
        for (int a = 0; a < KUm; ++a) {
            sv[a] = V::zero();
#pragma unroll
            for (int b = 0; b < KUm; ++b) H[a][b] = (a == b) ? 1.0 : 0.0;
        }
#pragma unroll
        for (int q = RP - 1; q >= 0; --q) {
            const int i = i0 + STEP * q;
#pragma unroll
            for (int a = 0; a < KU; ++a)
                uc[q][a] = PL ? ((i + (a + 1) * STEP < n) ? ft[(KL + 1 + a) * NR * LW + i * LW + sw] : 0.0) : tu[a * NS + i + c];
            dc[q] = PL ? ((i < n) ? 1.0 / ft[KL * NR * LW + i * LW + sw] : 0.0) : td[i + c];
            T y = xs[i * LW + sw];
            double hn[KUm];
#pragma unroll
            for (int b = 0; b < KUm; ++b) hn[b] = 0.0;
#pragma unroll
            for (int a = 0; a < KU; ++a) {
                y = V::fnma(uc[q][a], sv[a], y);
#pragma unroll
                for (int b = 0; b < KU; ++b) hn[b] = fma(-uc[q][a], H[a][b], hn[b]);
            }
            y = V::mul(dc[q], y);
#pragma unroll
            for (int a = KU - 1; a >= 1; --a) {
                sv[a] = sv[a - 1];
#pragma unroll
                for (int b = 0; b < KU; ++b) H[a][b] = H[a - 1][b];
            }
            if (KU > 0) {
                sv[0] = y;
#pragma unroll
                for (int b = 0; b < KU; ++b) H[0][b] = hn[b] * dc[q];
            } else {
                xs[i * LW + sw] = y;          // diagonal system: done
            }
        }
        if (KU > 0) {
            AMap<KUm, NC> m;
#pragma unroll
            for (int a = 0; a < KUm; ++a) {
#pragma unroll
                for (int b = 0; b < KUm; ++b) m.A[a * KUm + b] = H[a][b];
#pragma unroll
                for (int v = 0; v < NC; ++v) m.b[v * KUm + a] = V::get(sv[a], v);
            }
            double s[NC * KUm];
            amap_seg_scan<KUm, NC, NCH, true>(m, s, c);
#pragma unroll
            for (int a = 0; a < KUm; ++a)
#pragma unroll
                for (int v = 0; v < NC; ++v) V::set(sv[a], v, s[v * KUm + a]);
#pragma unroll
            for (int q = RP - 1; q >= 0; --q) {
                const int i = i0 + STEP * q;
                T y = xs[i * LW + sw];
#pragma unroll
                for (int a = 0; a < KU; ++a) y = V::fnma(uc[q][a], sv[a], y);
                y = V::mul(dc[q], y);
#pragma unroll
                for (int a = KU - 1; a >= 1; --a) sv[a] = sv[a - 1];
                sv[0] = y;
                xs[i * LW + sw] = y;
            }
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < NR * LW; idx += NT) {
        const int i = idx / LW, ll = idx % LW;
        if (i < n && cc0 + ll < d.inner) ub[(int64_t)i * rs + ll] = xs[i * LW + ((ll + i / RC) % LW)];
    }
}

template <class T, int KL, int KU, int STEP, int LW, int RP, int NCH, int MINB, bool PL = false>
static int launch_banded_tile(BandTileDesc d, cudaStream_t st) {
    constexpr int NR = STEP * RP * NCH, NS = NR + NCH;
    constexpr int KLm = KL > 0 ? KL : 1, KUm = KU > 0 ? KU : 1;
    constexpr size_t smem = (size_t)NR * LW * sizeof(T) +
                            (PL ? (size_t)(KL + 1 + KU) * NR * LW * sizeof(double) : (size_t)(KLm + KUm + 1) * NS * sizeof(double));
    static_assert(smem <= 227 * 1024, "tile does not fit in shared memory");
    static bool attr_set = false;
    if (!attr_set) {
        SHB_CUDA(cudaFuncSetAttribute(banded_tile_kernel<T, KL, KU, STEP, LW, RP, NCH, MINB, PL>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    d.tiles_per_outer = (d.inner + LW - 1) / LW;
    const int64_t grid = d.outer * d.tiles_per_outer;
    banded_tile_kernel<T, KL, KU, STEP, LW, RP, NCH, MINB, PL><<<(unsigned)grid, STEP * LW * NCH, smem, st>>>(d);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

template <class T, int KL, int KU, int STEP>
static int launch_banded_tile_n(const BandTileDesc &d, cudaStream_t st) {
    const int rows = (d.n + STEP - 1) / STEP;      // rows per parity class
    constexpr bool wide = (KL + KU >= 3);
    if (rows <= 64) return launch_banded_tile<T, KL, KU, STEP, 8, 8, 8, 2>(d, st);
    if (rows <= 128) return launch_banded_tile<T, KL, KU, STEP, 8, 8, 16, 2>(d, st);
    if (rows <= 256) {
        if (wide) return launch_banded_tile<T, KL, KU, STEP, 4, 8, 32, 2>(d, st);
        return launch_banded_tile<T, KL, KU, STEP, 8, 16, 16, 2>(d, st);
    }
    if (rows <= 512) return launch_banded_tile<T, KL, KU, STEP, 4, 16, 32, 2>(d, st);
    return SHB_ERR_UNSUPPORTED;
}

// per-line factors: tiles sized so that two CTAs fit an SM where possible
template <class T, int KL, int KU, int STEP>
static int launch_banded_tile_pl(const BandTileDesc &d, cudaStream_t st) {
    const int rows = (d.n + STEP - 1) / STEP;
    constexpr bool wide = (KL + KU >= 3);
    if (rows <= 64) return launch_banded_tile<T, KL, KU, STEP, 8, 8, 8, 2, true>(d, st);
    if (rows <= 128) return launch_banded_tile<T, KL, KU, STEP, 8, 8, 16, 2, true>(d, st);
    if (rows <= 256) return launch_banded_tile<T, KL, KU, STEP, 4, 8, 32, 2, true>(d, st);
    if (rows <= 512) {
        if (wide) return SHB_ERR_UNSUPPORTED;
        return launch_banded_tile<T, KL, KU, STEP, 4, 16, 32, 1, true>(d, st);
    }
    return SHB_ERR_UNSUPPORTED;
}

int banded_tile_launch(int kl, int ku, int step, void *u, const double *fac, int64_t outer, int64_t inner, int n,
                       int n_u, int is_complex, cudaStream_t st, int per_line) {
    static const bool off = getenv("SHB_BANDED_TILE") && atoi(getenv("SHB_BANDED_TILE")) == 0;
    if (off || inner < 4) return SHB_ERR_UNSUPPORTED;
    BandTileDesc d;
    d.u = u; d.fac = fac; d.outer = outer; d.inner = inner; d.n = n; d.n_u = n_u;
    if (per_line) {
        // measured (B200, profiles/): staging the per-line factor arrays in shared memory next to the data
        // (8 B cp.async, 4-line tiles) is SLOWER than the thread-per-line kernel (512^3 TDMA 1.41 vs 1.04 ms,
        // 256^3 PDMA 0.21 vs 0.17 ms): opt-in only
        static const bool pl_on = getenv("SHB_BANDED_TILE_PL") && atoi(getenv("SHB_BANDED_TILE_PL")) == 1;
        if (!pl_on) return SHB_ERR_UNSUPPORTED;
#define SHB_BTP(KLv, KUv)                                                                                  \
    if (kl == KLv && ku == KUv) {                                                                           \
        if (step == 2) return is_complex ? launch_banded_tile_pl<cplx, KLv, KUv, 2>(d, st)                  \
                                         : launch_banded_tile_pl<double, KLv, KUv, 2>(d, st);               \
        return is_complex ? launch_banded_tile_pl<cplx, KLv, KUv, 1>(d, st)                                 \
                          : launch_banded_tile_pl<double, KLv, KUv, 1>(d, st);                              \
    }
        SHB_BTP(1, 1) SHB_BTP(2, 2) SHB_BTP(1, 2)
#undef SHB_BTP
        return SHB_ERR_UNSUPPORTED;
    }
#define SHB_BT(KLv, KUv)                                                                                   \
    if (kl == KLv && ku == KUv) {                                                                           \
        if (step == 2) return is_complex ? launch_banded_tile_n<cplx, KLv, KUv, 2>(d, st)                   \
                                         : launch_banded_tile_n<double, KLv, KUv, 2>(d, st);                \
        return is_complex ? launch_banded_tile_n<cplx, KLv, KUv, 1>(d, st)                                  \
                          : launch_banded_tile_n<double, KLv, KUv, 1>(d, st);                               \
    }
    SHB_BT(1, 1) SHB_BT(2, 2) SHB_BT(1, 2) SHB_BT(0, 1) SHB_BT(0, 2)
#undef SHB_BT
    return SHB_ERR_UNSUPPORTED;
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

int helm_rows_pad(int n) { return n <= 128 ? 128 : n <= 256 ? 256 : n <= 512 ? 512 : ((n + 1023) / 1024) * 1024; }

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
    if (variant < 10) {
        rc = is_complex ? launch_helm_tile_any<cplx>(d, st) : launch_helm_tile_any<double>(d, st);
        if (rc == SHB_ERR_UNSUPPORTED) set_error("helmholtz_solve_lean: line length n=%d not covered", n);
        return rc;
    }
    if (is_complex) {
        if (variant == 11 && n <= 512) rc = launch_helm<cplx, 4, 8, 512>(d, st);
        else rc = launch_helm<cplx, 4, 4, 512>(d, st);
    } else {
        rc = (n <= 512) ? launch_helm<double, 4, 8, 512>(d, st) : launch_helm<double, 4, 4, 512>(d, st);
    }
    if (rc == SHB_ERR_UNSUPPORTED) set_error("helmholtz_solve_lean: line length n=%d not covered", n);
    return rc;
}

}  // extern "C"
