// linesolve.cuh -- building blocks of the chunk-parallel line solvers (linesolve.cu).
//
// A banded / structured triangular solve along a line is a linear recurrence
//     s_i = A_i s_{i-1} + b_i          (s: a small state vector, A_i from the factors, b_i from the data)
// which is serial in i but associative in the affine maps (A_i, b_i).  A line is cut into chunks of
// a few rows, one thread per chunk: every thread composes the map of its chunk with the state
// entering the chunk unknown (pass 1), the chunk maps of a line are scanned (this file), and the
// thread then runs the recurrence over its rows from the now-known entering state (pass 2).  The
// whole line stays in registers between the forward and the backward sweep: the data is read from
// HBM once and written once, and a line is worked on by n / rows-per-thread threads instead of one.
#pragma once
#include "common.cuh"

namespace shb {

// affine map on NB state vectors of dimension D sharing the matrix A:  s_v -> A s_v + b_v
template <int D, int NB> struct AMap {
    double A[D * D];
    double b[NB * D];
    static constexpr int E = D * D + NB * D;
};

template <int D, int NB> __device__ __forceinline__ void amap_identity(AMap<D, NB> &m) {
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) m.A[i * D + j] = (i == j) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < NB * D; ++i) m.b[i] = 0.0;
}

// second o first (apply `first`, then `second`)
template <int D, int NB>
__device__ __forceinline__ AMap<D, NB> amap_compose(const AMap<D, NB> &second, const AMap<D, NB> &first) {
    AMap<D, NB> r;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) a = fma(second.A[i * D + k], first.A[k * D + j], a);
            r.A[i * D + j] = a;
        }
#pragma unroll
    for (int v = 0; v < NB; ++v)
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double a = second.b[v * D + i];
#pragma unroll
            for (int k = 0; k < D; ++k) a = fma(second.A[i * D + k], first.b[v * D + k], a);
            r.b[v * D + i] = a;
        }
    return r;
}

template <int D, int NB> __device__ __forceinline__ AMap<D, NB> amap_shfl_up(const AMap<D, NB> &m, int off) {
    AMap<D, NB> r;
#pragma unroll
    for (int i = 0; i < D * D; ++i) r.A[i] = __shfl_up_sync(0xffffffffu, m.A[i], off);
#pragma unroll
    for (int i = 0; i < NB * D; ++i) r.b[i] = __shfl_up_sync(0xffffffffu, m.b[i], off);
    return r;
}

// Shared-memory layout of the chunk maps of a CTA: element e of the map of (pair, chunk) at
// maps[(e * NP + pair) * ldc + chunk], ldc >= C (padded so that the lines of a warp fall into
// different banks).  A "pair" is one independent recurrence: (line, parity class).
// NOTE: the element address is formed as (pair*ldc + c) + e*(NP*ldc) with the element stride in ONE
// variable; nvcc 12.9 mis-compiles the form ((e*NP + pair)*ldc + c) in an unrolled loop (its
// straight-line strength reduction chains the e-offsets: element 2 lands on element 3's slot).
template <int D, int NB>
__device__ __forceinline__ void amap_store(double *maps, int NP, int ldc, int pair, int c, const AMap<D, NB> &m) {
    const int es = NP * ldc;
    double *q = maps + (pair * ldc + c);
#pragma unroll
    for (int i = 0; i < D * D; ++i) q[i * es] = m.A[i];
#pragma unroll
    for (int i = 0; i < NB * D; ++i) q[(D * D + i) * es] = m.b[i];
}

template <int D, int NB>
__device__ __forceinline__ void amap_load(const double *maps, int NP, int ldc, int pair, int c, AMap<D, NB> &m) {
    const int es = NP * ldc;
    const double *q = maps + (pair * ldc + c);
#pragma unroll
    for (int i = 0; i < D * D; ++i) m.A[i] = q[i * es];
#pragma unroll
    for (int i = 0; i < NB * D; ++i) m.b[i] = q[(D * D + i) * es];
}

// entry i of the state of (pair, chunk c): same layout as the maps
__device__ __forceinline__ double &amap_state(double *states, int NP, int ldc, int pair, int c, int i) {
    const int es = NP * ldc;
    return states[(pair * ldc + c) + i * es];
}

// Exclusive scan of the chunk maps of every pair, starting from the zero state: on return
// amap_state(states, ..., pair, c, i), i < NB*D, is the state ENTERING chunk c.  `desc`: the recurrence
// runs from the last chunk to the first (back substitution).  One warp per pair (round robin): a lane
// folds CL consecutive chunks (C <= 32*CL; positions beyond C act as the identity), the lane totals go
// through a Kogge-Stone scan on shuffles.  All threads of the CTA must call this; the caller
// synchronises before (maps complete) and after.
template <int D, int NB, int CL>
__device__ __forceinline__ void amap_scan(const double *maps, double *states, int NP, int C, int ldc, bool desc) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int pair = warp; pair < NP; pair += nw) {
        AMap<D, NB> tot;
#pragma unroll
        for (int j = 0; j < CL; ++j) {
            const int pos = lane * CL + j;
            const int pc = pos < C ? pos : C - 1;                 // clamped: always a valid slot
            AMap<D, NB> m;
            amap_load(maps, NP, ldc, pair, desc ? C - 1 - pc : pc, m);
            if (pos >= C) amap_identity(m);
            tot = (j == 0) ? m : amap_compose(m, tot);
        }
        AMap<D, NB> inc = tot;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const AMap<D, NB> o = amap_shfl_up(inc, off);
            const AMap<D, NB> cmb = amap_compose(inc, o);
            if (lane >= off) inc = cmb;
        }
        const AMap<D, NB> ex = amap_shfl_up(inc, 1);
        double s[NB * D];
#pragma unroll
        for (int i = 0; i < NB * D; ++i) s[i] = (lane == 0) ? 0.0 : ex.b[i];
#pragma unroll
        for (int j = 0; j < CL; ++j) {
            const int pos = lane * CL + j;
            if (pos < C) {
                const int c = desc ? C - 1 - pos : pos;
#pragma unroll
                for (int i = 0; i < NB * D; ++i) amap_state(states, NP, ldc, pair, c, i) = s[i];
            }
            if (j + 1 < CL) {
                const int pc = pos < C ? pos : C - 1;
                AMap<D, NB> m;                                    // (re-read: cheaper than CL maps in registers)
                amap_load(maps, NP, ldc, pair, desc ? C - 1 - pc : pc, m);
                double t[NB * D];
#pragma unroll
                for (int v = 0; v < NB; ++v)
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        double a = m.b[v * D + i];
#pragma unroll
                        for (int k = 0; k < D; ++k) a = fma(m.A[i * D + k], s[v * D + k], a);
                        t[v * D + i] = a;
                    }
#pragma unroll
                for (int i = 0; i < NB * D; ++i) s[i] = t[i];
            }
        }
    }
}

// scalar / complex helpers: a value is NC doubles
template <class T> struct Val;
template <> struct Val<double> {
    static constexpr int NC = 1;
    static __device__ __forceinline__ double zero() { return 0.0; }
    static __device__ __forceinline__ double get(const double &x, int) { return x; }
    static __device__ __forceinline__ void set(double &x, int, double v) { x = v; }
    static __device__ __forceinline__ double fnma(double a, double x, double y) { return fma(-a, x, y); }
    static __device__ __forceinline__ double mul(double a, double x) { return a * x; }
    static __device__ __forceinline__ double add(double x, double y) { return x + y; }
};
template <> struct Val<cplx> {
    static constexpr int NC = 2;
    static __device__ __forceinline__ cplx zero() { return make_double2(0.0, 0.0); }
    static __device__ __forceinline__ double get(const cplx &x, int c) { return c ? x.y : x.x; }
    static __device__ __forceinline__ void set(cplx &x, int c, double v) { if (c) x.y = v; else x.x = v; }
    static __device__ __forceinline__ cplx fnma(double a, cplx x, cplx y) {
        return make_double2(fma(-a, x.x, y.x), fma(-a, x.y, y.y));
    }
    static __device__ __forceinline__ cplx mul(double a, cplx x) { return make_double2(a * x.x, a * x.y); }
    static __device__ __forceinline__ cplx add(cplx x, cplx y) { return make_double2(x.x + y.x, x.y + y.y); }
};

}  // namespace shb

namespace shb {

// Kogge-Stone scan of the chunk maps held by the W lanes of one aligned lane segment (W = number of
// chunks of a line, a power of two <= 32): no shared memory, no CTA barrier.  On return s[] is the state
// ENTERING this lane's chunk from the zero state; DESC: the recurrence runs from lane W-1 down to 0.
template <int D, int NB, int W, bool DESC>
__device__ __forceinline__ void amap_seg_scan(AMap<D, NB> m, double (&s)[NB * D], int c) {
#pragma unroll
    for (int off = 1; off < W; off <<= 1) {
        AMap<D, NB> o;
#pragma unroll
        for (int i = 0; i < D * D; ++i)
            o.A[i] = DESC ? __shfl_down_sync(0xffffffffu, m.A[i], off, W) : __shfl_up_sync(0xffffffffu, m.A[i], off, W);
#pragma unroll
        for (int i = 0; i < NB * D; ++i)
            o.b[i] = DESC ? __shfl_down_sync(0xffffffffu, m.b[i], off, W) : __shfl_up_sync(0xffffffffu, m.b[i], off, W);
        const AMap<D, NB> cmb = amap_compose(m, o);
        const bool has = DESC ? (c + off < W) : (c >= off);
        if (has) m = cmb;
    }
    const bool first = DESC ? (c == W - 1) : (c == 0);
#pragma unroll
    for (int i = 0; i < NB * D; ++i) {
        const double e = DESC ? __shfl_down_sync(0xffffffffu, m.b[i], 1, W) : __shfl_up_sync(0xffffffffu, m.b[i], 1, W);
        s[i] = first ? 0.0 : e;
    }
}

__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gsrc, bool pred) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_8(void *smem_dst, const void *gsrc, bool pred) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

}  // namespace shb
