// fft_fast.cu -- register-resident Stockham FFT kernels for the headline sizes.
//
// A line of M = R0*R1[*R2] complex points is transformed by P threads; each
// thread runs M/(Ri*P) radix-Ri butterflies per stage with the operands in
// registers.  Stage 0 reads its operands straight from global
// memory (fused zero-padding / Hermitian pre-twiddle), the last stage writes
// straight to global memory (fused truncation / Nyquist rules / scaling / real
// split), so data crosses shared memory only for the (stages-1) exchanges.
//
//   kernel A (fft_strided_kernel): transform axis is strided ([outer][M][inner],
//     inner > 1).  A CTA owns T consecutive columns; lanes run along the
//     contiguous `inner` index so every global access is a T*16 B segment and
//     the [pos][column] shared layout is bank-conflict free by construction.
//   kernel B (fft_contig_kernel): lines are contiguous (inner == 1): C2C, R2C
//     (half-length complex FFT + split) and C2R.  Lanes run along the line;
//     shared positions are skewed (pos + pos/8) against bank conflicts.
//
// Reference semantics: see fft_generic.cu (same C ABI, same results).
#include "fft_common.cuh"
#include "butterfly.cuh"
#include <stdlib.h>

namespace shb {

constexpr int cmax(int a, int b) { return a > b ? a : b; }

// A schedule: M = R0*R1[*R2] computed by P threads per line.  Stage i has
// NBi = M/Ri butterflies; a thread runs Bi = NBi/P of them (NBi a multiple of
// P) or at most one (NBi < P, threads j >= NBi idle in that stage).
template <int M_, int R0_, int R1_, int R2_, int P_>
struct Sched {
    static constexpr int M = M_, R0 = R0_, R1 = R1_, R2 = R2_, P = P_;
    static_assert(R0_ * R1_ * R2_ == M_, "radices must multiply to M");
    static constexpr int NST = (R2_ > 1) ? 3 : 2;
    static constexpr int NB0 = M / R0, NB1 = M / R1, NB2 = (R2_ > 1) ? M / R2 : 0;
    static constexpr int bpt(int nb) { return nb >= P_ ? nb / P_ : 1; }
    static_assert(NB0 < P_ || NB0 % P_ == 0, "stage 0 butterflies vs threads");
    static_assert(NB1 < P_ || NB1 % P_ == 0, "stage 1 butterflies vs threads");
    static_assert(NB2 < P_ || NB2 % P_ == 0, "stage 2 butterflies vs threads");
    static constexpr int B0 = bpt(NB0), B1 = bpt(NB1), B2 = (R2_ > 1) ? bpt(NB2) : 0;
    static constexpr int EMAX = cmax(B0 * R0, cmax(B1 * R1, B2 * R2));
    static constexpr int RL = (NST == 3) ? R2 : R1;     // last stage: radix,
    static constexpr int NBL = (NST == 3) ? NB2 : NB1;  //   butterflies per line,
    static constexpr int BL = (NST == 3) ? B2 : B1;     //   butterflies per thread
};

// Stages after the operands of stage 0 are in v: butterfly q of this thread
// (line butterfly b = j + q*P) has operand m at v[q*R0+m] = x[b + m*NB0].  On
// return v[q*RL+m] = X[b + m*NBL].  `S` maps a line position to shared memory
// for this thread's line, `Y` synchronises the threads that share the line.
// w[m] = w1^m for m = 1..R-1 with multiplication depth <= log2(R) (each entry stays within a few ulp)
template <int R> __device__ __forceinline__ void twiddle_powers(cplx w1, cplx *w) {
    w[1] = w1;
#pragma unroll
    for (int m = 2; m < R; ++m) w[m] = cmul(w[m / 2], w[m - m / 2]);
}

template <class SC, bool INV, bool CHAIN, class S, class Y>
__device__ __forceinline__ void fft_stages(cplx *v, const int j, S s, Y sync, const cplx *__restrict__ tw) {
    constexpr int R0 = SC::R0, R1 = SC::R1, R2 = SC::R2, P = SC::P;
    if (j < SC::NB0) {
#pragma unroll
        for (int q = 0; q < SC::B0; ++q) {
            Bfly<R0, INV>::run(v + q * R0);
#pragma unroll
            for (int m = 0; m < R0; ++m) s((j + q * P) * R0 + m) = v[q * R0 + m];
        }
    }
    sync();
    if (j < SC::NB1) {
#pragma unroll
        for (int q = 0; q < SC::B1; ++q) {
            const int b = j + q * P;
#pragma unroll
            for (int m = 0; m < R1; ++m) v[q * R1 + m] = s(b + m * SC::NB1);
        }
#pragma unroll
        for (int q = 0; q < SC::B1; ++q) {
            const int k = (j + q * P) % R0;
            if constexpr (CHAIN && R1 > 2) {
                cplx w[R1];
                twiddle_powers<R1>(__ldg(&tw[k * R2]), w);
#pragma unroll
                for (int m = 1; m < R1; ++m)
                    v[q * R1 + m] = INV ? cmulc(v[q * R1 + m], w[m]) : cmul(v[q * R1 + m], w[m]);
            } else {
#pragma unroll
                for (int m = 1; m < R1; ++m) {
                    const cplx w = __ldg(&tw[k * m * R2]);
                    v[q * R1 + m] = INV ? cmulc(v[q * R1 + m], w) : cmul(v[q * R1 + m], w);
                }
            }
            Bfly<R1, INV>::run(v + q * R1);
        }
    }
    if constexpr (SC::NST == 3) {
        sync();
        if (j < SC::NB1) {
#pragma unroll
            for (int q = 0; q < SC::B1; ++q) {
                const int b = j + q * P;
                const int k = b % R0;
                const int j0 = (b / R0) * (R0 * R1) + k;
#pragma unroll
                for (int m = 0; m < R1; ++m) s(j0 + m * R0) = v[q * R1 + m];
            }
        }
        sync();
        if (j < SC::NB2) {
#pragma unroll
            for (int q = 0; q < SC::B2; ++q) {
                const int b = j + q * P;
#pragma unroll
                for (int m = 0; m < R2; ++m) v[q * R2 + m] = s(b + m * SC::NB2);
            }
#pragma unroll
            for (int q = 0; q < SC::B2; ++q) {
                const int b = j + q * P;
                if constexpr (CHAIN && R2 > 2) {
                    cplx w[R2];
                    twiddle_powers<R2>(__ldg(&tw[b]), w);
#pragma unroll
                    for (int m = 1; m < R2; ++m)
                        v[q * R2 + m] = INV ? cmulc(v[q * R2 + m], w[m]) : cmul(v[q * R2 + m], w[m]);
                } else {
#pragma unroll
                    for (int m = 1; m < R2; ++m) {
                        const cplx w = __ldg(&tw[b * m]);
                        v[q * R2 + m] = INV ? cmulc(v[q * R2 + m], w) : cmul(v[q * R2 + m], w);
                    }
                }
                Bfly<R2, INV>::run(v + q * R2);
            }
        }
    }
}

// a * exp(-2 pi i STEP*m/48) for the unrolled loop index m (resolved at compile time after unrolling)
template <int STEP> __device__ __forceinline__ cplx root48_rt(cplx a, int m) {
    switch (m) {
    case 0: return a;
    case 1: return mul_root48<STEP * 1, false>(a);
    case 2: return mul_root48<STEP * 2, false>(a);
    case 3: return mul_root48<STEP * 3, false>(a);
    case 4: return mul_root48<STEP * 4, false>(a);
    case 5: return mul_root48<STEP * 5, false>(a);
    case 6: return mul_root48<STEP * 6, false>(a);
    case 7: return mul_root48<STEP * 7, false>(a);
    case 8: return mul_root48<STEP * 8, false>(a);
    case 9: return mul_root48<STEP * 9, false>(a);
    case 10: return mul_root48<STEP * 10, false>(a);
    default: return mul_root48<STEP * 11, false>(a);
    }
}

// threads sharing a line: a warp (P == 32), a few warps (named barrier) or the CTA
template <int P> struct GroupSync {
    int id;
    __device__ __forceinline__ void operator()() const {
        if constexpr (P == 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(P) : "memory");
    }
};
struct CtaSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

struct FastDesc {
    const void *in;
    void *out;
    int64_t outer, inner, nlines;
    int n_fft, N, n_in, n_out, flags;
    double scale;
    const cplx *tw, *tw2;
    int pf = 0;       // fft_contig_kernel: L2 prefetch distance in CTAs (0: off)
    // fft_strided_pad3x_kernel<..., CURL>: wavenumber tables of the three axes, extent of the last axis
    const double *k0 = nullptr, *k1 = nullptr, *k2 = nullptr;
    int curl_n2 = 1;
    // peer-chunked addressing along the transform axis (fused all-to-all pack / unpack):
    // row r of outer index o lives at (r / cr) * cs + ((o * cr + r % cr) * inner + c); cr == 0: off
    int in_cr, out_cr;
    int64_t in_cs, out_cs;
    // direct peer stores over NVLink: chunk p of the output goes to out_peer[p] (the receive
    // buffer of rank p, already offset to the slot of this rank) instead of out + p*out_cs
    cplx *out_peer[8];
    int use_peer;
    // generalised line addressing of the CH variants of fft_strided_kernel: line -> (a, b) = (line / cw,
    // line % cw); element r of the line at a*as + r*rs + b + c0 (chunked: p*cs + a*as + (r - p*cr)*rs
    // + b + c0).  The standard [outer][n][inner] array is cw = inner, c0 = 0, rs = inner,
    // as = n*inner (chunked: cr*inner); a column sub-range [c0, c0+cw) of it, or lines picked out of
    // a flattened (j, c) inner index, are other settings of the same five numbers -- what the
    // column-chunked (overlapped) slab exchange launches.
    int64_t cw, c0, in_as, in_rs, out_as, out_rs;
    int64_t cvalid;   // columns b >= cvalid of a line family are padding (tiles aligned to a padded pitch)
};

// zero-padding rule of the C2C backward transform as a branch-free gather:
// row r of the padded line comes from row `src` of the stored line times f.
struct PadMap {
    int N, M, h;
    bool padded, evenN, dd;
    __device__ __forceinline__ PadMap(int N_, int M_, int flags)
        : N(N_), M(M_), h(N_ / 2), padded(N_ < M_), evenN(N_ % 2 == 0), dd((flags & 1) != 0) {}
    __device__ __forceinline__ bool map(int r, int &src, double &f) const {
        src = r;
        f = 1.0;
        if (padded) {
            const bool lo = (r <= h), hi = (r >= M - h);
            src = lo ? r : r - M + N;
            if (evenN && (r == h || r == M - h)) f = 0.5;
            return lo || hi;
        }
        return !(dd && r >= N / 3 && r <= N - N / 3);
    }
};

// ---- kernel A ---------------------------------------------------------------------
template <class SC, int MODE, int T, bool CH>
__global__ void __launch_bounds__(SC::P *T, (SC::P * T <= 512 && SC::EMAX <= 8 ? 2 : 1))
fft_strided_kernel(const __grid_constant__ FastDesc d) {
    constexpr bool INV = (MODE == FFT_C2C_BWD);
    constexpr int M = SC::M, P = SC::P;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *sm = reinterpret_cast<cplx *>(smem_raw);
    const int c = threadIdx.x % T;
    const int j = threadIdx.x / T;
    const int64_t line = (int64_t)blockIdx.x * T + c;
    const int64_t dv = CH ? d.cw : d.inner;
    bool valid_ = line < d.nlines;
    const int64_t o = valid_ ? line / dv : 0, cc = valid_ ? line % dv : 0;
    if (CH) valid_ = valid_ && cc < d.cvalid;
    const bool valid = valid_;
    const cplx *__restrict__ inb = reinterpret_cast<const cplx *>(d.in);
    cplx *__restrict__ outb = reinterpret_cast<cplx *>(d.out);
    const int64_t ib = CH ? o * d.in_as + cc + d.c0 : 0, ob = CH ? o * d.out_as + cc + d.c0 : 0;
    auto in_at = [&](int r) -> const cplx & {
        if (CH) {
            if (d.in_cr) {
                const int p = r / d.in_cr;
                return inb[p * d.in_cs + ib + (r - p * d.in_cr) * d.in_rs];
            }
            return inb[ib + r * d.in_rs];
        }
        return inb[(o * d.n_in + r) * d.inner + cc];
    };
    auto out_at = [&](int r) -> cplx & {
        if (CH) {
            if (d.out_cr) {
                const int p = r / d.out_cr;
                const int64_t off = ob + (r - p * d.out_cr) * d.out_rs;
                if (d.use_peer) return d.out_peer[p][off];
                return outb[p * d.out_cs + off];
            }
            return outb[ob + r * d.out_rs];
        }
        return outb[(o * d.n_out + r) * d.inner + cc];
    };
    const int N = d.N;
    const bool padded = (N < M);
    const bool evenN = (N % 2 == 0);
    const int h = N / 2;
    const PadMap pm(N, M, d.flags);

    cplx v[SC::EMAX];
    if (j < SC::NB0) {
        // branch-free operand fetch: all loads are issued back to back.  The un-padded backward
        // transform (no 3/2-rule gather, no 2/3-rule zeroing) takes the forward transform's plain
        // path: the row mapping below costs the chunked-input variant 50 % (latency-bound loads)
        if (MODE != FFT_C2C_BWD || (!pm.padded && !pm.dd)) {
#pragma unroll
            for (int q = 0; q < SC::B0; ++q) {
#pragma unroll
                for (int m = 0; m < SC::R0; ++m) {
                    const int r = (j + q * P) + m * SC::NB0;
                    cplx x = make_double2(0.0, 0.0);
                    if (valid) x = in_at(r);
                    v[q * SC::R0 + m] = x;
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < SC::B0; ++q) {
#pragma unroll
                for (int m = 0; m < SC::R0; ++m) {
                    const int r = (j + q * P) + m * SC::NB0;
                    int src = r;
                    double f = 1.0;
                    const bool ok = pm.map(r, src, f) && valid;
                    cplx x = make_double2(0.0, 0.0);
                    if (ok) x = in_at(src);
                    v[q * SC::R0 + m] = make_double2(x.x * f, x.y * f);
                }
            }
        }
    }
    auto s = [&](int pos) -> cplx & { return sm[pos * T + c]; };
    fft_stages<SC, INV, true>(v, j, s, CtaSync(), d.tw);

    const double scale = d.scale;
    if (MODE == FFT_C2C_FWD && padded && evenN) {
        // fold X[M-h] into row h: pass it through shared memory (one value per column)
        __syncthreads();
        if (j < SC::NBL) {
#pragma unroll
            for (int q = 0; q < SC::BL; ++q)
#pragma unroll
                for (int m = 0; m < SC::RL; ++m)
                    if ((j + q * P) + m * SC::NBL == M - h) sm[c] = v[q * SC::RL + m];
        }
        __syncthreads();
    }
    if (j < SC::NBL && valid) {
#pragma unroll
        for (int q = 0; q < SC::BL; ++q) {
#pragma unroll
            for (int m = 0; m < SC::RL; ++m) {
                const int r = (j + q * P) + m * SC::NBL;
                const cplx x0 = v[q * SC::RL + m];
                if (MODE == FFT_C2C_BWD || !padded) {
                    out_at(r) = cscale(x0, scale);
                } else {
                    if (r <= h) {
                        cplx x = x0;
                        if (evenN && r == h) x = cadd(x, sm[c]);
                        out_at(r) = cscale(x, scale);
                    } else if (r > M - h || (!evenN && r == M - h)) {
                        out_at(r - M + N) = cscale(x0, scale);
                    }
                }
            }
        }
    }
}

// ---- kernel A, software-pipelined variant ------------------------------------------------
// For the long lines (768 ... 1536) one CTA fills an SM (>= 768 threads, ~100 KB of exchange
// buffer), so CTAs cannot overlap each other's load / compute / store phases.  This persistent
// variant overlaps them itself: the stage-0 operands of the NEXT tile are fetched with cp.async
// (16 B per copy, zero-fill for the padded rows) into a per-thread staging area in shared memory
// while the current tile is transformed and stored.  Each thread only ever reads the slots it
// filled, so cp.async.wait_group is the only synchronisation the staging area needs.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, bool pred) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}

template <class SC, int MODE, int T, bool CH>
__global__ void __launch_bounds__(SC::P *T, (SC::P * T <= 512 ? 2 : 1)) fft_strided_pipe_kernel(const __grid_constant__ FastDesc d,
                                                                       const int64_t ntiles) {
    constexpr bool INV = (MODE == FFT_C2C_BWD);
    constexpr int M = SC::M, P = SC::P, E0 = SC::B0 * SC::R0;
    constexpr int NT = (SC::NB0 < SC::P ? SC::NB0 : SC::P) * T;   // threads that own stage-0 operands
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *sm = reinterpret_cast<cplx *>(smem_raw);          // exchange buffer [M][T]
    cplx *stg = sm + (size_t)M * T;                         // staging [E0][NT] (= M*T slots)
    const int c = threadIdx.x % T;
    const int j = threadIdx.x / T;
    const cplx *__restrict__ inb = reinterpret_cast<const cplx *>(d.in);
    cplx *__restrict__ outb = reinterpret_cast<cplx *>(d.out);
    const int N = d.N;
    const bool padded = (N < M);
    const bool evenN = (N % 2 == 0);
    const int h = N / 2;
    const PadMap pm(N, M, d.flags);
    const double scale = d.scale;

    auto issue = [&](int64_t tile) {
        const int64_t line = tile * T + c;
        const bool valid = line < d.nlines;
        const int64_t o = valid ? line / d.inner : 0, cc = valid ? line % d.inner : 0;
        if (j < SC::NB0) {
#pragma unroll
            for (int q = 0; q < SC::B0; ++q) {
#pragma unroll
                for (int m = 0; m < SC::R0; ++m) {
                    const int r = (j + q * P) + m * SC::NB0;
                    bool ok = valid;
                    int src = r;
                    double f = 1.0;
                    if (MODE == FFT_C2C_BWD) ok = pm.map(r, src, f) && valid;
                    const cplx *g = inb;
                    if (ok) {
                        if (CH && d.in_cr) {
                            const int p = src / d.in_cr;
                            g = inb + (p * d.in_cs + ((o * d.in_cr + (src - p * d.in_cr)) * d.inner + cc));
                        } else {
                            g = inb + ((o * d.n_in + src) * d.inner + cc);
                        }
                    }
                    cp_async16(&stg[(size_t)(q * SC::R0 + m) * NT + threadIdx.x], g, ok);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int64_t tile = blockIdx.x;
    if (tile < ntiles) issue(tile);
    while (tile < ntiles) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        cplx v[SC::EMAX];
        if (j < SC::NB0) {
#pragma unroll
            for (int q = 0; q < SC::B0; ++q) {
#pragma unroll
                for (int m = 0; m < SC::R0; ++m) {
                    cplx x = stg[(size_t)(q * SC::R0 + m) * NT + threadIdx.x];
                    if (MODE == FFT_C2C_BWD && padded && evenN) {
                        const int r = (j + q * P) + m * SC::NB0;
                        if (r == h || r == M - h) x = cscale(x, 0.5);
                    }
                    v[q * SC::R0 + m] = x;
                }
            }
        }
        const int64_t next = tile + gridDim.x;
        if (next < ntiles) issue(next);

        const int64_t line = tile * T + c;
        const bool valid = line < d.nlines;
        const int64_t o = valid ? line / d.inner : 0, cc = valid ? line % d.inner : 0;
        auto out_at = [&](int r) -> cplx & {
            if (CH && d.out_cr) {
                const int p = r / d.out_cr;
                const int64_t off = (o * d.out_cr + (r - p * d.out_cr)) * d.inner + cc;
                if (d.use_peer) return d.out_peer[p][off];
                return outb[p * d.out_cs + off];
            }
            return outb[(o * d.n_out + r) * d.inner + cc];
        };
        auto s = [&](int pos) -> cplx & { return sm[pos * T + c]; };
        fft_stages<SC, INV, true>(v, j, s, CtaSync(), d.tw);

        if (MODE == FFT_C2C_FWD && padded && evenN) {
            __syncthreads();
            if (j < SC::NBL) {
#pragma unroll
                for (int q = 0; q < SC::BL; ++q)
#pragma unroll
                    for (int m = 0; m < SC::RL; ++m)
                        if ((j + q * P) + m * SC::NBL == M - h) sm[c] = v[q * SC::RL + m];
            }
            __syncthreads();
        }
        if (j < SC::NBL && valid) {
#pragma unroll
            for (int q = 0; q < SC::BL; ++q) {
#pragma unroll
                for (int m = 0; m < SC::RL; ++m) {
                    const int r = (j + q * P) + m * SC::NBL;
                    const cplx x0 = v[q * SC::RL + m];
                    if (MODE == FFT_C2C_BWD || !padded) {
                        out_at(r) = cscale(x0, scale);
                    } else {
                        if (r <= h) {
                            cplx x = x0;
                            if (evenN && r == h) x = cadd(x, sm[c]);
                            out_at(r) = cscale(x, scale);
                        } else if (r > M - h || (!evenN && r == M - h)) {
                            out_at(r - M + N) = cscale(x0, scale);
                        }
                    }
                }
            }
        }
        __syncthreads();     // exchange buffer is reused by the next tile
        tile = next;
    }
}

// ---- kernel A for the Chebyshev transforms ---------------------------------------------
// DCT-II / DCT-III of length M on the complex FFT core (one pass for complex data), with the
// quadrature scaling, the Shen stencil and zero-padding fused:
//   DCT2 (scalar product):  v[p] = x[2p] | x[2(M-1-p)+1];  V = FFT(v);
//        S_k = scale (w_k V_k + conj(w_k) V_{M-k}),  w_k = exp(-i pi k / 2M);
//        out_k = sum_i sten_i[k] S_{k+soff_i}   (k < dim),  0 for dim <= k < n_spec
//   DCT3 (evaluation):      c_j = sum_i sten_i[j-soff_i] u_{j-soff_i} (+ boundary lift);
//        U_0 = c_0, U_k = 0.5 w_k (c_k + i c_{M-k});  y = FFT(U);  x[2p] = y[p], x[2p+1] = y[M-1-p]
// Reference semantics: chebyshev/bases.py:315-353, spectralbase.py:1462-1474, 1616-1634.
struct DctFastDesc {
    const void *in;
    void *out;
    int64_t outer, inner, nlines;
    int n_spec, dim, is_real, nsten, nbc;
    int soff[DCT_MAX_STENCIL];
    const double *sten[DCT_MAX_STENCIL];
    const double *bc_lift;
    double scale;
    const cplx *tw, *tw3;
};

// NS: number of stencil diagonals known at compile time (1..3: the orthogonal, Dirichlet / Neumann and
// biharmonic bases), 0: read it from the descriptor.  With the runtime loop the stencil sections were
// 50-60 % of the instructions the kernel executes (ncu source counters, profiles/r2_*): branches and
// index arithmetic around two or three multiply-adds.
template <class SC, int MODE, int T, int NS>
__global__ void __launch_bounds__(SC::P *T, (SC::P * T <= 512 && SC::EMAX <= 8 ? 2 : 1))
dct_strided_kernel(const __grid_constant__ DctFastDesc d) {
    constexpr int M = SC::M, P = SC::P;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *sm = reinterpret_cast<cplx *>(smem_raw);
    const int c = threadIdx.x % T;
    const int j = threadIdx.x / T;
    const int64_t line = (int64_t)blockIdx.x * T + c;
    const bool valid = line < d.nlines;
    const int64_t o = valid ? line / d.inner : 0, cc = valid ? line % d.inner : 0;
    const int n_in = (MODE == FFT_DCT2) ? M : d.n_spec;
    const int n_out = (MODE == FFT_DCT2) ? d.n_spec : M;
    auto fetch = [&](int r) -> cplx {
        const int64_t idx = (o * n_in + r) * d.inner + cc;
        if (d.is_real) return make_double2(reinterpret_cast<const double *>(d.in)[idx], 0.0);
        return reinterpret_cast<const cplx *>(d.in)[idx];
    };
    auto put = [&](int r, cplx x) {
        const int64_t idx = (o * n_out + r) * d.inner + cc;
        if (d.is_real) reinterpret_cast<double *>(d.out)[idx] = x.x;
        else reinterpret_cast<cplx *>(d.out)[idx] = x;
    };
    auto s = [&](int pos) -> cplx & { return sm[pos * T + c]; };

    cplx v[SC::EMAX];
    if (MODE == FFT_DCT2) {
        if (j < SC::NB0) {
#pragma unroll
            for (int q = 0; q < SC::B0; ++q)
#pragma unroll
                for (int m = 0; m < SC::R0; ++m) {
                    const int pos = (j + q * P) + m * SC::NB0;
                    const int src = (pos < (M + 1) / 2) ? 2 * pos : 2 * (M - 1 - pos) + 1;
                    v[q * SC::R0 + m] = valid ? fetch(src) : make_double2(0.0, 0.0);
                }
        }
    } else {
        // 1. stored coefficients -> shared (rows beyond the dofs / boundary rows are zero)
        if (j < SC::NB0) {
#pragma unroll
            for (int q = 0; q < SC::B0; ++q)
#pragma unroll
                for (int m = 0; m < SC::R0; ++m) {
                    const int pos = (j + q * P) + m * SC::NB0;
                    cplx x = make_double2(0.0, 0.0);
                    if (valid && pos < d.dim + d.nbc && pos < d.n_spec) x = fetch(pos);
                    s(pos) = x;
                }
        }
        __syncthreads();
        // 2. c = K^T u (+ boundary lift) evaluated at pos AND at M - pos straight from the staged
        //    coefficients (twice the stencil arithmetic, but no publish / re-read round trip), then
        //    U_0 = c_0, U_k = 0.5 w_k (c_k + i c_{M-k})
        auto coeff = [&](int pos) -> cplx {
            cplx acc = make_double2(0.0, 0.0);
            for (int i = 0; i < d.nsten; ++i) {
                const int k = pos - d.soff[i];
                if (k >= 0 && k < d.dim) {
                    const double a = d.sten[i] ? __ldg(&d.sten[i][k]) : 1.0;
                    const cplx x = s(k);
                    acc.x = fma(a, x.x, acc.x);
                    acc.y = fma(a, x.y, acc.y);
                }
            }
            if (d.nbc > 0 && pos < 4) {
                for (int b = 0; b < d.nbc; ++b) {
                    if (d.dim + b >= M) break;
                    const double a = __ldg(&d.bc_lift[b * 4 + pos]);
                    const cplx x = s(d.dim + b);
                    acc.x = fma(a, x.x, acc.x);
                    acc.y = fma(a, x.y, acc.y);
                }
            }
            return acc;
        };
        if (j < SC::NB0) {
#pragma unroll
            for (int q = 0; q < SC::B0; ++q)
#pragma unroll
                for (int m = 0; m < SC::R0; ++m) {
                    const int pos = (j + q * P) + m * SC::NB0;
                    const cplx ck = coeff(pos);
                    if (pos > 0) {
                        const cplx cm = coeff(M - pos);
                        const cplx w = __ldg(&d.tw3[pos]);
                        const cplx sum = make_double2(ck.x - cm.y, ck.y + cm.x);
                        v[q * SC::R0 + m] = cscale(cmul(w, sum), 0.5);
                    } else {
                        v[q * SC::R0 + m] = ck;
                    }
                }
        }
        __syncthreads();
    }

    fft_stages<SC, false, true>(v, j, s, CtaSync(), d.tw);

    if (MODE == FFT_DCT3) {
        if (j < SC::NBL && valid) {
#pragma unroll
            for (int q = 0; q < SC::BL; ++q)
#pragma unroll
                for (int m = 0; m < SC::RL; ++m) {
                    const int pos = (j + q * P) + m * SC::NBL;
                    const int row = (pos < (M + 1) / 2) ? 2 * pos : 2 * (M - 1 - pos) + 1;
                    put(row, cscale(v[q * SC::RL + m], d.scale));
                }
        }
        return;
    }
    // DCT2 epilogue
    __syncthreads();
    if (j < SC::NBL) {
#pragma unroll
        for (int q = 0; q < SC::BL; ++q)
#pragma unroll
            for (int m = 0; m < SC::RL; ++m) s((j + q * P) + m * SC::NBL) = v[q * SC::RL + m];
    }
    __syncthreads();
    if (j < SC::NBL) {
#pragma unroll
        for (int q = 0; q < SC::BL; ++q)
#pragma unroll
            for (int m = 0; m < SC::RL; ++m) {
                const int k = (j + q * P) + m * SC::NBL;
                const cplx vk = v[q * SC::RL + m], vm = s(k == 0 ? 0 : M - k);
                const cplx w = __ldg(&d.tw3[k]);
                v[q * SC::RL + m] = cscale(cadd(cmul(w, vk), cmulc(vm, w)), d.scale);
            }
    }
    const bool plain = (d.nsten == 1 && d.soff[0] == 0 && d.sten[0] == nullptr);
    if (!plain) {
        __syncthreads();
        if (j < SC::NBL) {
#pragma unroll
            for (int q = 0; q < SC::BL; ++q)
#pragma unroll
                for (int m = 0; m < SC::RL; ++m) s((j + q * P) + m * SC::NBL) = v[q * SC::RL + m];
        }
        __syncthreads();
    }
    if (j < SC::NBL && valid) {
        const int nst = NS > 0 ? NS : d.nsten;
        int so[NS > 0 ? NS : DCT_MAX_STENCIL];
        const double *sp[NS > 0 ? NS : DCT_MAX_STENCIL];
#pragma unroll
        for (int i = 0; i < (NS > 0 ? NS : DCT_MAX_STENCIL); ++i) { so[i] = d.soff[i]; sp[i] = d.sten[i]; }
        const int dim = d.dim, nsp = d.n_spec;
#pragma unroll
        for (int q = 0; q < SC::BL; ++q)
#pragma unroll
            for (int m = 0; m < SC::RL; ++m) {
                const int k = (j + q * P) + m * SC::NBL;
                if (k >= nsp) continue;
                cplx acc = make_double2(0.0, 0.0);
                if (k < dim) {
                    if (plain) acc = v[q * SC::RL + m];
                    else {
#pragma unroll
                        for (int i = 0; i < (NS > 0 ? NS : DCT_MAX_STENCIL); ++i) {
                            if (i < nst) {
                                const int jj = k + so[i];
                                if (jj < M) {
                                    const double a = sp[i] ? __ldg(sp[i] + k) : 1.0;
                                    const cplx x = s(jj);
                                    acc.x = fma(a, x.x, acc.x);
                                    acc.y = fma(a, x.y, acc.y);
                                }
                            }
                        }
                    }
                }
                put(k, acc);
            }
    }
}

template <class SC, int MODE, int T, int NS>
static int launch_dct_ns(const DctFastDesc &d, cudaStream_t st) {
    constexpr size_t smem = (size_t)SC::M * T * sizeof(cplx);
    static bool attr_set = false;
    if (!attr_set) {
        SHB_CUDA(cudaFuncSetAttribute(dct_strided_kernel<SC, MODE, T, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        attr_set = true;
    }
    const int64_t ntiles = (d.nlines + T - 1) / T;
    dct_strided_kernel<SC, MODE, T, NS><<<(unsigned)ntiles, SC::P * T, smem, st>>>(d);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

template <class SC, int MODE, int T>
static int launch_dct(const DctFastDesc &d, cudaStream_t st) {
    // DCT-III keeps the runtime stencil loop: unrolled variants of its prologue (recompute, publish /
    // re-read, stencil straight from global memory) all measured slower on B200 (0.83 - 0.99 vs 0.78 ms)
    if (MODE == FFT_DCT3) return launch_dct_ns<SC, MODE, T, 0>(d, st);
    switch (d.nsten) {
    case 1: return launch_dct_ns<SC, MODE, T, 1>(d, st);
    case 2: return launch_dct_ns<SC, MODE, T, 2>(d, st);
    case 3: return launch_dct_ns<SC, MODE, T, 3>(d, st);
    default: return launch_dct_ns<SC, MODE, T, 0>(d, st);
    }
}

#define SHB_DCT_CASES(MODE)                                                                          \
    switch (M) {                                                                                     \
    case 64:   return launch_dct<Sched<64, 8, 8, 1, 8>, MODE, 8>(d, st);                             \
    case 128:  return launch_dct<Sched<128, 8, 4, 4, 32>, MODE, 8>(d, st);                           \
    case 256:  return launch_dct<Sched<256, 8, 8, 4, 32>, MODE, 8>(d, st);                           \
    case 512:  return launch_dct<Sched<512, 8, 8, 8, 64>, MODE, 8>(d, st);                           \
    case 1024: return launch_dct<Sched<1024, 16, 8, 8, 128>, MODE, 4>(d, st);                        \
    case 96:   return launch_dct<Sched<96, 8, 12, 1, 12>, MODE, 8>(d, st);                           \
    case 192:  return launch_dct<Sched<192, 4, 8, 6, 48>, MODE, 8>(d, st);                           \
    case 384:  return launch_dct<Sched<384, 8, 8, 6, 64>, MODE, 8>(d, st);                           \
    case 768:  return launch_dct<Sched<768, 8, 8, 12, 96>, MODE, 8>(d, st);                          \
    case 1536: return launch_dct<Sched<1536, 8, 16, 12, 192>, MODE, 4>(d, st);                       \
    default: return SHB_ERR_UNSUPPORTED;                                                             \
    }

static int dct2_fast(int M, const DctFastDesc &d, cudaStream_t st) { SHB_DCT_CASES(FFT_DCT2) }
static int dct3_fast(int M, const DctFastDesc &d, cudaStream_t st) { SHB_DCT_CASES(FFT_DCT3) }

int dct_fast_launch(int mode, const void *in, void *out, int64_t outer, int64_t inner, int n_phys, int n_spec,
                    double scale, const DctArgs &a, cudaStream_t st) {
    if (inner <= 1) return SHB_ERR_UNSUPPORTED;
    static const int sizes[] = {64, 128, 256, 512, 1024, 96, 192, 384, 768, 1536};
    bool ok = false;
    for (int x : sizes) ok = ok || (x == n_phys);
    if (!ok) return SHB_ERR_UNSUPPORTED;
    DctFastDesc d;
    d.in = in; d.out = out; d.outer = outer; d.inner = inner; d.nlines = outer * inner;
    if (d.nlines == 0) return SHB_OK;
    d.n_spec = n_spec; d.dim = a.dim; d.is_real = a.is_real; d.nsten = a.nsten; d.nbc = a.nbc;
    d.bc_lift = a.bc_lift; d.scale = scale;
    for (int i = 0; i < DCT_MAX_STENCIL; ++i) { d.soff[i] = a.soff[i]; d.sten[i] = a.sten[i]; }
    d.tw = twiddle_table(0, n_phys);
    d.tw3 = twiddle_table(2, n_phys);
    if (!d.tw || !d.tw3) { set_error("dct: twiddle table allocation failed"); return SHB_ERR_NOMEM; }
    return mode == FFT_DCT2 ? dct2_fast(n_phys, d, st) : dct3_fast(n_phys, d, st);
}

// ---- kernel A, 3/2-rule zero-padded backward transform split by output residue -------------
// Inverse DFT of length M = 3 M' of a line that holds N = 2 M' modes (rows M'+1 .. 2M'-1 of the
// padded line are zero).  Writing the output index as n = 3 n1 + n2 turns it into three
// independent M'-point inverse DFTs, one per residue n2, of the pre-twiddled sequence
//     Z_n2[k'] = w^(n2 k') (X[k'] + f X[M'+k']),  w = e^(2 pi i / M),  f = e^(4 pi i n2 / 3)
// (k' = 0 picks up the two halves of the split Nyquist mode: f = 1, -1/2, -1/2).  Each residue is
// its own 256-thread CTA with the register-resident M'-point schedule (3 CTAs per SM instead of one
// 768-thread CTA), the all-zero third of the padded line is never touched, and the input tile is
// re-read from L2 by the two sibling CTAs.  Reference semantics: fourier/bases.py:579-593.
template <class SC, int T, bool CH>
__global__ void __launch_bounds__(SC::P *T, (SC::P * T <= 256 ? 3 : 1)) fft_strided_pad3_kernel(const __grid_constant__ FastDesc d) {
    constexpr int MP = SC::M, P = SC::P;          // M' and threads per line
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *sm = reinterpret_cast<cplx *>(smem_raw);
    const int c = threadIdx.x % T;
    const int j = threadIdx.x / T;
    const int n2 = blockIdx.x % 3;
    const int64_t line = (int64_t)(blockIdx.x / 3) * T + c;
    const bool valid = line < d.nlines;
    const int64_t o = valid ? line / d.inner : 0, cc = valid ? line % d.inner : 0;
    const cplx *__restrict__ inb = reinterpret_cast<const cplx *>(d.in);
    cplx *__restrict__ outb = reinterpret_cast<cplx *>(d.out);
    auto in_at = [&](int r) -> const cplx & {
        if (CH && d.in_cr) {
            const int p = r / d.in_cr;
            return inb[p * d.in_cs + ((o * d.in_cr + (r - p * d.in_cr)) * d.inner + cc)];
        }
        return inb[(o * d.n_in + r) * d.inner + cc];
    };
    auto out_at = [&](int r) -> cplx & {
        if (CH && d.out_cr) {
            const int p = r / d.out_cr;
            const int64_t off = (o * d.out_cr + (r - p * d.out_cr)) * d.inner + cc;
            if (d.use_peer) return d.out_peer[p][off];
            return outb[p * d.out_cs + off];
        }
        return outb[(o * d.n_out + r) * d.inner + cc];
    };
    // f = e^(4 pi i n2 / 3)
    const double fr = (n2 == 0) ? 1.0 : -0.5;
    const double fi = (n2 == 0) ? 0.0 : (n2 == 1 ? -0.86602540378443864676372317075294 : 0.86602540378443864676372317075294);

    cplx v[SC::EMAX];
    if (j < SC::NB0) {
        cplx xa[SC::B0 * SC::R0], xb[SC::B0 * SC::R0];
#pragma unroll
        for (int q = 0; q < SC::B0; ++q)
#pragma unroll
            for (int m = 0; m < SC::R0; ++m) {
                const int k = (j + q * P) + m * SC::NB0;
                xa[q * SC::R0 + m] = valid ? in_at(k) : make_double2(0.0, 0.0);
                xb[q * SC::R0 + m] = valid ? in_at(k + MP) : make_double2(0.0, 0.0);
            }
#pragma unroll
        for (int q = 0; q < SC::B0; ++q)
#pragma unroll
            for (int m = 0; m < SC::R0; ++m) {
                const int k = (j + q * P) + m * SC::NB0;
                const cplx a = xa[q * SC::R0 + m], b = xb[q * SC::R0 + m];
                cplx z;
                if (k == 0) {
                    const double f0 = (n2 == 0) ? 1.0 : -0.5;
                    z = make_double2(fma(f0, b.x, a.x), fma(f0, b.y, a.y));
                } else {
                    z = make_double2(a.x + fr * b.x - fi * b.y, a.y + fr * b.y + fi * b.x);
                    if (n2 > 0) z = cmulc(z, __ldg(&d.tw2[n2 * k]));      // * e^(+2 pi i n2 k / M)
                }
                v[q * SC::R0 + m] = z;
            }
    }
    auto s = [&](int pos) -> cplx & { return sm[pos * T + c]; };
    fft_stages<SC, true, true>(v, j, s, CtaSync(), d.tw);
    const double scale = d.scale;
    if (j < SC::NBL && valid) {
#pragma unroll
        for (int q = 0; q < SC::BL; ++q)
#pragma unroll
            for (int m = 0; m < SC::RL; ++m) {
                const int n1 = (j + q * P) + m * SC::NBL;
                out_at(3 * n1 + n2) = cscale(v[q * SC::RL + m], scale);
            }
    }
}

// CURL (Navier-Stokes, first pass of the padded backward of [u_hat; curl u_hat]): the input holds the THREE
// velocity components only; a CTA owns one column tile and produces its six fields one after the other:
// the three velocity components, then fields 3..5 formed on the fly as i k x u_hat from two velocity
// lines and the wavenumber tables (demo/NavierStokes.py:37).  The velocity tile comes from HBM once -- the
// loads of the vorticity fields hit L2, the CTA read the same lines microseconds earlier -- and the
// vorticity spectrum is never written to or read from HBM.  (Six CTAs per tile, one per field, run
// together but each missed L2 on its own: no gain measured.)
template <class SC, int T, bool CH, bool CURL = false>
__global__ void __launch_bounds__(SC::P *T, (SC::P * T <= 256 ? 2 : 1)) fft_strided_pad3x_kernel(const __grid_constant__ FastDesc d) {
    static_assert(!(CH && CURL), "the fused curl runs on the plain layout");
    constexpr int MP = SC::M, P = SC::P;          // M' and threads per line
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *sm = reinterpret_cast<cplx *>(smem_raw);
    const int c = threadIdx.x % T;
    const int j = threadIdx.x / T;
    const int64_t line = (int64_t)blockIdx.x * T + c;
    const bool valid = CURL ? line < d.inner : line < d.nlines;
    const cplx *__restrict__ inb = reinterpret_cast<const cplx *>(d.in);
    cplx *__restrict__ outb = reinterpret_cast<cplx *>(d.out);
    const int64_t cc = CURL ? (valid ? line : 0) : (valid ? line % d.inner : 0);
    double kcol[3] = {0.0, 0.0, 0.0};
    if (CURL) {
        kcol[1] = __ldg(&d.k1[cc / d.curl_n2]);
        kcol[2] = __ldg(&d.k2[cc % d.curl_n2]);
    }
    auto s = [&](int pos) -> cplx & { return sm[pos * T + c]; };
    const double scale = d.scale;
#pragma unroll 1
    for (int fld = 0; fld < (CURL ? 6 : 1); ++fld) {
        const int64_t o = CURL ? (int64_t)fld : (valid ? line / d.inner : 0);
        // CURL: field 3 + ci = i (k_a u_b - k_b u_a), a = ci + 1, b = ci + 2 (mod 3); k_0 varies along the line
        const int ci = fld >= 3 ? fld - 3 : 0;
        const int fa = (ci + 1) % 3, fb = (ci + 2) % 3;
        auto in_at = [&](int r) -> cplx {
            if (CH && d.in_cr) {
                const int p = r / d.in_cr;
                return inb[p * d.in_cs + ((o * d.in_cr + (r - p * d.in_cr)) * d.inner + cc)];
            }
            if (CURL && fld >= 3) {
                const double k0r = __ldg(&d.k0[r]);
                const double ka = fa == 0 ? k0r : kcol[fa], kb = fb == 0 ? k0r : kcol[fb];
                const cplx ub = inb[((int64_t)fb * d.n_in + r) * d.inner + cc];
                const cplx ua = inb[((int64_t)fa * d.n_in + r) * d.inner + cc];
                const cplx t = make_double2(ka * ub.x - kb * ua.x, ka * ub.y - kb * ua.y);
                return make_double2(-t.y, t.x);
            }
            return inb[(o * d.n_in + r) * d.inner + cc];
        };
        auto out_at = [&](int r) -> cplx & {
            if (CH && d.out_cr) {
                const int p = r / d.out_cr;
                const int64_t off = (o * d.out_cr + (r - p * d.out_cr)) * d.inner + cc;
                if (d.use_peer) return d.out_peer[p][off];
                return outb[p * d.out_cs + off];
            }
            return outb[(o * d.n_out + r) * d.inner + cc];
        };
        // every thread keeps its 2 x EMAX inputs in registers and runs the three residue transforms
        // (n2 = 0, 1, 2) back to back: the inputs are read from HBM/L2 once instead of three times
        cplx xa[SC::B0 * SC::R0], xb[SC::B0 * SC::R0];
        if (j < SC::NB0) {
#pragma unroll
            for (int q = 0; q < SC::B0; ++q)
#pragma unroll
                for (int m = 0; m < SC::R0; ++m) {
                    const int k = (j + q * P) + m * SC::NB0;
                    xa[q * SC::R0 + m] = valid ? in_at(k) : make_double2(0.0, 0.0);
                    xb[q * SC::R0 + m] = valid ? in_at(k + MP) : make_double2(0.0, 0.0);
                }
        }
#pragma unroll 1
        for (int n2 = 0; n2 < 3; ++n2) {
            // f = e^(4 pi i n2 / 3)
            const double fr = (n2 == 0) ? 1.0 : -0.5;
            const double fi = (n2 == 0) ? 0.0 : (n2 == 1 ? -0.86602540378443864676372317075294 : 0.86602540378443864676372317075294);
            cplx v[SC::EMAX];
            if (j < SC::NB0) {
#pragma unroll
                for (int q = 0; q < SC::B0; ++q)
#pragma unroll
                    for (int m = 0; m < SC::R0; ++m) {
                        const int k = (j + q * P) + m * SC::NB0;
                        const cplx a = xa[q * SC::R0 + m], b = xb[q * SC::R0 + m];
                        cplx z;
                        if (k == 0) {
                            z = make_double2(fma(fr, b.x, a.x), fma(fr, b.y, a.y));
                        } else {
                            z = make_double2(a.x + fr * b.x - fi * b.y, a.y + fr * b.y + fi * b.x);
                            if (n2 > 0) z = cmulc(z, __ldg(&d.tw2[n2 * k]));      // * e^(+2 pi i n2 k / M)
                        }
                        v[q * SC::R0 + m] = z;
                    }
            }
            fft_stages<SC, true, true>(v, j, s, CtaSync(), d.tw);
            if (j < SC::NBL && valid) {
#pragma unroll
                for (int q = 0; q < SC::BL; ++q)
#pragma unroll
                    for (int m = 0; m < SC::RL; ++m) {
                        const int n1 = (j + q * P) + m * SC::NBL;
                        out_at(3 * n1 + n2) = cscale(v[q * SC::RL + m], scale);
                    }
            }
            __syncthreads();      // shared exchange buffer is reused by the next residue
        }
    }
}

template <class SC, int T>
static int launch_pad3(FastDesc d, cudaStream_t st) {
    constexpr size_t smem = (size_t)SC::M * T * sizeof(cplx);
    d.tw = twiddle_table(0, SC::M);
    d.tw2 = twiddle_table(0, 3 * SC::M);         // e^(-2 pi i k / M): conjugated in the kernel
    if (!d.tw || !d.tw2) { set_error("fft: twiddle table allocation failed"); return SHB_ERR_NOMEM; }
    const int64_t ntiles = (d.nlines + T - 1) / T;
    // default: one CTA runs the three residues over register-held inputs (inputs read once);
    // SHB_PAD3_SPLIT=1 selects the one-residue-per-CTA variant (inputs re-read through L2)
    static const bool split = getenv("SHB_PAD3_SPLIT") != nullptr;
    const bool ch = d.in_cr || d.out_cr;
    if (!split) {
        if (ch) {
            static bool a1 = false;
            if (!a1) { SHB_CUDA(cudaFuncSetAttribute(fft_strided_pad3x_kernel<SC, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); a1 = true; }
            fft_strided_pad3x_kernel<SC, T, true><<<(unsigned)ntiles, SC::P * T, smem, st>>>(d);
        } else {
            static bool a0 = false;
            if (!a0) { SHB_CUDA(cudaFuncSetAttribute(fft_strided_pad3x_kernel<SC, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); a0 = true; }
            fft_strided_pad3x_kernel<SC, T, false><<<(unsigned)ntiles, SC::P * T, smem, st>>>(d);
        }
    } else if (ch) {
        static bool a1 = false;
        if (!a1) { SHB_CUDA(cudaFuncSetAttribute(fft_strided_pad3_kernel<SC, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); a1 = true; }
        fft_strided_pad3_kernel<SC, T, true><<<(unsigned)(3 * ntiles), SC::P * T, smem, st>>>(d);
    } else {
        static bool a0 = false;
        if (!a0) { SHB_CUDA(cudaFuncSetAttribute(fft_strided_pad3_kernel<SC, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); a0 = true; }
        fft_strided_pad3_kernel<SC, T, false><<<(unsigned)(3 * ntiles), SC::P * T, smem, st>>>(d);
    }
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

template <class SC, int T>
static int launch_pad3_curl(FastDesc d, cudaStream_t st) {
    constexpr size_t smem = (size_t)SC::M * T * sizeof(cplx);
    d.tw = twiddle_table(0, SC::M);
    d.tw2 = twiddle_table(0, 3 * SC::M);
    if (!d.tw || !d.tw2) { set_error("fft: twiddle table allocation failed"); return SHB_ERR_NOMEM; }
    static bool a0 = false;
    if (!a0) { SHB_CUDA(cudaFuncSetAttribute(fft_strided_pad3x_kernel<SC, T, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); a0 = true; }
    const int64_t ntiles = (d.inner + T - 1) / T;
    fft_strided_pad3x_kernel<SC, T, false, true><<<(unsigned)ntiles, SC::P * T, smem, st>>>(d);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

// out[0..2] = padded backward of u_hat[0..2] along the leading axis, out[3..5] = the same of i k x u_hat
// (3/2 rule, N -> n_fft = 3N/2; u_hat: [3][N][inner], out: [6][n_fft][inner], inner = n1 * n2)
int fft_pad3_curl_launch(const void *u_hat, void *out, int64_t inner, int n2, int n_fft, int N, const double *k0,
                         const double *k1, const double *k2, double scale, cudaStream_t st) {
    if (N % 2 || 2 * n_fft != 3 * N || n2 < 1 || inner % n2) return SHB_ERR_UNSUPPORTED;
    FastDesc d;
    d.in_cr = d.out_cr = 0; d.use_peer = 0; d.in_cs = d.out_cs = 0;
    for (int p = 0; p < 8; ++p) d.out_peer[p] = nullptr;
    d.in = u_hat; d.out = out; d.outer = 6; d.inner = inner; d.nlines = 6 * inner;
    d.n_fft = n_fft; d.N = N; d.n_in = N; d.n_out = n_fft; d.flags = 0; d.scale = scale;
    d.cw = inner; d.cvalid = inner; d.c0 = 0; d.in_rs = inner; d.out_rs = inner;
    d.in_as = (int64_t)N * inner; d.out_as = (int64_t)n_fft * inner;
    d.k0 = k0; d.k1 = k1; d.k2 = k2; d.curl_n2 = n2;
    if (inner == 0) return SHB_OK;
    switch (N / 2) {
    case 64:  return launch_pad3_curl<Sched<64, 8, 8, 1, 8>, 8>(d, st);
    case 128: return launch_pad3_curl<Sched<128, 8, 4, 4, 32>, 8>(d, st);
    case 256: return launch_pad3_curl<Sched<256, 8, 8, 4, 32>, 8>(d, st);
    case 512: return launch_pad3_curl<Sched<512, 8, 8, 8, 64>, 8>(d, st);
    default: return SHB_ERR_UNSUPPORTED;
    }
}

static const bool g_no_pad3 = getenv("SHB_FFT_NO_PAD3") != nullptr;

// 3/2-rule backward with N == 2M/3, N even, no dealias_direct: residue-split kernel
static int strided_pad3(int M, const FastDesc &d, cudaStream_t st) {
    switch (M / 3) {
    case 64:  return launch_pad3<Sched<64, 8, 8, 1, 8>, 8>(d, st);
    case 128: return launch_pad3<Sched<128, 8, 4, 4, 32>, 8>(d, st);
    case 256: return launch_pad3<Sched<256, 8, 8, 4, 32>, 8>(d, st);      // 16-column tiles (512 threads, one CTA/SM): 0.68 / 1.09 vs 0.61 / 0.91 ms
    case 512: return launch_pad3<Sched<512, 8, 8, 8, 64>, 8>(d, st);
    default: return SHB_ERR_UNSUPPORTED;
    }
}


// ---- kernel B ---------------------------------------------------------------------
__device__ __forceinline__ int skew(int pos) { return pos + (pos >> 3); }

template <class SC, int MODE, int TL>
__global__ void __launch_bounds__(SC::P *TL) fft_contig_kernel(const __grid_constant__ FastDesc d) {
    constexpr bool INV = (MODE == FFT_C2C_BWD || MODE == FFT_C2R);
    constexpr int M = SC::M, P = SC::P;
    static_assert(P % 32 == 0, "a line is owned by whole warps");
    static_assert(P == 32 || TL <= 15, "one named barrier per line");
    constexpr int LP = M + (M >> 3) + 2;        // shared pitch of one line
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int j = threadIdx.x % P;
    const int l = threadIdx.x / P;
    cplx *sm = reinterpret_cast<cplx *>(smem_raw) + (size_t)l * LP;
    const int64_t line = (int64_t)blockIdx.x * TL + l;
    const bool valid = line < d.nlines;
    if (threadIdx.x == 0 && d.pf > 0) {
        // the (contiguous) input lines of the CTA `pf` blocks ahead are pulled into L2 by one bulk prefetch
        const int64_t l2 = ((int64_t)blockIdx.x + d.pf) * TL;
        if (l2 + TL <= d.nlines) {
            constexpr unsigned esz = (MODE == FFT_R2C) ? 8 : 16;
            const char *src = reinterpret_cast<const char *>(d.in) + l2 * d.n_in * esz;
            const unsigned bytes = TL * d.n_in * esz;
            if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && bytes % 16 == 0)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
        }
    }
    const int N = d.N;
    const int n_fft = d.n_fft;
    const bool padded = (N < n_fft);
    const bool evenN = (N % 2 == 0);
    const int h = N / 2;
    const double scale = d.scale;
    auto s = [&](int pos) -> cplx & { return sm[skew(pos)]; };
    const GroupSync<P> sync{l + 1};

    cplx v[SC::EMAX];
    if (MODE == FFT_C2C_FWD || MODE == FFT_C2C_BWD) {
        const cplx *__restrict__ in = reinterpret_cast<const cplx *>(d.in) + line * d.n_in;
        const PadMap pm(N, M, d.flags);
        if (j < SC::NB0) {
#pragma unroll
            for (int q = 0; q < SC::B0; ++q) {
#pragma unroll
                for (int m = 0; m < SC::R0; ++m) {
                    const int r = (j + q * P) + m * SC::NB0;
                    bool ok = valid;
                    int src = r;
                    double f = 1.0;
                    if (MODE == FFT_C2C_BWD) ok = pm.map(r, src, f) && valid;
                    cplx x = make_double2(0.0, 0.0);
                    if (ok) x = in[src];
                    v[q * SC::R0 + m] = (MODE == FFT_C2C_BWD) ? make_double2(x.x * f, x.y * f) : x;
                }
            }
        }
    } else if (MODE == FFT_R2C) {
        const double2 *__restrict__ in = reinterpret_cast<const double2 *>(
            reinterpret_cast<const double *>(d.in) + line * (int64_t)n_fft);
        if (j < SC::NB0) {
#pragma unroll
            for (int q = 0; q < SC::B0; ++q)
#pragma unroll
                for (int m = 0; m < SC::R0; ++m)
                    v[q * SC::R0 + m] = valid ? in[(j + q * P) + m * SC::NB0] : make_double2(0.0, 0.0);
        }
    } else {   // FFT_C2R: Y[n] = (X[n]+conj X[M-n]) + i e^{+i pi n/M} (X[n]-conj X[M-n])
        // X[n] and X[M-n] are both read straight from global memory: the mirrored
        // read is coalesced too (descending addresses) and hits L1/L2.
        const cplx *__restrict__ in = reinterpret_cast<const cplx *>(d.in) + line * d.n_in;
        const int ns = h + 1;
        const bool dd = (d.flags & 1) && !padded;
        if (j < SC::NB0) {
#pragma unroll
            for (int q = 0; q < SC::B0; ++q) {
                cplx wq = make_double2(1.0, 0.0);
                // operands are fetched in two batches to bound the live registers (2 x R0/2 pairs)
                constexpr int HB = (SC::R0 + 1) / 2;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    cplx xa[HB], xb[HB];
#pragma unroll
                    for (int mm = 0; mm < HB; ++mm) {
                        const int m = half * HB + mm;
                        if (m < SC::R0) {
                            const int n = (j + q * P) + m * SC::NB0;
                            const int nm = M - n;
                            xa[mm] = (valid && n < ns) ? in[n] : make_double2(0.0, 0.0);
                            xb[mm] = (valid && nm < ns) ? in[nm] : make_double2(0.0, 0.0);
                        }
                    }
#pragma unroll
                    for (int mm = 0; mm < HB; ++mm) {
                        const int m = half * HB + mm;
                        if (m < SC::R0) {
                            const int n = (j + q * P) + m * SC::NB0;
                            const int nm = M - n;
                            cplx xk = xa[mm], xc = xb[mm];
                            if (dd && n >= N / 3) xk = make_double2(0.0, 0.0);
                            if (dd && nm >= N / 3) xc = make_double2(0.0, 0.0);
                            if (padded && evenN && n == ns - 1) xk = make_double2(0.5 * xk.x, 0.0);
                            if (padded && evenN && nm == ns - 1) xc = make_double2(0.5 * xc.x, 0.0);
                            if (n == 0) { xk.y = 0.0; xc.y = 0.0; }      // DC and Nyquist (X[M]) are real
                            xc.y = -xc.y;
                            const cplx sum = cadd(xk, xc), df = csub(xk, xc);
                            cplx w;                                          // e^{-i pi n / M}
                            if constexpr (48 % (2 * SC::R0) == 0) {
                                if (m == 0) wq = __ldg(&d.tw2[j + q * P]);
                                w = root48_rt<48 / (2 * SC::R0)>(wq, m);
                            } else {
                                w = __ldg(&d.tw2[n]);
                            }
                            const cplx iw = make_double2(w.y, w.x);        // i * conj(w)
                            v[q * SC::R0 + m] = cadd(sum, cmul(iw, df));
                        }
                    }
                }
            }
        }
    }

    fft_stages<SC, INV, true>(v, j, s, sync, d.tw);

    if (MODE == FFT_C2C_BWD || (MODE == FFT_C2C_FWD && !padded)) {
        cplx *__restrict__ out = reinterpret_cast<cplx *>(d.out) + line * d.n_out;
        if (j < SC::NBL && valid) {
#pragma unroll
            for (int q = 0; q < SC::BL; ++q)
#pragma unroll
                for (int m = 0; m < SC::RL; ++m)
                    out[(j + q * P) + m * SC::NBL] = cscale(v[q * SC::RL + m], scale);
        }
    } else if (MODE == FFT_C2C_FWD) {
        cplx *__restrict__ out = reinterpret_cast<cplx *>(d.out) + line * d.n_out;
        if (evenN) {
            sync();
            if (j < SC::NBL) {
#pragma unroll
                for (int q = 0; q < SC::BL; ++q)
#pragma unroll
                    for (int m = 0; m < SC::RL; ++m)
                        if ((j + q * P) + m * SC::NBL == M - h) sm[0] = v[q * SC::RL + m];
            }
            sync();
        }
        if (j < SC::NBL && valid) {
#pragma unroll
            for (int q = 0; q < SC::BL; ++q) {
#pragma unroll
                for (int m = 0; m < SC::RL; ++m) {
                    const int r = (j + q * P) + m * SC::NBL;
                    const cplx x0 = v[q * SC::RL + m];
                    if (r <= h) {
                        cplx x = x0;
                        if (evenN && r == h) x = cadd(x, sm[0]);
                        out[r] = cscale(x, scale);
                    } else if (r > M - h || (!evenN && r == M - h)) {
                        out[r - M + N] = cscale(x0, scale);
                    }
                }
            }
        }
    } else if (MODE == FFT_R2C) {
        // X[k] = E + e^{-i pi k/M} O,  E = (Z[k]+conj Z[M-k])/2,  O = -i (Z[k]-conj Z[M-k])/2
        cplx *__restrict__ out = reinterpret_cast<cplx *>(d.out) + line * d.n_out;
        sync();
        if (j < SC::NBL) {
#pragma unroll
            for (int q = 0; q < SC::BL; ++q)
#pragma unroll
                for (int m = 0; m < SC::RL; ++m) s((j + q * P) + m * SC::NBL) = v[q * SC::RL + m];
        }
        sync();
        if (j < SC::NBL && valid) {
            const int ns = h + 1;
#pragma unroll
            for (int q = 0; q < SC::BL; ++q) {
                cplx wq = make_double2(1.0, 0.0);
#pragma unroll
                for (int m = 0; m < SC::RL; ++m) {
                    const int k = (j + q * P) + m * SC::NBL;
                    if (k < ns) {
                        const cplx zk = v[q * SC::RL + m];
                        const cplx zc = cconj(s(k == 0 ? 0 : M - k));
                        const cplx e = cscale(cadd(zk, zc), 0.5);
                        const cplx df = cscale(csub(zk, zc), 0.5);
                        const cplx o_ = make_double2(df.y, -df.x);
                        cplx w;
                        if constexpr (48 % (2 * SC::RL) == 0) {
                            if (m == 0) wq = __ldg(&d.tw2[j + q * P]);
                            w = root48_rt<48 / (2 * SC::RL)>(wq, m);
                        } else {
                            w = __ldg(&d.tw2[k]);
                        }
                        cplx x = cadd(e, cmul(w, o_));
                        if (padded && evenN && k == h) x = make_double2(2.0 * x.x, 0.0);
                        out[k] = cscale(x, scale);
                    }
                }
            }
            if (j == 0 && ns - 1 == M) {    // un-padded Nyquist: X[M] = Re Z0 - Im Z0
                const cplx z0 = v[0];
                out[M] = make_double2((z0.x - z0.y) * scale, 0.0);
            }
        }
    } else {   // FFT_C2R
        double2 *__restrict__ out = reinterpret_cast<double2 *>(
            reinterpret_cast<double *>(d.out) + line * (int64_t)n_fft);
        if (j < SC::NBL && valid) {
#pragma unroll
            for (int q = 0; q < SC::BL; ++q)
#pragma unroll
                for (int m = 0; m < SC::RL; ++m)
                    out[(j + q * P) + m * SC::NBL] = cscale(v[q * SC::RL + m], scale);
        }
    }
}

// ---- kernel B, C2R fused with the cross product (Navier-Stokes nonlinear term) --------------
// out = irfft(in) x b  on [3][L][n] arrays: a CTA transforms the three components of G spatial
// lines, exchanges them through shared memory and writes W_c = U_{c+1} b_{c+2} - U_{c+2} b_{c+1}
// directly, so the velocity never goes to HBM in physical space and the separate cross-product
// pass (read 6, write 3 arrays of the padded size) disappears
// (demo/NavierStokes.py:39-40: `U = U_hat.backward(U); W[:] = np.cross(U, curl_, axis=0)`).
template <class SC, int G, bool DUAL>
__global__ void __launch_bounds__(SC::P * (DUAL ? 6 : 3) * G) fft_c2r_cross_kernel(const __grid_constant__ FastDesc d,
                                                                                    const double *__restrict__ bvec,
                                                                                    const int64_t L) {
    // DUAL: `bvec` is a second SPECTRAL vector field (same layout as d.in); both fields are
    // transformed here (six lines per spatial line) and neither physical field reaches HBM.
    constexpr int M = SC::M, P = SC::P;
    constexpr int NSL = DUAL ? 6 : 3;       // line slots per spatial line
    static_assert(P % 32 == 0, "a line is owned by whole warps");
    static_assert(P == 32 || NSL * G <= 15, "one named barrier per line");
    constexpr int LP = M + (M >> 3) + 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int j = threadIdx.x % P;
    const int l = threadIdx.x / P;          // slot: NSL*g + 3*field + comp
    const int g = l / NSL, fld = (l % NSL) / 3, comp = l % 3;
    cplx *sm_all = reinterpret_cast<cplx *>(smem_raw);
    cplx *sm = sm_all + (size_t)l * LP;
    const int64_t sp = (int64_t)blockIdx.x * G + g;      // spatial line
    const bool valid = sp < L;
    const int64_t line = (int64_t)comp * L + sp;
    const int N = d.N;
    const int n_fft = d.n_fft;
    const bool padded = (N < n_fft);
    const bool evenN = (N % 2 == 0);
    const int h = N / 2;
    const double scale = d.scale;
    auto s = [&](int pos) -> cplx & { return sm[skew(pos)]; };
    const GroupSync<P> sync{l + 1};

    cplx v[SC::EMAX];
    {
        const cplx *__restrict__ in = (DUAL && fld ? reinterpret_cast<const cplx *>(bvec)
                                                   : reinterpret_cast<const cplx *>(d.in)) + (valid ? line : 0) * d.n_in;
        const int ns = h + 1;
        const bool dd = (d.flags & 1) && !padded;
        if (j < SC::NB0) {
#pragma unroll
            for (int q = 0; q < SC::B0; ++q) {
                cplx wq = make_double2(1.0, 0.0);
                constexpr int HB = (SC::R0 + 1) / 2;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    cplx xa[HB], xb[HB];
#pragma unroll
                    for (int mm = 0; mm < HB; ++mm) {
                        const int m = half * HB + mm;
                        if (m < SC::R0) {
                            const int n = (j + q * P) + m * SC::NB0;
                            const int nm = M - n;
                            xa[mm] = (valid && n < ns) ? in[n] : make_double2(0.0, 0.0);
                            xb[mm] = (valid && nm < ns) ? in[nm] : make_double2(0.0, 0.0);
                        }
                    }
#pragma unroll
                    for (int mm = 0; mm < HB; ++mm) {
                        const int m = half * HB + mm;
                        if (m < SC::R0) {
                            const int n = (j + q * P) + m * SC::NB0;
                            const int nm = M - n;
                            cplx xk = xa[mm], xc = xb[mm];
                            if (dd && n >= N / 3) xk = make_double2(0.0, 0.0);
                            if (dd && nm >= N / 3) xc = make_double2(0.0, 0.0);
                            if (padded && evenN && n == ns - 1) xk = make_double2(0.5 * xk.x, 0.0);
                            if (padded && evenN && nm == ns - 1) xc = make_double2(0.5 * xc.x, 0.0);
                            if (n == 0) { xk.y = 0.0; xc.y = 0.0; }
                            xc.y = -xc.y;
                            const cplx sum = cadd(xk, xc), df = csub(xk, xc);
                            cplx w;
                            if constexpr (48 % (2 * SC::R0) == 0) {
                                if (m == 0) wq = __ldg(&d.tw2[j + q * P]);
                                w = root48_rt<48 / (2 * SC::R0)>(wq, m);
                            } else {
                                w = __ldg(&d.tw2[n]);
                            }
                            const cplx iw = make_double2(w.y, w.x);
                            v[q * SC::R0 + m] = cadd(sum, cmul(iw, df));
                        }
                    }
                }
            }
        }
    }
    fft_stages<SC, true, true>(v, j, s, sync, d.tw);

    // publish this component's real line (as pairs) for the other two components of the same g
    __syncthreads();
    if (j < SC::NBL) {
        const double sc = (DUAL && fld) ? 1.0 : scale;
#pragma unroll
        for (int q = 0; q < SC::BL; ++q)
#pragma unroll
            for (int m = 0; m < SC::RL; ++m) s((j + q * P) + m * SC::NBL) = cscale(v[q * SC::RL + m], sc);
    }
    __syncthreads();
    if constexpr (DUAL) {
        // out[comp] = a[c1] b[c2] - a[c2] b[c1], a = field 0, b = field 1, all six lines in shared
        // memory; the two slots of a component each write half of its output line
        if (valid) {
            const int c1 = (comp + 1) % 3, c2 = (comp + 2) % 3;
            const cplx *a1 = sm_all + (size_t)(NSL * g + c1) * LP;
            const cplx *a2 = sm_all + (size_t)(NSL * g + c2) * LP;
            const cplx *w1p = sm_all + (size_t)(NSL * g + 3 + c1) * LP;
            const cplx *w2p = sm_all + (size_t)(NSL * g + 3 + c2) * LP;
            double2 *__restrict__ out = reinterpret_cast<double2 *>(
                reinterpret_cast<double *>(d.out) + line * (int64_t)n_fft);
            for (int idx = j; idx < M / 2; idx += P) {
                const int n = fld * (M / 2) + idx;
                const double2 x1 = a1[skew(n)], x2 = a2[skew(n)];
                const double2 w1 = w1p[skew(n)], w2 = w2p[skew(n)];
                out[n] = make_double2(x1.x * w2.x - x2.x * w1.x, x1.y * w2.y - x2.y * w1.y);
            }
        }
        return;
    }
    if (j < SC::NBL && valid) {
        const int c1 = (comp + 1) % 3, c2 = (comp + 2) % 3;
        const cplx *u1 = sm_all + (size_t)(3 * g + c1) * LP;
        const cplx *u2 = sm_all + (size_t)(3 * g + c2) * LP;
        const double2 *__restrict__ b1 = reinterpret_cast<const double2 *>(bvec + ((int64_t)c1 * L + sp) * n_fft);
        const double2 *__restrict__ b2 = reinterpret_cast<const double2 *>(bvec + ((int64_t)c2 * L + sp) * n_fft);
        double2 *__restrict__ out = reinterpret_cast<double2 *>(
            reinterpret_cast<double *>(d.out) + line * (int64_t)n_fft);
#pragma unroll
        for (int q = 0; q < SC::BL; ++q)
#pragma unroll
            for (int m = 0; m < SC::RL; ++m) {
                const int n = (j + q * P) + m * SC::NBL;
                const double2 x1 = u1[skew(n)], x2 = u2[skew(n)];
                const double2 w1 = b1[n], w2 = b2[n];
                out[n] = make_double2(x1.x * w2.x - x2.x * w1.x, x1.y * w2.y - x2.y * w1.y);
            }
    }
}

template <class SC, int G, bool DUAL>
static int launch_c2r_cross(const FastDesc &d, const double *b, int64_t L, cudaStream_t st) {
    constexpr int NSL = DUAL ? 6 : 3;
    constexpr size_t smem = (size_t)(SC::M + (SC::M >> 3) + 2) * NSL * G * sizeof(cplx);
    static bool attr_set = false;
    if (!attr_set) {
        SHB_CUDA(cudaFuncSetAttribute(fft_c2r_cross_kernel<SC, G, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        attr_set = true;
    }
    const int64_t ntiles = (L + G - 1) / G;
    fft_c2r_cross_kernel<SC, G, DUAL><<<(unsigned)ntiles, SC::P * NSL * G, smem, st>>>(d, b, L);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}


// distance (in CTAs' worth of work) of the L2 bulk prefetch of the C2R kernels below; SHB_PF_DIST=0: off
static int l2_prefetch_distance(int dflt) {
    static const int v = getenv("SHB_PF_DIST") ? atoi(getenv("SHB_PF_DIST")) : -1;
    return v < 0 ? dflt : v;
}

// ---- kernel B', 3/2-rule lines of two spectral vector fields: two real lines per complex transform --
// The fields a and b (N/2+1 Hermitian modes per line, N = 2 MP, n_fft = 3 MP) are transformed as the
// complex lines z_c = a_c + i b_c: spectrum Z[k] = A[k] + i B[k], Z[-k] = conj A[k] + i conj B[k],
// which is the 2 MP -> 3 MP complex padding of fft_strided_pad3x_kernel on contiguous lines (three
// residue transforms of MP points over register-held inputs).  The half-length pre-twiddle of the
// real transform disappears, three transforms of 3 MP points replace six of 3 MP / 2 + split, and the
// cross product is one fused multiply-add per output:
//     (a x b)_{c+2} = a_c b_{c+1} - a_{c+1} b_c = Im(conj(z_c) z_{c+1}).
// One warp per component, one CTA (three warps) per spatial line; every residue's values are handed to
// the neighbouring warp through the exchange buffer, the three output lines are assembled in shared
// memory and leave as 16-byte stores.
template <class SC>
__global__ void __launch_bounds__(96, 5) fft_c2r_cross3_kernel(const cplx *__restrict__ A, const cplx *__restrict__ B,
                                                               double *__restrict__ out, const int64_t L,
                                                               const cplx *__restrict__ tw, const cplx *__restrict__ tw3,
                                                               const double scale, const int pf_dist) {
    constexpr int MP = SC::M, P = SC::P;
    static_assert(P == 32, "one warp per line");
    constexpr int NIN = MP + 1, NF = 3 * MP;
    constexpr int LP = MP + (MP >> 3) + 2;
    constexpr int EM = SC::B0 * SC::R0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *ex_all = reinterpret_cast<cplx *>(smem_raw);
    double *ob_all = reinterpret_cast<double *>(ex_all + 3 * LP);
    const int j = threadIdx.x & 31, c = threadIdx.x >> 5;
    const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
    cplx *ex = ex_all + c * LP;
    const cplx *exn = ex_all + c1 * LP;
    double *ob = ob_all + c2 * NF;
    const int64_t sp = blockIdx.x;
    auto s = [&](int pos) -> cplx & { return ex[skew(pos)]; };
    struct WarpSync { __device__ __forceinline__ void operator()() const { __syncwarp(); } };
    // the lines of the CTA scheduled `pf_dist` spatial lines later are pulled into L2 now
    if (j == 0 && pf_dist > 0 && sp + pf_dist < L) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(A + ((int64_t)c * L + sp + pf_dist) * NIN), "n"(NIN * 16) : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(B + ((int64_t)c * L + sp + pf_dist) * NIN), "n"(NIN * 16) : "memory");
    }

    cplx xa[EM], xb[EM];
    if (j < SC::NB0) {
        const cplx *__restrict__ a = A + ((int64_t)c * L + sp) * NIN;
        const cplx *__restrict__ b = B + ((int64_t)c * L + sp) * NIN;
        cplx la[EM], lb[EM], ma[EM], mb[EM];
#pragma unroll
        for (int q = 0; q < SC::B0; ++q)
#pragma unroll
            for (int m = 0; m < SC::R0; ++m) {
                const int k = (j + q * P) + m * SC::NB0;
                la[q * SC::R0 + m] = __ldg(a + k);
                lb[q * SC::R0 + m] = __ldg(b + k);
                ma[q * SC::R0 + m] = __ldg(a + (MP - k));
                mb[q * SC::R0 + m] = __ldg(b + (MP - k));
            }
#pragma unroll
        for (int q = 0; q < SC::B0; ++q)
#pragma unroll
            for (int m = 0; m < SC::R0; ++m) {
                const int e = q * SC::R0 + m;
                const int k = (j + q * P) + m * SC::NB0;
                if (k == 0) {       // zero mode: imaginary parts dropped; slot MP: the split Nyquist mode
                    xa[e] = make_double2(la[e].x, lb[e].x);
                    xb[e] = make_double2(ma[e].x, mb[e].x);
                } else {
                    xa[e] = make_double2(la[e].x - lb[e].y, la[e].y + lb[e].x);
                    xb[e] = make_double2(ma[e].x + mb[e].y, mb[e].x - ma[e].y);
                }
            }
    }
#pragma unroll 1
    for (int n2 = 0; n2 < 3; ++n2) {
        // f = e^(4 pi i n2 / 3): the modes k - MP seen from output residue n2
        const double fr = (n2 == 0) ? 1.0 : -0.5;
        const double fi = (n2 == 0) ? 0.0 : (n2 == 1 ? -0.86602540378443864676372317075294 : 0.86602540378443864676372317075294);
        cplx v[SC::EMAX];
        if (j < SC::NB0) {
#pragma unroll
            for (int q = 0; q < SC::B0; ++q)
#pragma unroll
                for (int m = 0; m < SC::R0; ++m) {
                    const int k = (j + q * P) + m * SC::NB0;
                    const cplx a = xa[q * SC::R0 + m], b = xb[q * SC::R0 + m];
                    cplx z;
                    if (k == 0) {
                        z = make_double2(fma(fr, b.x, a.x), fma(fr, b.y, a.y));
                    } else {
                        z = make_double2(a.x + fr * b.x - fi * b.y, a.y + fr * b.y + fi * b.x);
                        if (n2 > 0) z = cmulc(z, __ldg(&tw3[n2 * k]));        // * e^(+2 pi i n2 k / (3 MP))
                    }
                    v[q * SC::R0 + m] = z;
                }
        }
        fft_stages<SC, true, true>(v, j, s, WarpSync(), tw);
        __syncwarp();
        if (j < SC::NBL) {
#pragma unroll
            for (int q = 0; q < SC::BL; ++q)
#pragma unroll
                for (int m = 0; m < SC::RL; ++m) s((j + q * P) + m * SC::NBL) = v[q * SC::RL + m];
        }
        __syncthreads();
        if (j < SC::NBL) {
#pragma unroll
            for (int q = 0; q < SC::BL; ++q)
#pragma unroll
                for (int m = 0; m < SC::RL; ++m) {
                    const int n1 = (j + q * P) + m * SC::NBL;
                    const cplx z = v[q * SC::RL + m], zn = exn[skew(n1)];
                    const double o = z.x * zn.y - z.y * zn.x;
                    ob[3 * n1 + n2] = (scale == 1.0) ? o : o * scale;
                }
        }
        if (n2 == 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // `ob` is read by the bulk-copy engine
        __syncthreads();      // the exchange buffers are rewritten by the next residue
    }
    // output line c (assembled by warp c+1) leaves as one bulk copy shared -> global: no LDS / STG wavefronts
    if (j == 0) {
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(ob_all + c * NF);
        double *dst = out + ((int64_t)c * L + sp) * NF;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "n"(NF * 8) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

template <class SC>
static int launch_c2r_cross3(const void *a, const void *b, void *out, int64_t L, double scale, cudaStream_t st) {
    constexpr int MP = SC::M;
    constexpr size_t smem = (size_t)3 * (MP + (MP >> 3) + 2) * sizeof(cplx) + (size_t)9 * MP * sizeof(double);
    const cplx *tw = twiddle_table(0, MP), *tw3 = twiddle_table(0, 3 * MP);
    if (!tw || !tw3) { set_error("fft: twiddle table allocation failed"); return SHB_ERR_NOMEM; }
    static bool attr_set = false;
    if (!attr_set) {
        SHB_CUDA(cudaFuncSetAttribute(fft_c2r_cross3_kernel<SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    fft_c2r_cross3_kernel<SC><<<(unsigned)L, 96, smem, st>>>(reinterpret_cast<const cplx *>(a), reinterpret_cast<const cplx *>(b),
                                                            reinterpret_cast<double *>(out), L, tw, tw3, scale, l2_prefetch_distance(384));
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

// ---- kernel B'', plain C2R on 3/2-rule lines: the same two-lines-per-complex-transform scheme for
// independent lines.  A warp owns the adjacent lines 2p and 2p+1 (z = x_2p + i x_2p+1), runs the three
// residue transforms, splits real / imaginary parts into the two output lines in shared memory and
// sends them off as one bulk copy; no CTA-wide barrier.  An odd line count pairs the last line with zero.
template <class SC, int W>
__global__ void __launch_bounds__(32 * W, 15 / W) fft_c2r_pair3_kernel(const cplx *__restrict__ X, double *__restrict__ out,
                                                                      const int64_t nlines, const cplx *__restrict__ tw,
                                                                      const cplx *__restrict__ tw3, const double scale,
                                                                      const int pf_dist) {
    constexpr int MP = SC::M, P = SC::P;
    static_assert(P == 32, "one warp per line pair");
    constexpr int NIN = MP + 1, NF = 3 * MP;
    constexpr int LP = MP + (MP >> 3) + 2;
    constexpr int EM = SC::B0 * SC::R0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int j = threadIdx.x & 31, w = threadIdx.x >> 5;
    cplx *ex = reinterpret_cast<cplx *>(smem_raw) + w * LP;
    double *ob = reinterpret_cast<double *>(reinterpret_cast<cplx *>(smem_raw) + W * LP) + w * 2 * NF;
    const int64_t p = (int64_t)blockIdx.x * W + w;
    if (2 * p >= nlines) return;
    const bool has_b = 2 * p + 1 < nlines;
    auto s = [&](int pos) -> cplx & { return ex[skew(pos)]; };
    struct WarpSync { __device__ __forceinline__ void operator()() const { __syncwarp(); } };
    // the lines of a warp scheduled `pf_dist` pairs later are pulled into L2 now (one bulk prefetch)
    if (j == 0 && pf_dist > 0 && 2 * (p + pf_dist) + 1 < nlines)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(X + 2 * (p + pf_dist) * NIN), "n"(2 * NIN * 16) : "memory");

    cplx xa[EM], xb[EM];
    if (j < SC::NB0) {
        const cplx *__restrict__ a = X + 2 * p * NIN;
        const cplx *__restrict__ b = has_b ? a + NIN : a;
        const double bm = has_b ? 1.0 : 0.0;
        cplx la[EM], lb[EM], ma[EM], mb[EM];
#pragma unroll
        for (int q = 0; q < SC::B0; ++q)
#pragma unroll
            for (int m = 0; m < SC::R0; ++m) {
                const int k = (j + q * P) + m * SC::NB0;
                la[q * SC::R0 + m] = __ldg(a + k);
                lb[q * SC::R0 + m] = __ldg(b + k);
                ma[q * SC::R0 + m] = __ldg(a + (MP - k));
                mb[q * SC::R0 + m] = __ldg(b + (MP - k));
            }
#pragma unroll
        for (int q = 0; q < SC::B0; ++q)
#pragma unroll
            for (int m = 0; m < SC::R0; ++m) {
                const int e = q * SC::R0 + m;
                const int k = (j + q * P) + m * SC::NB0;
                const cplx B1 = cscale(lb[e], bm), B2 = cscale(mb[e], bm);
                if (k == 0) {       // zero mode: imaginary parts dropped; slot MP: the split Nyquist mode
                    xa[e] = make_double2(la[e].x, B1.x);
                    xb[e] = make_double2(ma[e].x, B2.x);
                } else {
                    xa[e] = make_double2(la[e].x - B1.y, la[e].y + B1.x);
                    xb[e] = make_double2(ma[e].x + B2.y, B2.x - ma[e].y);
                }
            }
    }
#pragma unroll 1
    for (int n2 = 0; n2 < 3; ++n2) {
        const double fr = (n2 == 0) ? 1.0 : -0.5;
        const double fi = (n2 == 0) ? 0.0 : (n2 == 1 ? -0.86602540378443864676372317075294 : 0.86602540378443864676372317075294);
        cplx v[SC::EMAX];
        if (j < SC::NB0) {
#pragma unroll
            for (int q = 0; q < SC::B0; ++q)
#pragma unroll
                for (int m = 0; m < SC::R0; ++m) {
                    const int k = (j + q * P) + m * SC::NB0;
                    const cplx a = xa[q * SC::R0 + m], b = xb[q * SC::R0 + m];
                    cplx z;
                    if (k == 0) {
                        z = make_double2(fma(fr, b.x, a.x), fma(fr, b.y, a.y));
                    } else {
                        z = make_double2(a.x + fr * b.x - fi * b.y, a.y + fr * b.y + fi * b.x);
                        if (n2 > 0) z = cmulc(z, __ldg(&tw3[n2 * k]));
                    }
                    v[q * SC::R0 + m] = z;
                }
        }
        fft_stages<SC, true, true>(v, j, s, WarpSync(), tw);
        if (j < SC::NBL) {
#pragma unroll
            for (int q = 0; q < SC::BL; ++q)
#pragma unroll
                for (int m = 0; m < SC::RL; ++m) {
                    const int n1 = (j + q * P) + m * SC::NBL;
                    const cplx z = (scale == 1.0) ? v[q * SC::RL + m] : cscale(v[q * SC::RL + m], scale);
                    ob[3 * n1 + n2] = z.x;
                    ob[NF + 3 * n1 + n2] = z.y;
                }
        }
        __syncwarp();      // the exchange buffer is rewritten by the next residue
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (j == 0) {
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(ob);
        double *dst = out + 2 * p * NF;
        const uint32_t bytes = (has_b ? 2 : 1) * NF * 8;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

template <class SC>
static int launch_c2r_pair3(const void *in, void *out, int64_t nlines, double scale, cudaStream_t st) {
    constexpr int MP = SC::M, W = 3;
    constexpr size_t smem = (size_t)W * ((MP + (MP >> 3) + 2) * sizeof(cplx) + (size_t)6 * MP * sizeof(double));
    const cplx *tw = twiddle_table(0, MP), *tw3 = twiddle_table(0, 3 * MP);
    if (!tw || !tw3) { set_error("fft: twiddle table allocation failed"); return SHB_ERR_NOMEM; }
    static bool attr_set = false;
    if (!attr_set) {
        SHB_CUDA(cudaFuncSetAttribute(fft_c2r_pair3_kernel<SC, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int64_t pairs = (nlines + 1) / 2;
    fft_c2r_pair3_kernel<SC, W><<<(unsigned)((pairs + W - 1) / W), 32 * W, smem, st>>>(
        reinterpret_cast<const cplx *>(in), reinterpret_cast<double *>(out), nlines, tw, tw3, scale, l2_prefetch_distance(1024));
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

// 3/2-rule C2R lines (n_fft = 3 N / 2, N a multiple of 4, no dealias_direct): SHB_ERR_UNSUPPORTED when not covered
static int c2r_pair3(const void *in, void *out, int64_t nlines, int n_fft, int N, double scale, int flags, cudaStream_t st) {
    static const bool off = getenv("SHB_NO_PAIR3") != nullptr;
    if (off || (flags & 1) || N % 4 || 2 * n_fft != 3 * N || (reinterpret_cast<uintptr_t>(out) & 15) || nlines > 0x3fffffff)
        return SHB_ERR_UNSUPPORTED;
    switch (N / 2) {
    case 32:  return launch_c2r_pair3<Sched<32, 8, 4, 1, 32>>(in, out, nlines, scale, st);
    case 64:  return launch_c2r_pair3<Sched<64, 8, 8, 1, 32>>(in, out, nlines, scale, st);
    case 128: return launch_c2r_pair3<Sched<128, 8, 4, 4, 32>>(in, out, nlines, scale, st);
    case 256: return launch_c2r_pair3<Sched<256, 8, 8, 4, 32>>(in, out, nlines, scale, st);
    }
    return SHB_ERR_UNSUPPORTED;
}

// flags bit 0: dealias_direct; bit 1: `b` is a spectral vector field (both fields transformed here)
int fft_c2r_cross_launch(const void *in, const void *b, void *out, int64_t L, int n_fft, int N, double scale,
                         int flags, cudaStream_t st) {
    if (n_fft % 2) return SHB_ERR_UNSUPPORTED;
    const int M = n_fft / 2;
    const bool dual = flags & 2;
    FastDesc d;
    d.in = in; d.out = out; d.outer = 3 * L; d.inner = 1; d.nlines = 3 * L;
    d.n_fft = n_fft; d.N = N; d.flags = flags & 1; d.scale = scale;
    d.n_in = N / 2 + 1; d.n_out = n_fft;
    d.in_cr = d.out_cr = 0; d.in_cs = d.out_cs = 0; d.use_peer = 0;
    for (int p = 0; p < 8; ++p) d.out_peer[p] = nullptr;
    if (L == 0) return SHB_OK;
    // both fields spectral on 3/2-rule lines: the two-lines-per-complex-transform kernel (SHB_NO_CROSS3=1: off)
    static const bool no_cross3 = getenv("SHB_NO_CROSS3") != nullptr;
    if (dual && !(flags & 1) && !no_cross3 && N % 4 == 0 && 2 * n_fft == 3 * N && L <= 0x7fffffff &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        switch (N / 2) {
        case 32:  return launch_c2r_cross3<Sched<32, 8, 4, 1, 32>>(in, b, out, L, scale, st);
        case 64:  return launch_c2r_cross3<Sched<64, 8, 8, 1, 32>>(in, b, out, L, scale, st);
        case 128: return launch_c2r_cross3<Sched<128, 8, 4, 4, 32>>(in, b, out, L, scale, st);
        case 256: return launch_c2r_cross3<Sched<256, 8, 8, 4, 32>>(in, b, out, L, scale, st);
        }
    }
    static const int sizes[] = {32, 64, 128, 256, 384, 512, 768};
    bool ok = false;
    for (int x : sizes) ok = ok || (x == M);
    if (!ok) return SHB_ERR_UNSUPPORTED;
    d.tw = twiddle_table(0, M);
    d.tw2 = twiddle_table(1, M);
    if (!d.tw || !d.tw2) { set_error("fft: twiddle table allocation failed"); return SHB_ERR_NOMEM; }
    const double *bb = reinterpret_cast<const double *>(b);
    if (dual) {
        switch (M) {
        case 32:  return launch_c2r_cross<Sched<32, 8, 4, 1, 32>, 1, true>(d, bb, L, st);
        case 64:  return launch_c2r_cross<Sched<64, 8, 8, 1, 32>, 1, true>(d, bb, L, st);
        case 128: return launch_c2r_cross<Sched<128, 8, 4, 4, 32>, 1, true>(d, bb, L, st);
        case 256: return launch_c2r_cross<Sched<256, 8, 8, 4, 32>, 1, true>(d, bb, L, st);
        case 384: return launch_c2r_cross<Sched<384, 8, 8, 6, 64>, 1, true>(d, bb, L, st);
        case 512: return launch_c2r_cross<Sched<512, 8, 8, 8, 64>, 1, true>(d, bb, L, st);
        case 768: return launch_c2r_cross<Sched<768, 12, 8, 8, 96>, 1, true>(d, bb, L, st);
        }
        return SHB_ERR_UNSUPPORTED;
    }
    switch (M) {
    case 32:  return launch_c2r_cross<Sched<32, 8, 4, 1, 32>, 2, false>(d, bb, L, st);
    case 64:  return launch_c2r_cross<Sched<64, 8, 8, 1, 32>, 2, false>(d, bb, L, st);
    case 128: return launch_c2r_cross<Sched<128, 8, 4, 4, 32>, 2, false>(d, bb, L, st);
    case 256: return launch_c2r_cross<Sched<256, 8, 8, 4, 32>, 2, false>(d, bb, L, st);
    case 384: return launch_c2r_cross<Sched<384, 8, 8, 6, 64>, 2, false>(d, bb, L, st);
    case 512: return launch_c2r_cross<Sched<512, 8, 8, 8, 64>, 1, false>(d, bb, L, st);
    case 768: return launch_c2r_cross<Sched<768, 12, 8, 8, 96>, 1, false>(d, bb, L, st);
    }
    return SHB_ERR_UNSUPPORTED;
}

// ---- host-side dispatch ---------------------------------------------------------------
template <class SC, int MODE, int T, bool CH>
static int launch_strided_ch(const FastDesc &d, cudaStream_t st) {
    constexpr size_t smem = (size_t)SC::M * T * sizeof(cplx);

    static bool attr_set = false;
    if (!attr_set) {
        SHB_CUDA(cudaFuncSetAttribute(fft_strided_kernel<SC, MODE, T, CH>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int64_t ntiles = (d.nlines + T - 1) / T;
    fft_strided_kernel<SC, MODE, T, CH><<<(unsigned)ntiles, SC::P * T, smem, st>>>(d);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

template <class SC, int MODE, int T, bool CH>
static int launch_strided_pipe(const FastDesc &d, cudaStream_t st) {
    constexpr size_t smem = 2 * (size_t)SC::M * T * sizeof(cplx);
    static_assert(smem <= 227 * 1024, "pipelined kernel does not fit in shared memory");
    static bool attr_set = false;
    if (!attr_set) {
        SHB_CUDA(cudaFuncSetAttribute(fft_strided_pipe_kernel<SC, MODE, T, CH>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int64_t ntiles = (d.nlines + T - 1) / T;
    const int per_sm = (smem <= 110 * 1024 && SC::P * T <= 512) ? 2 : 1;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    const int64_t grid = ntiles < cap ? ntiles : cap;
    fft_strided_pipe_kernel<SC, MODE, T, CH><<<(unsigned)grid, SC::P * T, smem, st>>>(d, ntiles);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

static const bool g_no_pipe = getenv("SHB_FFT_NO_PIPE") != nullptr;
static const bool g_peer_t16 = getenv("SHB_PEER_T8") == nullptr;   // 256 B peer stores: +6% over 128 B on NVLink 5
static const bool g_t4 = getenv("SHB_T4") != nullptr;

template <class SC, int MODE, int T>
static int launch_strided(const FastDesc &d, cudaStream_t st) {
    // one CTA per SM anyway (>= 768 threads): use the persistent cp.async-pipelined variant
    constexpr bool big = (SC::P * T > 512) && (2 * (size_t)SC::M * T * sizeof(cplx) <= 227 * 1024);
    if constexpr (big) {
        if (!g_no_pipe) {
            if (d.in_cr || d.out_cr) return launch_strided_pipe<SC, MODE, T, true>(d, st);
            return launch_strided_pipe<SC, MODE, T, false>(d, st);
        }
    }
    if (d.in_cr || d.out_cr) return launch_strided_ch<SC, MODE, T, true>(d, st);
    return launch_strided_ch<SC, MODE, T, false>(d, st);
}

template <class SC, int MODE, int TL>
static int launch_contig(FastDesc d, cudaStream_t st) {
    constexpr size_t smem = (size_t)(SC::M + (SC::M >> 3) + 2) * TL * sizeof(cplx);
    {
        // about SHB_PF_CONTIG_MB (default 8) megabytes of input ahead of the running CTAs; 0: off
        static const double mb = getenv("SHB_PF_CONTIG_MB") ? atof(getenv("SHB_PF_CONTIG_MB")) : 8.0;
        const double per_cta = (double)TL * d.n_in * (MODE == FFT_R2C ? 8 : 16);
        d.pf = mb > 0 ? (int)(mb * 1048576.0 / per_cta) : 0;
    }
    static bool attr_set = false;
    if (!attr_set) {
        SHB_CUDA(cudaFuncSetAttribute(fft_contig_kernel<SC, MODE, TL>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int64_t ntiles = (d.nlines + TL - 1) / TL;
    fft_contig_kernel<SC, MODE, TL><<<(unsigned)ntiles, SC::P * TL, smem, st>>>(d);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

// Schedules.  Backward (zero-padded input) puts the widest radix first so the
// all-zero operands are never loaded; forward (truncated output) puts it last.
#define SHB_STRIDED_CASES(MODE, FIRSTWIDE)                                                          \
    switch (M) {                                                                                    \
    case 64:   return launch_strided<Sched<64, 8, 8, 1, 8>, MODE, 8>(d, st);                        \
    case 128:  return launch_strided<Sched<128, 8, 4, 4, 32>, MODE, 8>(d, st);                      \
    case 256:  return launch_strided<Sched<256, 8, 8, 4, 32>, MODE, 8>(d, st);                      \
    case 512:  if (d.use_peer && g_peer_t16) return launch_strided_ch<Sched<512, 8, 8, 8, 64>, MODE, 16, true>(d, st); \
               return launch_strided<Sched<512, 8, 8, 8, 64>, MODE, 8>(d, st);                      \
    case 1024: return launch_strided<Sched<1024, 16, 8, 8, 128>, MODE, 4>(d, st);                   \
    case 96:   return FIRSTWIDE ? launch_strided<Sched<96, 12, 8, 1, 12>, MODE, 8>(d, st)           \
                                : launch_strided<Sched<96, 8, 12, 1, 12>, MODE, 8>(d, st);          \
    case 192:  return FIRSTWIDE ? launch_strided<Sched<192, 6, 8, 4, 48>, MODE, 8>(d, st)           \
                                : launch_strided<Sched<192, 4, 8, 6, 48>, MODE, 8>(d, st);          \
    case 384:  return FIRSTWIDE ? launch_strided<Sched<384, 6, 8, 8, 64>, MODE, 8>(d, st)           \
                                : launch_strided<Sched<384, 8, 8, 6, 64>, MODE, 8>(d, st);          \
    case 768:  if (g_t4) return FIRSTWIDE ? launch_strided_pipe<Sched<768, 12, 8, 8, 96>, MODE, 4, false>(d, st) \
                                : launch_strided_pipe<Sched<768, 8, 8, 12, 96>, MODE, 4, false>(d, st);         \
               return FIRSTWIDE ? launch_strided<Sched<768, 12, 8, 8, 96>, MODE, 8>(d, st)          \
                                : launch_strided<Sched<768, 8, 8, 12, 96>, MODE, 8>(d, st);         \
    case 1536: return FIRSTWIDE ? launch_strided<Sched<1536, 12, 16, 8, 192>, MODE, 4>(d, st)       \
                                : launch_strided<Sched<1536, 8, 16, 12, 192>, MODE, 4>(d, st);      \
    case 1152: return launch_strided<Sched<1152, 12, 12, 8, 144>, MODE, 4>(d, st);                  \
    default: return SHB_ERR_UNSUPPORTED;                                                            \
    }

#define SHB_CONTIG_CASES(MODE)                                                                      \
    switch (M) {                                                                                    \
    case 32:   return launch_contig<Sched<32, 8, 4, 1, 32>, MODE, 8>(d, st);                        \
    case 64:   return launch_contig<Sched<64, 8, 8, 1, 32>, MODE, 8>(d, st);                        \
    case 128:  return launch_contig<Sched<128, 8, 4, 4, 32>, MODE, 8>(d, st);                       \
    case 256:  return launch_contig<Sched<256, 8, 8, 4, 32>, MODE, 8>(d, st);                       \
    case 512:  return launch_contig<Sched<512, 8, 8, 8, 64>, MODE, 4>(d, st);                       \
    case 1024: return launch_contig<Sched<1024, 16, 8, 8, 128>, MODE, 2>(d, st);                    \
    case 96:   return launch_contig<Sched<96, 12, 8, 1, 32>, MODE, 8>(d, st);                       \
    case 192:  return launch_contig<Sched<192, 6, 8, 4, 64>, MODE, 4>(d, st);                       \
    case 384:  return launch_contig<Sched<384, 8, 8, 6, 64>, MODE, 4>(d, st);                       \
    case 768:  return launch_contig<Sched<768, 12, 8, 8, 96>, MODE, 4>(d, st);                      \
    case 576:  return launch_contig<Sched<576, 6, 12, 8, 96>, MODE, 4>(d, st);                      \
    default: return SHB_ERR_UNSUPPORTED;                                                            \
    }

static int strided_fwd(int M, const FastDesc &d, cudaStream_t st) { SHB_STRIDED_CASES(FFT_C2C_FWD, false) }
static int strided_bwd(int M, const FastDesc &d, cudaStream_t st) { SHB_STRIDED_CASES(FFT_C2C_BWD, true) }
static int contig_c2c_fwd(int M, const FastDesc &d, cudaStream_t st) { SHB_CONTIG_CASES(FFT_C2C_FWD) }
static int contig_c2c_bwd(int M, const FastDesc &d, cudaStream_t st) { SHB_CONTIG_CASES(FFT_C2C_BWD) }
static int contig_r2c(int M, const FastDesc &d, cudaStream_t st) { SHB_CONTIG_CASES(FFT_R2C) }
static int contig_c2r(int M, const FastDesc &d, cudaStream_t st) { SHB_CONTIG_CASES(FFT_C2R) }

int fft_fast_launch(int mode, const void *in, void *out, int64_t outer, int64_t inner,
                    int n_fft, int N, double scale, int flags, cudaStream_t st, int in_cr, int out_cr,
                    void *const *peer_out, int npeer) {
    FastDesc d;
    d.in_cr = in_cr; d.out_cr = out_cr;
    d.use_peer = (peer_out != nullptr && npeer > 0);
    for (int p = 0; p < 8; ++p) d.out_peer[p] = (d.use_peer && p < npeer) ? reinterpret_cast<cplx *>(peer_out[p]) : nullptr;
    d.in_cs = (int64_t)outer * in_cr * inner; d.out_cs = (int64_t)outer * out_cr * inner;
    d.in = in; d.out = out; d.outer = outer; d.inner = inner; d.nlines = outer * inner;
    d.n_fft = n_fft; d.N = N; d.flags = flags; d.scale = scale;
    if (d.nlines == 0) return SHB_OK;
    {
        const int n_in = (mode == FFT_C2C_BWD) ? N : n_fft, n_out = (mode == FFT_C2C_BWD) ? n_fft : N;
        d.cw = inner; d.cvalid = inner; d.c0 = 0; d.in_rs = inner; d.out_rs = inner;
        d.in_as = (int64_t)(in_cr ? in_cr : n_in) * inner;
        d.out_as = (int64_t)(out_cr ? out_cr : n_out) * inner;
    }
    const int ns = N / 2 + 1;
    int M;
    switch (mode) {
    case FFT_C2C_FWD: M = n_fft; d.n_in = n_fft; d.n_out = N; break;
    case FFT_C2C_BWD: M = n_fft; d.n_in = N; d.n_out = n_fft; break;
    case FFT_R2C: if (n_fft % 2 || inner != 1) return SHB_ERR_UNSUPPORTED;
                  M = n_fft / 2; d.n_in = n_fft; d.n_out = ns; break;
    case FFT_C2R: if (n_fft % 2 || inner != 1) return SHB_ERR_UNSUPPORTED;
                  if (!in_cr && !out_cr && !d.use_peer) {
                      const int rc = c2r_pair3(in, out, d.nlines, n_fft, N, scale, flags, st);
                      if (rc != SHB_ERR_UNSUPPORTED) return rc;
                  }
                  M = n_fft / 2; d.n_in = ns; d.n_out = n_fft; break;
    default: return SHB_ERR_UNSUPPORTED;
    }
    // cheap support test before touching the twiddle cache
    auto supported = [&](int m, bool strided) {
        static const int a[] = {64, 128, 256, 512, 1024, 96, 192, 384, 768, 1536, 1152};
        static const int b[] = {32, 64, 128, 256, 512, 1024, 96, 192, 384, 768, 576};
        if (strided) { for (int x : a) if (x == m) return true; }
        else { for (int x : b) if (x == m) return true; }
        return false;
    };
    const bool strided = inner > 1;
    if (!supported(M, strided)) return SHB_ERR_UNSUPPORTED;
    if ((in_cr || out_cr) && !strided) return SHB_ERR_UNSUPPORTED;
    d.tw = twiddle_table(0, M);
    d.tw2 = (mode == FFT_R2C || mode == FFT_C2R) ? twiddle_table(1, M) : nullptr;
    if (!d.tw || ((mode == FFT_R2C || mode == FFT_C2R) && !d.tw2)) {
        set_error("fft: twiddle table allocation failed");
        return SHB_ERR_NOMEM;
    }
    if (strided && mode == FFT_C2C_BWD && !g_no_pad3 && M % 3 == 0 && 3 * N == 2 * M && N % 2 == 0 && !(flags & 1)) {
        const int rc = strided_pad3(M, d, st);
        if (rc != SHB_ERR_UNSUPPORTED) return rc;
    }
    if (strided) return mode == FFT_C2C_FWD ? strided_fwd(M, d, st) : strided_bwd(M, d, st);
    switch (mode) {
    case FFT_C2C_FWD: return contig_c2c_fwd(M, d, st);
    case FFT_C2C_BWD: return contig_c2c_bwd(M, d, st);
    case FFT_R2C: return contig_r2c(M, d, st);
    default: return contig_c2r(M, d, st);
    }
}

// C2C along a strided axis with fully general line addressing (FastDesc::cw ... out_rs): the
// column-chunked slab exchange.  Only the sizes served by fft_strided_kernel (not its persistent
// pipelined variant) are covered.
int fft_fast_launch_general(int mode, const void *in, void *out, int64_t na, int64_t ncols, int64_t ncols_valid, int64_t c0, int64_t in_as,
                            int64_t in_rs, int64_t out_as, int64_t out_rs, int n_fft, int N, double scale, int flags,
                            cudaStream_t st, int in_cr, int64_t in_cs, int out_cr, int64_t out_cs,
                            void *const *peer_out, int npeer) {
    if (mode != FFT_C2C_FWD && mode != FFT_C2C_BWD) return SHB_ERR_UNSUPPORTED;
    FastDesc d;
    d.in_cr = in_cr; d.out_cr = out_cr; d.in_cs = in_cs; d.out_cs = out_cs;
    d.use_peer = (peer_out != nullptr && npeer > 0);
    for (int p = 0; p < 8; ++p) d.out_peer[p] = (d.use_peer && p < npeer) ? reinterpret_cast<cplx *>(peer_out[p]) : nullptr;
    d.in = in; d.out = out; d.outer = na; d.inner = ncols; d.nlines = na * ncols;
    d.n_fft = n_fft; d.N = N; d.flags = flags; d.scale = scale;
    d.cw = ncols; d.cvalid = ncols_valid; d.c0 = c0; d.in_as = in_as; d.in_rs = in_rs; d.out_as = out_as; d.out_rs = out_rs;
    if (d.nlines == 0) return SHB_OK;
    const int M = n_fft;
    d.n_in = (mode == FFT_C2C_BWD) ? N : n_fft; d.n_out = (mode == FFT_C2C_BWD) ? n_fft : N;
    d.tw = twiddle_table(0, M);
    d.tw2 = nullptr;
    if (!d.tw) { set_error("fft: twiddle table allocation failed"); return SHB_ERR_NOMEM; }
    using S64 = Sched<64, 8, 8, 1, 8>;
    using S128 = Sched<128, 8, 4, 4, 32>;
    using S256 = Sched<256, 8, 8, 4, 32>;
    using S512 = Sched<512, 8, 8, 8, 64>;
    using S1024 = Sched<1024, 16, 8, 8, 128>;
#define SHB_GEN_CASE(MM, SCHED, TT)                                                                       \
    case MM: return mode == FFT_C2C_FWD ? launch_strided_ch<SCHED, FFT_C2C_FWD, TT, true>(d, st)          \
                                        : launch_strided_ch<SCHED, FFT_C2C_BWD, TT, true>(d, st);
    switch (M) {
        SHB_GEN_CASE(64, S64, 8)
        SHB_GEN_CASE(128, S128, 8)
        SHB_GEN_CASE(256, S256, 8)
        case 512:
            if (d.use_peer && g_peer_t16)
                return mode == FFT_C2C_FWD ? launch_strided_ch<S512, FFT_C2C_FWD, 16, true>(d, st)
                                           : launch_strided_ch<S512, FFT_C2C_BWD, 16, true>(d, st);
            return mode == FFT_C2C_FWD ? launch_strided_ch<S512, FFT_C2C_FWD, 8, true>(d, st)
                                       : launch_strided_ch<S512, FFT_C2C_BWD, 8, true>(d, st);
        SHB_GEN_CASE(1024, S1024, 4)
    default: return SHB_ERR_UNSUPPORTED;
    }
#undef SHB_GEN_CASE
}

}  // namespace shb
