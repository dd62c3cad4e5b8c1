// gemm.cu -- fp64 tensor-core contraction along one axis:
//     out[o][i][c] = sum_j A[i][j] * u[o][j][c]        A real (M x K), u real or complex
// used for the Legendre / Jacobi Vandermonde transforms (operator A = stencil x
// Vandermonde x quadrature weights, built once on the host).
//
// The data array is viewed as real with W = inner * (complex ? 2 : 1) contiguous
// columns, so one kernel serves real and complex fields (A is real).  Each CTA owns a
// 128 x 64 output tile of one outer index and walks K in chunks of 32 through a cp.async
// shared-memory pipeline (16-byte copies where the rows are 16 B aligned); 8 warps each hold a
// 32 x 32 sub-tile as 2 x 4 m16n8k16 DMMA accumulators (mma.sync f64 -- tcgen05 has no fp64 kind;
// ptxas lowers the shape to eight DMMA.8x8x4, the only fp64 tensor instruction of sm_100a;
// SHB_GEMM_K4=1 issues the m8n8k4 shape directly).  Two stages of 55 KB keep TWO CTAs on an SM:
// measured at the 256^3 Legendre shape 29.9 TFLOP/s (0.93 of cuBLAS DGEMM on the same shape); three
// stages (one CTA per SM) 25.1, three stages of K = 16 28.9.
//
// Reference semantics restated: shenfun/spectralbase.py:473-483 (evaluate_expansion_all,
// 'vandermonde': np.dot(P, u_hat) along the axis) and :539-548 (scalar product).
#include "common.cuh"
#include <stdlib.h>

namespace shb {

constexpr int GM_BM = 128, GM_BN = 64, GM_BK = 32, GM_THREADS = 256;
constexpr int GM_AS = GM_BK + 4;     // A tile row pitch (doubles): pitch % 16 == 4 -> conflict-free fragments
constexpr int GM_XS = GM_BN + 4;     // X tile row pitch: pitch % 16 == 4 (the 16 lanes of a half-warp hit 16 distinct 8-byte banks)
constexpr int GM_STAGES = 2;
constexpr size_t GM_SMEM = GM_STAGES * ((size_t)GM_BM * GM_AS + (size_t)GM_BK * GM_XS) * sizeof(double);

struct GemmDesc {
    const double *A;       // [M][K] row-major
    const double *X;
    double *C;
    int M, K;
    int64_t W;             // real columns
    int64_t nbatch;
    int64_t xs_batch, xs_k, xs_w;   // element strides of X
    int64_t cs_batch, cs_m, cs_w;   // element strides of C
};

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// m16n8k16: A fragment a[i] = A[g + 8*(i%2)][t + 4*(i/2)], B fragment b[i] = B[t + 4*i][g],
// C fragment c[0..1] = C[g][2t + {0,1}], c[2..3] = C[g+8][2t + {0,1}]   (g = lane/4, t = lane%4)
__device__ __forceinline__ void dmma16(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                 "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
                   "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__device__ __forceinline__ void cp_async16g(void *smem_dst, const void *gsrc, bool pred) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc, bool pred) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}

// CTA tile 128 x 64, K chunks of 32 through a GM_STAGES-deep cp.async pipeline; 8 warps (4 x 2), each
// holding a 32 x 32 sub-tile.
template <bool K16>
__global__ void __launch_bounds__(GM_THREADS, 2) gemm_axis_kernel(const __grid_constant__ GemmDesc d) {
    extern __shared__ __align__(16) unsigned char gm_smem[];
    double *sA = reinterpret_cast<double *>(gm_smem);                       // [stage][BM][AS]
    double *sX = sA + (size_t)GM_STAGES * GM_BM * GM_AS;                    // [stage][BK][XS]
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1;          // 0..3 : 32 rows each
    const int wn = warp & 1;           // 0..1 : 32 cols each
    const int64_t w0 = (int64_t)blockIdx.x * GM_BN;
    const int m0 = blockIdx.y * GM_BM;
    const int64_t batch = blockIdx.z;
    const double *X = d.X + batch * d.xs_batch;
    double *C = d.C + batch * d.cs_batch;

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const bool w_fast = (d.xs_w == 1);
    // 16-byte copies: rows of A (pitch K) and of X (pitch xs_k) start 16 B aligned and hold an even number of doubles
    const bool a16 = (d.K % 2 == 0) && ((reinterpret_cast<uintptr_t>(d.A) & 15) == 0);
    const bool x16 = w_fast && (d.xs_k % 2 == 0) && (d.xs_batch % 2 == 0) && (d.W % 2 == 0) &&
                     ((reinterpret_cast<uintptr_t>(d.X) & 15) == 0);
    auto load_tiles = [&](int buf, int k0) {
        double *a = sA + (size_t)buf * GM_BM * GM_AS;
        double *x = sX + (size_t)buf * GM_BK * GM_XS;
        // A tile: 128 x 32 doubles, k fastest (A is row-major)
        if (a16) {
#pragma unroll
            for (int r = 0; r < (GM_BM * GM_BK) / (2 * GM_THREADS); ++r) {
                const int e = tid + r * GM_THREADS;
                const int kk = (e % (GM_BK / 2)) * 2, mm = e / (GM_BK / 2);
                const int gm = m0 + mm, gk = k0 + kk;
                const bool ok = (gm < d.M && gk < d.K);
                cp_async16g(&a[mm * GM_AS + kk], ok ? &d.A[(int64_t)gm * d.K + gk] : d.A, ok);
            }
        } else {
#pragma unroll
            for (int r = 0; r < (GM_BM * GM_BK) / GM_THREADS; ++r) {
                const int e = tid + r * GM_THREADS;
                const int kk = e % GM_BK, mm = e / GM_BK;
                const int gm = m0 + mm, gk = k0 + kk;
                const bool ok = (gm < d.M && gk < d.K);
                cp_async8(&a[mm * GM_AS + kk], ok ? &d.A[(int64_t)gm * d.K + gk] : d.A, ok);
            }
        }
        // X tile: 32 x 64 doubles; fastest index follows the unit stride
        if (x16) {
#pragma unroll
            for (int r = 0; r < (GM_BK * GM_BN) / (2 * GM_THREADS); ++r) {
                const int e = tid + r * GM_THREADS;
                const int ww = (e % (GM_BN / 2)) * 2, kk = e / (GM_BN / 2);
                const int gk = k0 + kk;
                const int64_t gw = w0 + ww;
                const bool ok = (gk < d.K && gw < d.W);
                cp_async16g(&x[kk * GM_XS + ww], ok ? &X[(int64_t)gk * d.xs_k + gw] : X, ok);
            }
        } else {
#pragma unroll
            for (int r = 0; r < (GM_BK * GM_BN) / GM_THREADS; ++r) {
                const int e = tid + r * GM_THREADS;
                int kk, ww;
                if (w_fast) { ww = e % GM_BN; kk = e / GM_BN; } else { kk = e % GM_BK; ww = e / GM_BK; }
                const int gk = k0 + kk;
                const int64_t gw = w0 + ww;
                const bool ok = (gk < d.K && gw < d.W);
                cp_async8(&x[kk * GM_XS + ww], ok ? &X[(int64_t)gk * d.xs_k + gw * d.xs_w] : X, ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int nk = (d.K + GM_BK - 1) / GM_BK;
    // three stages: chunk kt+2 is in flight while chunk kt is multiplied (one barrier per chunk)
    // one commit group per chunk, always (empty groups past the end keep the wait count uniform)
#pragma unroll
    for (int p = 0; p < GM_STAGES - 1; ++p) {
        if (p < nk) load_tiles(p, p * GM_BK);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt % GM_STAGES;
        asm volatile("cp.async.wait_group %0;" ::"n"(GM_STAGES - 2) : "memory");
        __syncthreads();               // chunk kt landed for everyone; chunk kt-1 has been consumed by everyone
        if (kt + GM_STAGES - 1 < nk) load_tiles((kt + GM_STAGES - 1) % GM_STAGES, (kt + GM_STAGES - 1) * GM_BK);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        const double *a = sA + (size_t)buf * GM_BM * GM_AS;
        const double *x = sX + (size_t)buf * GM_BK * GM_XS;
        const int g = lane >> 2, t = lane & 3;
        if (K16) {
            // acc[i][j] (i: 8-row block, j: 8-column block) doubles as the m16n8 accumulator of row blocks (2i', 2i'+1)
#pragma unroll
            for (int ks = 0; ks < GM_BK; ks += 16) {
                double af[2][8], bf[4][4];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        af[i][e] = a[(wm * 32 + i * 16 + g + 8 * (e & 1)) * GM_AS + ks + t + 4 * (e >> 1)];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) bf[j][e] = x[(ks + t + 4 * e) * GM_XS + wn * 32 + j * 8 + g];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        double c[4] = {acc[2 * i][j][0], acc[2 * i][j][1], acc[2 * i + 1][j][0], acc[2 * i + 1][j][1]};
                        dmma16(c, af[i], bf[j]);
                        acc[2 * i][j][0] = c[0]; acc[2 * i][j][1] = c[1];
                        acc[2 * i + 1][j][0] = c[2]; acc[2 * i + 1][j][1] = c[3];
                    }
            }
        } else {
            // fragments of step ks+4 are fetched before the 16 DMMAs of step ks are issued
            double af[2][4], bf[2][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[0][i] = a[(wm * 32 + i * 8 + g) * GM_AS + t];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[0][j] = x[t * GM_XS + wn * 32 + j * 8 + g];
#pragma unroll
            for (int ks = 0; ks < GM_BK; ks += 4) {
                const int cur = (ks >> 2) & 1, nxt = cur ^ 1;
                if (ks + 4 < GM_BK) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) af[nxt][i] = a[(wm * 32 + i * 8 + g) * GM_AS + ks + 4 + t];
#pragma unroll
                    for (int j = 0; j < 4; ++j) bf[nxt][j] = x[(ks + 4 + t) * GM_XS + wn * 32 + j * 8 + g];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
            }
        }
    }
    // epilogue: lane holds C[row = lane/4][col = (lane%4)*2 + {0,1}] of each 8x8 tile
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + wm * 32 + i * 8 + (lane >> 2);
        if (gm >= d.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t gw = w0 + wn * 32 + j * 8 + (lane & 3) * 2;
            if (d.cs_w == 1 && gw + 1 < d.W && ((((int64_t)gm * d.cs_m + gw) & 1) == 0) &&
                ((reinterpret_cast<uintptr_t>(C) & 15) == 0)) {
                *reinterpret_cast<double2 *>(&C[(int64_t)gm * d.cs_m + gw]) = make_double2(acc[i][j][0], acc[i][j][1]);
            } else {
                if (gw < d.W) C[(int64_t)gm * d.cs_m + gw * d.cs_w] = acc[i][j][0];
                if (gw + 1 < d.W) C[(int64_t)gm * d.cs_m + (gw + 1) * d.cs_w] = acc[i][j][1];
            }
        }
    }
}

// ---- parity-split contraction -----------------------------------------------------------------------
// Legendre / Jacobi(a,a) / Chebyshev functions are even or odd and their quadrature points are symmetric,
// so an operator matrix of these families satisfies either
//   FWD (values -> coefficients):  A[i][K-1-j] = (-1)^i A[i][j]   ->  even rows act on s_j = u_j + u_{K-1-j},
//                                                                     odd rows on d_j = u_j - u_{K-1-j},
//   BWD (coefficients -> values):  A[M-1-i][k] = (-1)^k A[i][k]   ->  e_i = sum_{k even}, o_i = sum_{k odd},
//                                                                     out_i = e_i + o_i, out_{M-1-i} = e_i - o_i
// and the contraction is two products of half the size: HALF the tensor-core work of the dense product.
// The butterflies cost nothing extra here: FWD forms s, d in registers when the B fragments are read
// from the two staged row tiles (rows j and their mirrors), BWD adds / subtracts the two accumulator
// sets in the epilogue.  `A2` holds the two half operators [2][Mh][Kp] (Kp even, zero padded), built
// on the host by ops.gemm_axis after testing the structure numerically.
// CTA tile: 64 rows of EACH set x 64 columns, 16 contraction indices of each set per stage.
constexpr int GP_BM = 64, GP_BN = 64, GP_BK = 16, GP_AS = GP_BK + 4, GP_XS = GP_BN + 4, GP_STAGES = 3;
constexpr size_t GP_SMEM = GP_STAGES * 2 * ((size_t)GP_BM * GP_AS + (size_t)GP_BK * GP_XS) * sizeof(double);

struct ParityDesc {
    const double *A2;
    const double *X;
    double *C;
    int M, K, Mh, Kh, Kp;
    int tri;               // MODE 2: +1 upper triangular halves (A2[p][i][k] = 0 for k < i), -1 lower, 0 full
    int64_t W, xs_batch, xs_k, cs_batch, cs_m;
};

// MODE 0: row-mirror structure (coefficients -> values), 1: column-mirror structure (values -> coefficients),
// 2: checkerboard operator (A[i][j] = 0 for i + j odd: the Legendre <-> Chebyshev connection matrices): even rows
//    act on even columns, odd rows on odd columns, no butterflies; triangular halves skip their zero k-tiles.
template <int MODE>
__global__ void __launch_bounds__(GM_THREADS, 2) gemm_axis_parity_kernel(const __grid_constant__ ParityDesc d) {
    constexpr bool FWD = (MODE == 1);
    extern __shared__ __align__(16) unsigned char gm_smem[];
    double *sA = reinterpret_cast<double *>(gm_smem);                           // [stage][set][BM][AS]
    double *sX = sA + (size_t)GP_STAGES * 2 * GP_BM * GP_AS;                    // [stage][set][BK][XS]
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;      // 16 rows of each set x 32 columns per warp
    const int g = lane >> 2, t = lane & 3;
    const int64_t w0 = (int64_t)blockIdx.x * GP_BN;
    const int m0 = blockIdx.y * GP_BM;
    const double *X = d.X + (int64_t)blockIdx.z * d.xs_batch;
    double *C = d.C + (int64_t)blockIdx.z * d.cs_batch;

    double acc[2][4][4];
#pragma unroll
    for (int s_ = 0; s_ < 2; ++s_)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[s_][j][e] = 0.0;

    auto load_tiles = [&](int buf, int k0) {
        double *a = sA + (size_t)buf * 2 * GP_BM * GP_AS;
        double *x = sX + (size_t)buf * 2 * GP_BK * GP_XS;
#pragma unroll
        for (int r = 0; r < (2 * GP_BM * GP_BK) / (2 * GM_THREADS); ++r) {
            const int e = tid + r * GM_THREADS;
            const int set = e / (GP_BM * GP_BK / 2), e2 = e % (GP_BM * GP_BK / 2);
            const int kk = (e2 % (GP_BK / 2)) * 2, mm = e2 / (GP_BK / 2);
            const int gm = m0 + mm, gk = k0 + kk;
            const bool ok = (gm < d.Mh && gk < d.Kp);
            cp_async16g(&a[(set * GP_BM + mm) * GP_AS + kk], ok ? &d.A2[((int64_t)set * d.Mh + gm) * d.Kp + gk] : d.A2, ok);
        }
#pragma unroll
        for (int r = 0; r < (2 * GP_BK * GP_BN) / (2 * GM_THREADS); ++r) {
            const int e = tid + r * GM_THREADS;
            const int set = e / (GP_BK * GP_BN / 2), e2 = e % (GP_BK * GP_BN / 2);
            const int ww = (e2 % (GP_BN / 2)) * 2, kk = e2 / (GP_BN / 2);
            const int gk = k0 + kk;
            const int row = FWD ? (set == 0 ? gk : d.K - 1 - gk) : 2 * gk + set;      // MODE 0 and 2: even / odd rows
            const int64_t gw = w0 + ww;
            const bool ok = (gk < d.Kh && row < d.K && gw < d.W);
            cp_async16g(&x[(set * GP_BK + kk) * GP_XS + ww], ok ? &X[(int64_t)row * d.xs_k + gw] : X, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int nk = (d.Kh + GP_BK - 1) / GP_BK, kbeg = 0;
    if (MODE == 2 && d.tri > 0) kbeg = m0 / GP_BK;                                     // zero tiles left of the diagonal
    if (MODE == 2 && d.tri < 0) { const int e = (m0 + GP_BM + GP_BK - 1) / GP_BK; nk = e < nk ? e : nk; }
#pragma unroll
    for (int p = 0; p < GP_STAGES - 1; ++p) {
        if (kbeg + p < nk) load_tiles(p, (kbeg + p) * GP_BK);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kt = kbeg; kt < nk; ++kt) {
        const int buf = (kt - kbeg) % GP_STAGES;
        asm volatile("cp.async.wait_group %0;" ::"n"(GP_STAGES - 2) : "memory");
        __syncthreads();
        if (kt + GP_STAGES - 1 < nk) load_tiles((kt - kbeg + GP_STAGES - 1) % GP_STAGES, (kt + GP_STAGES - 1) * GP_BK);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        const double *a = sA + (size_t)buf * 2 * GP_BM * GP_AS;
        const double *x0 = sX + (size_t)buf * 2 * GP_BK * GP_XS;
        const double *x1 = x0 + GP_BK * GP_XS;
        double af[2][8];
#pragma unroll
        for (int s_ = 0; s_ < 2; ++s_)
#pragma unroll
            for (int e = 0; e < 8; ++e)
                af[s_][e] = a[(s_ * GP_BM + wm * 16 + g + 8 * (e & 1)) * GP_AS + t + 4 * (e >> 1)];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double b0[4], b1[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const double p0 = x0[(t + 4 * e) * GP_XS + wn * 32 + j * 8 + g];
                const double p1 = x1[(t + 4 * e) * GP_XS + wn * 32 + j * 8 + g];
                b0[e] = FWD ? p0 + p1 : p0;
                b1[e] = FWD ? p0 - p1 : p1;
            }
            dmma16(acc[0][j], af[0], b0);
            dmma16(acc[1][j], af[1], b1);
        }
    }
    // epilogue: c[0..1] = C[g][2t + {0,1}], c[2..3] = C[g + 8][2t + {0,1}] of each 16 x 8 tile
    const bool c16 = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (d.cs_m % 2 == 0);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = m0 + wm * 16 + g + 8 * h;       // row of the half problem
        if (i >= d.Mh) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t gw = w0 + wn * 32 + j * 8 + t * 2;
            const double e0 = acc[0][j][2 * h], e1 = acc[0][j][2 * h + 1];
            const double o0 = acc[1][j][2 * h], o1 = acc[1][j][2 * h + 1];
            constexpr bool SPLIT = (MODE != 0);      // outputs are the even / odd rows themselves
            const int ra = SPLIT ? 2 * i : i, rb = SPLIT ? 2 * i + 1 : d.M - 1 - i;
            const double a0 = SPLIT ? e0 : e0 + o0, a1 = SPLIT ? e1 : e1 + o1;
            const double b0 = SPLIT ? o0 : e0 - o0, b1 = SPLIT ? o1 : e1 - o1;
            if (c16 && gw + 1 < d.W) {
                if (ra < d.M) *reinterpret_cast<double2 *>(&C[(int64_t)ra * d.cs_m + gw]) = make_double2(a0, a1);
                if (rb < d.M) *reinterpret_cast<double2 *>(&C[(int64_t)rb * d.cs_m + gw]) = make_double2(b0, b1);
            } else {
                if (ra < d.M && gw < d.W) C[(int64_t)ra * d.cs_m + gw] = a0;
                if (ra < d.M && gw + 1 < d.W) C[(int64_t)ra * d.cs_m + gw + 1] = a1;
                if (rb < d.M && gw < d.W) C[(int64_t)rb * d.cs_m + gw] = b0;
                if (rb < d.M && gw + 1 < d.W) C[(int64_t)rb * d.cs_m + gw + 1] = b1;
            }
        }
    }
}

}  // namespace shb

using namespace shb;

extern "C" {

int shb_gemm_axis(const double *d_A, const void *d_u, void *d_out, int M, int K, int64_t outer, int64_t inner,
                  int is_complex, void *stream) {
    SHB_REQUIRE(d_A && d_u && d_out && M >= 1 && K >= 1 && outer >= 0 && inner >= 0, "bad arguments");
    if (d_u == d_out) { set_error("gemm_axis: in-place is not supported"); return SHB_ERR_ARG; }
    if (outer * inner == 0) return SHB_OK;
    GemmDesc d;
    d.A = d_A; d.X = (const double *)d_u; d.C = (double *)d_out; d.M = M; d.K = K;
    const int64_t W = inner * (is_complex ? 2 : 1);
    if (W > 1) {
        d.W = W; d.nbatch = outer;
        d.xs_batch = (int64_t)K * W; d.xs_k = W; d.xs_w = 1;
        d.cs_batch = (int64_t)M * W; d.cs_m = W; d.cs_w = 1;
    } else {
        // lines are contiguous: out[o][i] = sum_j A[i][j] u[o][j]  ->  columns are the lines
        d.W = outer; d.nbatch = 1;
        d.xs_batch = 0; d.xs_k = 1; d.xs_w = K;
        d.cs_batch = 0; d.cs_m = 1; d.cs_w = M;
    }
    if (d.nbatch > 65535) { set_error("gemm_axis: outer extent %lld too large", (long long)d.nbatch); return SHB_ERR_UNSUPPORTED; }
    static const bool k4 = getenv("SHB_GEMM_K4") != nullptr;
    // (set on every call: the attribute is per device, and the call costs well under a microsecond)
    SHB_CUDA(cudaFuncSetAttribute(gemm_axis_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GM_SMEM));
    SHB_CUDA(cudaFuncSetAttribute(gemm_axis_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GM_SMEM));
    dim3 grid((unsigned)((d.W + GM_BN - 1) / GM_BN), (unsigned)((M + GM_BM - 1) / GM_BM), (unsigned)d.nbatch);
    if (k4) gemm_axis_kernel<false><<<grid, GM_THREADS, GM_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(d);
    else gemm_axis_kernel<true><<<grid, GM_THREADS, GM_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(d);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

/* Parity-split contraction (see gemm_axis_parity_kernel): A2 = the two half operators [2][Mh][Kp] built by the
 * caller; forward != 0: column-mirror structure (K even), else row-mirror structure (M even).  Needs inner*(complex?2:1)
 * > 1 and even, 16-byte aligned arrays; SHB_ERR_UNSUPPORTED otherwise (use shb_gemm_axis). */
int shb_gemm_axis_parity(const double *d_A2, const void *d_u, void *d_out, int M, int K, int Kp, int forward,
                         int64_t outer, int64_t inner, int is_complex, void *stream) {
    const int mode = forward & 3, tri = (forward & 4) ? 1 : ((forward & 8) ? -1 : 0);
    forward = (mode == 1);
    SHB_REQUIRE(d_A2 && d_u && d_out && M >= 1 && K >= 1 && Kp >= 2 && Kp % 2 == 0 && outer >= 0 && inner >= 0, "bad arguments");
    if (d_u == d_out) { set_error("gemm_axis_parity: in-place is not supported"); return SHB_ERR_ARG; }
    if (outer * inner == 0) return SHB_OK;
    const int64_t W = inner * (is_complex ? 2 : 1);
    if (W < 2 || W % 2 || outer > 65535 || (mode == 1 && K % 2) || (mode == 0 && M % 2) || mode == 3 ||
        ((reinterpret_cast<uintptr_t>(d_A2) | reinterpret_cast<uintptr_t>(d_u) | reinterpret_cast<uintptr_t>(d_out)) & 15))
        return SHB_ERR_UNSUPPORTED;
    ParityDesc d;
    d.A2 = d_A2; d.X = (const double *)d_u; d.C = (double *)d_out; d.M = M; d.K = K; d.Kp = Kp;
    d.Mh = mode == 0 ? M / 2 : (M + 1) / 2;
    d.Kh = mode == 1 ? K / 2 : (K + 1) / 2;
    d.tri = mode == 2 ? tri : 0;
    if (d.Kh > Kp) { set_error("gemm_axis_parity: operator pitch %d < %d", Kp, d.Kh); return SHB_ERR_ARG; }
    d.W = W; d.xs_batch = (int64_t)K * W; d.xs_k = W; d.cs_batch = (int64_t)M * W; d.cs_m = W;
    SHB_CUDA(cudaFuncSetAttribute(gemm_axis_parity_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GP_SMEM));
    SHB_CUDA(cudaFuncSetAttribute(gemm_axis_parity_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GP_SMEM));
    SHB_CUDA(cudaFuncSetAttribute(gemm_axis_parity_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GP_SMEM));
    dim3 grid((unsigned)((W + GP_BN - 1) / GP_BN), (unsigned)((d.Mh + GP_BM - 1) / GP_BM), (unsigned)outer);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (mode == 1) gemm_axis_parity_kernel<1><<<grid, GM_THREADS, GP_SMEM, st>>>(d);
    else if (mode == 2) gemm_axis_parity_kernel<2><<<grid, GM_THREADS, GP_SMEM, st>>>(d);
    else gemm_axis_parity_kernel<0><<<grid, GM_THREADS, GP_SMEM, st>>>(d);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

}  // extern "C"
