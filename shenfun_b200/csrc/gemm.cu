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
constexpr int GM_XS = GM_BN + 8;     // X tile row pitch: pitch % 16 == 8
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

}  // extern "C"
