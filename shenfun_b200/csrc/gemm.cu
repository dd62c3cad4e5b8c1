// gemm.cu -- fp64 tensor-core contraction along one axis:
//     out[o][i][c] = sum_j A[i][j] * u[o][j][c]        A real (M x K), u real or complex
// used for the Legendre / Jacobi Vandermonde transforms (operator A = stencil x
// Vandermonde x quadrature weights, built once on the host).
//
// The data array is viewed as real with W = inner * (complex ? 2 : 1) contiguous
// columns, so one kernel serves real and complex fields (A is real).  Each CTA owns a
// 64 x 64 output tile of one outer index and walks K in chunks of 16 through a
// double-buffered shared-memory pipeline; 8 warps each hold a 32 x 16 sub-tile as
// 4 x 2 m8n8k4 DMMA accumulators (mma.sync f64 -- tcgen05 has no fp64 kind).
//
// Reference semantics restated: shenfun/spectralbase.py:473-483 (evaluate_expansion_all,
// 'vandermonde': np.dot(P, u_hat) along the axis) and :539-548 (scalar product).
#include "common.cuh"

namespace shb {

constexpr int GM_BM = 64, GM_BN = 64, GM_BK = 16, GM_THREADS = 256;
constexpr int GM_APAD = 4;    // A tile stored [BM][BK+4]  (conflict-free fragment reads)
constexpr int GM_XPAD = 8;    // X tile stored [BK][BN+8]

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

__global__ void __launch_bounds__(GM_THREADS) gemm_axis_kernel(const __grid_constant__ GemmDesc d) {
    __shared__ double sA[2][GM_BM][GM_BK + GM_APAD];
    __shared__ double sX[2][GM_BK][GM_BN + GM_XPAD];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2;          // 0..1 : 32 rows each
    const int wn = warp & 3;           // 0..3 : 16 cols each
    const int64_t w0 = (int64_t)blockIdx.x * GM_BN;
    const int m0 = blockIdx.y * GM_BM;
    const int64_t batch = blockIdx.z;
    const double *X = d.X + batch * d.xs_batch;
    double *C = d.C + batch * d.cs_batch;

    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const bool w_fast = (d.xs_w == 1);
    auto load_tiles = [&](int buf, int k0) {
        // A tile: 64 x 16 = 1024 elements, 4 per thread, k fastest (A row-major)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int e = tid + r * GM_THREADS;
            const int kk = e % GM_BK, mm = e / GM_BK;
            const int gm = m0 + mm, gk = k0 + kk;
            sA[buf][mm][kk] = (gm < d.M && gk < d.K) ? __ldg(&d.A[(int64_t)gm * d.K + gk]) : 0.0;
        }
        // X tile: 16 x 64 elements, 4 per thread; fastest index follows the unit stride
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int e = tid + r * GM_THREADS;
            int kk, ww;
            if (w_fast) { ww = e % GM_BN; kk = e / GM_BN; } else { kk = e % GM_BK; ww = e / GM_BK; }
            const int gk = k0 + kk;
            const int64_t gw = w0 + ww;
            sX[buf][kk][ww] = (gk < d.K && gw < d.W) ? X[(int64_t)gk * d.xs_k + gw * d.xs_w] : 0.0;
        }
    };

    const int nk = (d.K + GM_BK - 1) / GM_BK;
    load_tiles(0, 0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles(buf ^ 1, (kt + 1) * GM_BK);
#pragma unroll
        for (int ks = 0; ks < GM_BK; ks += 4) {
            double a[4], b[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[buf][wm * 32 + i * 8 + (lane >> 2)][ks + (lane & 3)];
#pragma unroll
            for (int j = 0; j < 2; ++j) b[j] = sX[buf][ks + (lane & 3)][wn * 16 + j * 8 + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }
    // epilogue: lane holds C[row = lane/4][col = (lane%4)*2 + {0,1}] of each 8x8 tile
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + wm * 32 + i * 8 + (lane >> 2);
        if (gm >= d.M) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int64_t gw = w0 + wn * 16 + j * 8 + (lane & 3) * 2;
            if (gw < d.W) C[(int64_t)gm * d.cs_m + gw * d.cs_w] = acc[i][j][0];
            if (gw + 1 < d.W) C[(int64_t)gm * d.cs_m + (gw + 1) * d.cs_w] = acc[i][j][1];
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
    dim3 grid((unsigned)((d.W + GM_BN - 1) / GM_BN), (unsigned)((M + GM_BM - 1) / GM_BM), (unsigned)d.nbatch);
    gemm_axis_kernel<<<grid, GM_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

}  // extern "C"
