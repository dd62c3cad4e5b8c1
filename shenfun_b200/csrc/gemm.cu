// gemm.cu -- fp64 tensor-core contraction along one axis:
//     out[o][i][c] = sum_j A[i][j] * u[o][j][c]        A real (M x K), u real or complex
// used for the Legendre / Jacobi Vandermonde transforms (operator A = stencil x
// Vandermonde x quadrature weights, built once on the host).
//
// The data array is viewed as real with W = inner * (complex ? 2 : 1) contiguous
// columns, so one kernel serves real and complex fields (A is real).  Each CTA owns a
// 128 x 64 output tile of one outer index and walks K in chunks of 32 through a two-stage
// cp.async shared-memory pipeline; 8 warps each hold a 32 x 32 sub-tile as 4 x 4 m8n8k4 DMMA
// accumulators (mma.sync f64 -- tcgen05 has no fp64 kind).
//
// Reference semantics restated: shenfun/spectralbase.py:473-483 (evaluate_expansion_all,
// 'vandermonde': np.dot(P, u_hat) along the axis) and :539-548 (scalar product).
#include "common.cuh"

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

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc, bool pred) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}

// CTA tile 128 x 64, K chunks of 32 through a two-stage cp.async pipeline; 8 warps (4 x 2), each
// holding a 32 x 32 sub-tile as 4 x 4 m8n8k4 DMMA accumulators (16 DMMA per 8 fragment loads).
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
    auto load_tiles = [&](int buf, int k0) {
        double *a = sA + (size_t)buf * GM_BM * GM_AS;
        double *x = sX + (size_t)buf * GM_BK * GM_XS;
        // A tile: 128 x 32 doubles, 16 per thread, k fastest (A is row-major)
#pragma unroll
        for (int r = 0; r < (GM_BM * GM_BK) / GM_THREADS; ++r) {
            const int e = tid + r * GM_THREADS;
            const int kk = e % GM_BK, mm = e / GM_BK;
            const int gm = m0 + mm, gk = k0 + kk;
            const bool ok = (gm < d.M && gk < d.K);
            cp_async8(&a[mm * GM_AS + kk], ok ? &d.A[(int64_t)gm * d.K + gk] : d.A, ok);
        }
        // X tile: 32 x 64 doubles, 8 per thread; fastest index follows the unit stride
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
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int nk = (d.K + GM_BK - 1) / GM_BK;
    load_tiles(0, 0);
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) {
            load_tiles(buf ^ 1, (kt + 1) * GM_BK);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double *a = sA + (size_t)buf * GM_BM * GM_AS;
        const double *x = sX + (size_t)buf * GM_BK * GM_XS;
#pragma unroll
        for (int ks = 0; ks < GM_BK; ks += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = a[(wm * 32 + i * 8 + (lane >> 2)) * GM_AS + ks + (lane & 3)];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = x[(ks + (lane & 3)) * GM_XS + wn * 32 + j * 8 + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
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
    static bool attr_set = false;
    if (!attr_set) {
        SHB_CUDA(cudaFuncSetAttribute(gemm_axis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GM_SMEM));
        attr_set = true;
    }
    dim3 grid((unsigned)((d.W + GM_BN - 1) / GM_BN), (unsigned)((M + GM_BM - 1) / GM_BM), (unsigned)d.nbatch);
    gemm_axis_kernel<<<grid, GM_THREADS, GM_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(d);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

}  // extern "C"
