// common.cuh -- shared helpers for libshenb200 (sm_100a only)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/shenb200.h"

namespace shb {

typedef double2 cplx;

// ---- error plumbing (no exceptions cross the C ABI) -------------------------
void set_error(const char *fmt, ...);
int  cuda_fail(cudaError_t e, const char *what);

#define SHB_CUDA(call)                                         \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return shb::cuda_fail(e__, #call); \
    } while (0)

#define SHB_REQUIRE(cond, msg)                                  \
    do {                                                        \
        if (!(cond)) { shb::set_error("%s: %s", __func__, msg); return SHB_ERR_ARG; } \
    } while (0)

// ---- device-resident tables cached per (device, kind, n) ---------------------
// kind 0: W_n^k = exp(-2 pi i k / n), k = 0..n-1       (n entries)
// kind 1: exp(-2 pi i k / (2 n)),      k = 0..n         (n+1 entries; real-FFT split)
// kind 2: exp(-i pi k / (2 n)),        k = 0..n         (n+1 entries; DCT quarter-wave)
const cplx *twiddle_table(int kind, int n);

int sm_count();

// ---- complex arithmetic -------------------------------------------------------
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
// multiply by -i (forward) or +i (inverse)
template <bool INV> __device__ __forceinline__ cplx mul_mi(cplx a) {
    return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

}  // namespace shb
