// fft_common.cuh -- descriptor shared by the generic and fast FFT kernels
#pragma once
#include "common.cuh"

namespace shb {

enum FftMode {
    FFT_C2C_FWD = 0,
    FFT_C2C_BWD = 1,
    FFT_R2C = 2,        // even n_fft: half-length complex FFT + split
    FFT_C2R = 3,
    FFT_R2C_FULL = 4,   // odd n_fft: full complex FFT of the real line
    FFT_C2R_FULL = 5,
    FFT_DCT2 = 6,       // Chebyshev scalar product: DCT-II * scale, then stencil K
    FFT_DCT3 = 7        // Chebyshev evaluation: stencil K^T, then 0.5*DCT-III + c0/2
};

constexpr int DCT_MAX_STENCIL = 6;

constexpr int FFT_MAX_STAGES = 24;

struct FftDesc {
    const void *in;
    void *out;
    int64_t outer, inner;
    int n_fft;      // physical (possibly padded) transform length
    int N;          // un-padded basis size (spectral length N or N/2+1)
    int M;          // complex FFT length run by the core
    int n_in;       // rows along the axis in the input array
    int n_out;      // rows along the axis in the output array
    int T;          // lines per CTA
    int PT;         // shared-memory pitch (complex elements) of one row
    int nstages;
    int radix[FFT_MAX_STAGES];
    int flags;      // bit0: dealias_direct
    double scale;
    const cplx *tw;   // exp(-2 pi i k / M)
    const cplx *tw2;  // exp(-2 pi i k / (2M)), k = 0..M
    // --- DCT / Shen-stencil modes ---
    const cplx *tw3;  // exp(-i pi k / (2M)), k = 0..M
    int is_real;      // data arrays are real (double) instead of complex
    int dim;          // degrees of freedom along the axis (rows that carry the stencil)
    int nsten;        // stencil diagonals: phi_k = sum_i sten[i][k] T_{k+soff[i]}
    int soff[DCT_MAX_STENCIL];
    const double *sten[DCT_MAX_STENCIL];   // NULL = all ones
    int nbc;                               // boundary dofs stored in rows dim..dim+nbc-1
    const double *bc_lift;                 // [nbc][4]: ortho coefficients T_0..T_3 of each boundary function
};

struct DctArgs {
    int is_real, dim, nsten;
    int soff[DCT_MAX_STENCIL];
    const double *sten[DCT_MAX_STENCIL];
    int nbc;
    const double *bc_lift;
};

// out = irfft(in) x b on [3][L][n] arrays (C2R of the three components fused with the cross product)
int fft_c2r_cross_launch(const void *in, const void *b, void *out, int64_t L, int n_fft, int N, double scale,
                         int flags, cudaStream_t st);

// register-resident DCT kernels (strided axes, headline sizes); SHB_ERR_UNSUPPORTED otherwise
int dct_fast_launch(int mode, const void *in, void *out, int64_t outer, int64_t inner, int n_phys, int n_spec,
                    double scale, const DctArgs &a, cudaStream_t st);

int dct_generic_launch(int mode, const void *in, void *out, int64_t outer, int64_t inner, int n_phys,
                       int n_spec, double scale, const DctArgs &a, cudaStream_t st);

int fft_generic_launch(int mode, const void *in, void *out, int64_t outer, int64_t inner,
                       int n_fft, int N, double scale, int flags, cudaStream_t st);

// returns SHB_ERR_UNSUPPORTED when no specialised kernel covers the case
int fft_pad3_curl_launch(const void *u_hat, void *out, int64_t inner, int n2, int n_fft, int N, const double *k0,
                         const double *k1, const double *k2, double scale, cudaStream_t st);
int fft_fast_launch(int mode, const void *in, void *out, int64_t outer, int64_t inner,
                    int n_fft, int N, double scale, int flags, cudaStream_t st, int in_cr = 0, int out_cr = 0,
                    void *const *peer_out = nullptr, int npeer = 0);
int fft_fast_launch_general(int mode, const void *in, void *out, int64_t na, int64_t ncols, int64_t ncols_valid, int64_t c0,
                            int64_t in_as,
                            int64_t in_rs, int64_t out_as, int64_t out_rs, int n_fft, int N, double scale, int flags,
                            cudaStream_t st, int in_cr, int64_t in_cs, int out_cr, int64_t out_cs,
                            void *const *peer_out, int npeer);

}  // namespace shb
