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
    FFT_C2R_FULL = 5
};

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
};

int fft_generic_launch(int mode, const void *in, void *out, int64_t outer, int64_t inner,
                       int n_fft, int N, double scale, int flags, cudaStream_t st);

// returns SHB_ERR_UNSUPPORTED when no specialised kernel covers the case
int fft_fast_launch(int mode, const void *in, void *out, int64_t outer, int64_t inner,
                    int n_fft, int N, double scale, int flags, cudaStream_t st);

}  // namespace shb
