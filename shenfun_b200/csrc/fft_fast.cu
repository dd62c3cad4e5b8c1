// fft_fast.cu -- register-resident Stockham kernels for the headline sizes
// (placeholder until the first parity run; see fft_generic.cu)
#include "fft_common.cuh"

namespace shb {

int fft_fast_launch(int, const void *, void *, int64_t, int64_t, int, int, double, int, cudaStream_t) {
    return SHB_ERR_UNSUPPORTED;
}

}  // namespace shb
