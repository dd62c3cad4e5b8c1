// dct_api.cu -- C ABI for the Chebyshev axis transforms (DCT-II / DCT-III built on
// the complex FFT core, with scaling and the Shen stencil fused).
#include "fft_common.cuh"
#include <stdlib.h>

using namespace shb;

static int dct_dispatch(int mode, const void *in, void *out, int64_t outer, int64_t inner, int n_phys, int n_spec,
                        double scale, const DctArgs &a, cudaStream_t st) {
    static const bool force_generic = getenv("SHB_FFT_FORCE_GENERIC") != nullptr;
    if (!force_generic && in != out) {
        int rc = dct_fast_launch(mode, in, out, outer, inner, n_phys, n_spec, scale, a, st);
        if (rc != SHB_ERR_UNSUPPORTED) return rc;
    }
    return dct_generic_launch(mode, in, out, outer, inner, n_phys, n_spec, scale, a, st);
}

static int fill(DctArgs &a, int is_real, int dim, int nsten, const int *soff, const double *const *d_sten,
                int nbc, const double *d_bc_lift) {
    if (nsten < 1 || nsten > DCT_MAX_STENCIL || !soff) { set_error("dct: bad stencil"); return SHB_ERR_ARG; }
    a.is_real = is_real; a.dim = dim; a.nsten = nsten; a.nbc = nbc; a.bc_lift = d_bc_lift;
    for (int i = 0; i < DCT_MAX_STENCIL; ++i) {
        a.soff[i] = i < nsten ? soff[i] : 0;
        a.sten[i] = (i < nsten && d_sten) ? d_sten[i] : nullptr;
    }
    return SHB_OK;
}

extern "C" {

int shb_cheb_scalar_product(const void *d_in, void *d_out, int64_t outer, int64_t inner, int n_phys, int n_spec,
                            double scale, int is_real, int dim, int nsten, const int *soff,
                            const double *const *d_sten, void *stream) {
    if (!d_in || !d_out || n_phys < 1 || n_spec < 1 || dim > n_spec || dim > n_phys) {
        set_error("cheb_scalar_product: bad arguments"); return SHB_ERR_ARG;
    }
    DctArgs a;
    int rc = fill(a, is_real, dim, nsten, soff, d_sten, 0, nullptr);
    if (rc) return rc;
    return dct_dispatch(FFT_DCT2, d_in, d_out, outer, inner, n_phys, n_spec, scale, a,
                              reinterpret_cast<cudaStream_t>(stream));
}

int shb_cheb_backward(const void *d_in, void *d_out, int64_t outer, int64_t inner, int n_phys, int n_spec,
                      int is_real, int dim, int nsten, const int *soff, const double *const *d_sten, int nbc,
                      const double *d_bc_lift, void *stream) {
    if (!d_in || !d_out || n_phys < 1 || n_spec < 1 || dim > n_spec || dim > n_phys || (nbc > 0 && !d_bc_lift)) {
        set_error("cheb_backward: bad arguments"); return SHB_ERR_ARG;
    }
    DctArgs a;
    int rc = fill(a, is_real, dim, nsten, soff, d_sten, nbc, d_bc_lift);
    if (rc) return rc;
    return dct_dispatch(FFT_DCT3, d_in, d_out, outer, inner, n_phys, n_spec, 1.0, a,
                              reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
