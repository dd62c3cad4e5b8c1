// fft_api.cu -- C ABI entry points for the Fourier axis transforms
#include "fft_common.cuh"
#include <stdlib.h>

using namespace shb;

static int dispatch(int mode, const void *in, void *out, int64_t outer, int64_t inner,
                    int n_fft, int N, double scale, int flags, void *stream) {
    if (!in || !out) { set_error("fft: null pointer"); return SHB_ERR_ARG; }
    if (outer < 0 || inner < 0 || n_fft < 1 || N < 1 || N > n_fft) {
        set_error("fft: bad sizes outer=%lld inner=%lld n_fft=%d N=%d", (long long)outer,
                  (long long)inner, n_fft, N);
        return SHB_ERR_ARG;
    }
    if ((flags & 1) && N != n_fft) { set_error("fft: dealias_direct requires N == n_fft"); return SHB_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (in == out && !(N == n_fft && (mode == FFT_C2C_FWD || mode == FFT_C2C_BWD))) {
        set_error("fft: in-place only for unpadded c2c"); return SHB_ERR_ARG;
    }
    static const bool force_generic = getenv("SHB_FFT_FORCE_GENERIC") != nullptr;
    if (!force_generic) {
        int rc = fft_fast_launch(mode, in, out, outer, inner, n_fft, N, scale, flags, st);
        if (rc != SHB_ERR_UNSUPPORTED) return rc;
    }
    return fft_generic_launch(mode, in, out, outer, inner, n_fft, N, scale, flags, st);
}

extern "C" {

int shb_fft_c2r_cross(const void *d_in, const void *d_b, void *d_out, int64_t nlines, int n_fft, int N,
                      double scale, int flags, void *stream) {
    if (!d_in || !d_b || !d_out || d_b == d_out || nlines < 0 || n_fft < 1 || N < 1 || N > n_fft) {
        set_error("fft_c2r_cross: bad arguments");
        return SHB_ERR_ARG;
    }
    int rc = fft_c2r_cross_launch(d_in, d_b, d_out, nlines, n_fft, N, scale, flags,
                                  reinterpret_cast<cudaStream_t>(stream));
    if (rc == SHB_ERR_UNSUPPORTED) set_error("fft_c2r_cross: no fused kernel for n_fft=%d", n_fft);
    return rc;
}

int shb_fft_c2c_backward_curl(const void *d_u_hat, void *d_out, int64_t inner, int n2, int n_fft, int N,
                              const double *d_k0, const double *d_k1, const double *d_k2, double scale, void *stream) {
    if (!d_u_hat || !d_out || d_u_hat == d_out || !d_k0 || !d_k1 || !d_k2 || inner < 0 || n2 < 1 || n_fft < 1 || N < 1) {
        set_error("fft_c2c_backward_curl: bad arguments");
        return SHB_ERR_ARG;
    }
    int rc = fft_pad3_curl_launch(d_u_hat, d_out, inner, n2, n_fft, N, d_k0, d_k1, d_k2, scale,
                                  reinterpret_cast<cudaStream_t>(stream));
    if (rc == SHB_ERR_UNSUPPORTED) set_error("fft_c2c_backward_curl: no fused kernel for N=%d -> n_fft=%d", N, n_fft);
    return rc;
}

int shb_fft_c2c_peer(int backward, const void *d_in, void *const *d_peer_out, int nranks, int64_t outer,
                     int64_t inner, int n_fft, int N, double scale, int flags, int in_chunk_rows,
                     int out_chunk_rows, void *stream) {
    if (!d_in || !d_peer_out || nranks < 1 || nranks > 8 || outer < 0 || inner < 0 || n_fft < 1 || N < 1 ||
        N > n_fft || in_chunk_rows < 0 || out_chunk_rows <= 0) {
        set_error("fft_c2c_peer: bad arguments");
        return SHB_ERR_ARG;
    }
    const int n_in = backward ? N : n_fft, n_out = backward ? n_fft : N;
    if ((in_chunk_rows && n_in % in_chunk_rows) || n_out % out_chunk_rows || n_out / out_chunk_rows != nranks) {
        set_error("fft_c2c_peer: chunk rows must split the axis into one block per rank");
        return SHB_ERR_ARG;
    }
    int rc = fft_fast_launch(backward ? FFT_C2C_BWD : FFT_C2C_FWD, d_in, d_peer_out[0], outer, inner, n_fft, N, scale,
                             flags, reinterpret_cast<cudaStream_t>(stream), in_chunk_rows, out_chunk_rows,
                             d_peer_out, nranks);
    if (rc == SHB_ERR_UNSUPPORTED) set_error("fft_c2c_chunked: no fused kernel for n_fft=%d inner=%lld", n_fft, (long long)inner);
    return rc;
}

int shb_fft_c2c_lines(int backward, const void *d_in, void *d_out, void *const *d_peer_out, int nranks, int64_t na,
                      int64_t ncols, int64_t ncols_valid, int64_t c0, int64_t in_as, int64_t in_rs, int64_t out_as, int64_t out_rs,
                      int n_fft, int N, double scale, int flags, int in_chunk_rows, int64_t in_chunk_stride,
                      int out_chunk_rows, int64_t out_chunk_stride, void *stream) {
    if (!d_in || (!d_out && !d_peer_out) || na < 0 || ncols < 0 || ncols_valid < 0 || ncols_valid > ncols || c0 < 0 || n_fft < 1 || N < 1 || N > n_fft ||
        in_chunk_rows < 0 || out_chunk_rows < 0 || nranks < 0 || nranks > 8 || (d_peer_out && out_chunk_rows <= 0)) {
        set_error("fft_c2c_lines: bad arguments");
        return SHB_ERR_ARG;
    }
    const int n_in = backward ? N : n_fft, n_out = backward ? n_fft : N;
    if ((in_chunk_rows && n_in % in_chunk_rows) || (out_chunk_rows && n_out % out_chunk_rows) ||
        (d_peer_out && n_out / out_chunk_rows != nranks)) {
        set_error("fft_c2c_lines: chunk rows must split the axis into equal blocks (one per rank for peer stores)");
        return SHB_ERR_ARG;
    }
    int rc = fft_fast_launch_general(backward ? FFT_C2C_BWD : FFT_C2C_FWD, d_in, d_peer_out ? d_peer_out[0] : d_out, na,
                                     ncols, ncols_valid, c0, in_as, in_rs, out_as, out_rs, n_fft, N, scale, flags,
                                     reinterpret_cast<cudaStream_t>(stream), in_chunk_rows, in_chunk_stride,
                                     out_chunk_rows, out_chunk_stride, d_peer_out, d_peer_out ? nranks : 0);
    if (rc == SHB_ERR_UNSUPPORTED) set_error("fft_c2c_lines: no register-resident kernel for n_fft=%d", n_fft);
    return rc;
}

int shb_fft_c2c_chunked(int backward, const void *d_in, void *d_out, int64_t outer, int64_t inner, int n_fft,
                        int N, double scale, int flags, int in_chunk_rows, int out_chunk_rows, void *stream) {
    if (!d_in || !d_out || d_in == d_out || outer < 0 || inner < 0 || n_fft < 1 || N < 1 || N > n_fft ||
        in_chunk_rows < 0 || out_chunk_rows < 0) {
        set_error("fft_c2c_chunked: bad arguments");
        return SHB_ERR_ARG;
    }
    const int n_in = backward ? N : n_fft, n_out = backward ? n_fft : N;
    if ((in_chunk_rows && n_in % in_chunk_rows) || (out_chunk_rows && n_out % out_chunk_rows)) {
        set_error("fft_c2c_chunked: chunk rows must divide the axis length");
        return SHB_ERR_ARG;
    }
    int rc = fft_fast_launch(backward ? FFT_C2C_BWD : FFT_C2C_FWD, d_in, d_out, outer, inner, n_fft, N, scale, flags,
                             reinterpret_cast<cudaStream_t>(stream), in_chunk_rows, out_chunk_rows);
    if (rc == SHB_ERR_UNSUPPORTED) set_error("fft_c2c_chunked: no fused kernel for n_fft=%d inner=%lld", n_fft, (long long)inner);
    return rc;
}

int shb_fft_c2c_forward(const void *d_in, void *d_out, int64_t outer, int64_t inner,
                        int n_fft, int N, double scale, void *stream) {
    return dispatch(FFT_C2C_FWD, d_in, d_out, outer, inner, n_fft, N, scale, 0, stream);
}

int shb_fft_c2c_backward(const void *d_in, void *d_out, int64_t outer, int64_t inner,
                         int n_fft, int N, double scale, int flags, void *stream) {
    return dispatch(FFT_C2C_BWD, d_in, d_out, outer, inner, n_fft, N, scale, flags, stream);
}

int shb_fft_r2c_forward(const void *d_in, void *d_out, int64_t outer, int64_t inner,
                        int n_fft, int N, double scale, void *stream) {
    return dispatch(FFT_R2C, d_in, d_out, outer, inner, n_fft, N, scale, 0, stream);
}

int shb_fft_c2r_backward(const void *d_in, void *d_out, int64_t outer, int64_t inner,
                         int n_fft, int N, double scale, int flags, void *stream) {
    return dispatch(FFT_C2R, d_in, d_out, outer, inner, n_fft, N, scale, flags, stream);
}

}  // extern "C"
