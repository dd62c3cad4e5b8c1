// plan_api.cu -- N-D Fourier transform plans behind the C ABI (single device).
//
// The plan is what TensorProductSpace.__init__/configure_backwards builds in the reference
// (shenfun/tensorproductspace.py:314-541): the order of the axes, the per-axis transform kind and
// the padded (3/2-rule) physical sizes.  Executing it is ForwardTransform.__call__ / BackwardTransform.
// __call__ (:205-259, :46-96): one pass per axis, last entry of `axes` first for the forward
// direction, each pass a single launch of the axis kernels with pad / truncate / normalise fused.
// The plan owns nothing on the device except the library's cached twiddle tables; the caller passes
// a workspace of shb_plan_workspace_bytes() bytes that the passes ping-pong through.
#include "common.cuh"
#include "../../include/shenb200.h"

#include <new>

using namespace shb;

struct shb_plan {
    shb_plan_desc d;
    int64_t max_elems;      // largest intermediate array, in complex elements per component
};

static int64_t axis_len(const shb_plan_desc &d, int ax, bool spectral) {
    if (!spectral) return d.padded[ax];
    return d.kind[ax] == SHB_AXIS_R2C ? d.shape[ax] / 2 + 1 : d.shape[ax];
}

extern "C" {

int shb_plan_create(const shb_plan_desc *desc, shb_plan **plan) {
    SHB_REQUIRE(desc && plan, "null pointer");
    const shb_plan_desc &d = *desc;
    SHB_REQUIRE(d.ndim >= 1 && d.ndim <= SHB_PLAN_MAXDIM, "ndim must be 1..4");
    bool seen[SHB_PLAN_MAXDIM] = {false, false, false, false};
    int nreal = 0;
    for (int i = 0; i < d.ndim; ++i) {
        const int ax = d.axes[i];
        SHB_REQUIRE(ax >= 0 && ax < d.ndim && !seen[ax], "axes must be a permutation of 0..ndim-1");
        seen[ax] = true;
        SHB_REQUIRE(d.shape[i] >= 1 && d.padded[i] >= d.shape[i] && d.padded[i] < (1ll << 31), "bad sizes");
        SHB_REQUIRE(d.kind[i] == SHB_AXIS_C2C || d.kind[i] == SHB_AXIS_R2C, "axis kind must be C2C or R2C");
        SHB_REQUIRE(!(d.dealias_direct && d.padded[i] != d.shape[i]), "dealias_direct requires no padding");
        nreal += d.kind[i] == SHB_AXIS_R2C;
    }
    // tensorproductspace.py:365-366, 382-385: at most one real-to-complex axis and it is transformed first
    SHB_REQUIRE(nreal <= 1, "at most one R2C axis");
    if (nreal == 1) SHB_REQUIRE(d.kind[d.axes[d.ndim - 1]] == SHB_AXIS_R2C, "the R2C axis must be last in `axes`");
    shb_plan *p = new (std::nothrow) shb_plan;
    if (!p) { set_error("plan: out of host memory"); return SHB_ERR_NOMEM; }
    p->d = d;
    // intermediates of the forward sweep: after k passes, the k axes done hold their spectral length
    int64_t mx = 0;
    bool done[SHB_PLAN_MAXDIM] = {false, false, false, false};
    for (int k = d.ndim - 1; k >= 0; --k) {
        done[d.axes[k]] = true;
        int64_t n = 1;
        for (int ax = 0; ax < d.ndim; ++ax) n *= axis_len(d, ax, done[ax]);
        if (n > mx) mx = n;
    }
    p->max_elems = mx;
    *plan = p;
    return SHB_OK;
}

int shb_plan_destroy(shb_plan *plan) {
    delete plan;
    return SHB_OK;
}

int64_t shb_plan_workspace_bytes(const shb_plan *plan, int ncomp) {
    if (!plan || ncomp < 1) return 0;
    // two ping-pong buffers of the largest intermediate (complex128); none needed for one axis
    return plan->d.ndim == 1 ? 0 : 2 * plan->max_elems * ncomp * (int64_t)sizeof(cplx);
}

int64_t shb_plan_array_elems(const shb_plan *plan, int spectral) {
    if (!plan) return 0;
    int64_t n = 1;
    for (int ax = 0; ax < plan->d.ndim; ++ax) n *= axis_len(plan->d, ax, spectral != 0);
    return n;
}

static int run(const shb_plan *plan, bool forward, const void *d_in, void *d_out, void *d_work, int ncomp,
               void *stream) {
    SHB_REQUIRE(plan && d_in && d_out && ncomp >= 1, "bad arguments");
    const shb_plan_desc &d = plan->d;
    SHB_REQUIRE(d.ndim == 1 || d_work, "workspace required (shb_plan_workspace_bytes)");
    bool spec[SHB_PLAN_MAXDIM];
    for (int ax = 0; ax < SHB_PLAN_MAXDIM; ++ax) spec[ax] = !forward;     // state of every axis before the sweep
    cplx *buf[2] = {reinterpret_cast<cplx *>(d_work),
                    reinterpret_cast<cplx *>(d_work) + plan->max_elems * ncomp};
    const void *cur = d_in;
    int which = 0;
    for (int step = 0; step < d.ndim; ++step) {
        const int ax = forward ? d.axes[d.ndim - 1 - step] : d.axes[step];
        int64_t outer = ncomp, inner = 1;
        for (int a = 0; a < ax; ++a) outer *= axis_len(d, a, spec[a]);
        for (int a = ax + 1; a < d.ndim; ++a) inner *= axis_len(d, a, spec[a]);
        void *dst = (step == d.ndim - 1) ? d_out : static_cast<void *>(buf[which]);
        which ^= 1;
        const int n_fft = (int)d.padded[ax], N = (int)d.shape[ax];
        int rc;
        if (forward) {
            const double scale = 1.0 / (double)n_fft;       // fourier/bases.py:138-142 (normalised by the padded length)
            rc = d.kind[ax] == SHB_AXIS_R2C ? shb_fft_r2c_forward(cur, dst, outer, inner, n_fft, N, scale, stream)
                                            : shb_fft_c2c_forward(cur, dst, outer, inner, n_fft, N, scale, stream);
        } else {
            const int flags = d.dealias_direct ? 1 : 0;
            rc = d.kind[ax] == SHB_AXIS_R2C ? shb_fft_c2r_backward(cur, dst, outer, inner, n_fft, N, 1.0, flags, stream)
                                            : shb_fft_c2c_backward(cur, dst, outer, inner, n_fft, N, 1.0, flags, stream);
        }
        if (rc != SHB_OK) return rc;
        spec[ax] = forward;
        cur = dst;
    }
    return SHB_OK;
}

int shb_plan_forward(const shb_plan *plan, const void *d_in, void *d_out, void *d_work, int ncomp, void *stream) {
    return run(plan, true, d_in, d_out, d_work, ncomp, stream);
}

int shb_plan_backward(const shb_plan *plan, const void *d_in, void *d_out, void *d_work, int ncomp, void *stream) {
    return run(plan, false, d_in, d_out, d_work, ncomp, stream);
}

}  // extern "C"
