// C wrapper (ours) around the reference's own fastgl.cpp so that the oracle can call
// the reference's Gauss-Legendre node generator through ctypes.  Built by
// oracle/build_ref.sh from the sources where they lie under /root/reference; the
// resulting shared object goes to oracle/_ref/ (git-ignored, travels with gpurun).
#include "fastgl.h"

extern "C" void fastgl_pairs(int n, double *x, double *w) {
    // same ordering as shenfun/legendre/fastgl/fastgl_wrap.pyx:22-28 (ascending nodes)
    for (int i = 0; i < n; ++i) {
        fastgl::QuadPair p = fastgl::GLPair((size_t)n, (size_t)(n - i));
        x[i] = p.x();
        w[i] = p.weight;
    }
}
