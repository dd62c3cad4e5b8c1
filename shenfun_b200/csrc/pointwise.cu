// pointwise.cu -- memory-bound element-wise kernels of the Navier-Stokes path:
// cross product in physical space, spectral curl, pressure projection fused
// with the Runge-Kutta stage update, Nyquist masking and generic RK4 axpy.
//
// Reference semantics restated:
//   shenfun/optimization/cython/cross.pyx:25-40 (cross3D),
//   shenfun/optimization/cython/applymask.pyx:15-169 (apply_mask),
//   demo/NavierStokes.py:35-44 (NonlinearRHS: curl, cross, div-free projection),
//   shenfun/utilities/integrators.py:587-599 (RK4 stage algebra),
//   shenfun/fourier/matrices.py:149-176 (diagonal Fourier matrices (ik)^n).
#include "common.cuh"

namespace shb {

constexpr int PW_THREADS = 256;

static inline int pw_blocks(int64_t n, int per_thread = 1) {
    int64_t b = (n + (int64_t)PW_THREADS * per_thread - 1) / ((int64_t)PW_THREADS * per_thread);
    int64_t cap = (int64_t)sm_count() * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// c = a x b, real, two points per thread (double2 loads)
__global__ void __launch_bounds__(PW_THREADS) cross3d_real_kernel(
    const double2 *__restrict__ a, const double2 *__restrict__ b, double2 *__restrict__ c, int64_t n2) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 a0 = a[i], a1 = a[n2 + i], a2 = a[2 * n2 + i];
        const double2 b0 = b[i], b1 = b[n2 + i], b2 = b[2 * n2 + i];
        double2 c0, c1, c2;
        c0.x = a1.x * b2.x - a2.x * b1.x;  c0.y = a1.y * b2.y - a2.y * b1.y;
        c1.x = a2.x * b0.x - a0.x * b2.x;  c1.y = a2.y * b0.y - a0.y * b2.y;
        c2.x = a0.x * b1.x - a1.x * b0.x;  c2.y = a0.y * b1.y - a1.y * b0.y;
        c[i] = c0; c[n2 + i] = c1; c[2 * n2 + i] = c2;
    }
}

__global__ void __launch_bounds__(PW_THREADS) cross3d_real_tail_kernel(
    const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ c, int64_t n, int64_t start) {
    for (int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double a0 = a[i], a1 = a[n + i], a2 = a[2 * n + i];
        const double b0 = b[i], b1 = b[n + i], b2 = b[2 * n + i];
        c[i] = a1 * b2 - a2 * b1;
        c[n + i] = a2 * b0 - a0 * b2;
        c[2 * n + i] = a0 * b1 - a1 * b0;
    }
}

__global__ void __launch_bounds__(PW_THREADS) cross3d_cplx_kernel(
    const cplx *__restrict__ a, const cplx *__restrict__ b, cplx *__restrict__ c, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const cplx a0 = a[i], a1 = a[n + i], a2 = a[2 * n + i];
        const cplx b0 = b[i], b1 = b[n + i], b2 = b[2 * n + i];
        c[i] = csub(cmul(a1, b2), cmul(a2, b1));
        c[n + i] = csub(cmul(a2, b0), cmul(a0, b2));
        c[2 * n + i] = csub(cmul(a0, b1), cmul(a1, b0));
    }
}

// out = i k x u  on [3][n0][n1][n2] complex
__global__ void __launch_bounds__(PW_THREADS) curl_hat_kernel(
    const cplx *__restrict__ u, cplx *__restrict__ out, const double *__restrict__ k0,
    const double *__restrict__ k1, const double *__restrict__ k2, int n0, int n1, int n2) {
    const int64_t n = (int64_t)n0 * n1 * n2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int i2 = (int)(i % n2);
        const int64_t r = i / n2;
        const int i1 = (int)(r % n1);
        const int i0 = (int)(r / n1);
        const double kx = __ldg(&k0[i0]), ky = __ldg(&k1[i1]), kz = __ldg(&k2[i2]);
        const cplx ux = u[i], uy = u[n + i], uz = u[2 * n + i];
        // (k x u) then times i:  i*(a+ib) = (-b, a)
        const cplx cx = make_double2(ky * uz.x - kz * uy.x, ky * uz.y - kz * uy.y);
        const cplx cy = make_double2(kz * ux.x - kx * uz.x, kz * ux.y - kx * uz.y);
        const cplx cz = make_double2(kx * uy.x - ky * ux.x, kx * uy.y - ky * ux.y);
        out[i] = make_double2(-cx.y, cx.x);
        out[n + i] = make_double2(-cy.y, cy.x);
        out[2 * n + i] = make_double2(-cz.y, cz.x);
    }
}

// Fused: pressure projection + viscous term + RK4 stage update.
//   dU   <- W_hat - k (k.W_hat)/k^2 - nu k^2 u_hat        (demo/NavierStokes.py:41-43, integrators.py:592-593)
//   P    <- i (k.W_hat)/k^2                                (optional)
//   U1   <- U1 + a_dt*dU ;  u_hat <- U0 + b_dt*dU (stages 0..2)  or  u_hat <- U1 (last stage)
__global__ void __launch_bounds__(PW_THREADS) ns_rk_stage_kernel(
    cplx *__restrict__ dU, cplx *__restrict__ u_hat, const cplx *__restrict__ U0, cplx *__restrict__ U1,
    cplx *__restrict__ P, cplx *__restrict__ curl_out, const double *__restrict__ k0,
    const double *__restrict__ k1, const double *__restrict__ k2, int n0, int n1, int n2, double nu,
    double a_dt, double b_dt, int last) {
    const int64_t n = (int64_t)n0 * n1 * n2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int i2 = (int)(i % n2);
        const int64_t r = i / n2;
        const int i1 = (int)(r % n1);
        const int i0 = (int)(r / n1);
        const double kx = __ldg(&k0[i0]), ky = __ldg(&k1[i1]), kz = __ldg(&k2[i2]);
        const double kk = kx * kx + ky * ky + kz * kz;
        const double ikk = kk > 0.0 ? 1.0 / kk : 0.0;
        cplx w[3] = {dU[i], dU[n + i], dU[2 * n + i]};
        const cplx kd = make_double2(kx * w[0].x + ky * w[1].x + kz * w[2].x,
                                     kx * w[0].y + ky * w[1].y + kz * w[2].y);
        const cplx p = make_double2(-kd.y * ikk, kd.x * ikk);       // i*kd/k^2
        if (P) P[i] = p;
        const double kv[3] = {kx, ky, kz};
        const double lin = -nu * kk;
        cplx un[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const cplx uh = u_hat[c * n + i];
            // dU_c + i k_c P  =  dU_c + (-k_c p.y, k_c p.x)
            cplx d = make_double2(w[c].x - kv[c] * p.y, w[c].y + kv[c] * p.x);
            d.x = fma(lin, uh.x, d.x);
            d.y = fma(lin, uh.y, d.y);
            if (!(last & 2)) dU[c * n + i] = d;      // bit 1 of `last`: the caller does not need dU back
            cplx u1 = U1[c * n + i];
            u1.x = fma(a_dt, d.x, u1.x);
            u1.y = fma(a_dt, d.y, u1.y);
            U1[c * n + i] = u1;
            if (last & 1) {
                un[c] = u1;
            } else {
                const cplx u0 = U0[c * n + i];
                un[c] = make_double2(fma(b_dt, d.x, u0.x), fma(b_dt, d.y, u0.y));
            }
            u_hat[c * n + i] = un[c];
        }
        if (curl_out) {
            // curl of the updated velocity for the next stage: i k x u
            const cplx cx = make_double2(ky * un[2].x - kz * un[1].x, ky * un[2].y - kz * un[1].y);
            const cplx cy = make_double2(kz * un[0].x - kx * un[2].x, kz * un[0].y - kx * un[2].y);
            const cplx cz = make_double2(kx * un[1].x - ky * un[0].x, kx * un[1].y - ky * un[0].y);
            curl_out[i] = make_double2(-cx.y, cx.x);
            curl_out[n + i] = make_double2(-cy.y, cy.x);
            curl_out[2 * n + i] = make_double2(-cz.y, cz.x);
        }
    }
}

// generic RK4 stage: dU += L*u_hat (L real, broadcast over ncomp); U1 += a_dt dU; u_hat = U0 + b_dt dU | U1
__global__ void __launch_bounds__(PW_THREADS) rk4_stage_kernel(
    cplx *__restrict__ dU, cplx *__restrict__ u_hat, const cplx *__restrict__ U0, cplx *__restrict__ U1,
    const double *__restrict__ L, const cplx *__restrict__ Lc, int64_t n, int ncomp, double a_dt, double b_dt,
    int last) {
    const int64_t tot = n * ncomp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
        cplx d = dU[i];
        const cplx uh = u_hat[i];
        if (L) {
            const double l = L[i % n];
            d.x = fma(l, uh.x, d.x);
            d.y = fma(l, uh.y, d.y);
            dU[i] = d;
        } else if (Lc) {
            // complex diagonal operator (odd derivatives: KdV's -u_xxx, Ginzburg-Landau)
            const cplx l = Lc[i % n];
            d.x = fma(l.x, uh.x, fma(-l.y, uh.y, d.x));
            d.y = fma(l.x, uh.y, fma(l.y, uh.x, d.y));
            dU[i] = d;
        }
        cplx u1 = U1[i];
        u1.x = fma(a_dt, d.x, u1.x);
        u1.y = fma(a_dt, d.y, u1.y);
        U1[i] = u1;
        if (last) u_hat[i] = u1;
        else {
            const cplx u0 = U0[i];
            u_hat[i] = make_double2(fma(b_dt, d.x, u0.x), fma(b_dt, d.y, u0.y));
        }
    }
}

// u *= mask (mask real, same shape as one component)
__global__ void __launch_bounds__(PW_THREADS) apply_mask_kernel(cplx *__restrict__ u, const double *__restrict__ mask,
                                                                int64_t n, int ncomp) {
    const int64_t tot = n * ncomp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
        const double m = mask[i % n];
        cplx x = u[i];
        u[i] = make_double2(x.x * m, x.y * m);
    }
}

// out = mask * sum_j coef_j T_j  (+ up to two plain copies save_dst = T_idx riding along): the
// right-hand-side accumulation of one implicit-explicit Runge-Kutta stage
// (shenfun/utilities/integrators.py:692-696, 803-812: rhs = (u0, v) + dt sum_j b_j N_j + dt sum_j a_j L_j,
// then the Nyquist mask), one pass over the arrays instead of a copy and an axpy per term.
// Arrays are addressed as doubles (a complex array is 2n doubles; the mask is per complex entry).
constexpr int LC_MAXT = 10;
struct LinComb {
    const double *t[LC_MAXT];
    double c[LC_MAXT];
    int nterms;
    const double *mask;       // [n_mask] or NULL
    int64_t n_mask;
    int cplx_shift;           // 1: two doubles per mask entry
    int times_i;              // complex data: multiply the result by i (Fourier derivatives of odd order)
    double *save_dst[2];
    int save_idx[2];
    int nsave;
};
__global__ void __launch_bounds__(PW_THREADS) lincomb_kernel(double *__restrict__ out, const LinComb p, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        // times i: (x + i y) i = -y + i x, i.e. double 2e takes -(double 2e+1), double 2e+1 takes double 2e
        const int64_t is = p.times_i ? (i ^ 1) : i;
#pragma unroll
        for (int j = 0; j < LC_MAXT; ++j) {
            if (j < p.nterms) {
                const double v = p.t[j][is];
                acc = fma(p.c[j], v, acc);
                if (p.nsave > 0 && p.save_idx[0] == j) p.save_dst[0][i] = v;
                if (p.nsave > 1 && p.save_idx[1] == j) p.save_dst[1][i] = v;
            }
        }
        if (p.mask) acc *= p.mask[(i >> p.cplx_shift) % p.n_mask];
        if (p.times_i && !(i & 1)) acc = -acc;
        out[i] = acc;
    }
}

// H_i = sum_j u_j g[3 i + j]: the convective term u . grad u in physical space from the nine
// gradient components (demo/ChannelFlow.py:213-215: three products and two sums per component
// with numpy temporaries; here twelve arrays are read once and three written once).
struct Grad9 { const double *p[9]; };
__global__ void __launch_bounds__(PW_THREADS) dot_grad3d_kernel(const double *__restrict__ u, const Grad9 g,
                                                                double *__restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double u0 = u[i], u1 = u[n + i], u2 = u[2 * n + i];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double h = u0 * __ldg(&g.p[3 * c][i]);
            h = fma(u1, __ldg(&g.p[3 * c + 1][i]), h);
            h = fma(u2, __ldg(&g.p[3 * c + 2][i]), h);
            out[c * n + i] = h;
        }
    }
}

}  // namespace shb

using namespace shb;

extern "C" {

int shb_cross3d(const void *d_a, const void *d_b, void *d_c, int64_t n, int is_complex, void *stream) {
    SHB_REQUIRE(d_a && d_b && d_c && n >= 0, "bad arguments");
    if (n == 0) return SHB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (is_complex) {
        cross3d_cplx_kernel<<<pw_blocks(n), PW_THREADS, 0, st>>>(
            (const cplx *)d_a, (const cplx *)d_b, (cplx *)d_c, n);
    } else if (n % 2 == 0) {
        cross3d_real_kernel<<<pw_blocks(n / 2), PW_THREADS, 0, st>>>(
            (const double2 *)d_a, (const double2 *)d_b, (double2 *)d_c, n / 2);
    } else {
        cross3d_real_tail_kernel<<<pw_blocks(n), PW_THREADS, 0, st>>>(
            (const double *)d_a, (const double *)d_b, (double *)d_c, n, 0);
    }
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

int shb_dot_grad3d(const double *d_u, const double *const *d_grad, double *d_out, int64_t n, void *stream) {
    SHB_REQUIRE(d_u && d_grad && d_out && n >= 0, "bad arguments");
    Grad9 g;
    for (int k = 0; k < 9; ++k) {
        SHB_REQUIRE(d_grad[k] != nullptr, "null gradient component");
        g.p[k] = d_grad[k];
    }
    if (n == 0) return SHB_OK;
    dot_grad3d_kernel<<<pw_blocks(n), PW_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_u, g, d_out, n);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

int shb_curl_hat(const void *d_u, void *d_out, const double *d_k0, const double *d_k1, const double *d_k2,
                 int n0, int n1, int n2, void *stream) {
    SHB_REQUIRE(d_u && d_out && d_k0 && d_k1 && d_k2, "null pointer");
    const int64_t n = (int64_t)n0 * n1 * n2;
    if (n == 0) return SHB_OK;
    curl_hat_kernel<<<pw_blocks(n), PW_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        (const cplx *)d_u, (cplx *)d_out, d_k0, d_k1, d_k2, n0, n1, n2);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

int shb_ns_rk_stage(void *d_dU, void *d_u_hat, const void *d_U0, void *d_U1, void *d_P, void *d_curl,
                    const double *d_k0, const double *d_k1, const double *d_k2, int n0, int n1, int n2,
                    double nu, double a_dt, double b_dt, int last, void *stream) {
    SHB_REQUIRE(d_dU && d_u_hat && d_U0 && d_U1 && d_k0 && d_k1 && d_k2, "null pointer");
    const int64_t n = (int64_t)n0 * n1 * n2;
    if (n == 0) return SHB_OK;
    ns_rk_stage_kernel<<<pw_blocks(n), PW_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        (cplx *)d_dU, (cplx *)d_u_hat, (const cplx *)d_U0, (cplx *)d_U1, (cplx *)d_P, (cplx *)d_curl, d_k0, d_k1, d_k2,
        n0, n1, n2, nu, a_dt, b_dt, last);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

int shb_rk4_stage(void *d_dU, void *d_u_hat, const void *d_U0, void *d_U1, const double *d_L,
                  int64_t n, int ncomp, double a_dt, double b_dt, int last, void *stream) {
    SHB_REQUIRE(d_dU && d_u_hat && d_U0 && d_U1 && n >= 0 && ncomp >= 1, "bad arguments");
    if (n == 0) return SHB_OK;
    rk4_stage_kernel<<<pw_blocks(n * ncomp), PW_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        (cplx *)d_dU, (cplx *)d_u_hat, (const cplx *)d_U0, (cplx *)d_U1, d_L, nullptr, n, ncomp, a_dt, b_dt, last);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

int shb_rk4_stage_cplx(void *d_dU, void *d_u_hat, const void *d_U0, void *d_U1, const void *d_L,
                       int64_t n, int ncomp, double a_dt, double b_dt, int last, void *stream) {
    SHB_REQUIRE(d_dU && d_u_hat && d_U0 && d_U1 && d_L && n >= 0 && ncomp >= 1, "bad arguments");
    if (n == 0) return SHB_OK;
    rk4_stage_kernel<<<pw_blocks(n * ncomp), PW_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        (cplx *)d_dU, (cplx *)d_u_hat, (const cplx *)d_U0, (cplx *)d_U1, nullptr, (const cplx *)d_L, n, ncomp,
        a_dt, b_dt, last);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

int shb_lincomb(void *d_out, int nterms, const void *const *d_terms, const double *coefs, const double *d_mask,
                int64_t n_mask, int64_t n, int is_complex, int nsave, const int *save_idx, void *const *d_save_dst,
                void *stream) {
    SHB_REQUIRE(d_out && d_terms && coefs && nterms >= 1 && nterms <= LC_MAXT && n >= 0 && nsave >= 0 && nsave <= 2 &&
                (nsave == 0 || (save_idx && d_save_dst)) && (!d_mask || n_mask > 0), "bad arguments");
    if (n == 0) return SHB_OK;
    LinComb p;
    p.nterms = nterms;
    for (int j = 0; j < LC_MAXT; ++j) {
        p.t[j] = j < nterms ? reinterpret_cast<const double *>(d_terms[j]) : nullptr;
        p.c[j] = j < nterms ? coefs[j] : 0.0;
        SHB_REQUIRE(j >= nterms || p.t[j], "null term");
    }
    p.mask = d_mask; p.n_mask = d_mask ? n_mask : 1; p.cplx_shift = is_complex ? 1 : 0;
    p.times_i = (is_complex & 2) ? 1 : 0;
    SHB_REQUIRE(!p.times_i || ((is_complex & 1) && nsave == 0), "times-i needs complex data and no save slots");
    p.nsave = nsave;
    for (int k = 0; k < 2; ++k) {
        p.save_idx[k] = k < nsave ? save_idx[k] : -1;
        p.save_dst[k] = k < nsave ? reinterpret_cast<double *>(d_save_dst[k]) : nullptr;
        SHB_REQUIRE(k >= nsave || (p.save_dst[k] && p.save_idx[k] >= 0 && p.save_idx[k] < nterms), "bad save slot");
    }
    const int64_t nd = is_complex ? 2 * n : n;
    lincomb_kernel<<<pw_blocks(nd), PW_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<double *>(d_out), p, nd);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

// Velocity components from f = du/dx and the wall-normal vorticity g under the divergence constraint
// (demo/ChannelFlow.py:235-239):  v = i (a f + b g),  w = i (b f - a g),  a = ky/(ky^2+kz^2), b = kz/(...)
// broadcast over the wall-normal index.  One pass instead of eight element-wise launches.
__global__ void __launch_bounds__(PW_THREADS) kmm_vw_kernel(const cplx *__restrict__ f, const cplx *__restrict__ g,
                                                            const double *__restrict__ a, const double *__restrict__ b,
                                                            cplx *__restrict__ v, cplx *__restrict__ w, int64_t n,
                                                            int64_t n12) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i % n12;
        const double A = __ldg(&a[k]), B = __ldg(&b[k]);
        const cplx F = f[i], G = g[i];
        const cplx t1 = make_double2(fma(A, F.x, B * G.x), fma(A, F.y, B * G.y));
        const cplx t2 = make_double2(fma(B, F.x, -A * G.x), fma(B, F.y, -A * G.y));
        v[i] = make_double2(-t1.y, t1.x);
        w[i] = make_double2(-t2.y, t2.x);
    }
}

int shb_kmm_vw(const void *d_f, const void *d_g, const double *d_a, const double *d_b, void *d_v, void *d_w, int64_t n,
               int64_t n12, void *stream) {
    SHB_REQUIRE(d_f && d_g && d_a && d_b && d_v && d_w && n >= 0 && n12 >= 1, "bad arguments");
    if (n == 0) return SHB_OK;
    kmm_vw_kernel<<<pw_blocks(n), PW_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        (const cplx *)d_f, (const cplx *)d_g, d_a, d_b, (cplx *)d_v, (cplx *)d_w, n, n12);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

// out = (i)^times_i * mask * in on complex arrays, mask real [n_mask] repeated along the leading index
// (n = rows * n_mask): a Fourier derivative in spectral space.  The grid carries (column block, row), so
// there is no per-element modulo.
__global__ void __launch_bounds__(PW_THREADS) scale_mask_kernel(cplx *__restrict__ out, const cplx *__restrict__ in,
                                                                const double *__restrict__ mask, int64_t n_mask,
                                                                int times_i) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_mask) return;
    const int64_t i = (int64_t)blockIdx.y * n_mask + c;
    const double m = __ldg(&mask[c]);
    const cplx x = in[i];
    out[i] = times_i ? make_double2(-m * x.y, m * x.x) : make_double2(m * x.x, m * x.y);
}

int shb_scale_mask(void *d_out, const void *d_in, const double *d_mask, int64_t n_mask, int64_t n, int times_i,
                   void *stream) {
    SHB_REQUIRE(d_out && d_in && d_mask && n_mask >= 1 && n >= 0 && n % n_mask == 0, "bad arguments");
    if (n == 0) return SHB_OK;
    const int64_t rows = n / n_mask;
    SHB_REQUIRE(rows <= 65535, "leading extent too large");
    dim3 grid((unsigned)((n_mask + PW_THREADS - 1) / PW_THREADS), (unsigned)rows);
    scale_mask_kernel<<<grid, PW_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        (cplx *)d_out, (const cplx *)d_in, d_mask, n_mask, times_i);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

int shb_apply_mask(void *d_u, const double *d_mask, int64_t n, int ncomp, void *stream) {
    SHB_REQUIRE(d_u && d_mask && n >= 0 && ncomp >= 1, "bad arguments");
    if (n == 0) return SHB_OK;
    apply_mask_kernel<<<pw_blocks(n * ncomp), PW_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        (cplx *)d_u, d_mask, n, ncomp);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

}  // extern "C"
