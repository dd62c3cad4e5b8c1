// fft_generic.cu -- any-length batched 1-D FFT along the middle axis of
// [outer][n][inner], shared-memory Stockham autosort with a run-time radix
// schedule (8,4,2,3,5 specialised butterflies + O(r^2) butterfly for other
// primes), with the Fourier-basis pad / truncate / Nyquist rules and the
// normalisation fused into the load and store phases.
//
// This is the *coverage* path (every n, every axis).  The headline sizes
// (powers of two and 3*2^k) are served by the register-resident kernels in
// fft_fast.cu; both share the C ABI entry points in fft_api.cu.
//
// Reference semantics restated here (no code shared):
//   shenfun/fourier/bases.py:131-149 (forward = FFT, truncate, *1/N)
//   shenfun/fourier/bases.py:404-435 (R2C pad/truncate, Nyquist rules)
//   shenfun/fourier/bases.py:570-598 (C2C pad/truncate)
#include "fft_common.cuh"
#include "butterfly.cuh"

namespace shb {

// ---- one Stockham stage over a tile held in shared memory ----------------------
// src/dst: [M][PT] complex, this thread works on column c with P threads per line.
template <int R, bool INV>
__device__ __forceinline__ void stage_fixed(const cplx *__restrict__ src, cplx *__restrict__ dst,
                                            const cplx *__restrict__ tw, int M, int Ns, int PT,
                                            int c, int t, int P) {
    const int nb = M / R;
    const int tstep = M / (Ns * R);
    for (int b = t; b < nb; b += P) {
        const int k = b % Ns;
        cplx v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = src[(size_t)(b + m * nb) * PT + c];
        if (Ns > 1) {
#pragma unroll
            for (int m = 1; m < R; ++m) {
                cplx w = __ldg(&tw[k * m * tstep]);
                v[m] = INV ? cmulc(v[m], w) : cmul(v[m], w);
            }
        }
        Bfly<R, INV>::run(v);
        const int j0 = (b / Ns) * Ns * R + k;
#pragma unroll
        for (int m = 0; m < R; ++m) dst[(size_t)(j0 + m * Ns) * PT + c] = v[m];
    }
}

// generic prime radix: O(R^2) butterfly, twiddles folded into one table index
template <bool INV>
__device__ __forceinline__ void stage_any(const cplx *__restrict__ src, cplx *__restrict__ dst,
                                          const cplx *__restrict__ tw, int M, int R, int Ns, int PT,
                                          int c, int t, int P) {
    const int nb = M / R;
    const int tstep = M / (Ns * R);
    for (int b = t; b < nb; b += P) {
        const int k = b % Ns;
        const int j0 = (b / Ns) * Ns * R + k;
        for (int q = 0; q < R; ++q) {
            double ax = 0.0, ay = 0.0;
            for (int m = 0; m < R; ++m) {
                cplx x = src[(size_t)(b + m * nb) * PT + c];
                long long idx = ((long long)k * m * tstep + (long long)((m * q) % R) * nb) % M;
                cplx w = __ldg(&tw[idx]);
                cplx p = INV ? cmulc(x, w) : cmul(x, w);
                ax += p.x;
                ay += p.y;
            }
            dst[(size_t)(j0 + q * Ns) * PT + c] = make_double2(ax, ay);
        }
    }
}

// Runs all stages; returns the buffer holding the result (natural order).
template <bool INV>
__device__ cplx *fft_core_smem(cplx *buf0, cplx *buf1, const FftDesc &d, int c, int t, int P,
                               bool active) {
    cplx *src = buf0, *dst = buf1;
    int Ns = 1;
    for (int s = 0; s < d.nstages; ++s) {
        const int R = d.radix[s];
        if (active) {
            switch (R) {
            case 8: stage_fixed<8, INV>(src, dst, d.tw, d.M, Ns, d.PT, c, t, P); break;
            case 4: stage_fixed<4, INV>(src, dst, d.tw, d.M, Ns, d.PT, c, t, P); break;
            case 2: stage_fixed<2, INV>(src, dst, d.tw, d.M, Ns, d.PT, c, t, P); break;
            case 3: stage_fixed<3, INV>(src, dst, d.tw, d.M, Ns, d.PT, c, t, P); break;
            case 5: stage_fixed<5, INV>(src, dst, d.tw, d.M, Ns, d.PT, c, t, P); break;
            default: stage_any<INV>(src, dst, d.tw, d.M, R, Ns, d.PT, c, t, P); break;
            }
        }
        __syncthreads();
        cplx *tmp = src; src = dst; dst = tmp;
        Ns *= R;
    }
    return src;
}

// ---- the kernel ------------------------------------------------------------------
// MODE selects the load (prologue) and store (epilogue) behaviour.
template <int MODE>
__global__ void __launch_bounds__(256) fft_generic_kernel(const __grid_constant__ FftDesc d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *buf0 = reinterpret_cast<cplx *>(smem_raw);
    cplx *buf1 = buf0 + (size_t)d.M * d.PT;

    constexpr bool INV = (MODE == FFT_C2C_BWD || MODE == FFT_C2R || MODE == FFT_C2R_FULL);
    constexpr bool DCT = (MODE == FFT_DCT2 || MODE == FFT_DCT3);
    const int T = d.T, PT = d.PT, M = d.M;
    const int P = blockDim.x / T;
    const int64_t line0 = (int64_t)blockIdx.x * T;
    const int64_t nlines = d.outer * d.inner;
    const int nl = (int)min((int64_t)T, nlines - line0);   // lines in this tile
    const bool contig = (d.inner == 1);

    const int n_in = d.n_in, n_out = d.n_out, N = d.N, n_fft = d.n_fft;
    const bool padded = (N < n_fft);
    const bool evenN = (N % 2 == 0);

    // -------- prologue: global -> buf0 (rows 0..M-1) -----------------------------
    // iteration space: (row, line) with the fast index chosen for coalescing
    const int rows_load = M;
    for (int idx = threadIdx.x; !DCT && idx < rows_load * T; idx += blockDim.x) {
        int r, l;
        if (contig) { r = idx % rows_load; l = idx / rows_load; }
        else        { l = idx % T;         r = idx / T; }
        cplx val = make_double2(0.0, 0.0);
        if (l < nl) {
            const int64_t line = line0 + l;
            const int64_t o = line / d.inner, cc = line % d.inner;
            const int64_t base = o * (int64_t)n_in * d.inner + cc;   // element (o, 0, cc)
            if (MODE == FFT_C2C_FWD) {
                const cplx *in = reinterpret_cast<const cplx *>(d.in);
                val = in[base + (int64_t)r * d.inner];
            } else if (MODE == FFT_C2C_BWD) {
                const cplx *in = reinterpret_cast<const cplx *>(d.in);
                if (!padded) {
                    val = in[base + (int64_t)r * d.inner];
                    if ((d.flags & 1) && r >= N / 3 && r <= N - N / 3) val = make_double2(0.0, 0.0);
                } else {
                    const int h = N / 2;
                    if (r <= h) {
                        val = in[base + (int64_t)r * d.inner];
                        if (evenN && r == h) val = cscale(val, 0.5);
                    } else if (r >= n_fft - h) {
                        val = in[base + (int64_t)(r - n_fft + N) * d.inner];
                        if (evenN && r == n_fft - h) val = cscale(val, 0.5);
                    }
                }
            } else if (MODE == FFT_R2C) {
                const double *in = reinterpret_cast<const double *>(d.in);
                val.x = in[base + (int64_t)(2 * r) * d.inner];
                val.y = in[base + (int64_t)(2 * r + 1) * d.inner];
            } else if (MODE == FFT_R2C_FULL) {
                const double *in = reinterpret_cast<const double *>(d.in);
                val.x = in[base + (int64_t)r * d.inner];
            } else if (MODE == FFT_C2R) {
                // Y[k] = (X[k] + conj X[M-k]) + i e^{+2 pi i k/n_fft} (X[k] - conj X[M-k])
                const cplx *in = reinterpret_cast<const cplx *>(d.in);
                const int ns = N / 2 + 1;            // stored modes
                auto getX = [&](int k) -> cplx {
                    cplx x = make_double2(0.0, 0.0);
                    if (k < ns) {
                        x = in[base + (int64_t)k * d.inner];
                        if ((d.flags & 1) && !padded && k >= N / 3) x = make_double2(0.0, 0.0);
                        if (padded && evenN && k == ns - 1) x = make_double2(0.5 * x.x, 0.0);
                    }
                    if (k == 0 || 2 * k == n_fft) x.y = 0.0;   // c2r ignores these
                    return x;
                };
                cplx xk = getX(r), xc = cconj(getX(M - r));
                cplx s = cadd(xk, xc), df = csub(xk, xc);
                cplx w = cconj(__ldg(&d.tw2[r]));                 // e^{+2 pi i r / n_fft}
                cplx iw = make_double2(-w.y, w.x);                // i * w
                val = cadd(s, cmul(iw, df));
            } else if (MODE == FFT_C2R_FULL) {
                const cplx *in = reinterpret_cast<const cplx *>(d.in);
                const int ns = N / 2 + 1;
                auto getX = [&](int k) -> cplx {
                    cplx x = make_double2(0.0, 0.0);
                    if (k < ns) {
                        x = in[base + (int64_t)k * d.inner];
                        if ((d.flags & 1) && !padded && k >= N / 3) x = make_double2(0.0, 0.0);
                        if (padded && evenN && k == ns - 1) x = make_double2(0.5 * x.x, 0.0);
                    }
                    if (k == 0 || 2 * k == n_fft) x.y = 0.0;
                    return x;
                };
                val = (2 * r <= n_fft) ? getX(r) : cconj(getX(n_fft - r));
            }
        }
        buf0[(size_t)r * PT + l] = val;
    }
    if (DCT) {
        // generic element fetch of the (real or complex) input line
        auto fetch = [&](int64_t base, int r) -> cplx {
            if (d.is_real) return make_double2(reinterpret_cast<const double *>(d.in)[base + (int64_t)r * d.inner], 0.0);
            return reinterpret_cast<const cplx *>(d.in)[base + (int64_t)r * d.inner];
        };
        if (MODE == FFT_DCT2) {
            // even-odd permutation: v[n] = x[2n], v[M-1-n] = x[2n+1]
            for (int idx = threadIdx.x; idx < M * T; idx += blockDim.x) {
                int r, l;
                if (contig) { r = idx % M; l = idx / M; } else { l = idx % T; r = idx / T; }
                cplx val = make_double2(0.0, 0.0);
                if (l < nl) {
                    const int64_t line = line0 + l;
                    const int64_t o = line / d.inner, cc = line % d.inner;
                    val = fetch(o * (int64_t)n_in * d.inner + cc, r);
                }
                const int pos = (r & 1) ? (M - 1 - (r >> 1)) : (r >> 1);
                buf0[(size_t)pos * PT + l] = val;
            }
        } else {
            // DCT3: stored coefficients -> buf0; stencil transpose (to_ortho) c_j -> buf1;
            // U_0 = c_0, U_k = 0.5 w_k (c_k + i c_{M-k}) -> buf0
            for (int idx = threadIdx.x; idx < M * T; idx += blockDim.x) {
                int r, l;
                if (contig) { r = idx % M; l = idx / M; } else { l = idx % T; r = idx / T; }
                cplx val = make_double2(0.0, 0.0);
                if (l < nl && r < d.dim + d.nbc && r < n_in) {
                    const int64_t line = line0 + l;
                    const int64_t o = line / d.inner, cc = line % d.inner;
                    val = fetch(o * (int64_t)n_in * d.inner + cc, r);
                }
                buf0[(size_t)r * PT + l] = val;
            }
            __syncthreads();
            for (int idx = threadIdx.x; idx < M * T; idx += blockDim.x) {
                const int l = idx % T, jr = idx / T;
                cplx acc = make_double2(0.0, 0.0);
                for (int i = 0; i < d.nsten; ++i) {
                    const int k = jr - d.soff[i];
                    if (k >= 0 && k < d.dim) {
                        const double a = d.sten[i] ? __ldg(&d.sten[i][k]) : 1.0;
                        const cplx x = buf0[(size_t)k * PT + l];
                        acc.x = fma(a, x.x, acc.x);
                        acc.y = fma(a, x.y, acc.y);
                    }
                }
                if (d.nbc > 0 && jr < 4) {
                    for (int b = 0; b < d.nbc; ++b) {
                        if (d.dim + b >= M) break;
                        const double a = __ldg(&d.bc_lift[b * 4 + jr]);
                        const cplx x = buf0[(size_t)(d.dim + b) * PT + l];
                        acc.x = fma(a, x.x, acc.x);
                        acc.y = fma(a, x.y, acc.y);
                    }
                }
                buf1[(size_t)jr * PT + l] = acc;
            }
            __syncthreads();
            for (int idx = threadIdx.x; idx < M * T; idx += blockDim.x) {
                const int l = idx % T, k = idx / T;
                cplx u;
                if (k == 0) u = buf1[l];
                else {
                    const cplx ck = buf1[(size_t)k * PT + l], cm = buf1[(size_t)(M - k) * PT + l];
                    const cplx w = __ldg(&d.tw3[k]);
                    const cplx sum = make_double2(ck.x - cm.y, ck.y + cm.x);     // c_k + i c_{M-k}
                    u = cscale(cmul(w, sum), 0.5);
                }
                buf0[(size_t)k * PT + l] = u;
            }
        }
    }
    __syncthreads();

    // -------- FFT -------------------------------------------------------------------
    const int c = threadIdx.x % T;
    const int t = threadIdx.x / T;
    cplx *res = fft_core_smem<INV>(buf0, buf1, d, c, t, P, true);

    if (DCT) {
        cplx *oth = (res == buf0) ? buf1 : buf0;
        if (MODE == FFT_DCT2) {
            // S_k = scale * (w_k V_k + conj(w_k) V_{M-k})  -> oth ; then out_k = sum_i sten_i[k] S_{k+soff_i}
            for (int idx = threadIdx.x; idx < M * T; idx += blockDim.x) {
                const int l = idx % T, k = idx / T;
                const cplx vk = res[(size_t)k * PT + l], vm = res[(size_t)((M - k) % M) * PT + l];
                const cplx w = __ldg(&d.tw3[k]);
                cplx sres = cadd(cmul(w, vk), cmulc(vm, w));
                oth[(size_t)k * PT + l] = cscale(sres, d.scale);
            }
            __syncthreads();
            for (int idx = threadIdx.x; idx < n_out * T; idx += blockDim.x) {
                int r, l;
                if (contig) { r = idx % n_out; l = idx / n_out; } else { l = idx % T; r = idx / T; }
                if (l >= nl) continue;
                cplx acc = make_double2(0.0, 0.0);
                if (r < d.dim) {
                    for (int i = 0; i < d.nsten; ++i) {
                        const int jj = r + d.soff[i];
                        if (jj >= 0 && jj < M) {
                            const double a = d.sten[i] ? __ldg(&d.sten[i][r]) : 1.0;
                            const cplx x = oth[(size_t)jj * PT + l];
                            acc.x = fma(a, x.x, acc.x);
                            acc.y = fma(a, x.y, acc.y);
                        }
                    }
                }
                const int64_t line = line0 + l;
                const int64_t o = line / d.inner, cc = line % d.inner;
                const int64_t oidx = o * (int64_t)n_out * d.inner + cc + (int64_t)r * d.inner;
                if (d.is_real) reinterpret_cast<double *>(d.out)[oidx] = acc.x;
                else reinterpret_cast<cplx *>(d.out)[oidx] = acc;
            }
        } else {
            // un-permute: out[2n] = y[n], out[2n+1] = y[M-1-n]
            for (int idx = threadIdx.x; idx < M * T; idx += blockDim.x) {
                int r, l;
                if (contig) { r = idx % M; l = idx / M; } else { l = idx % T; r = idx / T; }
                if (l >= nl) continue;
                const int pos = (r & 1) ? (M - 1 - (r >> 1)) : (r >> 1);
                const cplx y = cscale(res[(size_t)pos * PT + l], d.scale);
                const int64_t line = line0 + l;
                const int64_t o = line / d.inner, cc = line % d.inner;
                const int64_t oidx = o * (int64_t)n_out * d.inner + cc + (int64_t)r * d.inner;
                if (d.is_real) reinterpret_cast<double *>(d.out)[oidx] = y.x;
                else reinterpret_cast<cplx *>(d.out)[oidx] = y;
            }
        }
        return;
    }

    // -------- epilogue: result -> global ----------------------------------------------
    const double scale = d.scale;
    int rows_store;
    if (MODE == FFT_C2R) rows_store = M;          // each row writes 2 reals
    else rows_store = n_out;
    for (int idx = threadIdx.x; idx < rows_store * T; idx += blockDim.x) {
        int r, l;
        if (contig) { r = idx % rows_store; l = idx / rows_store; }
        else        { l = idx % T;          r = idx / T; }
        if (l >= nl) continue;
        const int64_t line = line0 + l;
        const int64_t o = line / d.inner, cc = line % d.inner;
        const int64_t obase = o * (int64_t)n_out * d.inner + cc;
        if (MODE == FFT_C2C_FWD) {
            cplx *out = reinterpret_cast<cplx *>(d.out);
            cplx v;
            if (!padded) v = res[(size_t)r * PT + l];
            else {
                const int h = N / 2;
                v = make_double2(0.0, 0.0);
                if (r <= h) v = res[(size_t)r * PT + l];
                if (r >= N - h) v = cadd(v, res[(size_t)(n_fft - N + r) * PT + l]);
            }
            out[obase + (int64_t)r * d.inner] = cscale(v, scale);
        } else if (MODE == FFT_C2C_BWD) {
            cplx *out = reinterpret_cast<cplx *>(d.out);
            out[obase + (int64_t)r * d.inner] = cscale(res[(size_t)r * PT + l], scale);
        } else if (MODE == FFT_R2C) {
            // X[k] = E + W^k O, E = (Z[k] + conj Z[M-k])/2, O = -i (Z[k] - conj Z[M-k])/2
            cplx *out = reinterpret_cast<cplx *>(d.out);
            const int k = r;
            cplx zk = res[(size_t)(k % M) * PT + l];
            cplx zc = cconj(res[(size_t)((M - k) % M) * PT + l]);
            cplx e = cscale(cadd(zk, zc), 0.5);
            cplx df = cscale(csub(zk, zc), 0.5);
            cplx o_ = make_double2(df.y, -df.x);
            cplx w = __ldg(&d.tw2[k]);
            cplx x = cadd(e, cmul(w, o_));
            if (2 * k == n_fft) x.y = 0.0;
            if (padded && evenN && k == N / 2) x = make_double2(2.0 * x.x, 0.0);
            out[obase + (int64_t)k * d.inner] = cscale(x, scale);
        } else if (MODE == FFT_R2C_FULL) {
            cplx *out = reinterpret_cast<cplx *>(d.out);
            cplx x = res[(size_t)r * PT + l];
            if (padded && evenN && r == N / 2) x = make_double2(2.0 * x.x, 0.0);
            out[obase + (int64_t)r * d.inner] = cscale(x, scale);
        } else if (MODE == FFT_C2R) {
            double *out = reinterpret_cast<double *>(d.out);
            cplx z = res[(size_t)r * PT + l];
            out[obase + (int64_t)(2 * r) * d.inner] = z.x * scale;
            out[obase + (int64_t)(2 * r + 1) * d.inner] = z.y * scale;
        } else if (MODE == FFT_C2R_FULL) {
            double *out = reinterpret_cast<double *>(d.out);
            out[obase + (int64_t)r * d.inner] = res[(size_t)r * PT + l].x * scale;
        }
    }
}

template <int MODE>
static int launch_mode(const FftDesc &d, int threads, size_t smem, int64_t ntiles, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        SHB_CUDA(cudaFuncSetAttribute(fft_generic_kernel<MODE>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    fft_generic_kernel<MODE><<<(unsigned)ntiles, threads, smem, st>>>(d);
    SHB_CUDA(cudaGetLastError());
    return SHB_OK;
}

static int factorize(int M, int *radix) {
    int n = 0;
    const int pref[] = {8, 4, 2, 3, 5};
    for (int p : pref)
        while (M % p == 0 && M > 1) { radix[n++] = p; M /= p; if (n >= FFT_MAX_STAGES) return -1; }
    for (int p = 7; M > 1; p += 2)
        while (M % p == 0) { radix[n++] = p; M /= p; if (n >= FFT_MAX_STAGES) return -1; }
    return n;
}

int fft_generic_launch(int mode, const void *in, void *out, int64_t outer, int64_t inner,
                       int n_fft, int N, double scale, int flags, cudaStream_t st) {
    FftDesc d;
    d.in = in; d.out = out; d.outer = outer; d.inner = inner;
    d.n_fft = n_fft; d.N = N; d.scale = scale; d.flags = flags;
    const int ns = N / 2 + 1;
    switch (mode) {
    case FFT_C2C_FWD: d.M = n_fft; d.n_in = n_fft; d.n_out = N; break;
    case FFT_C2C_BWD: d.M = n_fft; d.n_in = N; d.n_out = n_fft; break;
    case FFT_R2C:
        if (n_fft % 2) { mode = FFT_R2C_FULL; d.M = n_fft; } else d.M = n_fft / 2;
        d.n_in = n_fft; d.n_out = ns; break;
    case FFT_C2R:
        if (n_fft % 2) { mode = FFT_C2R_FULL; d.M = n_fft; } else d.M = n_fft / 2;
        d.n_in = ns; d.n_out = n_fft; break;
    default: set_error("fft_generic_launch: bad mode %d", mode); return SHB_ERR_ARG;
    }
    if (outer * inner == 0) return SHB_OK;
    d.nstages = factorize(d.M, d.radix);
    if (d.nstages < 0) { set_error("fft: too many factors in n=%d", d.M); return SHB_ERR_UNSUPPORTED; }
    if (d.M == 1) d.nstages = 0;
    d.tw = twiddle_table(0, d.M);
    d.tw2 = (mode == FFT_R2C || mode == FFT_C2R) ? twiddle_table(1, d.M) : nullptr;
    if (!d.tw || ((mode == FFT_R2C || mode == FFT_C2R) && !d.tw2)) {
        set_error("fft: twiddle table allocation failed"); return SHB_ERR_NOMEM;
    }
    // tile geometry: T lines per CTA, 256 threads, two ping-pong buffers
    const size_t smem_max = 200 * 1024;
    int T = 16;
    while (T > 1 && 2 * (size_t)d.M * (T | 1) * sizeof(cplx) > smem_max) T >>= 1;
    const int64_t nlines = outer * inner;
    while (T > 1 && (int64_t)T / 2 >= nlines) T >>= 1;
    d.T = T; d.PT = (T | 1);
    const size_t smem = 2 * (size_t)d.M * d.PT * sizeof(cplx);
    if (smem > 227 * 1024) { set_error("fft: n=%d too long for the shared-memory path", n_fft); return SHB_ERR_UNSUPPORTED; }
    const int threads = 256;
    const int64_t ntiles = (nlines + T - 1) / T;
    switch (mode) {
    case FFT_C2C_FWD: return launch_mode<FFT_C2C_FWD>(d, threads, smem, ntiles, st);
    case FFT_C2C_BWD: return launch_mode<FFT_C2C_BWD>(d, threads, smem, ntiles, st);
    case FFT_R2C: return launch_mode<FFT_R2C>(d, threads, smem, ntiles, st);
    case FFT_C2R: return launch_mode<FFT_C2R>(d, threads, smem, ntiles, st);
    case FFT_R2C_FULL: return launch_mode<FFT_R2C_FULL>(d, threads, smem, ntiles, st);
    case FFT_C2R_FULL: return launch_mode<FFT_C2R_FULL>(d, threads, smem, ntiles, st);
    }
    return SHB_ERR_ARG;
}

int dct_generic_launch(int mode, const void *in, void *out, int64_t outer, int64_t inner, int n_phys,
                       int n_spec, double scale, const DctArgs &a, cudaStream_t st) {
    FftDesc d;
    d.in = in; d.out = out; d.outer = outer; d.inner = inner;
    d.n_fft = n_phys; d.N = n_spec; d.scale = scale; d.flags = 0;
    d.M = n_phys;
    if (mode == FFT_DCT2) { d.n_in = n_phys; d.n_out = n_spec; }
    else if (mode == FFT_DCT3) { d.n_in = n_spec; d.n_out = n_phys; }
    else { set_error("dct: bad mode %d", mode); return SHB_ERR_ARG; }
    if (outer * inner == 0) return SHB_OK;
    d.is_real = a.is_real; d.dim = a.dim; d.nsten = a.nsten; d.nbc = a.nbc; d.bc_lift = a.bc_lift;
    for (int i = 0; i < DCT_MAX_STENCIL; ++i) { d.soff[i] = a.soff[i]; d.sten[i] = a.sten[i]; }
    d.nstages = factorize(d.M, d.radix);
    if (d.nstages < 0) { set_error("dct: too many factors in n=%d", d.M); return SHB_ERR_UNSUPPORTED; }
    if (d.M == 1) d.nstages = 0;
    d.tw = twiddle_table(0, d.M);
    d.tw2 = nullptr;
    d.tw3 = twiddle_table(2, d.M);
    if (!d.tw || !d.tw3) { set_error("dct: twiddle table allocation failed"); return SHB_ERR_NOMEM; }
    const size_t smem_max = 200 * 1024;
    int T = 16;
    while (T > 1 && 2 * (size_t)d.M * (T | 1) * sizeof(cplx) > smem_max) T >>= 1;
    const int64_t nlines = outer * inner;
    while (T > 1 && (int64_t)T / 2 >= nlines) T >>= 1;
    d.T = T; d.PT = (T | 1);
    const size_t smem = 2 * (size_t)d.M * d.PT * sizeof(cplx);
    if (smem > 227 * 1024) { set_error("dct: n=%d too long for the shared-memory path", n_phys); return SHB_ERR_UNSUPPORTED; }
    const int64_t ntiles = (nlines + T - 1) / T;
    if (mode == FFT_DCT2) return launch_mode<FFT_DCT2>(d, 256, smem, ntiles, st);
    return launch_mode<FFT_DCT3>(d, 256, smem, ntiles, st);
}

}  // namespace shb
