"""Thin Python wrappers over the C ABI: torch tensors in, kernels launched on
torch's current stream.  Every function fails loudly when libshenb200.so or a
CUDA device is missing -- there is no eager/CPU fallback on the product path.
"""
import numpy as np
import torch

from . import _lib

LAUNCHES = 0   # number of libshenb200 kernel launches issued (bench.py reads it)


def plain(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) else t


def device():
    if not torch.cuda.is_available():
        raise _lib.ShenB200Error('no CUDA device: shenfun_b200 has no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def as_device_tensor(x):
    if isinstance(x, torch.Tensor):
        t = plain(x)
    else:
        a = np.ascontiguousarray(x)
        t = torch.from_numpy(a if a.flags.writeable else a.copy())
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise _lib.ShenB200Error('no CUDA device: shenfun_b200 has no CPU fallback')
        t = t.cuda()
    return t


def _req(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.ShenB200Error(f'{name}: expected a CUDA tensor (no CPU fallback)')
    if t.dtype != dtype:
        raise TypeError(f'{name}: expected {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise ValueError(f'{name}: tensor must be C-contiguous')
    return t.data_ptr()


def _call(fn_name, *args):
    global LAUNCHES
    L = _lib.load()
    _lib.check(getattr(L, fn_name)(*args))
    LAUNCHES += 1


C128 = torch.complex128
F64 = torch.float64


def fft_c2c_forward(u, out, outer, inner, n_fft, N, scale):
    _call('shb_fft_c2c_forward', _req(u, C128, 'in'), _req(out, C128, 'out'),
          outer, inner, n_fft, N, scale, _lib.stream_ptr())


def fft_c2c_backward(uh, out, outer, inner, n_fft, N, scale=1.0, dealias_direct=False):
    _call('shb_fft_c2c_backward', _req(uh, C128, 'in'), _req(out, C128, 'out'),
          outer, inner, n_fft, N, scale, int(dealias_direct), _lib.stream_ptr())


def fft_c2c_chunked(backward, u, out, outer, inner, n_fft, N, scale, dealias_direct=False, in_chunk=0, out_chunk=0):
    """C2C transform reading from / writing to the peer-chunked layout of the slab all-to-all."""
    _call('shb_fft_c2c_chunked', int(backward), _req(u, C128, 'in'), _req(out, C128, 'out'), outer, inner, n_fft, N,
          scale, int(dealias_direct), int(in_chunk), int(out_chunk), _lib.stream_ptr())


def fft_c2c_peer(backward, u, peer_ptrs, outer, inner, n_fft, N, scale, dealias_direct=False, in_chunk=0,
                 out_chunk=0):
    """C2C transform whose output blocks are stored directly into the peers' receive buffers."""
    import ctypes
    arr = (ctypes.c_void_p*len(peer_ptrs))(*[int(p) for p in peer_ptrs])
    _call('shb_fft_c2c_peer', int(backward), _req(u, C128, 'in'), arr, len(peer_ptrs), outer, inner, n_fft, N,
          scale, int(dealias_direct), int(in_chunk), int(out_chunk), _lib.stream_ptr())


def fft_c2c_lines(backward, u, out, peer_ptrs, na, ncols, c0, in_as, in_rs, out_as, out_rs, n_fft, N, scale,
                  dealias_direct=False, in_chunk=0, in_chunk_stride=0, out_chunk=0, out_chunk_stride=0, ncols_valid=None):
    """C2C transform of the lines (a, b), a < na, b < ncols, with explicit strides (shenb200.h:
    shb_fft_c2c_lines); output to `out` or, when `peer_ptrs` is given, into the peers' buffers."""
    import ctypes
    arr = None
    nr = 0
    if peer_ptrs is not None:
        arr = (ctypes.c_void_p*len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        nr = len(peer_ptrs)
    _call('shb_fft_c2c_lines', int(backward), _req(u, C128, 'in'), None if out is None else _req(out, C128, 'out'), arr,
          nr, na, ncols, ncols if ncols_valid is None else ncols_valid, c0, in_as, in_rs, out_as, out_rs, n_fft, N, scale,
          int(dealias_direct), int(in_chunk),
          int(in_chunk_stride), int(out_chunk), int(out_chunk_stride), _lib.stream_ptr())


def fft_c2r_cross(uh, b, out, nlines, n_fft, N, scale=1.0, dealias_direct=False):
    """out = irfft(uh) x b for (3, nlines, .) arrays (C2R fused with the cross product).  A complex
    `b` is a second spectral field, transformed in the same kernel: out = irfft(uh) x irfft(b)."""
    spectral_b = b.dtype == C128
    _call('shb_fft_c2r_cross', _req(uh, C128, 'in'), _req(b, C128 if spectral_b else F64, 'b'), _req(out, F64, 'out'),
          nlines, n_fft, N, scale, int(dealias_direct) | (2 if spectral_b else 0), _lib.stream_ptr())


def fft_c2c_backward_curl(u_hat, out, inner, n2, n_fft, N, k0, k1, k2, scale=1.0):
    """out[0:3] = 3/2-rule padded backward C2C of u_hat (3, N, inner) along the leading axis, out[3:6] the
    same of i k x u_hat formed on the fly (demo/NavierStokes.py:37-39); out: (6, n_fft, inner)."""
    _call('shb_fft_c2c_backward_curl', _req(u_hat, C128, 'u_hat'), _req(out, C128, 'out'), inner, n2, n_fft, N,
          _req(k0, F64, 'k0'), _req(k1, F64, 'k1'), _req(k2, F64, 'k2'), scale, _lib.stream_ptr())


def fft_r2c_forward(u, out, outer, inner, n_fft, N, scale):
    _call('shb_fft_r2c_forward', _req(u, F64, 'in'), _req(out, C128, 'out'),
          outer, inner, n_fft, N, scale, _lib.stream_ptr())


def fft_c2r_backward(uh, out, outer, inner, n_fft, N, scale=1.0, dealias_direct=False):
    _call('shb_fft_c2r_backward', _req(uh, C128, 'in'), _req(out, F64, 'out'),
          outer, inner, n_fft, N, scale, int(dealias_direct), _lib.stream_ptr())


def cross3d(a, b, c):
    """c = a x b on (3, ...) arrays (utilities/__init__.py:177-210 `cross`)."""
    assert a.shape == b.shape == c.shape and a.shape[0] == 3
    n = a[0].numel()
    dt = a.dtype
    _call('shb_cross3d', _req(a, dt, 'a'), _req(b, dt, 'b'), _req(c, dt, 'c'), n, int(dt.is_complex),
          _lib.stream_ptr())
    return c


def dot_grad3d(u, grads, out):
    """out[i] = sum_j u[j]*grads[3*i+j] on real (3, ...) arrays and nine (...) arrays."""
    assert len(grads) == 9 and u.shape[0] == 3 and out.shape == u.shape
    n = u[0].numel()
    _call('shb_dot_grad3d', _req(u, F64, 'u'), _ptr_array(list(grads)), _req(out, F64, 'out'), n, _lib.stream_ptr())
    return out


def curl_hat(u_hat, out, k0, k1, k2):
    n0, n1, n2 = u_hat.shape[1:]
    _call('shb_curl_hat', _req(u_hat, C128, 'u_hat'), _req(out, C128, 'out'), _req(k0, F64, 'k0'),
          _req(k1, F64, 'k1'), _req(k2, F64, 'k2'), n0, n1, n2, _lib.stream_ptr())
    return out


def ns_rk_stage(dU, u_hat, U0, U1, P, k0, k1, k2, nu, a_dt, b_dt, last, keep_dU=True, curl_out=None):
    n0, n1, n2 = u_hat.shape[1:]
    _call('shb_ns_rk_stage', _req(dU, C128, 'dU'), _req(u_hat, C128, 'u_hat'), _req(U0, C128, 'U0'),
          _req(U1, C128, 'U1'), None if P is None else _req(P, C128, 'P'),
          None if curl_out is None else _req(curl_out, C128, 'curl_out'), _req(k0, F64, 'k0'),
          _req(k1, F64, 'k1'), _req(k2, F64, 'k2'), n0, n1, n2, float(nu), float(a_dt), float(b_dt),
          int(bool(last)) | (0 if keep_dU else 2), _lib.stream_ptr())


def rk4_stage(dU, u_hat, U0, U1, L, a_dt, b_dt, last):
    ncomp = 1
    n = u_hat.numel()
    if L is not None:
        n = L.numel()
        ncomp = u_hat.numel()//n
    if L is not None and L.is_complex():
        _call('shb_rk4_stage_cplx', _req(dU, C128, 'dU'), _req(u_hat, C128, 'u_hat'), _req(U0, C128, 'U0'),
              _req(U1, C128, 'U1'), _req(L, C128, 'L'), n, ncomp, float(a_dt), float(b_dt), int(last),
              _lib.stream_ptr())
        return
    _call('shb_rk4_stage', _req(dU, C128, 'dU'), _req(u_hat, C128, 'u_hat'), _req(U0, C128, 'U0'),
          _req(U1, C128, 'U1'), None if L is None else _req(L, F64, 'L'), n, ncomp, float(a_dt),
          float(b_dt), int(last), _lib.stream_ptr())


def lincomb(out, terms, coefs, mask=None, saves=(), times_i=False):
    """out = mask * sum_j coefs[j]*terms[j] in one pass; `saves` = [(term index, destination)] also
    copies up to two terms out unchanged (the IMEX-RK stage right-hand side); `times_i` multiplies the
    (complex) result by i."""
    import ctypes
    dt = out.dtype
    n = out.numel()
    tp = (ctypes.c_void_p*len(terms))(*[_req(t, dt, 'term') for t in terms])
    for t in terms:
        assert t.numel() == n
    cf = (ctypes.c_double*len(terms))(*[float(c) for c in coefs])
    si = (ctypes.c_int*max(len(saves), 1))(*[int(i) for i, _ in saves])
    sd = (ctypes.c_void_p*max(len(saves), 1))(*[_req(d, dt, 'save') for _, d in saves])
    _call('shb_lincomb', _req(out, dt, 'out'), len(terms), tp, cf, None if mask is None else _req(mask, F64, 'mask'),
          0 if mask is None else mask.numel(), n, int(dt.is_complex) | (2 if times_i else 0), len(saves), si, sd,
          _lib.stream_ptr())
    return out


def kmm_vw(f, g, a, b, v, w):
    """v = i(a f + b g), w = i(b f - a g); a, b real, broadcast over the leading index
    (demo/ChannelFlow.py:235-239)."""
    n12 = a.numel()
    assert b.numel() == n12 and f.numel() % n12 == 0 and g.numel() == f.numel() == v.numel() == w.numel()
    _call('shb_kmm_vw', _req(f, C128, 'f'), _req(g, C128, 'g'), _req(a, F64, 'a'), _req(b, F64, 'b'),
          _req(v, C128, 'v'), _req(w, C128, 'w'), f.numel(), n12, _lib.stream_ptr())


def scale_mask(out, inp, mask, times_i=False):
    """out = mask * inp (times i) on complex arrays; the real mask covers the trailing dimensions."""
    n_mask, n = mask.numel(), inp.numel()
    if n % n_mask or n//n_mask > 65535 or out.numel() != n:
        return lincomb(out, [inp], [1.0], mask=mask, times_i=times_i)
    _call('shb_scale_mask', _req(out, C128, 'out'), _req(inp, C128, 'in'), _req(mask, F64, 'mask'), n_mask, n,
          int(bool(times_i)), _lib.stream_ptr())
    return out


def apply_mask(u_hat, mask):
    n = mask.numel()
    _call('shb_apply_mask', _req(u_hat, C128, 'u_hat'), _req(mask, F64, 'mask'), n, u_hat.numel()//n,
          _lib.stream_ptr())
    return u_hat


def _ptr_array(tensors):
    """Host array of device pointers (None -> NULL) for the C ABI."""
    import ctypes
    arr = (ctypes.c_void_p*len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else _req(t, F64, 'table')
    return arr


def banded_factor(kl, ku, step, mats, scales, fac, outer, inner, n, pin_line=-1):
    pm = _ptr_array(mats)
    ps = _ptr_array(scales)
    _call('shb_banded_factor', kl, ku, step, len(mats), pm, ps, _req(fac, F64, 'fac'), outer, inner, n,
          pin_line, _lib.stream_ptr())


def banded_solve(kl, ku, step, u, fac, outer, inner, n, n_u, fac_shared):
    _call('shb_banded_solve', kl, ku, step, _req(u, u.dtype, 'u'), _req(fac, F64, 'fac'), outer, inner, n, n_u,
          int(fac_shared), int(u.dtype.is_complex), _lib.stream_ptr())


def banded_matvec(u, out, offsets, mats, scales, outer, inner, n_rows, n_cols, n_u, n_out, accumulate):
    import ctypes
    offs = (ctypes.c_int*len(offsets))(*[int(o) for o in offsets])
    pm = _ptr_array(mats)
    ps = _ptr_array(scales)
    assert u.dtype == out.dtype
    _call('shb_banded_matvec', _req(u, u.dtype, 'u'), _req(out, out.dtype, 'out'), len(mats), len(offsets), offs,
          pm, ps, outer, inner, n_rows, n_cols, n_u, n_out, int(accumulate), int(u.dtype.is_complex),
          _lib.stream_ptr())


def helmholtz_factor(A0, A2, A4, Bm2, B0, B2, alfa, beta, fac, outer, inner, n, n_fac=None, neumann=False):
    """`fac` is either one packed tensor [4][outer][n][inner] or the four arrays (d0, d1, d2, L),
    each [outer][n_fac][inner]."""
    parts = [fac[i] for i in range(4)] if isinstance(fac, torch.Tensor) else list(fac)
    n_fac = n if n_fac is None else n_fac
    _call('shb_helmholtz_factor', *[_req(t, F64, 'table') for t in (A0, A2, A4, Bm2, B0, B2, alfa, beta)],
          *[_req(t, F64, 'fac') for t in parts], outer, inner, n, n_fac, int(neumann), _lib.stream_ptr())


def helmholtz_solve(b, u, fac, outer, inner, n, n_u, n_fac=None, neumann=False, shared=False):
    assert b.dtype == u.dtype
    parts = [fac[i] for i in range(4)] if isinstance(fac, torch.Tensor) else list(fac)
    n_fac = n if n_fac is None else n_fac
    _call('shb_helmholtz_solve', _req(b, b.dtype, 'b'), _req(u, u.dtype, 'u'), *[_req(t, F64, 'fac') for t in parts],
          outer, inner, n, n_fac, n_u, int(u.dtype.is_complex), int(neumann) | (2 if shared else 0),
          _lib.stream_ptr())


def helmholtz_factor_lean(tabs, alfa, beta, invd0, outer, inner, n):
    """invd0[outer][n][inner] = 1/d0 of every line (tabs = A0, A2, A4, Bm2, B0, B2)."""
    _call('shb_helmholtz_factor_lean', *[_req(t, F64, 'table') for t in tabs], _req(alfa, F64, 'alfa'),
          _req(beta, F64, 'beta'), _req(invd0, F64, 'invd0'), outer, inner, n, _lib.stream_ptr())


def helmholtz_lean_tables(tabs, n):
    """The two interleaved diagonal tables of shb_helmholtz_solve_lean (see shenb200.h)."""
    A0, A2, A4, Bm2, B0, B2 = tabs
    rows = int(_lib.load().shb_helmholtz_lean_table_rows(n))
    tF = torch.zeros((rows, 2), dtype=F64, device=A0.device)
    tB = torch.zeros((rows, 2), dtype=F64, device=A0.device)
    tF[2:n, 0] = Bm2[:n-2]
    tF[:n, 1] = A4[:n]
    tB[:n, 0] = A2[:n]
    tB[:n, 1] = B2[:n]
    return tF, tB


def helmholtz_solve_lean(b, u, lean_tabs, alfa, beta, invd0, outer, inner, n, n_u):
    """Chunk-parallel Helmholtz solve from the single stored factor array."""
    assert b.dtype == u.dtype
    tF, tB = lean_tabs
    _call('shb_helmholtz_solve_lean', _req(b, b.dtype, 'b'), _req(u, u.dtype, 'u'), _req(tF, F64, 'tabF'),
          _req(tB, F64, 'tabB'), _req(alfa, F64, 'alfa'), _req(beta, F64, 'beta'), _req(invd0, F64, 'invd0'),
          outer, inner, n, n_u, int(u.dtype.is_complex), _lib.stream_ptr())


def _stencil_args(soff, sten):
    import ctypes
    offs = (ctypes.c_int*len(soff))(*[int(o) for o in soff])
    return offs, _ptr_array(sten)


def cheb_scalar_product(u, out, outer, inner, n_phys, n_spec, scale, dim, soff, sten):
    assert u.dtype == out.dtype
    offs, ps = _stencil_args(soff, sten)
    _call('shb_cheb_scalar_product', _req(u, u.dtype, 'in'), _req(out, out.dtype, 'out'), outer, inner, n_phys,
          n_spec, float(scale), int(not u.dtype.is_complex), dim, len(soff), offs, ps, _lib.stream_ptr())


def cheb_backward(uh, out, outer, inner, n_phys, n_spec, dim, soff, sten, nbc=0, bc_lift=None):
    assert uh.dtype == out.dtype
    offs, ps = _stencil_args(soff, sten)
    _call('shb_cheb_backward', _req(uh, uh.dtype, 'in'), _req(out, out.dtype, 'out'), outer, inner, n_phys,
          n_spec, int(not uh.dtype.is_complex), dim, len(soff), offs, ps, nbc,
          None if bc_lift is None else _req(bc_lift, F64, 'bc_lift'), _lib.stream_ptr())


_PARITY = {}        # operator data_ptr -> (operator (kept alive), packed halves | None, forward flag, Kp)


def _parity_halves(A):
    """Test a (M, K) operator for the parity structure of the symmetric families and pack its two half
    operators for `shb_gemm_axis_parity` (include/shenb200.h).  None when the structure is absent."""
    import numpy as np
    a = A.detach().cpu().numpy()
    M, K = a.shape
    tol = 1e-13*max(1.0, float(np.abs(a).max()))
    sgn_r = (-1.0)**np.arange(M)[:, None]
    sgn_c = (-1.0)**np.arange(K)[None, :]
    if K % 2 == 0 and K >= 4 and np.abs(a[:, ::-1]-sgn_r*a).max() <= tol:          # values -> coefficients
        Kh, Mh = K//2, (M+1)//2
        Kp = Kh+Kh % 2
        sym = 0.5*(a[:, :Kh]+sgn_r*a[:, ::-1][:, :Kh])
        h = np.zeros((2, Mh, Kp))
        h[0, :, :Kh] = sym[0::2]
        h[1, :M//2, :Kh] = sym[1::2]
        return as_device_tensor(h), 1, Kp
    if M % 2 == 0 and M >= 4 and np.abs(a[::-1, :]-sgn_c*a).max() <= tol:          # coefficients -> values
        Mh, Kh = M//2, (K+1)//2
        Kp = Kh+Kh % 2
        sym = 0.5*(a[:Mh]+sgn_c*a[::-1][:Mh])
        h = np.zeros((2, Mh, Kp))
        h[0, :, :Kh] = sym[:, 0::2]
        h[1, :, :K//2] = sym[:, 1::2]
        return as_device_tensor(h), 0, Kp
    if M >= 4 and K >= 4 and not a[0::2, 1::2].any() and not a[1::2, 0::2].any():  # checkerboard (i + j odd: zero)
        Mh, Kh = (M+1)//2, (K+1)//2
        Kp = Kh+Kh % 2
        h = np.zeros((2, Mh, Kp))
        h[0, :, :Kh] = a[0::2, 0::2]
        h[1, :M//2, :K//2] = a[1::2, 1::2]
        flag = 2
        if not np.tril(h[0], -1).any() and not np.tril(h[1], -1).any():
            flag |= 4
        elif not np.triu(h[0], 1).any() and not np.triu(h[1], 1).any():
            flag |= 8
        return as_device_tensor(h), flag, Kp
    return None


def gemm_axis(A, u, out, outer, inner):
    """out[o][i][c] = sum_j A[i][j] u[o][j][c] (A real (M, K) device matrix).  Operators with the parity
    structure of the symmetric families (tested once per operator on the host; do the first call outside
    CUDA-graph capture) go through the half-size contraction `shb_gemm_axis_parity`; SHB_GEMM_PARITY=0: off."""
    import os
    global LAUNCHES
    assert u.dtype == out.dtype
    M, K = A.shape
    cplx = int(u.dtype.is_complex)
    W = inner*(2 if cplx else 1)
    if W >= 2 and W % 2 == 0 and outer <= 65535 and os.environ.get('SHB_GEMM_PARITY', '1') != '0':
        key = (A.data_ptr(), M, K)
        ent = _PARITY.get(key)
        if ent is None or ent[0] is not A:
            if len(_PARITY) > 64:
                _PARITY.clear()
            ent = _PARITY[key] = (A, _parity_halves(A))
        if ent[1] is not None:
            A2, fwd, Kp = ent[1]
            L = _lib.load()
            rc = L.shb_gemm_axis_parity(_req(A2, F64, 'A2'), _req(u, u.dtype, 'u'), _req(out, out.dtype, 'out'), M, K, Kp,
                                        fwd, outer, inner, cplx, _lib.stream_ptr())
            if rc != _lib.SHB_ERR_UNSUPPORTED:
                _lib.check(rc)
                LAUNCHES += 1
                return
    _call('shb_gemm_axis', _req(A, F64, 'A'), _req(u, u.dtype, 'u'), _req(out, out.dtype, 'out'), M, K, outer, inner,
          cplx, _lib.stream_ptr())


def tail_matvec(u, out, r, t, scale, w, outer, inner, n_rows, n_cols, n_u, n_out, accumulate):
    """out_k (+)= scale * r_k * sum_{j >= k+w, same parity} t_j u_j (separable upper tail)."""
    assert u.dtype == out.dtype
    _call('shb_tail_matvec', _req(u, u.dtype, 'u'), _req(out, out.dtype, 'out'), _req(r, F64, 'r'), _req(t, F64, 't'),
          None if scale is None else _req(scale, F64, 'scale'), int(w), outer, inner, n_rows, n_cols, n_u, n_out,
          int(accumulate), int(u.dtype.is_complex), _lib.stream_ptr())


def helmholtz_matvec(v, c, tabs, alfa, beta, outer, inner, n, n_u):
    assert v.dtype == c.dtype
    _call('shb_helmholtz_matvec', _req(v, v.dtype, 'v'), _req(c, c.dtype, 'c'), *[_req(t, F64, 'table') for t in tabs],
          _req(alfa, F64, 'alfa'), _req(beta, F64, 'beta'), outer, inner, n, n_u, int(v.dtype.is_complex),
          _lib.stream_ptr())


def biharmonic_factor(tab, a, b, c, fac, outer, inner, n):
    _call('shb_biharmonic_factor', *[_req(t, F64, 'table') for t in (tab, a, b, c)], _req(fac, F64, 'fac'),
          outer, inner, n, _lib.stream_ptr())


def biharmonic_solve(b, u, fac, outer, inner, n, n_u):
    assert b.dtype == u.dtype
    _call('shb_biharmonic_solve', _req(b, b.dtype, 'b'), _req(u, u.dtype, 'u'), _req(fac, F64, 'fac'), outer, inner,
          n, n_u, int(u.dtype.is_complex), _lib.stream_ptr())


def biharmonic_matvec(v, c, tab, sa, sb, sc, outer, inner, n, n_u):
    assert v.dtype == c.dtype
    _call('shb_biharmonic_matvec', _req(v, v.dtype, 'v'), _req(c, c.dtype, 'c'),
          *[_req(t, F64, 'table') for t in (tab, sa, sb, sc)], outer, inner, n, n_u, int(v.dtype.is_complex),
          _lib.stream_ptr())
