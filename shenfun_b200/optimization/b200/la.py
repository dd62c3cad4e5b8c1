"""Line solvers under the reference's kernel names and positional signatures.

    X_LU(data)                    shenfun/la.py:331-820, optimization/cython/la.pyx:199-303
    X_Solve(u, data, axis)        la.pyx:462-611
    X_inner_solve(u, data)
    SolverGeneric1ND_solve_data(u, data, sol, naxes, is_zero_index)      la.pyx:371-426
    LU_Helmholtz / Solve_Helmholtz / Poisson_Solve_ADD                   la.pyx:611-973
    LU_Biharmonic / Biharmonic_factor_pr / Biharmonic_Solve              la.pyx:975-2121

`data` keeps the reference's scipy-DIA layout (`data[d, j]` = entry of column j on diagonal d),
so factors written by `X_LU` here are the same numbers the reference writes and the device
sweeps read that layout directly (shb_banded_solve, fac_shared = 2 / 3).
"""
import numpy as np
import torch

from ... import ops
from ._args import Staged, axis_view, host

# (kl, ku, step) and the diagonal offsets in `data` row order
_KIND = {
    'TDMA': (1, 1, 2), 'TDMA_O': (1, 1, 1), 'PDMA': (2, 2, 2), 'FDMA': (1, 2, 2),
    'TwoDMA': (0, 1, 2), 'ThreeDMA': (0, 2, 2), 'HeptaDMA': (2, 4, 2), 'DiagMA': (0, 0, 1),
}


def _offsets(kind):
    kl, ku, step = _KIND[kind]
    return [(d-kl)*step for d in range(kl+1+ku)]


def _dia_to_rows(data, kind):
    """Row-indexed table tab[d][i] = M[i, i+off_d] from scipy-DIA data[d][i+off_d]."""
    n = data.shape[-1]
    tab = torch.zeros_like(data)
    for d, off in enumerate(_offsets(kind)):
        if off >= 0:
            tab[..., d, :n-off] = data[..., d, off:]
        else:
            tab[..., d, -off:] = data[..., d, :n+off]
    return tab


def _rows_to_dia(tab, data, kind):
    n = data.shape[-1]
    for d, off in enumerate(_offsets(kind)):
        if off >= 0:
            data[..., d, off:] = tab[..., d, :n-off]
        else:
            data[..., d, :n+off] = tab[..., d, -off:]


def _lu(kind, data):
    st = Staged()
    dt = st(data, out=True, dtype=torch.float64)
    kl, ku, step = _KIND[kind]
    assert dt.ndim == 2 and dt.shape[0] == kl+1+ku, f'{kind}_LU: data must be ({kl+1+ku}, N)'
    n = dt.shape[1]
    tab = _dia_to_rows(dt, kind).contiguous()
    fac = torch.empty((kl+1+ku, 1, n, 1), dtype=torch.float64, device=dt.device)
    ops.banded_factor(kl, ku, step, [tab], [None], fac, 1, 1, n)
    _rows_to_dia(fac.view(kl+1+ku, n), dt, kind)
    st.finish()


def _solve(kind, u, data, axis=0):
    st = Staged()
    ut = st(u, out=True)
    dt = st(data, dtype=torch.float64)
    kl, ku, step = _KIND[kind]
    assert dt.ndim == 2 and dt.shape[0] == kl+1+ku, f'{kind}_Solve: data must be ({kl+1+ku}, N)'
    outer, n_u, inner = axis_view(ut.shape, axis)
    n = dt.shape[1]
    assert n <= n_u
    ops.banded_solve(kl, ku, step, ut, dt, outer, inner, n, n_u, 2)
    st.finish()
    return u


def _make(kind, has_lu=True):
    def LU(data):
        return _lu(kind, data)

    def Solve(u, data, axis=0):
        return _solve(kind, u, data, axis)

    def inner_solve(u, data):
        return _solve(kind, u, data, 0)
    LU.__name__ = LU.__qualname__ = f'{kind}_LU'
    Solve.__name__ = Solve.__qualname__ = f'{kind}_Solve'
    inner_solve.__name__ = inner_solve.__qualname__ = f'{kind}_inner_solve'
    return (LU if has_lu else None), Solve, inner_solve


TDMA_LU, TDMA_Solve, TDMA_inner_solve = _make('TDMA')
TDMA_O_LU, TDMA_O_Solve, TDMA_O_inner_solve = _make('TDMA_O')
PDMA_LU, PDMA_Solve, PDMA_inner_solve = _make('PDMA')
FDMA_LU, FDMA_Solve, FDMA_inner_solve = _make('FDMA')
HeptaDMA_LU, HeptaDMA_Solve, HeptaDMA_inner_solve = _make('HeptaDMA')
_, TwoDMA_Solve, TwoDMA_inner_solve = _make('TwoDMA', False)
_, ThreeDMA_Solve, ThreeDMA_inner_solve = _make('ThreeDMA', False)
_, DiagMA_Solve, DiagMA_inner_solve = _make('DiagMA', False)


def SolverGeneric1ND_solve_data(u, data, sol, naxes, is_zero_index):
    """Per-wavenumber banded solves: `data[i, j]` (2-D: `data[i]`) holds the DIA LU factors of
    the line with non-solve indices (i, j); `sol` is one of the `X_inner_solve` functions and
    is resolved by name like the Cython path does (la.pyx:307-325).  The (0, 0) line is left
    untouched when `is_zero_index` (la.py:1352-1360)."""
    name = getattr(sol, '__name__', str(sol))
    kind = name.replace('_inner_solve', '')
    if kind not in _KIND:
        raise NotImplementedError(f'SolverGeneric1ND_solve_data: no device kernel for {name}')
    kl, ku, step = _KIND[kind]
    st = Staged()
    ut = st(u, out=True)
    dt = st(data, dtype=torch.float64)
    assert ut.ndim in (2, 3) and dt.ndim == ut.ndim+1 and dt.shape[-2] == kl+1+ku
    outer, n_u, inner = axis_view(ut.shape, naxes)
    n = dt.shape[-1]
    assert outer*inner == int(np.prod(dt.shape[:-2])), 'data does not match the lines of u'
    keep = None
    if is_zero_index:
        line0 = ut.view(outer, n_u, inner)[0, :, 0]
        keep = line0.clone()
    ops.banded_solve(kl, ku, step, ut, dt, outer, inner, n, n_u, 3)
    if keep is not None:
        line0.copy_(keep)
    st.finish()
    return u


# ---- Chebyshev Helmholtz -------------------------------------------------------------------

def _lines(scale, shape, axis, device):
    """Per-line scale array [outer*inner] from a scalar or an array broadcastable against
    `shape` with extent 1 along `axis`."""
    s = list(shape)
    s[axis] = 1
    if isinstance(scale, torch.Tensor):
        t = ops.plain(scale).to(torch.float64)
        if t.ndim == 0:
            t = t.reshape([1]*len(s))
        return t.expand(s).contiguous().view(-1).to(device)
    a = np.asarray(host(scale), dtype=float)
    if a.ndim < len(s):
        a = a.reshape([1]*(len(s)-a.ndim)+list(a.shape))
    a = np.broadcast_to(a, s)
    return ops.as_device_tensor(np.ascontiguousarray(a).ravel())


def _helm_tables(A, B, n, neumann):
    """The six diagonals, zero-padded to n rows; Neumann: columns scaled by 1/k^2 as in
    la.pyx:737-745."""
    def diag(m, off):
        v = np.zeros(n)
        if off in m:
            d = np.real(host(m[off])).astype(float)
            d = np.broadcast_to(d, (n-abs(off),)) if d.size == 1 else d
            v[:len(d)] = d
        return v
    A0, A2, A4 = diag(A, 0), diag(A, 2), diag(A, 4)
    Bm2, B0, B2 = diag(B, -2), diag(B, 0), diag(B, 2)
    if neumann:
        k = np.arange(n, dtype=float)
        k[0] = 1.0
        k2 = k*k
        A0 /= k2
        B0 /= k2
        A2[:n-2] /= (np.arange(2, n, dtype=float))**2
        B2[:n-2] /= (np.arange(2, n, dtype=float))**2
        A4[:n-4] /= (np.arange(4, n, dtype=float))**2
        Bm2[1:n-2] /= (np.arange(1, n-2, dtype=float))**2
    return [ops.as_device_tensor(t) for t in (A0, A2, A4, Bm2, B0, B2)]


def LU_Helmholtz(A, B, A_s, B_s, neumann, d0, d1, d2, L, axis=0):
    """Factor alfa*A + beta*B for every line into d0, d1, d2, L (arrays of the solution's shape
    with N = A.shape[0] + 2 entries along `axis`, chebyshev/la.py:270-276)."""
    st = Staged()
    f = [st(x, out=True, dtype=torch.float64) for x in (d0, d1, d2, L)]
    n = int(np.size(host(A[0])))
    outer, n_fac, inner = axis_view(f[0].shape, axis)
    assert n_fac >= n
    alfa = _lines(A_s, f[0].shape, axis, f[0].device)
    beta = _lines(B_s, f[0].shape, axis, f[0].device)
    tabs = _helm_tables(A, B, n, bool(neumann))
    ops.helmholtz_factor(*tabs, alfa, beta, f, outer, inner, n, n_fac=n_fac, neumann=bool(neumann))
    st.finish()


def Solve_Helmholtz(b, u, neumann, d0, d1, d2, L, axis=0):
    st = Staged()
    ut = st(u, out=True)
    bt = ut if b is u else st(b, dtype=ut.dtype)
    f = [st(x, dtype=torch.float64) for x in (d0, d1, d2, L)]
    outer, n_fac, inner = axis_view(f[0].shape, axis)
    n_u = ut.shape[axis]
    n = n_fac-2
    ops.helmholtz_solve(bt, ut, f, outer, inner, n, n_u, n_fac=n_fac, neumann=bool(neumann))
    st.finish()
    return u


def Poisson_Solve_ADD(A, b, u, axis=0):
    """A u = b with A the Chebyshev Dirichlet stiffness matrix: upper triangular, every row
    constant from the +2 diagonal on (chebyshev/la.py:13-30, helmholtz.py:388-417)."""
    st = Staged()
    ut = st(u, out=True)
    bt = ut if b is u else st(b, dtype=ut.dtype)
    d = np.real(host(A[0])).astype(float)
    n = d.shape[0]
    d1 = np.zeros(n)
    d1[:n-2] = np.real(host(A[2]))
    scale = float(np.real(getattr(A, 'scale', 1.0)))
    fac = [ops.as_device_tensor(t) for t in (d*scale, d1*scale, d1*scale, np.zeros(n))]
    outer, n_u, inner = axis_view(ut.shape, axis)
    ops.helmholtz_solve(bt, ut, fac, outer, inner, n, n_u, n_fac=n, shared=True)
    st.finish()
    return u


# ---- Chebyshev biharmonic ------------------------------------------------------------------
# The reference keeps the factors split by parity: arrays of shape (2,) + shape with M = n/2 rows
# along axis+1 (chebyshev/la.py:488-500).  The device kernels use one packed table
# fac[7][outer][n][inner] = (L1, L2, D, U1, U2, FA, FB) with row i = 2m + parity, so these
# functions scatter to / gather from the reference arrays; a gathered table is cached against the
# tensors' version counters so that repeated solves do not re-pack.

_PACK_CACHE = []


def _bih_tab(n, sii, siu, siuu, ail, aii, aiu, bill, bil, bii, biu, biuu):
    k = np.arange(n, dtype=float)
    tab = np.zeros((13, n))

    def put(row, d, off):
        d = np.real(host(d)).astype(float)
        if off >= 0:
            tab[row, :len(d)] = d
        else:
            tab[row, -off:-off+len(d)] = d
    put(0, sii, 0); put(1, siu, 2); put(2, siuu, 4)
    tab[3] = 8*np.pi*k*(k+1)*(k+2)*(k+4)          # la.pyx:1218-1236 (pp, rr)
    tab[4] = 24*np.pi*(k+1)*(k+2)
    put(5, ail, -2); put(6, aii, 0); put(7, aiu, 2)
    put(8, bill, -4); put(9, bil, -2); put(10, bii, 0); put(11, biu, 2); put(12, biuu, 4)
    return ops.as_device_tensor(tab)


def _split_view(x, axis):
    """Reference factor array (2, ..., M, ...) -> view [2][outer][M][inner]."""
    shape = x.shape[1:]
    outer, M, inner = axis_view(shape, axis)
    return x.view(2, outer, M, inner), outer, M, inner


def _rows_from_split(x, axis, shift):
    """[outer][n][inner] table with row 2(m+shift)+p taken from x[p][m]."""
    v, outer, M, inner = _split_view(x, axis)
    n = 2*M
    t = torch.zeros((outer, M, 2, inner), dtype=torch.float64, device=x.device)
    if shift < M:
        t[:, shift:, :, :] = v[:, :, :M-shift, :].permute(1, 2, 0, 3)
    return t.view(outer, n, inner)


def _split_from_rows(rows, x, axis, shift):
    v, outer, M, inner = _split_view(x, axis)
    r = rows.view(outer, M, 2, inner)
    v.zero_()
    if shift < M:
        v[:, :, :M-shift, :] = r[:, shift:, :, :].permute(2, 0, 1, 3)


def LU_Biharmonic(a0, alfa, beta, sii, siu, siuu, ail, aii, aiu, bill, bil, bii, biu, biuu,
                  u0, u1, u2, l0, l1, axis=0):
    st = Staged()
    f = [st(x, out=True, dtype=torch.float64) for x in (u0, u1, u2, l0, l1)]
    n = int(np.size(host(sii)))
    shape = tuple(f[0].shape[1:])
    outer, M, inner = axis_view(shape, axis)
    if n != 2*M:
        raise NotImplementedError('LU_Biharmonic: odd number of unknowns')
    dev = f[0].device
    a = torch.full((outer*inner,), float(np.atleast_1d(host(a0)).item()), dtype=torch.float64, device=dev)
    b = _lines(alfa, shape, axis, dev)
    c = _lines(beta, shape, axis, dev)
    tab = _bih_tab(n, sii, siu, siuu, ail, aii, aiu, bill, bil, bii, biu, biuu)
    fac = torch.empty((7, outer, n, inner), dtype=torch.float64, device=dev)
    ops.biharmonic_factor(tab, a, b, c, fac, outer, inner, n)
    _split_from_rows(fac[2], f[0], axis, 0)      # u0 = D
    _split_from_rows(fac[3], f[1], axis, 0)      # u1
    _split_from_rows(fac[4], f[2], axis, 0)      # u2
    _split_from_rows(fac[0], f[3], axis, 1)      # l0[m] multiplies row m+1
    _split_from_rows(fac[1], f[4], axis, 2)      # l1[m] multiplies row m+2
    st.finish()


def Biharmonic_factor_pr(ak, bk, l0, l1, axis=0):
    """Eliminated coefficients of the rank-2 tail of S (la.pyx:1195-1233):
    a_k = 8 k(k+1)(k+2)(k+4) pi - l0_{k-1} a_{k-1} - l1_{k-2} a_{k-2}, same for b_k with
    24 (k+1)(k+2) pi; one vectorised step per row, all lines and both parities at once."""
    st = Staged()
    a = st(ak, out=True, dtype=torch.float64)
    b = st(bk, out=True, dtype=torch.float64)
    L0 = st(l0, dtype=torch.float64)
    L1 = st(l1, dtype=torch.float64)
    av, outer, M, inner = _split_view(a, axis)
    bv = _split_view(b, axis)[0]
    l0v = _split_view(L0, axis)[0]
    l1v = _split_view(L1, axis)[0]
    par = torch.arange(2, dtype=torch.float64, device=a.device).view(2, 1, 1)
    av.zero_()
    bv.zero_()
    # la.pyx:1221-1233 loops to l0.shape[0]+1-3, and l0 is as long as ak: rows 2 .. M-3 are filled
    for m in sorted(set(range(min(2, M))) | set(range(2, M-2))):
        kk = 2.0*m+par
        pa = 8*kk*(kk+1)*(kk+2)*(kk+4)*np.pi
        pb = 24*(kk+1)*(kk+2)*np.pi
        ra = pa.expand(2, outer, inner).clone()
        rb = pb.expand(2, outer, inner).clone()
        if m >= 1:
            ra -= l0v[:, :, m-1, :]*av[:, :, m-1, :]
            rb -= l0v[:, :, m-1, :]*bv[:, :, m-1, :]
        if m >= 2:
            ra -= l1v[:, :, m-2, :]*av[:, :, m-2, :]
            rb -= l1v[:, :, m-2, :]*bv[:, :, m-2, :]
        av[:, :, m, :] = ra
        bv[:, :, m, :] = rb
    st.finish()


def _bih_pack(f, a0, axis, originals):
    """Packed factor table for the seven reference arrays.  Cached per caller tensor OBJECT (weak
    references, so a recycled address can never alias) and invalidated by torch's in-place version
    counters; host (numpy) arguments are staged afresh each call and always re-packed."""
    import weakref
    u0, u1, u2, l0, l1, ak, bk = f
    cacheable = all(isinstance(t, torch.Tensor) for t in originals)
    if cacheable:
        for refs, vers, ca0, cax, fac in list(_PACK_CACHE):
            if ca0 == float(a0) and cax == axis and all(r() is t for r, t in zip(refs, originals)) \
                    and vers == tuple(t._version for t in originals):
                return fac
    outer, M, inner = axis_view(tuple(u0.shape[1:]), axis)
    fac = torch.empty((7, outer, 2*M, inner), dtype=torch.float64, device=u0.device)
    fac[0] = _rows_from_split(l0, axis, 1)
    fac[1] = _rows_from_split(l1, axis, 2)
    fac[2] = _rows_from_split(u0, axis, 0)
    fac[3] = _rows_from_split(u1, axis, 0)
    fac[4] = _rows_from_split(u2, axis, 0)
    fac[5] = _rows_from_split(ak, axis, 0)*float(a0)
    fac[6] = _rows_from_split(bk, axis, 0)*float(a0)
    if cacheable:
        _PACK_CACHE[:] = [e for e in _PACK_CACHE if all(r() is not None for r in e[0])][-3:]
        _PACK_CACHE.append((tuple(weakref.ref(t) for t in originals), tuple(t._version for t in originals),
                            float(a0), axis, fac))
    return fac


def Biharmonic_Solve(b, u, u0, u1, u2, l0, l1, ak, bk, a0, axis=0):
    """Positional order of the call site chebyshev/la.py:547 (the declaration at :517 differs)."""
    st = Staged()
    ut = st(u, out=True)
    bt = ut if b is u else st(b, dtype=ut.dtype)
    f = [st(x, dtype=torch.float64) for x in (u0, u1, u2, l0, l1, ak, bk)]
    a0 = float(np.atleast_1d(host(a0)).item())
    fac = _bih_pack(f, a0, axis, (u0, u1, u2, l0, l1, ak, bk))
    outer, n_u, inner = axis_view(ut.shape, axis)
    n = fac.shape[2]
    ops.biharmonic_solve(bt, ut, fac, outer, inner, n, n_u)
    st.finish()
    return u
