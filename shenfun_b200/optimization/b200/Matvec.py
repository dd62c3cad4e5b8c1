"""Matrix-vector products under the reference's kernel names
(optimization/cython/Matvec.pyx:452-537 banded, :697-768 Helmholtz, :908-1000 biharmonic)."""
import numpy as np
import torch

from ... import ops
from ._args import Staged, axis_view, host
from .la import _lines, _helm_tables, _bih_tab


def _banded(v, b, axis, offsets, diags):
    st = Staged()
    bt = st(b, out=True)
    vt = st(v, dtype=bt.dtype)
    outer, n_u, inner = axis_view(vt.shape, axis)
    n = int(np.size(host(diags[len(diags)//2])))
    tab = np.zeros((len(offsets), n))
    for r, (off, d) in enumerate(zip(offsets, diags)):
        d = np.real(host(d)).astype(float)
        if off >= 0:
            tab[r, :n-off] = d[:n-off]
        else:
            tab[r, -off:] = d[:n+off]
    tab = ops.as_device_tensor(tab)
    ops.banded_matvec(vt, bt, offsets, [tab], [None], outer, inner, n, n, n_u, bt.shape[axis], False)
    st.finish()
    return b


def Tridiagonal_matvec(v, b, axis, ld, dd, ud):
    """b = M v, M with offsets (-2, 0, 2) (Matvec.pyx:452-500)."""
    return _banded(v, b, axis, (-2, 0, 2), (ld, dd, ud))


def Pentadiagonal_matvec(v, b, axis, ldd, ld, dd, ud, udd):
    """b = M v, M with offsets (-4, -2, 0, 2, 4) (Matvec.pyx:502-537)."""
    return _banded(v, b, axis, (-4, -2, 0, 2, 4), (ldd, ld, dd, ud, udd))


def Helmholtz_matvec(v, b, alfa, beta, A, B, axis=0):
    """b = (alfa*A + beta*B) v along `axis`, A with constant rows beyond +4 (Matvec.pyx:697-768)."""
    st = Staged()
    bt = st(b, out=True)
    vt = st(v, dtype=bt.dtype)
    outer, n_u, inner = axis_view(vt.shape, axis)
    n = int(np.size(host(A[0])))
    al = _lines(alfa, vt.shape, axis, vt.device)
    be = _lines(beta, vt.shape, axis, vt.device)
    tabs = _helm_tables(A, B, n, False)
    ops.helmholtz_matvec(vt, bt, tabs, al, be, outer, inner, n, n_u)
    st.finish()
    return b


def Biharmonic_matvec(v, b, a0, alfa, beta, sii, siu, siuu, ail, aii, aiu, bill, bil, bii, biu, biuu, axis=0):
    """b = (a0*S + alfa*A + beta*B) v on the ShenBiharmonic basis (Matvec.pyx:908-1000)."""
    st = Staged()
    bt = st(b, out=True)
    vt = st(v, dtype=bt.dtype)
    outer, n_u, inner = axis_view(vt.shape, axis)
    n = int(np.size(host(sii)))
    a = torch.full((outer*inner,), float(np.atleast_1d(host(a0)).item()), dtype=torch.float64, device=vt.device)
    al = _lines(alfa, vt.shape, axis, vt.device)
    be = _lines(beta, vt.shape, axis, vt.device)
    tab = _bih_tab(n, sii, siu, siuu, ail, aii, aiu, bill, bil, bii, biu, biuu)
    ops.biharmonic_matvec(vt, bt, tab, a, al, be, outer, inner, n, n_u)
    st.finish()
    return b
