"""Legendre <-> Chebyshev coefficient conversion and the discrete Legendre transform on the device.

Host-side mirror of `shenfun/legendre/dlt.py`: `leg2cheb` (:353-421), `cheb2leg` (:423-474), the
callable classes `Leg2Cheb` / `Cheb2Leg` (compiled in the reference, `optimization/cython/transforms.pyx:
528-660`), their FMM variants `FMMLeg2Cheb` (:976-1078) / `FMMCheb2Leg` (:1080-1189) and `DLT` (:20-206).

The connection matrices (Alpert & Rokhlin 1991) are upper triangular with a checkerboard pattern -- even
rows couple to even columns only -- so each conversion is two triangular products of half the size.  Here
they run on the fp64 tensor pipe (`shb_gemm_axis_parity`, checkerboard mode: N^2/4 multiply-adds per line,
zero k-tiles skipped) with the matrix built once per length on the host.  The reference compresses the
smooth off-diagonal part hierarchically (FMM) to reach O(N log N); at DMMA rates the direct triangular
product is the faster choice up to N of a few thousand, so `FMMLeg2Cheb` /
`FMMCheb2Leg` are the same operators under the reference's names and give the same numbers to rounding.
`DLT` is the Legendre transform of `shenfun_b200.legendre` (parity-split tensor-core contraction) behind
the reference's plan-like signature.
"""
import numpy as np
import torch
from scipy.special import gammaln

from . import ops
from .optimization.b200._args import Staged, axis_view

__all__ = ('Lambda', 'leg2cheb', 'cheb2leg', 'Leg2Cheb', 'Cheb2Leg', 'FMMLeg2Cheb', 'FMMCheb2Leg', 'DLT')

_MAT = {}


def Lambda(z):
    """Gamma(z + 1/2) / Gamma(z + 1) (dlt.py:344; transforms.pyx:482-526)."""
    z = np.asarray(z, dtype=float)
    return np.exp(gammaln(z+0.5)-gammaln(z+1))


def _l2c_matrix(N):
    """M[n, m] = (2/pi) Lambda((m-n)/2) Lambda((m+n)/2) for m >= n, m - n even; row 0 halved
    (dlt.py:396-410: `cc[:(N-n)] += a[n//2]*a[n//2:N-n//2]*cl[n:]`, `cc /= ck`)."""
    a = Lambda(np.arange(N, dtype=float))
    M = np.zeros((N, N))
    for n in range(0, N, 2):
        i = np.arange(N-n)
        M[i, i+n] = a[n//2]*a[n//2:N-n//2]
    M *= 2/np.pi
    M[0] *= 0.5
    return M


def _c2l_matrix(N):
    """Legendre coefficients from Chebyshev ones (dlt.py:459-470)."""
    k = np.arange(N, dtype=float)
    k[0] = 1
    a = 1/(2*Lambda(k)*k*(k+0.5))
    a[0] = 2/np.sqrt(np.pi)
    kf = k.copy()                       # vn = cc * k with k[0] = 1
    L = np.zeros((N, N))
    i = np.arange(N)
    L[i, i] = np.sqrt(np.pi)*a*kf
    for n in range(2, N, 2):
        dn = float(Lambda((n-2)/2))/n
        i = np.arange(N-n)
        L[i, i+n] = -dn*a[n//2:N-n//2]*kf[n:]
    k[0] = 0
    L *= (k+0.5)[:, None]
    return L


def _matrix(kind, N):
    key = (kind, N, torch.cuda.current_device() if torch.cuda.is_available() else -1)
    m = _MAT.get(key)
    if m is None:
        if kind == 'l2c':
            a = _l2c_matrix(N)
        elif kind == 'l2cT':
            # transpose=True (dlt.py:412-415): cx = cl/ck first, i.e. M^T with the scaling on the input side
            a = np.ascontiguousarray(_l2c_matrix(N).T)
        else:
            a = _c2l_matrix(N)
        if len(_MAT) > 32:
            _MAT.clear()
        m = _MAT[key] = ops.as_device_tensor(np.ascontiguousarray(a))
    return m


def _apply(kind, inp, out, axis):
    st = Staged()
    o = st(out, out=True)
    i = st(inp, dtype=o.dtype)
    if i.data_ptr() == o.data_ptr():
        i = i.clone()
    outer, N, inner = axis_view(o.shape, axis)
    assert i.shape == o.shape
    ops.gemm_axis(_matrix(kind, N), i, o, outer, inner)
    st.finish()
    return out


def _like(x):
    return torch.empty_like(ops.plain(x)) if isinstance(x, torch.Tensor) else np.zeros_like(x)


def leg2cheb(cl, cc=None, axis=0, transpose=False):
    """Chebyshev coefficients from Legendre coefficients along `axis` (dlt.py:353-421);
    `transpose=True` applies M^T (used by the reference's forward DLT)."""
    if cc is None:
        cc = _like(cl)
    return _apply('l2cT' if transpose else 'l2c', cl, cc, axis)


def cheb2leg(cc, cl=None, axis=0):
    """Legendre coefficients from Chebyshev coefficients along `axis` (dlt.py:423-474)."""
    if cl is None:
        cl = _like(cc)
    return _apply('c2l', cc, cl, axis)


class Leg2Cheb:
    """Callable conversion object with the reference's signature
    (`FMMLeg2Cheb(input, output_array=None, domains=None, levels=None, maxs=100, axis=0, use_direct=-1)`,
    dlt.py:1037-1048; `__call__(input_array, output_array=None, transpose=False)`, :1050-1086).  The
    hierarchy parameters are accepted and ignored: the product is direct on the tensor cores."""

    def __init__(self, input=None, output_array=None, domains=None, levels=None, diagonals=16, maxs=100,
                 axis=0, use_direct=-1, **kw):
        self.axis = axis
        self._output_array = output_array

    def __call__(self, input_array, output_array=None, transpose=False):
        out = output_array if output_array is not None else self._output_array
        if out is None or tuple(np.shape(out)) != tuple(np.shape(input_array)):
            out = _like(input_array)
        return leg2cheb(input_array, out, axis=self.axis, transpose=transpose)


class Cheb2Leg:
    """`FMMCheb2Leg` (dlt.py:1080-1189) as a direct product; same call signature."""

    def __init__(self, input=None, output_array=None, domains=None, levels=None, diagonals=16, maxs=100,
                 axis=0, use_direct=-1, **kw):
        self.axis = axis
        self._output_array = output_array

    def __call__(self, input_array, output_array=None):
        out = output_array if output_array is not None else self._output_array
        if out is None or tuple(np.shape(out)) != tuple(np.shape(input_array)):
            out = _like(input_array)
        return cheb2leg(input_array, out, axis=self.axis)


FMMLeg2Cheb = Leg2Cheb
FMMCheb2Leg = Cheb2Leg


class DLT:
    """Discrete Legendre transform with the reference's plan-like signature (dlt.py:20-206):
    `DLT(input_array, axes=(-1,), kind='backward' | 'forward' | 'scalar product', output_array=None)`;
    calling it transforms `input_array` (or the array given at call time) along the axis on N
    Legendre-Gauss points.  Results equal the reference's FFT-based DLT to rounding; the arithmetic is
    the parity-split contraction of `shenfun_b200.legendre.Orthogonal`."""

    def __init__(self, input_array, s=None, axes=(-1,), threads=1, kind='backward', flags=None, output_array=None):
        from .legendre import Orthogonal
        axis = axes[-1] if isinstance(axes, (tuple, list)) else int(axes)
        assert kind in ('forward', 'scalar product', 'backward')
        self.kind = kind
        self.axis = axis % np.ndim(input_array)
        self.N = np.shape(input_array)[self.axis]
        self.space = Orthogonal(self.N)
        self.input_array = input_array
        self.output_array = output_array if output_array is not None else _like(input_array)

    def __call__(self, input_array=None, output_array=None, **kw):
        inp = self.input_array if input_array is None else input_array
        out = self.output_array if output_array is None else output_array
        st = Staged()
        o = st(out, out=True)
        i = st(inp, dtype=o.dtype)
        if i.data_ptr() == o.data_ptr():
            i = i.clone()
        outer, N, inner = axis_view(o.shape, self.axis)
        if self.kind == 'backward':
            self.space._backward_axis(i, o, outer, inner)
        elif self.kind == 'scalar product':
            self.space._scalar_product_axis(i, o, outer, inner)
        else:
            self.space._forward_axis(i, o, outer, inner)
        st.finish()
        return out
