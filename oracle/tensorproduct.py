"""Oracle: N-D tensor-product forward / backward transforms, one axis at a time.

Restates shenfun/tensorproductspace.py:205-259 (`ForwardTransform.__call__`:
last axis in `axes` is transformed first) and :46-96 (`BackwardTransform`), for a
single process (transfers are identity).  Axis descriptors are tuples
    ('R2C'|'C2C', N, n_phys)                    Fourier
    ('CHEB', N, n_phys, spec)                   see oracle/chebyshev.py
    ('LEG',  N, n_phys, spec)                   see oracle/legendre.py
"""
import numpy as np
from . import fourier


def _fwd_axis(u, axis, d, scalar_product=False):
    kind = d[0]
    if kind == 'R2C':
        return fourier.r2c_forward(u, axis, d[1], d[2])
    if kind == 'C2C':
        return fourier.c2c_forward(u, axis, d[1], d[2])
    if kind == 'CHEB':
        from . import chebyshev
        return chebyshev.forward_axis(u, axis, d[3], scalar_product=scalar_product)
    if kind == 'LEG':
        from . import legendre
        return legendre.forward_axis(u, axis, d[3], scalar_product=scalar_product)
    raise ValueError(kind)


def _bwd_axis(uh, axis, d):
    kind = d[0]
    if kind == 'R2C':
        return fourier.c2r_backward(uh, axis, d[1], d[2])
    if kind == 'C2C':
        return fourier.c2c_backward(uh, axis, d[2])
    if kind == 'CHEB':
        from . import chebyshev
        return chebyshev.backward_axis(uh, axis, d[3])
    if kind == 'LEG':
        from . import legendre
        return legendre.backward_axis(uh, axis, d[3])
    raise ValueError(kind)


def forward(u, descs, axes=None, scalar_product=False, lead=0):
    """physical -> spectral; `lead` leading component axes are left alone."""
    nd = len(descs)
    axes = tuple(range(nd)) if axes is None else axes
    out = u
    for ax in reversed(axes):
        out = _fwd_axis(out, ax+lead, descs[ax], scalar_product)
    return out


def backward(uh, descs, axes=None, lead=0):
    nd = len(descs)
    axes = tuple(range(nd)) if axes is None else axes
    out = uh
    for ax in axes:
        out = _bwd_axis(out, ax+lead, descs[ax])
    return out


def fourier_poisson3d(N):
    """demo/fourier_poisson3D.py:19-45 restated: u = cos(4x) + sin(4y) + sin(6z) on (0, 2 pi)^3,
    f = div grad u on the quadrature mesh x_j = 2 pi j / N (fourier/bases.py:76).  The Fourier
    scalar product carries the 1/N normalisation (fourier/bases.py:157-162), so f_hat = (v, f) equals
    forward(f) and A = (v, div grad u) = -|k|^2; u_hat = f_hat / A with the (0,0,0) mode pinned to 0
    (matrixbase.py:1418-1424).  Returns (f_hat, u_hat)."""
    import numpy as np
    xs = [2*np.pi*np.arange(n)/n for n in N]
    X, Y, Z = np.meshgrid(*xs, indexing='ij')
    f = -16*np.cos(4*X)-16*np.sin(4*Y)-36*np.sin(6*Z)
    d = [('C2C', N[0], N[0]), ('C2C', N[1], N[1]), ('R2C', N[2], N[2])]
    f_hat = forward(f, d, scalar_product=True)
    k0 = np.fft.fftfreq(N[0], 1./N[0])[:, None, None]
    k1 = np.fft.fftfreq(N[1], 1./N[1])[None, :, None]
    k2 = np.fft.rfftfreq(N[2], 1./N[2])[None, None, :]
    A = -(k0**2+k1**2+k2**2)
    with np.errstate(divide='ignore', invalid='ignore'):
        u_hat = np.where(A == 0, 0, f_hat/A)
    return f_hat, u_hat
