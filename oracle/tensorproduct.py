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
