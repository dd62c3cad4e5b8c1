"""`FunctionSpace` factory (shenfun/forms/arguments.py:30-418)."""
import numpy as np

from . import fourier


def FunctionSpace(N, family='Fourier', bc=None, dtype='d', quad=None, domain=None, scaled=None,
                  padding_factor=1, basis=None, dealias_direct=False, coordinates=None, **kw):
    """Return a 1-D function space.

    Same arguments as the reference factory.  Families on the device path:
    Fourier ('F'), Chebyshev ('C'), Legendre ('L'), orthogonal Jacobi ('J', alpha=, beta=) and Chebyshev of the
    second kind ('U').
    """
    fam = family.lower()
    par = dict(padding_factor=padding_factor, dealias_direct=dealias_direct)
    if fam in ('fourier', 'f'):
        if domain is not None:
            par['domain'] = tuple(float(d) for d in domain)
        if np.dtype(dtype).char in 'FDG':
            return fourier.C2C(N, **par)
        return fourier.R2C(N, **par)
    if fam in ('chebyshev', 'c'):
        from . import chebyshev
        return chebyshev.make_space(N, bc=bc, quad=quad, domain=domain, dtype=dtype, scaled=scaled,
                                    basis=basis, **par)
    if fam in ('legendre', 'l'):
        from . import legendre
        return legendre.make_space(N, bc=bc, quad=quad, domain=domain, dtype=dtype, scaled=scaled,
                                   basis=basis, **par)
    if fam in ('chebyshevu', 'u'):
        from . import chebyshevu
        if domain is not None:
            par['domain'] = tuple(float(d) for d in domain)
        return chebyshevu.Orthogonal(N, quad=quad, bc=bc, dtype=dtype, **par)
    if fam in ('jacobi', 'j'):
        from . import jacobi
        if domain is not None:
            par['domain'] = tuple(float(d) for d in domain)
        return jacobi.Orthogonal(N, quad=quad, alpha=kw.get('alpha', 0), beta=kw.get('beta', 0), bc=bc, dtype=dtype, **par)
    raise NotImplementedError(f'family {family!r} is outside the device hot path')
