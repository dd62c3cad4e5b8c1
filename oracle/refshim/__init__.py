"""Activate the serial stand-ins so that /root/reference/shenfun imports here.

Besides mpi4py / mpi4py_fft this injects
  * shenfun.legendre.fastgl.fastgl_wrap -> ctypes wrapper of the reference's own
    fastgl.cpp (oracle/_ref/libfastgl.so, built by oracle/build_ref.sh);
  * shenfun.optimization.cython -> a stub that only carries placeholders for the
    FMM classes `Leg2Cheb` / `Cheb2Leg` that legendre/dlt.py instantiates while
    planning (the FMM 'fast' Legendre kind is out of scope);
    kernels are taken from the reference's numba module (optimization.mode='numba').
"""
import os
import sys
import types


class _Unavailable:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("serial shim: the reference's compiled FMM Legendre transform is not built")


def activate(reference='/root/reference'):
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    if reference not in sys.path:
        sys.path.insert(0, reference)
    import fastgl_wrap
    sys.modules['shenfun.legendre.fastgl.fastgl_wrap'] = fastgl_wrap
    stub = types.ModuleType('shenfun.optimization.cython')
    stub.Leg2Cheb = _Unavailable
    stub.Cheb2Leg = _Unavailable
    try:
        # Gamma(z+1/2)/Gamma(z+1) from the reference's own compiled transforms.pyx when oracle/_ref has it,
        # else the closed form the reference keeps as a comment (legendre/dlt.py:344)
        from oracle import refcy
        stub.Lambda = refcy.load('transforms').Lambda
    except Exception:  # noqa: BLE001
        import numpy as _np
        from scipy.special import gammaln as _gl
        stub.Lambda = lambda z: _np.exp(_gl(_np.asarray(z, dtype=float)+0.5)-_gl(_np.asarray(z, dtype=float)+1))
    sys.modules['shenfun.optimization.cython'] = stub
    import numpy as np
    if not hasattr(np, 'asscalar'):
        # removed in numpy 1.23; the reference's numba 1-D Helmholtz still calls it
        # (shenfun/optimization/numba/helmholtz.py:16)
        np.asscalar = lambda a: np.asarray(a).item()
    import shenfun
    shenfun.config['optimization']['mode'] = 'numba'
    _route_matvecs_away_from_cython(shenfun)
    return shenfun


def _route_matvecs_away_from_cython(shenfun):
    """Several hard-coded matrices default to format='cython' in their `matvec`
    (e.g. shenfun/chebyshev/matrices.py:966-990), i.e. to the compiled Matvec.pyx that is not
    built here.  Route the default to the reference's own alternative implementations of the same
    product: the numpy 'self' branch where the class has one, else the scipy 'csr' base class."""
    import functools
    import inspect
    from shenfun.chebyshev import matrices as cm
    from shenfun.legendre import matrices as lm
    for mod in (cm, lm):
        for name, cls in inspect.getmembers(mod, inspect.isclass):
            if 'matvec' not in cls.__dict__:
                continue
            orig = cls.__dict__['matvec']

            def wrapped(self, v, c, format=None, axis=0, _orig=orig):
                if format is None or format == 'cython':
                    format = 'self' if 'self' in getattr(self, '_matvec_methods', []) else 'csr'
                return _orig(self, v, c, format=format, axis=axis)
            setattr(cls, 'matvec', functools.wraps(orig)(wrapped))
