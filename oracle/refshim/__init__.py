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
    sys.modules['shenfun.optimization.cython'] = stub
    import shenfun
    shenfun.config['optimization']['mode'] = 'numba'
    return shenfun
