"""Stand-in for the reference's Cython wrapper shenfun/legendre/fastgl/fastgl_wrap.pyx:
same functions, backed by the reference's own fastgl.cpp compiled into
oracle/_ref/libfastgl.so (oracle/build_ref.sh).  Test infrastructure only."""
import ctypes
import os
import numpy as np

_so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), '_ref', 'libfastgl.so')
_lib = ctypes.CDLL(_so)
_lib.fastgl_pairs.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]


def leggauss(N):
    x = np.zeros(N)
    w = np.zeros(N)
    _lib.fastgl_pairs(int(N), x.ctypes.data, w.ctypes.data)
    return x, w


def getGLPair(N, k):
    x, w = leggauss(N)
    return x[N-k], w[N-k]
