"""mpi4py_fft.fftw surface on scipy.fft (test infrastructure only)."""
import numpy as np
import scipy.fft as sfft
from . import utilities  # noqa: F401
from .utilities import FFTW_MEASURE, FFTW_ESTIMATE, FFTW_PATIENT, FFTW_EXHAUSTIVE, \
    FFTW_DESTROY_INPUT, FFTW_PRESERVE_INPUT, FFTW_UNALIGNED, FFTW_WISDOM_ONLY

flag_dict = {key: val for key, val in vars(utilities).items() if key.startswith('FFTW_')}


def aligned(shape, n=32, dtype=np.dtype('d'), fill=None):
    a = np.empty(shape, dtype=np.dtype(dtype))
    if fill is not None:
        a.fill(fill)
    return a


def aligned_like(z, fill=None):
    return aligned(z.shape, dtype=z.dtype, fill=fill)


def export_wisdom(filename):
    pass


def import_wisdom(filename):
    pass


class _Plan:
    """Planned un-normalised transform with fixed input/output work arrays."""

    def __init__(self, func, input_array, output_array, s, axes, M):
        self._func = func
        self.input_array = input_array
        self.output_array = output_array
        self.s = s
        self.axes = axes
        self.M = M

    def get_normalization(self):
        return self.M

    def __call__(self, input_array=None, output_array=None, implicit=True,
                 normalize=False, **kw):
        # kw like normalise_idft=False are pyfftw-isms that mpi4py_fft swallows
        if input_array is not None:
            self.input_array[...] = input_array
        out = self._func(self.input_array)
        if normalize:
            out = out*self.M
        self.output_array[...] = out
        if output_array is not None:
            output_array[...] = self.output_array
            return output_array
        return self.output_array


def _sa(input_array, s, axes):
    axes = tuple(axes) if not isinstance(axes, int) else (axes,)
    axes = tuple(a % input_array.ndim for a in axes)
    return s, axes


def fftn(input_array, s=None, axes=(-1,), threads=1, flags=(FFTW_MEASURE,), output_array=None):
    s, axes = _sa(input_array, s, axes)
    if output_array is None:
        output_array = aligned_like(input_array)
    M = 1./np.prod(np.take(input_array.shape, axes))
    return _Plan(lambda u: sfft.fftn(u, axes=axes), input_array, output_array, s, axes, M)


def ifftn(input_array, s=None, axes=(-1,), threads=1, flags=(FFTW_MEASURE,), output_array=None):
    s, axes = _sa(input_array, s, axes)
    if output_array is None:
        output_array = aligned_like(input_array)
    n = np.prod(np.take(input_array.shape, axes))
    M = 1./n
    # FFTW backward c2c is un-normalised
    return _Plan(lambda u: sfft.ifftn(u, axes=axes, norm='forward'), input_array, output_array, s, axes, M)


def rfftn(input_array, s=None, axes=(-1,), threads=1, flags=(FFTW_MEASURE,), output_array=None):
    s, axes = _sa(input_array, s, axes)
    sh = list(input_array.shape)
    sh[axes[-1]] = sh[axes[-1]]//2+1
    if output_array is None:
        output_array = aligned(sh, dtype=np.result_type(input_array.dtype, np.complex64))
    M = 1./np.prod(np.take(input_array.shape, axes))
    return _Plan(lambda u: sfft.rfftn(u, axes=axes), input_array, output_array, s, axes, M)


def irfftn(input_array, s=None, axes=(-1,), threads=1, flags=(FFTW_MEASURE,), output_array=None):
    s, axes = _sa(input_array, s, axes)
    assert s is not None
    sh = list(input_array.shape)
    for a, n in zip(axes, s):
        sh[a] = n
    if output_array is None:
        output_array = aligned(sh, dtype=input_array.real.dtype)
    s = tuple(int(i) for i in s)
    M = 1./np.prod(s)
    return _Plan(lambda u: sfft.irfftn(u, s=s, axes=axes, norm='forward'), input_array, output_array, s, axes, M)


def _r2r(kindfun, input_array, s, axes, type, output_array):
    s, axes = _sa(input_array, s, axes)
    if output_array is None:
        output_array = aligned_like(input_array)
    # FFTW logical sizes: REDFT00 2(n-1), REDFT10/01 2n, RODFT00 2(n+1), RODFT10/01 2n
    M = 1.
    for a in axes:
        n = input_array.shape[a]
        if kindfun is sfft.dctn:
            M *= 1./(2*(n-1)) if type == 1 else 1./(2*n)
        else:
            M *= 1./(2*(n+1)) if type == 1 else 1./(2*n)
    return _Plan(lambda u: kindfun(u, type=type, axes=axes), input_array, output_array, s, axes, M)


def dctn(input_array, s=None, axes=(-1,), type=2, threads=1, flags=(FFTW_MEASURE,), output_array=None):
    return _r2r(sfft.dctn, input_array, s, axes, type, output_array)


def idctn(input_array, s=None, axes=(-1,), type=2, threads=1, flags=(FFTW_MEASURE,), output_array=None):
    return _r2r(sfft.dctn, input_array, s, axes, {1: 1, 2: 3, 3: 2}[type], output_array)


def dstn(input_array, s=None, axes=(-1,), type=2, threads=1, flags=(FFTW_MEASURE,), output_array=None):
    return _r2r(sfft.dstn, input_array, s, axes, type, output_array)


def idstn(input_array, s=None, axes=(-1,), type=2, threads=1, flags=(FFTW_MEASURE,), output_array=None):
    return _r2r(sfft.dstn, input_array, s, axes, {1: 1, 2: 3, 3: 2}[type], output_array)
