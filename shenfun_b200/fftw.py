"""`mpi4py_fft.fftw`-shaped plan objects on the device (boundary 2).

The reference's bases build their per-axis plans through this surface
(shenfun/fourier/bases.py:218-279, shenfun/chebyshev/bases.py:407-465):

    U = fftw.aligned(shape, dtype)                         -> work array
    fwd = fftw.rfftn(U, s=s, axes=axes, threads=, flags=)  -> plan, fwd.output_array
    bck = fftw.irfftn(V, s=s, axes=axes, output_array=U)
    fwd.get_normalization(); fwd(input_array=None, output_array=None, normalize=False)

Same names and semantics here (FFTW conventions: every transform un-normalised, r2c keeps
n//2+1 modes, REDFT10 / REDFT01 for `dctn` types 2 / 3), with CUDA tensors as the work arrays
and the axis passes running in libshenb200.  `threads` and `flags` are accepted and ignored.
"""
import numpy as np
import torch

from . import ops

FFTW_MEASURE, FFTW_DESTROY_INPUT, FFTW_UNALIGNED, FFTW_EXHAUSTIVE = 0, 1, 2, 8
FFTW_PRESERVE_INPUT, FFTW_PATIENT, FFTW_ESTIMATE, FFTW_WISDOM_ONLY = 16, 32, 64, 2097152
flag_dict = {k: v for k, v in list(globals().items()) if k.startswith('FFTW_')}

_DT = {np.dtype('d'): torch.float64, np.dtype('D'): torch.complex128}


def _torch_dtype(dtype):
    if isinstance(dtype, torch.dtype):
        if dtype not in (torch.float64, torch.complex128):
            raise TypeError('fp64 / complex128 only')
        return dtype
    dt = np.dtype(dtype)
    if dt not in _DT:
        raise TypeError(f'{dt}: the device build is fp64 / complex128 only')
    return _DT[dt]


def aligned(shape, n=32, dtype=np.dtype('d'), fill=None):
    """Work array (device memory is always aligned far beyond `n`)."""
    if not torch.cuda.is_available():
        raise ops._lib.ShenB200Error('no CUDA device: shenfun_b200 has no CPU fallback')
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    a = torch.empty(shape, dtype=_torch_dtype(dtype), device='cuda')
    if fill is not None:
        a.fill_(fill)
    return a


def aligned_like(z, fill=None):
    return aligned(tuple(z.shape), dtype=z.dtype, fill=fill)


def export_wisdom(filename):
    pass


def import_wisdom(filename):
    pass


def _view(shape, axis):
    outer = int(np.prod(shape[:axis], dtype=np.int64))
    inner = int(np.prod(shape[axis+1:], dtype=np.int64))
    return outer, inner


class _Plan:
    """Planned transform over fixed work arrays: a list of axis passes, ping-ponging through
    one scratch array when more than one axis is transformed."""

    def __init__(self, kind, input_array, output_array, s, axes, M, dct_type=None):
        self.kind, self.s, self.axes, self.M = kind, tuple(s), tuple(axes), M
        self.input_array = ops.plain(input_array)
        self.output_array = ops.plain(output_array)
        self._dct_type = dct_type
        self._tmp = None
        self._sten = None

    def get_normalization(self):
        return self.M

    # -- one axis pass ------------------------------------------------------------------
    def _pass(self, step, src, dst):
        what, axis, n = step
        if what == 'r2c':
            outer, inner = _view(src.shape, axis)
            ops.fft_r2c_forward(src, dst, outer, inner, n, n, 1.0)
        elif what == 'c2r':
            outer, inner = _view(dst.shape, axis)
            ops.fft_c2r_backward(src, dst, outer, inner, n, n, 1.0)
        elif what == 'fwd':
            outer, inner = _view(src.shape, axis)
            ops.fft_c2c_forward(src, dst, outer, inner, n, n, 1.0)
        elif what == 'bck':
            outer, inner = _view(src.shape, axis)
            ops.fft_c2c_backward(src, dst, outer, inner, n, n, 1.0)
        elif what == 'dct2':
            outer, inner = _view(src.shape, axis)
            ops.cheb_scalar_product(src, dst, outer, inner, n, n, 1.0, n, [0], [None])
        elif what == 'dct3':
            outer, inner = _view(src.shape, axis)
            if self._sten is None or self._sten.shape[0] != n:
                w = np.full(n, 2.0)
                w[0] = 1.0                  # REDFT01: X_0 + 2 sum_{j>=1} X_j cos(pi j (k+1/2)/n)
                self._sten = ops.as_device_tensor(w)
            ops.cheb_backward(src, dst, outer, inner, n, n, n, [0], [self._sten])

    def _steps(self):
        k, ax, s = self.kind, self.axes, self.s
        if k == 'fftn':
            return [('fwd', a, n) for a, n in zip(reversed(ax), reversed(s))]
        if k == 'ifftn':
            return [('bck', a, n) for a, n in zip(ax, s)]
        if k == 'rfftn':
            return [('r2c', ax[-1], s[-1])]+[('fwd', a, n) for a, n in zip(reversed(ax[:-1]), reversed(s[:-1]))]
        if k == 'irfftn':
            return [('bck', a, n) for a, n in zip(ax[:-1], s[:-1])]+[('c2r', ax[-1], s[-1])]
        if k == 'dctn':
            return [('dct2' if self._dct_type == 2 else 'dct3', a, n) for a, n in zip(ax, s)]
        raise AssertionError(k)

    def __call__(self, input_array=None, output_array=None, implicit=True, normalize=False, **kw):
        if input_array is not None:
            self.input_array.copy_(ops.as_device_tensor(input_array))
        steps = self._steps()
        cur = self.input_array
        for i, step in enumerate(steps):
            remaining = len(steps)-1-i
            if remaining == 0:
                dst = self.output_array
            else:
                # intermediate results live in spectral-shaped scratch; never overwrite `cur`
                shape = self.output_array.shape if self.kind != 'irfftn' else self.input_array.shape
                dtype = self.output_array.dtype if self.kind != 'irfftn' else self.input_array.dtype
                if self._tmp is None:
                    self._tmp = [torch.empty(shape, dtype=dtype, device=cur.device) for _ in range(2)]
                dst = self._tmp[i % 2]
            self._pass(step, cur, dst)
            cur = dst
        if normalize:
            self.output_array.mul_(self.M)
        if output_array is not None:
            ops.plain(output_array).copy_(self.output_array)
            return output_array
        return self.output_array


def _prep(input_array, s, axes):
    input_array = ops.plain(input_array)
    if not (isinstance(input_array, torch.Tensor) and input_array.is_cuda and input_array.is_contiguous()):
        raise ops._lib.ShenB200Error('plans are made over contiguous CUDA work arrays (use fftw.aligned)')
    axes = (axes,) if isinstance(axes, int) else tuple(axes)
    axes = tuple(a % input_array.ndim for a in axes)
    return input_array, axes


def fftn(input_array, s=None, axes=(-1,), threads=1, flags=(FFTW_MEASURE,), output_array=None):
    U, axes = _prep(input_array, s, axes)
    s = [U.shape[a] for a in axes]
    V = aligned_like(U) if output_array is None else output_array
    return _Plan('fftn', U, V, s, axes, 1.0/np.prod(s))


def ifftn(input_array, s=None, axes=(-1,), threads=1, flags=(FFTW_MEASURE,), output_array=None):
    U, axes = _prep(input_array, s, axes)
    s = [U.shape[a] for a in axes]
    V = aligned_like(U) if output_array is None else output_array
    return _Plan('ifftn', U, V, s, axes, 1.0/np.prod(s))


def rfftn(input_array, s=None, axes=(-1,), threads=1, flags=(FFTW_MEASURE,), output_array=None):
    U, axes = _prep(input_array, s, axes)
    s = [U.shape[a] for a in axes]
    if output_array is None:
        shape = list(U.shape)
        shape[axes[-1]] = s[-1]//2+1
        output_array = aligned(shape, dtype=torch.complex128)
    return _Plan('rfftn', U, output_array, s, axes, 1.0/np.prod(s))


def irfftn(input_array, s=None, axes=(-1,), threads=1, flags=(FFTW_MEASURE,), output_array=None):
    V, axes = _prep(input_array, s, axes)
    if s is None:
        s = [V.shape[a] for a in axes]
        s[-1] = 2*(s[-1]-1)
        if output_array is not None:
            s[-1] = ops.plain(output_array).shape[axes[-1]]
    s = list(s)
    if output_array is None:
        shape = list(V.shape)
        shape[axes[-1]] = s[-1]
        output_array = aligned(shape, dtype=torch.float64)
    return _Plan('irfftn', V, output_array, s, axes, 1.0/np.prod(s))


def dctn(input_array, s=None, axes=(-1,), type=2, threads=1, flags=(FFTW_MEASURE,), output_array=None):
    """REDFT10 (type 2) / REDFT01 (type 3), real or complex work arrays (the reference wraps
    complex data in DCTWrap, chebyshev/bases.py:100-115; here both parts go in one pass)."""
    if type not in (2, 3):
        raise NotImplementedError('dctn: types 2 and 3 (Gauss-Chebyshev points) only')
    U, axes = _prep(input_array, s, axes)
    s = [U.shape[a] for a in axes]
    V = aligned_like(U) if output_array is None else output_array
    return _Plan('dctn', U, V, s, axes, 1.0/np.prod([2*n for n in s]), dct_type=type)
