"""Thin Python wrappers over the C ABI: torch tensors in, kernels launched on
torch's current stream.  Every function fails loudly when libshenb200.so or a
CUDA device is missing -- there is no eager/CPU fallback on the product path.
"""
import numpy as np
import torch

from . import _lib

LAUNCHES = 0   # number of libshenb200 kernel launches issued (bench.py reads it)


def plain(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) else t


def as_device_tensor(x):
    if isinstance(x, torch.Tensor):
        t = plain(x)
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise _lib.ShenB200Error('no CUDA device: shenfun_b200 has no CPU fallback')
        t = t.cuda()
    return t


def _req(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.ShenB200Error(f'{name}: expected a CUDA tensor (no CPU fallback)')
    if t.dtype != dtype:
        raise TypeError(f'{name}: expected {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise ValueError(f'{name}: tensor must be C-contiguous')
    return t.data_ptr()


def _call(fn_name, *args):
    global LAUNCHES
    L = _lib.load()
    _lib.check(getattr(L, fn_name)(*args))
    LAUNCHES += 1


C128 = torch.complex128
F64 = torch.float64


def fft_c2c_forward(u, out, outer, inner, n_fft, N, scale):
    _call('shb_fft_c2c_forward', _req(u, C128, 'in'), _req(out, C128, 'out'),
          outer, inner, n_fft, N, scale, _lib.stream_ptr())


def fft_c2c_backward(uh, out, outer, inner, n_fft, N, scale=1.0, dealias_direct=False):
    _call('shb_fft_c2c_backward', _req(uh, C128, 'in'), _req(out, C128, 'out'),
          outer, inner, n_fft, N, scale, int(dealias_direct), _lib.stream_ptr())


def fft_r2c_forward(u, out, outer, inner, n_fft, N, scale):
    _call('shb_fft_r2c_forward', _req(u, F64, 'in'), _req(out, C128, 'out'),
          outer, inner, n_fft, N, scale, _lib.stream_ptr())


def fft_c2r_backward(uh, out, outer, inner, n_fft, N, scale=1.0, dealias_direct=False):
    _call('shb_fft_c2r_backward', _req(uh, C128, 'in'), _req(out, F64, 'out'),
          outer, inner, n_fft, N, scale, int(dealias_direct), _lib.stream_ptr())
