"""Argument staging for the by-name kernels."""
import numpy as np
import torch

from ... import ops
from ..._lib import ShenB200Error


class Staged:
    """Device views of a call's arguments.  CUDA tensors are used as they are; numpy arrays
    are copied to the device and, when marked `out`, copied back in place by `finish()`."""

    def __init__(self):
        self._wb = []

    def __call__(self, x, out=False, dtype=None):
        if isinstance(x, torch.Tensor):
            t = ops.plain(x)
            if not t.is_cuda:
                raise ShenB200Error('expected a CUDA tensor or a numpy array (no CPU fallback)')
            if not t.is_contiguous():
                raise ValueError('device arguments must be C-contiguous')
            if dtype is not None and t.dtype != dtype:
                if out:
                    raise TypeError(f'expected {dtype}, got {t.dtype}')
                t = t.to(dtype)
            return t
        a = np.asarray(x)
        if dtype is not None:
            a = a.astype({torch.float64: np.float64, torch.complex128: np.complex128}[dtype], copy=False)
        t = ops.as_device_tensor(np.ascontiguousarray(a))
        if out:
            self._wb.append((x, t))
        return t

    def finish(self):
        for x, t in self._wb:
            np.copyto(np.asarray(x), t.cpu().numpy().reshape(np.shape(x)))
        self._wb = []


def axis_view(shape, axis):
    outer = int(np.prod(shape[:axis], dtype=np.int64))
    inner = int(np.prod(shape[axis+1:], dtype=np.int64))
    return outer, int(shape[axis]), inner


def host(x):
    """Small coefficient arrays as float64 numpy on the host."""
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)
