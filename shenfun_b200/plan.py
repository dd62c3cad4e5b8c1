"""`FourierPlan`: the C-level N-D transform plan (`shb_plan_*` of include/shenb200.h) from Python.

The whole forward / backward transform of a Fourier tensor-product space is ONE call across the
C ABI (plan created once, per-axis launches issued by the library), which is the form a non-Python
host would use; `TensorProductSpace` drives the same axis kernels itself because it also handles
the polynomial bases and the multi-GPU exchange.
"""
import ctypes

import numpy as np
import torch

from . import _lib, ops


class FourierPlan:
    def __init__(self, shape, kinds, padded=None, axes=None, dealias_direct=False):
        """shape: modes per axis; kinds: 'C2C' / 'R2C' per axis; padded: quadrature points per axis
        (default: no padding); axes: backward transform order (default 0..ndim-1)."""
        nd = len(shape)
        padded = list(shape) if padded is None else list(padded)
        axes = list(range(nd)) if axes is None else list(axes)
        d = _lib.PlanDesc()
        d.ndim = nd
        for i in range(nd):
            d.shape[i], d.padded[i] = int(shape[i]), int(padded[i])
            d.kind[i] = 1 if str(kinds[i]).upper() == 'R2C' else 0
            d.axes[i] = int(axes[i])
        d.dealias_direct = int(bool(dealias_direct))
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        _lib.check(self._lib.shb_plan_create(ctypes.byref(d), ctypes.byref(self._h)))
        self.real = any(str(k).upper() == 'R2C' for k in kinds)
        self.physical_shape = tuple(int(p) for p in padded)
        self.spectral_shape = tuple(int(n)//2+1 if str(k).upper() == 'R2C' else int(n) for n, k in zip(shape, kinds))
        assert int(np.prod(self.physical_shape)) == self._lib.shb_plan_array_elems(self._h, 0)
        assert int(np.prod(self.spectral_shape)) == self._lib.shb_plan_array_elems(self._h, 1)
        self._work = None

    def _workspace(self, ncomp, device):
        nbytes = self._lib.shb_plan_workspace_bytes(self._h, ncomp)
        if self._work is None or self._work.numel() < nbytes or self._work.device != device:
            self._work = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)
        return self._work

    def _run(self, fn, src, dst, ncomp):
        ops._req(src, src.dtype, 'in')
        ops._req(dst, dst.dtype, 'out')
        w = self._workspace(ncomp, src.device)
        _lib.check(fn(self._h, src.data_ptr(), dst.data_ptr(), w.data_ptr(), ncomp, _lib.stream_ptr()))
        ops.LAUNCHES += len(self.physical_shape)
        return dst

    def forward(self, u, out=None, ncomp=1):
        lead = (ncomp,) if ncomp > 1 else ()
        if out is None:
            out = torch.empty(lead+self.spectral_shape, dtype=torch.complex128, device=u.device)
        assert u.dtype == (torch.float64 if self.real else torch.complex128)
        assert tuple(u.shape) == lead+self.physical_shape and tuple(out.shape) == lead+self.spectral_shape
        return self._run(self._lib.shb_plan_forward, u, out, ncomp)

    def backward(self, uh, out=None, ncomp=1):
        lead = (ncomp,) if ncomp > 1 else ()
        if out is None:
            out = torch.empty(lead+self.physical_shape, dtype=torch.float64 if self.real else torch.complex128,
                              device=uh.device)
        assert uh.dtype == torch.complex128
        assert tuple(uh.shape) == lead+self.spectral_shape and tuple(out.shape) == lead+self.physical_shape
        return self._run(self._lib.shb_plan_backward, uh, out, ncomp)

    def __del__(self):
        h = getattr(self, '_h', None)
        if h:
            self._lib.shb_plan_destroy(h)
            self._h = None
