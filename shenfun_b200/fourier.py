"""Fourier bases R2C / C2C on the device.

Host-side mirror of shenfun/fourier/bases.py (`FourierBase` :55-279, `R2C`
:283-497, `C2C` :500-660).  Transform arithmetic, 3/2-rule pad/truncate, Nyquist
fix-ups and the 1/N normalisation are one fused kernel per axis
(csrc/fft_*.cu) instead of FFTW + three numpy passes.
"""
import numpy as np
import torch

from .spectralbase import SpectralBase
from . import ops

__all__ = ('R2C', 'C2C')


class FourierBase(SpectralBase):
    def __init__(self, N, padding_factor=1., domain=(0., 2*np.pi), dtype=float,
                 dealias_direct=False, **kw):
        SpectralBase.__init__(self, N, quad='', padding_factor=padding_factor,
                              domain=domain, dtype=dtype, dealias_direct=dealias_direct)

    @staticmethod
    def family():
        return 'fourier'

    @staticmethod
    def boundary_condition():
        return 'Periodic'

    def reference_domain(self):
        return (0., 2*np.pi)

    def _ctor_kwargs(self):
        return dict(domain=self.domain, padding_factor=self.padding_factor,
                    dealias_direct=self.dealias_direct)

    def points_and_weights(self, N=None, map_true_domain=False, weighted=True):
        """fourier/bases.py:72-83"""
        if N is None:
            N = self.physical_len()
        points = np.arange(N, dtype=float)*2*np.pi/N
        if map_true_domain:
            points = self.map_true_domain(points)
        if weighted:
            return points, np.array([float(self.domain_factor())/N])
        return points, np.array([2*np.pi/N])

    def get_normalization(self):
        return 1./self.physical_len()

    def l2_norm_sq(self, i=None):
        return np.ones(self.N) if i is None else 1

    def is_orthogonal_basis(self):
        return True

    def spectral_matrix(self, d_test, d_trial, trial=None):
        """(phi_l^(d_trial), phi_k^(d_test))_w: diagonal (fourier/matrices.py:149-176,
        with the domain scaling of matrixbase.py:690-691)."""
        from .matrices import SparseMatrix
        k = self.wavenumbers(bcast=False, scaled=False, eliminate_highest_freq=False).astype(float)
        n = d_test+d_trial
        if n > 0:
            val = (1j*k)**d_trial*(-1j*k)**d_test
            if n % 2 == 0:
                val = val.real
            val = val*float(self.domain_factor())**n
        else:
            val = np.ones(len(k))
        return SparseMatrix({0: val}, (len(k), len(k)))

    def evaluate_basis_all(self, x):
        """E[p, k] = exp(i k_k X_p) on the reference domain (fourier/bases.py:87-103)."""
        X = np.atleast_1d(self.map_reference_domain(np.asarray(x, dtype=float)))
        k = np.asarray(self.wavenumbers(bcast=False), dtype=float)
        return np.exp(1j*np.outer(X, k))


class R2C(FourierBase):
    """Real-to-complex Fourier space (fourier/bases.py:283-497)."""

    def __init__(self, N, padding_factor=1., domain=(0., 2*np.pi), dealias_direct=False, **kw):
        FourierBase.__init__(self, N, padding_factor=padding_factor, domain=domain,
                             dtype=float, dealias_direct=dealias_direct)

    @staticmethod
    def short_name():
        return 'R2C'

    def wavenumbers(self, bcast=True, scaled=False, eliminate_highest_freq=False):
        """fourier/bases.py:334-342"""
        k = np.fft.rfftfreq(self.N, 1./self.N).astype(int)
        if self.N % 2 == 0 and eliminate_highest_freq:
            k[-1] = 0
        if scaled:
            k = k*float(self.domain_factor())
        return self.broadcast_to_ndims(k) if bcast else k

    def get_mask_nyquist(self, bcast=True):
        """fourier/bases.py:344-362"""
        if self.N % 2 != 0:
            return None
        f = np.ones(self.N//2+1, dtype=int)
        f[-1] = 0
        return self.broadcast_to_ndims(f) if bcast else f

    def slice(self):
        return slice(0, self.N//2+1)

    def spectral_len(self):
        return self.N//2+1

    def spectral_dtype_for(self, physical_dtype):
        return torch.complex128

    def _forward_axis(self, u, out, outer, inner, scalar_product=False):
        n = self.physical_len()
        ops.fft_r2c_forward(u, out, outer, inner, n, self.N, 1./n)

    def _backward_axis(self, uh, out, outer, inner):
        ops.fft_c2r_backward(uh, out, outer, inner, self.physical_len(), self.N, 1.0,
                             self.dealias_direct)


class C2C(FourierBase):
    """Complex-to-complex Fourier space (fourier/bases.py:500-660)."""

    def __init__(self, N, padding_factor=1., domain=(0., 2*np.pi), dealias_direct=False, **kw):
        FourierBase.__init__(self, N, padding_factor=padding_factor, domain=domain,
                             dtype=complex, dealias_direct=dealias_direct)

    @staticmethod
    def short_name():
        return 'C2C'

    def wavenumbers(self, bcast=True, scaled=False, eliminate_highest_freq=False):
        """fourier/bases.py:535-543"""
        k = np.fft.fftfreq(self.N, 1./self.N).astype(int)
        if self.N % 2 == 0 and eliminate_highest_freq:
            k[self.N//2] = 0
        if scaled:
            k = k*float(self.domain_factor())
        return self.broadcast_to_ndims(k) if bcast else k

    def get_mask_nyquist(self, bcast=True):
        if self.N % 2 != 0:
            return None
        f = np.ones(self.N, dtype=int)
        f[self.N//2] = 0
        return self.broadcast_to_ndims(f) if bcast else f

    def spectral_dtype_for(self, physical_dtype):
        return torch.complex128

    def inplace_ok(self):
        return not self.is_padded

    supports_chunked = True    # axis kernels can address the all-to-all send / receive layout

    def _forward_axis(self, u, out, outer, inner, scalar_product=False, in_chunk=0, out_chunk=0, peers=None):
        n = self.physical_len()
        if peers is not None:
            ops.fft_c2c_peer(False, u, peers, outer, inner, n, self.N, 1./n, False, in_chunk, out_chunk)
        elif in_chunk or out_chunk:
            ops.fft_c2c_chunked(False, u, out, outer, inner, n, self.N, 1./n, False, in_chunk, out_chunk)
        else:
            ops.fft_c2c_forward(u, out, outer, inner, n, self.N, 1./n)

    def _backward_axis(self, uh, out, outer, inner, in_chunk=0, out_chunk=0, peers=None):
        if peers is not None:
            ops.fft_c2c_peer(True, uh, peers, outer, inner, self.physical_len(), self.N, 1.0,
                             self.dealias_direct, in_chunk, out_chunk)
        elif in_chunk or out_chunk:
            ops.fft_c2c_chunked(True, uh, out, outer, inner, self.physical_len(), self.N, 1.0,
                                self.dealias_direct, in_chunk, out_chunk)
        else:
            ops.fft_c2c_backward(uh, out, outer, inner, self.physical_len(), self.N, 1.0,
                                 self.dealias_direct)
