"""Host-side 1-D function-space base class.

Mirrors the part of shenfun/spectralbase.py that the tensor-product hot path
touches: sizes and slices (:602-700), domain mapping (:845-900), the
forward = scalar_product -> inverse-mass / backward = pad -> evaluate skeleton
(:200-313).  All array work is delegated to device kernels through
`shenfun_b200.ops`; there are no numpy fallbacks.
"""
import copy
import numpy as np
import torch

from . import ops


class Domain(tuple):
    def __new__(cls, a, b):
        return super().__new__(cls, (a, b))


class SpectralBase:
    """Common bookkeeping for every 1-D basis."""

    def __init__(self, N, quad='', padding_factor=1, domain=(-1., 1.), dtype=float,
                 dealias_direct=False):
        self.N = int(N)
        self.quad = quad
        self.domain = tuple(domain)
        self._dtype = np.dtype(dtype)
        self.dealias_direct = bool(dealias_direct)
        # spectralbase.py:52-53: snap the factor so that the padded length is an integer
        self.padding_factor = float(np.floor(self.N*padding_factor)/self.N) if self.N > 0 else 1.0
        if self.dealias_direct:
            assert abs(self.padding_factor-1) < 1e-8, 'dealias_direct needs padding_factor=1'
        self.axis = 0
        self.tensorproductspace = None
        self._dealiased = {}

    # ---- identity ---------------------------------------------------------------
    @staticmethod
    def family():
        raise NotImplementedError

    @staticmethod
    def boundary_condition():
        return ''

    @property
    def is_composite_space(self):
        return 0

    @property
    def is_padded(self):
        return self.padding_factor > 1.+1e-8

    @property
    def rank(self):
        return 0

    @property
    def dimensions(self):
        return 1

    @property
    def dtype(self):
        return self._dtype

    def __len__(self):
        return 1

    def __getitem__(self, i):
        assert i == 0
        return self

    def __hash__(self):
        return hash((type(self).__name__, self.N, self.quad, self.padding_factor,
                     self.dealias_direct, self.domain))

    def __eq__(self, other):
        return type(self) is type(other) and hash(self) == hash(other)

    # ---- sizes --------------------------------------------------------------------
    def dim(self):
        """Number of degrees of freedom."""
        s = self.slice()
        return s.stop - s.start

    def dims(self):
        return (self.dim(),)

    def slice(self):
        return slice(0, self.N)

    def spectral_len(self):
        """Entries stored along the axis in spectral space."""
        return self.N

    def physical_len(self):
        """Quadrature points along the axis (padded)."""
        return int(np.floor(self.padding_factor*self.N))

    def shape(self, forward_output=True):
        return self.spectral_len() if forward_output else self.physical_len()

    def spectral_dtype_for(self, physical_dtype):
        return physical_dtype

    # ---- domain ---------------------------------------------------------------------
    def reference_domain(self):
        return (-1., 1.)

    def domain_length(self):
        return float(self.domain[1]-self.domain[0])

    def domain_factor(self):
        """spectralbase.py:873-887: reference length / true length."""
        a, b = self.domain
        c, d = self.reference_domain()
        L = float(b-a)
        R = float(d-c)
        if abs(L-R) < 1e-12:
            return 1.0
        return R/L

    def map_true_domain(self, x):
        a, b = self.domain
        c, d = self.reference_domain()
        if abs((b-a)-(d-c)) < 1e-12 and abs(a-c) < 1e-12:
            return x
        return a + (x-c)/self.domain_factor()

    def map_reference_domain(self, x):
        a, b = self.domain
        c, d = self.reference_domain()
        if abs((b-a)-(d-c)) < 1e-12 and abs(a-c) < 1e-12:
            return x
        return c + (x-a)*self.domain_factor()

    def points_and_weights(self, N=None, map_true_domain=False, weighted=True):
        raise NotImplementedError

    def mesh(self, bcast=True, map_true_domain=True):
        N = self.physical_len()
        x = self.points_and_weights(N)[0]
        if map_true_domain:
            x = self.map_true_domain(x)
        if bcast:
            x = self.broadcast_to_ndims(x)
        return x

    def broadcast_to_ndims(self, x):
        T = self.tensorproductspace
        nd = T.dimensions if T is not None else 1
        s = [np.newaxis]*nd
        s[self.axis] = slice(None)
        return x[tuple(s)]

    def wavenumbers(self, bcast=True, **kw):
        k = np.arange(self.spectral_len(), dtype=float)
        return self.broadcast_to_ndims(k) if bcast else k

    def get_mask_nyquist(self, bcast=True):
        return None

    # ---- derived spaces -------------------------------------------------------------
    def _ctor_kwargs(self):
        return dict(quad=self.quad, domain=self.domain, dtype=self._dtype,
                    padding_factor=self.padding_factor, dealias_direct=self.dealias_direct)

    def get_unplanned(self, **kw):
        args = self._ctor_kwargs()
        args.update(kw)
        return type(self)(self.N, **args)

    def get_dealiased(self, padding_factor=1.5, dealias_direct=False, **kw):
        return self.get_unplanned(padding_factor=padding_factor, dealias_direct=dealias_direct, **kw)

    def get_refined(self, N, **kw):
        args = self._ctor_kwargs()
        args.update(kw)
        return type(self)(int(N), **args)

    def get_orthogonal(self, **kw):
        return self.get_unplanned(**kw)

    # ---- device transforms along one axis of [outer][n][inner] --------------------
    # subclasses implement these three
    def _forward_axis(self, u, out, outer, inner, scalar_product=False):
        raise NotImplementedError

    def _backward_axis(self, uh, out, outer, inner):
        raise NotImplementedError

    def inplace_ok(self):
        """True if the axis kernels accept in == out (same shape and dtype)."""
        return False

    # ---- 1-D convenience API (spectralbase.py:200-313 for a single line) ----------
    def _alloc(self, forward_output, like):
        n = self.shape(forward_output)
        pd = torch.float64 if self._dtype.kind == 'f' else torch.complex128
        dt = self.spectral_dtype_for(pd) if forward_output else pd
        return torch.empty(n, dtype=dt, device=like.device)

    def forward(self, input_array, output_array=None, kind=None):
        u = ops.as_device_tensor(input_array)
        out = self._alloc(True, u) if output_array is None else output_array
        self._forward_axis(u.contiguous(), ops.plain(out), 1, 1)
        return out

    def scalar_product(self, input_array, output_array=None, kind=None):
        u = ops.as_device_tensor(input_array)
        out = self._alloc(True, u) if output_array is None else output_array
        self._forward_axis(u.contiguous(), ops.plain(out), 1, 1, scalar_product=True)
        return out

    def backward(self, input_array, output_array=None, kind=None):
        uh = ops.as_device_tensor(input_array)
        out = self._alloc(False, uh) if output_array is None else output_array
        self._backward_axis(uh.contiguous(), ops.plain(out), 1, 1)
        return out
