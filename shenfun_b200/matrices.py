"""1-D spectral matrices and their tensor products.

Host-side mirror of shenfun/matrixbase.py: `SparseMatrix` (dict of diagonals,
:20-520), `SpectralMatrix` (:534-800), `TPMatrix` (:1275-1500) and
`get_simplified_tpmatrices` (:1648-1677).  Assembly is one-off host work (numpy);
`matvec` / `solve` run on the device.
"""
import numpy as np
import torch

from . import ops

__all__ = ('SparseMatrix', 'TPMatrix', 'get_simplified_tpmatrices', 'Identity')


class SparseMatrix(dict):
    """Banded matrix as {offset: diagonal} plus a scalar scale (matrixbase.py:20-190)."""

    def __init__(self, d, shape, scale=1.0):
        dict.__init__(self, {int(k): np.atleast_1d(np.asarray(v)) for k, v in d.items()})
        self.shape = tuple(shape)
        self.scale = scale
        for k in list(self.keys()):
            n = self._diag_len(k)
            if self[k].shape[0] == 1 and n != 1:
                self[k] = np.full(n, self[k][0])
            assert self[k].shape[0] == n, (k, self[k].shape, n)

    def _diag_len(self, k):
        M, N = self.shape
        return min(M, N-k) if k >= 0 else min(M+k, N)

    def is_diagonal(self):
        return all(k == 0 for k in self.keys())

    def isidentity(self):
        return self.is_diagonal() and np.allclose(self.get(0, np.zeros(1)), 1.0)

    def scaled_diagonal(self, k=0):
        return self[k]*self.scale

    def diags(self):
        """Dense array (host) -- for tests and small problems."""
        M, N = self.shape
        A = np.zeros((M, N), dtype=np.result_type(*[v.dtype for v in self.values()], float))
        for k, v in self.items():
            if k >= 0:
                A[np.arange(len(v)), np.arange(len(v))+k] = v
            else:
                A[np.arange(len(v))-k, np.arange(len(v))] = v
        return A*self.scale

    def dia_data(self):
        """scipy DIA layout `(offsets, data[ndiag, N])` of the scaled matrix: upper
        diagonals right-aligned, lower left-aligned (la.py:347-349)."""
        M, N = self.shape
        offsets = sorted(self.keys())
        data = np.zeros((len(offsets), N), dtype=float)
        for r, k in enumerate(offsets):
            v = np.real(self[k])*self.scale
            if k >= 0:
                data[r, k:k+len(v)] = v
            else:
                data[r, :len(v)] = v
        return offsets, data


class Identity(SparseMatrix):
    def __init__(self, shape, scale=1.0):
        SparseMatrix.__init__(self, {0: np.ones(min(shape))}, shape, scale)


class TPMatrix:
    """Tensor product of 1-D matrices with a scalar scale (matrixbase.py:1275-1500)."""

    def __init__(self, mats, test_space, trial_space, scale=1.0, global_index=(0, 0)):
        self.mats = list(mats)
        self.space = test_space
        self.trialspace = trial_space
        self.scale = scale
        self.global_index = global_index
        self._diag = None

    @classmethod
    def assemble(cls, test_space, trial_space, d_test, d_trial, scale=1.0, global_index=(0, 0)):
        mats = []
        for ax, (tb, ub) in enumerate(zip(test_space.bases, trial_space.bases)):
            mats.append(tb.spectral_matrix(d_test[ax], d_trial[ax], ub))
        return cls(mats, test_space, trial_space, scale, global_index)

    @property
    def dimensions(self):
        return len(self.mats)

    def is_diagonal(self):
        return all(m.is_diagonal() for m in self.mats)

    def isidentity(self):
        return all(m.isidentity() for m in self.mats)

    def get_key(self):
        return tuple(tuple(sorted(m.keys())) for m in self.mats)

    def diagonal_array(self):
        """Device tensor of the local block of prod_ax diag_ax * scale (broadcast shape)."""
        if self._diag is None:
            assert self.is_diagonal()
            sl = self.space.local_slice(True)
            nd = self.dimensions
            arr = None
            for ax, m in enumerate(self.mats):
                v = m.scaled_diagonal(0)
                n = self.space.bases[ax].spectral_len()
                full = np.zeros(n, dtype=v.dtype)
                s = self.space.bases[ax].slice()
                full[s] = v
                shape = [1]*nd
                shape[ax] = sl[ax].stop-sl[ax].start
                part = full[sl[ax]].reshape(shape)
                arr = part if arr is None else arr*part
            arr = arr*self.scale
            if np.iscomplexobj(arr) and np.abs(arr.imag).max() == 0:
                arr = arr.real
            self._diag = ops.as_device_tensor(np.ascontiguousarray(arr))
        return self._diag

    def matvec(self, u, out, accumulate=False):
        """out (+)= A u along every axis."""
        u, out = ops.plain(u), ops.plain(out)
        if self.is_diagonal():
            d = self.diagonal_array()
            if accumulate:
                out.addcmul_(u, d.to(u.dtype) if d.dtype != u.dtype else d)
            else:
                torch.mul(u, d, out=out)
            return out
        from .la import tp_matvec
        return tp_matvec(self, u, out, accumulate)

    def solve(self, b, u=None, constraints=()):
        """matrixbase.py:1404-1435: diagonal branch divides with 1/0 -> 0."""
        if u is None:
            u = b
        if self.is_diagonal():
            inv = getattr(self, '_inv_diag', None)
            if inv is None:
                d = self.diagonal_array()
                inv = torch.where(d == 0, torch.zeros_like(d), 1.0/d).contiguous()
                self._inv_diag = inv                      # the reciprocal is formed once per matrix
            bt, ut = ops.plain(b), ops.plain(u)
            if (inv.dtype == torch.float64 and ut.dtype == torch.complex128 and ut.is_contiguous()
                    and ut.numel() % inv.numel() == 0 and tuple(ut.shape[ut.ndim-inv.ndim:]) == tuple(inv.shape)):
                if ut.data_ptr() != bt.data_ptr():
                    ut.copy_(bt)
                ops.apply_mask(ut, inv)                   # u *= 1/d in the library's point-wise kernel
            else:
                torch.mul(bt, inv, out=ut)
            return u
        from .la import SolverGeneric1ND
        return SolverGeneric1ND([self])(b, u, constraints=constraints)

    def __mul__(self, a):
        return TPMatrix(self.mats, self.space, self.trialspace, self.scale*a, self.global_index)

    __rmul__ = __mul__


class DiagonalTPMatrix(TPMatrix):
    """Sum of diagonal TPMatrices folded into one scale array (what
    `get_simplified_tpmatrices` leaves for an all-Fourier form)."""

    def __init__(self, mats):
        m0 = mats[0]
        TPMatrix.__init__(self, m0.mats, m0.space, m0.trialspace, 1.0, m0.global_index)
        d = None
        for m in mats:
            a = m.diagonal_array()
            d = a if d is None else d+a
        self._diag = d

    def is_diagonal(self):
        return True

    def isidentity(self):
        return True

    @property
    def scale_array(self):
        return self._diag


def get_simplified_tpmatrices(mats):
    """matrixbase.py:1648-1677: fold diagonal (Fourier) axes into scale arrays and add
    matrices with identical sparsity.  An all-diagonal form collapses to one matrix."""
    if isinstance(mats, TPMatrix):
        mats = [mats]
    if all(m.is_diagonal() for m in mats):
        groups = {}
        for m in mats:
            groups.setdefault(m.global_index, []).append(m)
        return [DiagonalTPMatrix(g) for g in groups.values()]
    return list(mats)


def mass_solve(T, rhs, out):
    from .la import mass_solve as _ms
    return _ms(T, rhs, out)
