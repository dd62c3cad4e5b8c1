"""Orthogonal Chebyshev spaces of the second kind U_n (shenfun/chebyshevu/bases.py:60-190,
`FunctionSpace(N, 'U')`): weight sqrt(1 - x^2), quadrature 'GU' (Gauss points of U_N) or 'GC' (Chebyshev-Gauss
points with the weights times 1 - x^2).  The reference transforms with a DST; here the transforms are the
parity-split tensor-core contraction of host-built operators, as for Legendre.  Orthogonal space only."""
import numpy as np

from . import legendre
from .polybase import BoundaryConditions

__all__ = ('Orthogonal',)


class Orthogonal(legendre.Orthogonal):
    def __init__(self, N, quad='GU', padding_factor=1, domain=(-1., 1.), dtype=float, dealias_direct=False,
                 bc=None, scaled=False, **kw):
        if bc is not None and (bc if isinstance(bc, BoundaryConditions) else BoundaryConditions(bc)).num_bcs():
            raise NotImplementedError('composite Chebyshev-U bases are outside the device hot path')
        quad = 'GU' if quad is None else quad
        if quad not in ('GU', 'GC'):
            raise NotImplementedError("Chebyshev-U quadrature: 'GU' or 'GC'")
        legendre.PolyBase.__init__(self, N, quad, padding_factor=padding_factor, domain=domain, dtype=dtype,
                                   dealias_direct=dealias_direct, bc=None, scaled=False)

    @staticmethod
    def family():
        return 'chebyshevu'

    @staticmethod
    def short_name():
        return 'U'

    def points_and_weights(self, N=None, map_true_domain=False, weighted=True):
        """chebyshevu/bases.py:116-128."""
        if N is None:
            N = self.physical_len()
        key = ('pw', N)
        if key not in self._tables:
            if self.quad == 'GU':
                x = np.cos((np.arange(N)+1)*np.pi/(N+1))
                w = np.full(N, np.pi/(N+1))*(1-x**2)
            else:
                x, w = np.polynomial.chebyshev.chebgauss(N)
                x = x.astype(float)
                w = (w*(1-x**2)).astype(float)
            self._tables[key] = (x, w)
        x, w = self._tables[key]
        if map_true_domain:
            x = self.map_true_domain(x)
        return x, w

    def _vandermonde(self, Np):
        return self.orthogonal_vandermonde(self.points_and_weights(Np)[0])

    def orthogonal_vandermonde(self, X):
        """V[j, n] = U_n(x_j): U_0 = 1, U_1 = 2x, U_{n+1} = 2x U_n - U_{n-1}."""
        X = np.asarray(X, dtype=float)
        V = np.zeros((X.shape[0], self.N))
        V[:, 0] = 1
        if self.N > 1:
            V[:, 1] = 2*X
        for n in range(2, self.N):
            V[:, n] = 2*X*V[:, n-1]-V[:, n-2]
        return V

    def ortho_norm_sq(self, N=None):
        """pi/2; on the Chebyshev-Gauss points the last mode has the discrete norm pi (bases.py:181-189)."""
        N = self.N if N is None else N
        h = np.full(N, np.pi/2)
        if self.quad == 'GC' and N == self.N:
            h[-1] *= 2
        return h

    def ortho_gram(self, N, d_test, d_trial):
        if d_test or d_trial:
            raise NotImplementedError('derivative matrices of the Chebyshev-U family are outside the device hot path')
        return np.diag(self.ortho_norm_sq(N).astype(np.longdouble)), np.longdouble(1)

    def ortho_deriv_matrix(self, N, k):
        raise NotImplementedError('derivative matrices of the Chebyshev-U family are outside the device hot path')

    def closed_form_matrix(self, d_test, d_trial, trial):
        return None

    def boundary_values(self, n, order, side):
        raise NotImplementedError
