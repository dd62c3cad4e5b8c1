"""Orthogonal Jacobi spaces P_n^(alpha, beta) on the Gauss-Jacobi points (shenfun/jacobi/bases.py:150-260,
`FunctionSpace(N, 'J', alpha=..., beta=...)`).

Same device path as the Legendre family: the transforms are contractions with host-built operator
matrices on the fp64 tensor pipe (`shb_gemm_axis`; for alpha == beta the operators have the even / odd
structure and take the half-size parity contraction).  Only the orthogonal space is provided -- the
reference's composite Jacobi bases (`CompactDirichlet`, `Phi1..4`, `Generic`) are outside the hot path.
"""
import numpy as np
from scipy.special import eval_jacobi, gammaln, roots_jacobi

from . import legendre
from .polybase import BoundaryConditions

__all__ = ('Orthogonal',)


class Orthogonal(legendre.Orthogonal):
    def __init__(self, N, quad='JG', alpha=0, beta=0, padding_factor=1, domain=(-1., 1.), dtype=float,
                 dealias_direct=False, bc=None, scaled=False, **kw):
        if bc is not None and (bc if isinstance(bc, BoundaryConditions) else BoundaryConditions(bc)).num_bcs():
            raise NotImplementedError('composite Jacobi bases are outside the device hot path')
        quad = 'JG' if quad is None else quad
        if quad != 'JG':
            raise NotImplementedError("Jacobi quadrature: 'JG' (Gauss-Jacobi)")
        self.alpha, self.beta = alpha, beta
        legendre.PolyBase.__init__(self, N, quad, padding_factor=padding_factor, domain=domain, dtype=dtype,
                                   dealias_direct=dealias_direct, bc=None, scaled=False)

    @staticmethod
    def family():
        return 'jacobi'

    @staticmethod
    def short_name():
        return 'J'

    def _ctor_kwargs(self):
        d = legendre.PolyBase._ctor_kwargs(self)
        d.update(alpha=self.alpha, beta=self.beta)
        return d

    def __hash__(self):
        return hash((legendre.PolyBase.__hash__(self), self.alpha, self.beta))

    def points_and_weights(self, N=None, map_true_domain=False, weighted=True):
        """jacobi/bases.py:183-193 (`scipy.special.roots_jacobi`, as the reference)."""
        if N is None:
            N = self.physical_len()
        key = ('pw', N)
        if key not in self._tables:
            self._tables[key] = roots_jacobi(N, self.alpha, self.beta)
        x, w = self._tables[key]
        if map_true_domain:
            x = self.map_true_domain(x)
        return x, w

    def _vandermonde(self, Np):
        x = self.points_and_weights(Np)[0]
        return self.orthogonal_vandermonde(x)

    def orthogonal_vandermonde(self, X):
        """V[j, n] = P_n^(alpha, beta)(x_j) (jacobi/bases.py:195-218, 'numpy' mode)."""
        X = np.asarray(X, dtype=float)
        V = np.zeros((X.shape[0], self.N))
        for n in range(self.N):
            V[:, n] = eval_jacobi(n, self.alpha, self.beta, X)
        return V

    def ortho_norm_sq(self, N=None):
        """h_n = 2^(a+b+1) G(n+a+1) G(n+b+1) / ((2n+a+b+1) n! G(n+a+b+1)) (jacobi/recursions.py `h`)."""
        N = self.N if N is None else N
        a, b = float(self.alpha), float(self.beta)
        n = np.arange(N, dtype=float)
        lg = (a+b+1)*np.log(2.0)+gammaln(n+a+1)+gammaln(n+b+1)-gammaln(n+1)-gammaln(n+a+b+1)
        h = np.exp(lg)/(2*n+a+b+1)
        if a+b+1 == 0:          # n = 0: Gamma(0) in the formula; the limit is the weight integral
            h[0] = np.exp((a+b+1)*np.log(2.0)+gammaln(a+1)+gammaln(b+1)-gammaln(a+b+2))
        return h

    def ortho_gram(self, N, d_test, d_trial):
        if d_test or d_trial:
            raise NotImplementedError('derivative matrices of the Jacobi family are outside the device hot path')
        return np.diag(self.ortho_norm_sq(N).astype(np.longdouble)), np.longdouble(1)

    def ortho_deriv_matrix(self, N, k):
        raise NotImplementedError('derivative matrices of the Jacobi family are outside the device hot path')

    def closed_form_matrix(self, d_test, d_trial, trial):
        return None

    def boundary_values(self, n, order, side):
        raise NotImplementedError
