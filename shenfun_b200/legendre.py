"""Legendre bases (Gauss-Legendre points) on the device.

Host-side mirror of shenfun/legendre/bases.py (`Orthogonal` :60-360, Shen stencils
:374,:502,:564).  The transforms use the 'vandermonde' definition
(spectralbase.py:473-483, 539-548): a dense contraction along the axis, run as an
fp64 tensor-core (DMMA) GEMM with the quadrature weights, the domain factor and the
Shen stencil folded into the (N x N) operator matrix -- the replacement for the
reference's default matrix-free 'recursive' kernels (transforms.pyx:175-323).
"""
import numpy as np

from .polybase import PolyBase
from . import ops

__all__ = ('Orthogonal', 'ShenDirichlet', 'ShenNeumann', 'ShenBiharmonic', 'make_space', 'leggauss')


def leggauss(N):
    """Gauss-Legendre nodes (ascending) and weights by Newton iteration in extended
    precision; agrees with the reference's fastgl (legendre/fastgl/fastgl.cpp, Bogaert's
    iteration-free formulas, accurate to a few ulp) to within that accuracy."""
    ld = np.longdouble
    k = np.arange(1, N+1, dtype=ld)
    # Tricomi initial guess
    x = np.cos(ld(np.pi)*(k-ld(0.25))/(ld(N)+ld(0.5)))*(1-ld(1)/(8*ld(N)**2)+ld(1)/(8*ld(N)**3))
    for _ in range(100):
        p0 = np.ones_like(x)
        p1 = x.copy()
        for n in range(2, N+1):
            p0, p1 = p1, ((2*n-1)*x*p1-(n-1)*p0)/n
        dp = N*(x*p1-p0)/(x*x-1)
        dx = p1/dp
        x = x-dx
        if np.abs(dx).max() < 1e-19:
            break
    p0 = np.ones_like(x)
    p1 = x.copy()
    for n in range(2, N+1):
        p0, p1 = p1, ((2*n-1)*x*p1-(n-1)*p0)/n
    dp = N*(x*p1-p0)/(x*x-1)
    w = 2/((1-x*x)*dp*dp)
    x = x[::-1]
    w = w[::-1]
    # enforce exact antisymmetry / symmetry
    x = (x-x[::-1])/2
    w = (w+w[::-1])/2
    return np.asarray(x, dtype=float), np.asarray(w, dtype=float)


def leglobatto(N):
    """Legendre-Gauss-Lobatto nodes (ascending, -1 and 1 included) and weights 2 / (N (N-1) L_{N-1}(x)^2):
    the interior nodes are the roots of L_{N-1}', by Newton iteration in extended precision (the reference
    iterates the same function in double precision, legendre/lobatto.py:7-46)."""
    ld = np.longdouble
    n = N-1
    j = np.arange(N, dtype=ld)
    x = -np.cos(ld(np.pi)*j/n)                       # Chebyshev-Lobatto start

    def pn(x):
        p0, p1 = np.ones_like(x), x.copy()
        for m in range(2, n+1):
            p0, p1 = p1, ((2*m-1)*x*p1-(m-1)*p0)/m
        return p0, p1                                 # L_{n-1}, L_n

    xi = x[1:-1]
    for _ in range(100):
        p0, p1 = pn(xi)
        d1 = n*(xi*p1-p0)/(xi*xi-1)                   # L_n'
        d2 = (2*xi*d1-n*(n+1)*p1)/(1-xi*xi)           # L_n'' from the Legendre equation
        dx = d1/d2
        xi = xi-dx
        if np.abs(dx).max() < 1e-19:
            break
    x[1:-1] = xi
    x = (x-x[::-1])/2
    w = 2/(N*n*pn(x)[1]**2)
    w = (w+w[::-1])/2
    return np.asarray(x, dtype=float), np.asarray(w, dtype=float)


class Orthogonal(PolyBase):
    def __init__(self, N, quad='LG', padding_factor=1, domain=(-1., 1.), dtype=float, dealias_direct=False,
                 bc=None, scaled=False, **kw):
        quad = 'LG' if quad is None else quad
        if quad not in ('LG', 'GL'):
            raise NotImplementedError("Legendre quadrature: 'LG' (Gauss) or 'GL' (Gauss-Lobatto)")
        PolyBase.__init__(self, N, quad, padding_factor=padding_factor, domain=domain, dtype=dtype,
                          dealias_direct=dealias_direct, bc=bc, scaled=scaled)

    @staticmethod
    def family():
        return 'legendre'

    def _stencil_scaling(self, k):
        if self.boundary_condition() == 'Dirichlet':
            return 1.0/np.sqrt(4.0*np.asarray(k, dtype=float)+6.0)     # legendre/bases.py:373-376
        return PolyBase._stencil_scaling(self, k)

    @staticmethod
    def short_name():
        return 'L'

    def points_and_weights(self, N=None, map_true_domain=False, weighted=True):
        if N is None:
            N = self.physical_len()
        key = ('pw', N)
        if key not in self._tables:
            self._tables[key] = leggauss(N) if self.quad == 'LG' else leglobatto(N)
        x, w = self._tables[key]
        if map_true_domain:
            x = self.map_true_domain(x)
        return x, w

    def orthogonal_vandermonde(self, X):
        return np.polynomial.legendre.legvander(X, self.N-1)

    def ortho_norm_sq(self, N=None):
        N = self.N if N is None else N
        return 2.0/(2*np.arange(N)+1.0)

    def _gl_norm(self, G):
        """Gauss-Lobatto points integrate L_{N-1}^2 to 2/(N-1), not 2/(2N-1); the reference puts that
        discrete norm into its mass matrices (legendre/bases.py:157-164 `l2_norm_sq`)."""
        if self.quad == 'GL' and self.N > 1:
            G = G.copy()
            G[self.N-1, self.N-1] = np.longdouble(2)/(self.N-1)
        return G

    def ortho_deriv_matrix(self, N, k):
        """L_j' = sum_{i<j, i+j odd} (2i+1) L_i."""
        D1 = np.zeros((N, N))
        j = np.arange(N)
        for i in range(N):
            cols = j[(j > i) & ((i+j) % 2 == 1)]
            D1[i, cols] = 2.0*i+1.0
        D = np.eye(N)
        for _ in range(k):
            D = D1 @ D
        return D

    def ortho_gram(self, N, d_test, d_trial):
        """(L_j^(b), L_i) in closed integer form where available (exact), extended precision."""
        ld = np.longdouble
        i = np.arange(N, dtype=ld)[:, None]
        j = np.arange(N, dtype=ld)[None, :]
        par = (np.arange(N)[:, None]+np.arange(N)[None, :]) % 2
        if d_trial == 0:
            G = np.diag(ld(2)/(2*np.arange(N, dtype=ld)+1))
            if d_test == 0 and N == self.N:
                G = self._gl_norm(G)
        elif d_trial == 1:
            G = np.where((j > i) & (par == 1), ld(2), ld(0))
        elif d_trial == 2:
            G = np.where((j > i) & (par == 0), j*(j+1)-i*(i+1), ld(0))
        else:
            G = (ld(2)/(2*np.arange(N, dtype=ld)+1))[:, None]*self._deriv_ld(N, d_trial)
        if d_test > 0:
            G = self._deriv_ld(N, d_test).T @ G
        return G, ld(1)

    def closed_form_matrix(self, d_test, d_trial, trial):
        """(phi_j^(4), phi_k) = 2 (2k+3)^2 (2k+5) delta_kj on the Legendre biharmonic basis (Shen 1994)."""
        if (d_test, d_trial) == (0, 4) and trial.boundary_condition() == 'Biharmonic' \
                and self.boundary_condition() == 'Biharmonic' and not self._scaled and not trial._scaled:
            k = np.arange(self.dim(), dtype=float)
            return {0: 2*(2*k+3)**2*(2*k+5)}
        return None

    def boundary_values(self, n, order, side):
        n = np.asarray(n, dtype=float)
        sgn = 1.0 if side == 'right' else -1.0
        if order == 0:
            return sgn**n
        if order == 1:
            return sgn**(n+1)*n*(n+1)/2
        if order == 2:
            return sgn**n*(n-1)*n*(n+1)*(n+2)/8
        if order == 3:
            return sgn**(n+1)*(n-2)*(n-1)*n*(n+1)*(n+2)*(n+3)/48
        raise NotImplementedError

    def _canonical_stencil(self, st, k):
        kind = self.boundary_condition()
        if kind == 'Dirichlet':
            return {0: np.ones(len(k), dtype=k.dtype), 2: -np.ones(len(k), dtype=k.dtype)}
        if kind == 'Neumann':
            return {0: np.ones(len(k), dtype=k.dtype), 2: -k*(k+1.0)/(k**2+5.0*k+6)}
        if kind == 'Biharmonic':
            return {0: np.ones(len(k), dtype=k.dtype), 2: (-4.0*k-10)/(2*k+7), 4: (2.0*k+3)/(2*k+7)}
        return st

    # ---- operator matrices of the two contractions -----------------------------------------------
    def _vandermonde(self, Np):
        """V[j, k] = L_k(x_j), k < N, by the three-term recurrence."""
        x = self.points_and_weights(Np)[0]
        N = self.N
        V = np.zeros((Np, N))
        V[:, 0] = 1
        if N > 1:
            V[:, 1] = x
        for n in range(2, N):
            V[:, n] = ((2*n-1)*x*V[:, n-1]-(n-1)*V[:, n-2])/n
        return V

    def _operators(self):
        """Forward (N x Np): rows k < dim = K (V^T W)/df, zero rows for the boundary dofs.
        Backward (Np x N): V K^T on the dofs plus the boundary functions on the bc rows."""
        if 'ops' not in self._tables:
            Np = self.physical_len()
            x, w = self.points_and_weights(Np)
            V = self._vandermonde(Np)
            K = self.stencil_matrix().diags()                 # (dim, N)
            dim = self.dim()
            F = np.zeros((self.N, Np))
            F[:dim] = K @ (V.T*w[None, :])/float(self.domain_factor())
            B = np.zeros((Np, self.N))
            B[:, :dim] = V @ K.T
            nbc = self.bcs.num_bcs()
            # homogeneous conditions: the boundary dofs are zero by definition, their columns stay zero and the
            # operator keeps the parity structure ops.gemm_axis exploits (half the tensor-core work)
            if nbc and not self.bcs.is_homogeneous():
                L = self.bc_basis()                           # (nbc, 4)
                B[:, dim:] = V[:, :4] @ L[:, :min(4, self.N)].T if self.N >= 4 else 0
            self._tables['ops'] = (ops.as_device_tensor(np.ascontiguousarray(F)),
                                   ops.as_device_tensor(np.ascontiguousarray(B)))
        return self._tables['ops']

    def _scalar_product_axis(self, u, out, outer, inner):
        F, _ = self._operators()
        ops.gemm_axis(F, u, out, outer, inner)

    def _backward_axis(self, uh, out, outer, inner):
        _, B = self._operators()
        ops.gemm_axis(B, uh, out, outer, inner)


def ShenDirichlet(N, bc=(0, 0), **kw):
    return Orthogonal(N, bc=bc, **kw)


def ShenNeumann(N, bc=None, **kw):
    return Orthogonal(N, bc={'left': {'N': 0}, 'right': {'N': 0}} if bc is None else bc, **kw)


def ShenBiharmonic(N, bc=(0, 0, 0, 0), **kw):
    return Orthogonal(N, bc=bc, **kw)


_BASIS_BCS = {'orthogonal': None, 'shendirichlet': (0, 0), 'shenneumann': {'left': {'N': 0}, 'right': {'N': 0}},
              'shenbiharmonic': (0, 0, 0, 0)}


def make_space(N, bc=None, quad=None, domain=None, dtype='d', scaled=None, basis=None, padding_factor=1,
               dealias_direct=False, **kw):
    if basis is not None:
        # forms/arguments.py:30-418: `basis` names a class; only the ones boundary conditions imply exist here
        name = str(basis).lower()
        if name not in _BASIS_BCS:
            raise NotImplementedError(f'basis={basis!r}: only {sorted(_BASIS_BCS)} are on the device path')
        if bc is None:
            bc = _BASIS_BCS[name]
    par = dict(quad=quad or 'LG', padding_factor=padding_factor, dealias_direct=dealias_direct,
               dtype=np.dtype(dtype), scaled=bool(scaled))
    if domain is not None:
        par['domain'] = tuple(float(d) for d in domain)
    return Orthogonal(N, bc=bc, **par)
