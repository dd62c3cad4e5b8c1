"""Chebyshev bases (first kind, Gauss-Chebyshev points) on the device.

Host-side mirror of shenfun/chebyshev/bases.py: `Orthogonal` :118-470 and the
composite bases defined by their stencils (`ShenDirichlet` :476-554, `ShenNeumann`
:937, `ShenBiharmonic` :1249, and the mixed Dirichlet/Neumann variants).  The fast
transforms are DCT-II/III built on the complex FFT kernels, with the quadrature
scaling, the Shen stencil recombination and zero-padding fused
(`shb_cheb_scalar_product`, `shb_cheb_backward`); complex data is transformed in one
pass.  quad='GC' (the reference default) takes that path; quad='GL' (Gauss-Lobatto points,
DCT-I in the reference, chebyshev/bases.py:331-338, 349-353) is evaluated as the same quadrature
sums through the fp64 tensor-core contraction (`shb_gemm_axis`) with the operator matrices built
once on the host -- the route the Legendre bases take.
"""
import numpy as np

from .polybase import PolyBase, BoundaryConditions
from . import ops

__all__ = ('Orthogonal', 'ShenDirichlet', 'ShenNeumann', 'ShenBiharmonic', 'CompositeBase', 'make_space')


class Orthogonal(PolyBase):
    def __init__(self, N, quad='GC', padding_factor=1, domain=(-1., 1.), dtype=float, dealias_direct=False,
                 bc=None, scaled=False, **kw):
        quad = 'GC' if quad is None else quad
        if quad not in ('GC', 'GL'):
            raise NotImplementedError("Chebyshev quad='GC' or 'GL' on the device path")
        PolyBase.__init__(self, N, quad, padding_factor=padding_factor, domain=domain, dtype=dtype,
                          dealias_direct=dealias_direct, bc=bc, scaled=scaled)

    @staticmethod
    def family():
        return 'chebyshev'

    @staticmethod
    def short_name():
        return 'T'

    def points_and_weights(self, N=None, map_true_domain=False, weighted=True):
        """chebyshev/bases.py:201-216: x_j = cos(pi (2j+1)/(2N)), w = pi/N."""
        if N is None:
            N = self.physical_len()
        if self.quad == 'GL':
            # chebyshev/bases.py:206-210, 224-234: x_j = cos(pi j/(N-1)), end weights halved
            import scipy.fft as sfft
            points = np.cos(np.arange(N)*np.pi/(N-1))
            if weighted:
                weights = np.full(N, np.pi/(N-1))
            else:
                d = np.zeros(N)
                k = 2*(1+np.arange((N-1)//2))
                d[::2] = (2.0/(N-1))/np.hstack((1.0, 1.0-k*k))
                weights = sfft.dct(d, type=1)
            weights[0] *= 0.5
            weights[-1] *= 0.5
            if map_true_domain:
                points = self.map_true_domain(points)
            return points, weights
        points, weights = np.polynomial.chebyshev.chebgauss(N)
        points = points.astype(float)
        weights = weights.astype(float)
        if not weighted:
            # chebyshev/bases.py:235-241: weights of int f dx (no Chebyshev weight) on the GC points,
            # w = DCT-III(d), d[::2] = (2/N) / (1, 1-k^2), k = 2, 4, ...
            import scipy.fft as sfft
            d = np.zeros(N)
            k = 2*(1+np.arange((N-1)//2))
            d[::2] = (2.0/N)/np.hstack((1.0, 1.0-k*k))
            weights = sfft.dct(d, type=3)
        if map_true_domain:
            points = self.map_true_domain(points)
        return points, weights

    def orthogonal_vandermonde(self, X):
        return np.polynomial.chebyshev.chebvander(X, self.N-1)

    def ortho_norm_sq(self, N=None):
        N = self.N if N is None else N
        h = np.full(N, np.pi/2)
        h[0] = np.pi
        if self.quad == 'GL' and N == self.N:
            h[-1] = np.pi           # discrete norm of T_{N-1} on the Lobatto points (bases.py:271-280)
        return h

    def ortho_deriv_matrix(self, N, k):
        """T_j' = sum_{i<j, i+j odd} (2/c_i) j T_i ; higher orders by exact integer products."""
        D1 = np.zeros((N, N))
        j = np.arange(N)
        for i in range(N):
            cols = j[(j > i) & ((i+j) % 2 == 1)]
            D1[i, cols] = 2.0*cols/(2.0 if i == 0 else 1.0)
        D = np.eye(N)
        for _ in range(k):
            D = D1 @ D
        return D

    def ortho_gram(self, N, d_test, d_trial):
        """(T_j^(b), T_k^(a))_w = (pi/2) * c_k-weighted integer matrix: pi/2 is pulled out."""
        ld = np.longdouble
        c = np.ones(N, dtype=ld)
        c[0] = 2
        if self.quad == 'GL' and N == self.N and d_test == 0 and d_trial == 0:
            c[-1] = 2               # mass matrix with the Lobatto quadrature (matrices.py BSDSDmat: d[0][-1] = 1.5 pi)
        G = c[:, None]*self._deriv_ld(N, d_trial)
        if d_test > 0:
            G = self._deriv_ld(N, d_test).T @ G
        return G, np.longdouble('1.5707963267948966192313216916397514')      # pi/2 to long-double precision

    def boundary_values(self, n, order, side):
        n = np.asarray(n, dtype=float)
        sgn = 1.0 if side == 'right' else -1.0
        if order == 0:
            return sgn**n
        if order == 1:
            return sgn**(n+1)*n**2
        if order == 2:
            return sgn**n*n**2*(n**2-1)/3
        if order == 3:
            return sgn**(n+1)*n**2*(n**2-1)*(n**2-4)/15
        raise NotImplementedError

    def _canonical_stencil(self, st, k):
        kind = self.boundary_condition()
        if kind == 'Dirichlet':
            return {0: np.ones(len(k), dtype=k.dtype), 2: -np.ones(len(k), dtype=k.dtype)}
        if kind == 'Neumann':
            return {0: np.ones(len(k), dtype=k.dtype), 2: -(k/(k+2.0))**2}
        if kind == 'Biharmonic':
            return {0: np.ones(len(k), dtype=k.dtype), 2: (-2.0*k-4)/(k+3), 4: (k+1.0)/(k+3)}
        return st

    def closed_form_matrix(self, d_test, d_trial, trial):
        """(phi_j^(4), phi_k)_w on the biharmonic basis (Shen 1995; the reference's SSBSBmat,
        chebyshev/matrices.py:1429-1462): upper triangular, every even off-diagonal follows
        8 (k+1)(k+2)(k(k+4) + 3(j+2)^2) pi / (j+3)."""
        if (d_test, d_trial) == (0, 4) and trial.boundary_condition() == 'Biharmonic' \
                and self.boundary_condition() == 'Biharmonic' and not self._scaled and not trial._scaled:
            n = self.dim()
            ki = np.arange(n, dtype=np.longdouble)
            pi = np.longdouble('3.14159265358979323846264338327950288')
            d = {0: np.asarray(pi*8*(ki+1)**2*(ki+2)*(ki+4), dtype=float)}
            for off in range(2, n, 2):
                k = ki[:-off]
                j = ki[off:]
                d[off] = np.asarray(8*(k+1)*(k+2)*(k*(k+4)+3*(j+2)**2)*pi/(j+3), dtype=float)
            return d
        return None

    # ---- device transforms ---------------------------------------------------------------------
    def _gl_operators(self):
        """Operator matrices of the Lobatto-point transforms.  Forward (N x Np): rows k < dim =
        K (V^T W)/df, zero rows for the boundary dofs.  Backward (Np x N): V K^T on the dofs plus
        the boundary functions on the bc rows.  V[j, k] = T_k(x_j) = cos(pi j k/(Np-1))."""
        if 'gl_ops' not in self._tables:
            F, B = self._gl_operators_host()
            self._tables['gl_ops'] = (ops.as_device_tensor(F), ops.as_device_tensor(B))
        return self._tables['gl_ops']

    def _gl_operators_host(self):
        Np = self.physical_len()
        ld = np.longdouble
        pi = ld('3.14159265358979323846264338327950288')
        jk = np.outer(np.arange(Np), np.arange(self.N)) % (2*(Np-1))       # exact integer phase
        V = np.cos(pi*jk.astype(ld)/(Np-1))
        w = np.full(Np, pi/(Np-1))
        w[0] /= 2
        w[-1] /= 2
        K = self.stencil_matrix().diags().astype(ld)             # (dim, N)
        dim = self.dim()
        F = np.zeros((self.N, Np), dtype=ld)
        F[:dim] = K @ (V.T*w[None, :])/ld(float(self.domain_factor()))
        B = np.zeros((Np, self.N), dtype=ld)
        B[:, :dim] = V @ K.T
        nbc = self.bcs.num_bcs()
        if nbc and self.N >= 4:
            L = self.bc_basis()                                   # (nbc, 4)
            B[:, dim:] = V[:, :4] @ L[:, :4].T.astype(ld)
        return np.ascontiguousarray(F.astype(np.float64)), np.ascontiguousarray(B.astype(np.float64))

    def _scalar_product_axis(self, u, out, outer, inner):
        if self.quad == 'GL':
            ops.gemm_axis(self._gl_operators()[0], u, out, outer, inner)
            return
        offs, arrs = self._device_stencil()
        n = self.physical_len()
        scale = np.pi/(2*float(self.domain_factor())*n)
        ops.cheb_scalar_product(u, out, outer, inner, n, self.N, scale, self.dim(), offs, arrs)

    def _backward_axis(self, uh, out, outer, inner):
        if self.quad == 'GL':
            ops.gemm_axis(self._gl_operators()[1], uh, out, outer, inner)
            return
        offs, arrs = self._device_stencil()
        nbc = self.bcs.num_bcs() if self.has_nonhomogeneous_bcs else 0
        lift = None
        if nbc:
            if 'bc_lift' not in self._tables:
                self._tables['bc_lift'] = ops.as_device_tensor(np.ascontiguousarray(self.bc_basis()))
            lift = self._tables['bc_lift']
        ops.cheb_backward(uh, out, outer, inner, self.physical_len(), self.N, self.dim(), offs, arrs, nbc, lift)


CompositeBase = Orthogonal


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
    par = dict(quad=quad or 'GC', padding_factor=padding_factor, dealias_direct=dealias_direct,
               dtype=np.dtype(dtype), scaled=bool(scaled))
    if domain is not None:
        par['domain'] = tuple(float(d) for d in domain)
    return Orthogonal(N, bc=bc, **par)
