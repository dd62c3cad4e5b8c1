"""Shared host logic of the polynomial (Chebyshev / Legendre) bases.

Mirrors, for the two families on the device path, the pieces of
shenfun/spectralbase.py that make a composite (Shen) basis out of an orthogonal one:
`BoundaryConditions` (:1677-1900, the subset used by the demos), the stencil
matrix `K` (`getCompositeBase.stencil_matrix` :1504-1526; here computed numerically
from the boundary conditions instead of through sympy), `to_ortho` / scalar product
recombination (:1462-1474, :1616-1634 -- fused into the device transforms), the
mass matrix and general 1-D matrices (:534-800 of matrixbase.py, assembled as
K_test G K_trial^T from closed-form Gram matrices of the orthogonal family), and the
inverse-mass solve of the forward transform (:560-577).
"""
import numpy as np
import torch

from .spectralbase import SpectralBase
from .matrices import SparseMatrix
from . import ops


class BoundaryConditions(dict):
    """{'left': {'D': a, 'N': b, 'N2': c}, 'right': {...}} with value lookup
    (spectralbase.py:1677-1900).  Tuples follow the reference's shorthand:
    (a, b) Dirichlet both ends, (a, b, c, d) clamped (u and u' at both ends)."""

    ORDER = {'D': 0, 'N': 1, 'N2': 2, 'N3': 3, 'N4': 4}

    def __init__(self, bc):
        d = {'left': {}, 'right': {}}
        if bc is None:
            pass
        elif isinstance(bc, (tuple, list)):
            if len(bc) == 2:
                if bc[0] is not None:
                    d['left']['D'] = bc[0]
                if bc[1] is not None:
                    d['right']['D'] = bc[1]
            elif len(bc) == 4:
                for side, (a, b) in (('left', (bc[0], bc[1])), ('right', (bc[2], bc[3]))):
                    if a is not None:
                        d[side]['D'] = a
                    if b is not None:
                        d[side]['N'] = b
            else:
                raise NotImplementedError('bc tuples of length 2 or 4')
        elif isinstance(bc, dict):
            for side in ('left', 'right'):
                for k, v in (bc.get(side) or {}).items():
                    d[side][k] = v
        dict.__init__(self, d)

    def conditions(self):
        """Ordered list of (side, derivative order, value) -- left first, low order first."""
        out = []
        for side in ('left', 'right'):
            for k in sorted(self[side], key=lambda s: self.ORDER[s]):
                out.append((side, self.ORDER[k], self[side][k]))
        return out

    def num_bcs(self):
        return len(self.conditions())

    def is_homogeneous(self):
        return all(isinstance(v, (int, float)) and v == 0 for (_, _, v) in self.conditions())

    def key(self):
        return tuple((s, o, str(v)) for s, o, v in self.conditions())


class PolyBase(SpectralBase):
    """Orthogonal polynomial basis or a composite basis built on it."""

    def __init__(self, N, quad, padding_factor=1, domain=(-1., 1.), dtype=float, dealias_direct=False,
                 bc=None, scaled=False):
        SpectralBase.__init__(self, N, quad=quad, padding_factor=padding_factor, domain=domain, dtype=dtype,
                              dealias_direct=dealias_direct)
        self.bcs = bc if isinstance(bc, BoundaryConditions) else BoundaryConditions(bc)
        self._scaled = scaled
        self._tables = {}
        self._mass_fac = None
        self._stencil_cache = {}

    # ---- to be provided by the family ------------------------------------------------
    def ortho_norm_sq(self, N=None):
        raise NotImplementedError

    def ortho_deriv_matrix(self, N, k):
        """D (N x N) with d^k P_j / dx^k = sum_i D[i, j] P_i."""
        raise NotImplementedError

    def boundary_values(self, n, order, side):
        """d^order P_n / dx^order at x = -1 ('left') or +1 ('right'), n an integer array."""
        raise NotImplementedError

    # ---- identity --------------------------------------------------------------------------
    def _ctor_kwargs(self):
        d = dict(quad=self.quad, domain=self.domain, dtype=self._dtype, padding_factor=self.padding_factor,
                 dealias_direct=self.dealias_direct, bc=self.bcs, scaled=self._scaled)
        return d

    def __hash__(self):
        return hash((type(self).__name__, self.N, self.quad, self.padding_factor, self.dealias_direct,
                     self.domain, self.bcs.key(), self._scaled))

    @property
    def is_orthogonal(self):
        return self.bcs.num_bcs() == 0

    def boundary_condition(self):
        c = self.bcs.conditions()
        if not c:
            return 'Apparently not'
        kinds = {(s, o) for s, o, _ in c}
        if kinds == {('left', 0), ('right', 0)}:
            return 'Dirichlet'
        if kinds == {('left', 1), ('right', 1)}:
            return 'Neumann'
        if kinds == {('left', 0), ('left', 1), ('right', 0), ('right', 1)}:
            return 'Biharmonic'
        return 'Mixed'

    def reference_domain(self):
        return (-1., 1.)

    def slice(self):
        return slice(0, self.N-self.bcs.num_bcs())

    def spectral_len(self):
        return self.N

    def get_orthogonal(self, **kw):
        args = self._ctor_kwargs()
        args.update(kw)
        args['bc'] = None
        return type(self)(self.N, **args)

    def get_testspace(self, kind='Galerkin', **kw):
        if kind.lower() in ('galerkin', 'g'):
            return self
        raise NotImplementedError('Petrov-Galerkin test spaces are outside the device hot path')

    # ---- stencil -------------------------------------------------------------------------------
    def stencil(self, N=None):
        """{offset: values[k], k < dim}: phi_k = sum_off stencil[off][k] P_{k+off}, found from
        the homogeneous boundary conditions with unit leading coefficient (the reference's
        `get_stencil_matrix`, spectralbase.py:1902-1990, solved numerically here)."""
        N = self.N if N is None else N
        if N in self._stencil_cache:
            return self._stencil_cache[N]
        conds = self.bcs.conditions()
        nbc = len(conds)
        dim = N-nbc
        k = np.arange(dim)
        if nbc == 0:
            st = {0: np.ones(dim)}
        else:
            # unknown a_1..a_nbc : sum_i a_i B_c(P_{k+i}) = -B_c(P_k)  for every condition c
            A = np.zeros((dim, nbc, nbc))
            rhs = np.zeros((dim, nbc))
            for c, (side, order, _) in enumerate(conds):
                rhs[:, c] = -self.boundary_values(k, order, side)
                for i in range(1, nbc+1):
                    A[:, c, i-1] = self.boundary_values(k+i, order, side)
            # scale rows for conditioning (values grow like k^(2*order))
            s = np.abs(A).max(axis=2, keepdims=True)
            s[s == 0] = 1
            a = np.linalg.solve(A/s, (rhs/s[:, :, 0])[..., None])[..., 0]
            st = {0: np.ones(dim)}
            for i in range(1, nbc+1):
                v = a[:, i-1]
                if np.abs(v).max() > 1e-13:
                    st[i] = v
            st = self._canonical_stencil(st, k)
        if self._scaled:
            sc = self._stencil_scaling(k)
            st = {o: v*sc for o, v in st.items()}
        self._stencil_cache[N] = st
        return st

    def _canonical_stencil(self, st, k):
        """Families override to substitute closed forms for the well-known bases (exact to the
        last bit, like the reference's sympy expressions)."""
        return st

    def stencil_ld(self):
        """Stencil in extended precision for matrix assembly: the closed forms evaluated in long
        double where the family knows them (the recombination K G K^T cancels O(N^7) entries of
        the 4th-derivative Gram matrix, so 1e-16 errors in K would show up as 1e-11 in the
        matrix); otherwise the float64 stencil."""
        ld = np.longdouble
        st = self.stencil()
        k = np.arange(self.dim(), dtype=ld)
        stl = self._canonical_stencil({o: np.asarray(v, dtype=ld) for o, v in st.items()}, k)
        out = {}
        for o in st:
            v = np.asarray(stl.get(o, st[o]), dtype=ld)
            # keep only closed forms that agree with the float64 stencil (sanity)
            out[o] = v if np.allclose(np.asarray(v, dtype=float), st[o], rtol=1e-14, atol=1e-15) else np.asarray(st[o], dtype=ld)
        return out

    def _stencil_scaling(self, k):
        """Row scaling of the stencil for `scaled=True`; families implement the bases that have one
        in the reference (Legendre ShenDirichlet: 1/sqrt(4k+6), legendre/bases.py:373-376)."""
        raise NotImplementedError(f'scaled=True is not available for the {self.family()} '
                                  f'{self.boundary_condition()} basis')

    def stencil_matrix(self, N=None):
        """K as a (dim x N) SparseMatrix (spectralbase.py:1504-1526)."""
        N = self.N if N is None else N
        st = self.stencil(N)
        dim = N-self.bcs.num_bcs()
        return SparseMatrix({o: v[:min(dim, N-o)] for o, v in st.items()}, (dim, N))

    def _device_stencil(self):
        """(offsets, device arrays) of the stencil for the transform kernels."""
        if 'stencil' not in self._tables:
            st = self.stencil()
            offs = sorted(st.keys())
            arrs = []
            for o in offs:
                v = st[o]
                arrs.append(None if np.all(v == 1.0) else ops.as_device_tensor(np.ascontiguousarray(v, dtype=float)))
            self._tables['stencil'] = (offs, arrs)
        return self._tables['stencil']

    # ---- boundary lifting ------------------------------------------------------------------------
    def bc_basis(self):
        """Ortho coefficients (nbc x 4) of the boundary functions: function b satisfies
        condition b with value 1 and the other conditions with value 0
        (the reference's BCGeneric space, chebyshev/bases.py & legendre/bases.py `BCGeneric`)."""
        conds = self.bcs.conditions()
        nbc = len(conds)
        if nbc == 0:
            return np.zeros((0, 4))
        # polynomials of the lowest degrees that can carry the conditions (pure Neumann
        # data cannot be lifted by a constant)
        m0 = min(order for (_, order, _) in conds)
        n = m0+np.arange(nbc)
        assert n.max() < 4, 'boundary lifting uses polynomials up to degree 3'
        A = np.zeros((nbc, nbc))
        for c, (side, order, _) in enumerate(conds):
            A[c] = self.boundary_values(n, order, side)
        C = np.linalg.solve(A, np.eye(nbc)).T       # row b: coefficients of function b
        out = np.zeros((nbc, 4))
        out[:, n] = C
        # derivatives scale with the domain: value of d^o u/dx^o = df^o d^o u/dX^o
        df = self.domain_factor()
        for b, (_, order, _) in enumerate(conds):
            out[b] /= df**order
        return out

    def evaluate_basis_all(self, x):
        """E[p, k] = phi_k(X_p) for the dofs k < dim and the boundary functions on the bc rows
        (spectralbase.py:1000-1035 `evaluate_basis_all`; composite: spectralbase.py:1616-1634)."""
        X = np.atleast_1d(self.map_reference_domain(np.asarray(x, dtype=float)))
        V = self.orthogonal_vandermonde(X)                       # (npoints, N)
        E = np.zeros((len(X), self.N))
        dim = self.dim()
        K = self.stencil_matrix().diags()                         # (dim, N)
        E[:, :dim] = V @ K.T
        nbc = self.bcs.num_bcs()
        if nbc and self.N >= 4:
            E[:, dim:dim+nbc] = V[:, :4] @ self.bc_basis()[:, :4].T
        return E

    @property
    def has_nonhomogeneous_bcs(self):
        return not self.bcs.is_homogeneous()

    # ---- matrices ------------------------------------------------------------------------------------
    def ortho_gram(self, N, d_test, d_trial):
        """(G, factor): factor*G[k, j] = (P_j^(d_trial), P_k^(d_test))_w on the reference domain.
        G is held in extended precision with any irrational constant pulled out into `factor`, so
        that the stencil recombination below (which cancels O(N^3) entries down to O(N)) loses
        nothing: for integer stencils the result is exact."""
        ld = np.longdouble
        h = self.ortho_norm_sq(N).astype(ld)
        G = h[:, None]*self._deriv_ld(N, d_trial)
        if d_test > 0:
            G = self._deriv_ld(N, d_test).T @ G
        return G, ld(1)

    def _deriv_ld(self, N, k):
        """k-th derivative matrix in extended precision (entries are integers / half-integers;
        float64 products are exact up to k = 2, beyond that the magnitudes exceed 2^53)."""
        ld = np.longdouble
        if k <= 2:
            return self.ortho_deriv_matrix(N, k).astype(ld)
        D1 = self.ortho_deriv_matrix(N, 1).astype(ld)
        D = self.ortho_deriv_matrix(N, 2).astype(ld)
        for _ in range(k-2):
            D = D1 @ D
        return D

    def spectral_matrix(self, d_test, d_trial, trial=None, tol=1e-9):
        """(phi_j^(d_trial), psi_k^(d_test))_w as a banded SparseMatrix (dim_test x dim_trial);
        domain scaling as in matrixbase.py:690-691."""
        trial = self if trial is None else trial
        assert trial.N == self.N and trial.family() == self.family()
        key = ('mat', d_test, d_trial, hash(trial))
        if key in self._tables:
            return self._tables[key]
        ld = np.longdouble
        N = self.N
        cf = self.closed_form_matrix(d_test, d_trial, trial)
        if cf is not None:
            scale = 1.0
            df = float(self.domain_factor())
            if df != 1:
                scale = df**(d_test+d_trial-1)
            self._tables[key] = SparseMatrix(cf, (self.dim(), trial.dim()), scale)
            return self._tables[key]
        G, fac = self.ortho_gram(N, d_test, d_trial)
        dim_t, dim_u = self.dim(), trial.dim()
        X = np.zeros((dim_t, N), dtype=ld)
        for off, v in self.stencil_ld().items():
            X += np.asarray(v, dtype=ld)[:, None]*G[off:off+dim_t, :]
        A = np.zeros((dim_t, dim_u), dtype=ld)
        for off, v in trial.stencil_ld().items():
            A += X[:, off:off+dim_u]*np.asarray(v, dtype=ld)[None, :]
        A = np.asarray(A*fac, dtype=float)
        scale = 1.0
        df = float(self.domain_factor())
        if df != 1:
            scale = df**(d_test+d_trial-1)
        amax = np.abs(A).max()
        d = {}
        M, Nn = A.shape
        for off in range(-(M-1), Nn):
            v = np.diagonal(A, off)
            if len(v) and np.abs(v).max() > tol*amax:
                d[off] = v.copy()
        if not d:
            d = {0: np.zeros(min(M, Nn))}
        self._tables[key] = SparseMatrix(d, A.shape, scale)
        return self._tables[key]

    def closed_form_matrix(self, d_test, d_trial, trial):
        """Families return {offset: diagonal} for matrices whose numerical assembly would lose
        accuracy (fourth-derivative matrices cancel O(N^7) terms), else None."""
        return None

    def mass_matrix(self):
        if 'mass' not in self._tables:
            self._tables['mass'] = self.spectral_matrix(0, 0)
        return self._tables['mass']

    def bc_mass(self):
        """(beta_b, phi_k)_w for the boundary functions: (dim x nbc), used by the forward
        transform to move the lifting to the right hand side (tensorproductspace.py:2091-2106)."""
        if 'bc_mass' not in self._tables:
            N = self.N
            G, fac = self.ortho_gram(N, 0, 0)
            G = np.asarray(G*fac, dtype=float)
            Kt = self.stencil_matrix().diags()
            L = np.zeros((self.bcs.num_bcs(), N))
            L[:, :4] = self.bc_basis()[:, :min(4, N)]
            self._tables['bc_mass'] = (Kt @ G @ L.T)/float(self.domain_factor())
        return self._tables['bc_mass']

    # ---- inverse mass ---------------------------------------------------------------------------------
    def _mass_solver(self):
        if self._mass_fac is None:
            from .la import BandedFactor
            self._mass_fac = BandedFactor.from_matrix(self.mass_matrix())
        return self._mass_fac

    def apply_inverse_mass(self, out, outer, inner):
        """In-place B^-1 along the axis (spectralbase.py:560-577) on [outer][N][inner]."""
        self._mass_solver().solve(out, outer, inner, self.N)

    def _line_bcs(self, outer, inner):
        """Boundary values per line for the current axis pass: [nbc][outer1][inner] tensors from the
        tensor-product space (functions of the other coordinates, already transformed along the axes
        that ran before this one; tensorproductspace.py:1909-2032), or None for a 1-D space / lines
        that all carry the same numbers."""
        T = self.tensorproductspace
        if T is None or getattr(T, 'dimensions', 1) < 2:
            return None
        return T._tensor_bcs(self)[0]

    def _set_bc_rows(self, out, outer, inner):
        """Write the boundary values into rows dim.. of every line (set_boundary_dofs, not final)."""
        conds = self.bcs.conditions()
        dim = self.dim()
        lb = self._line_bcs(outer, inner)
        if lb is not None:
            o1 = lb.shape[1]
            v = out.view(outer//o1, o1, self.N, inner)
            for b in range(len(conds)):
                v[:, :, dim+b, :] = lb[b].to(out.dtype)
            return
        v = out.view(outer, self.N, inner)
        for b, (_, _, val) in enumerate(conds):
            v[:, dim+b, :] = float(val)

    def _subtract_bc_mass(self, out, outer, inner):
        """rhs -= (beta_b, phi_k)_w * bc value of the line (`_add_mass_rhs`, tensorproductspace.py:2091-2106)."""
        conds = self.bcs.conditions()
        Mb = self.bc_mass()
        rows = np.nonzero(np.abs(Mb).max(axis=1) > 0)[0]
        lb = self._line_bcs(outer, inner)
        if lb is not None:
            o1 = lb.shape[1]
            v = out.view(outer//o1, o1, self.N, inner)
            for r in rows:
                for b in range(len(conds)):
                    if Mb[r, b] != 0:
                        v[:, :, r, :] -= float(Mb[r, b])*lb[b].to(out.dtype)
            return
        if all(float(v) == 0 for _, _, v in conds):
            return
        v = out.view(outer, self.N, inner)
        for r in rows:
            s = sum(Mb[r, b]*float(val) for b, (_, _, val) in enumerate(conds))
            if s != 0:
                v[:, r, :] -= s

    # ---- transforms: scalar product -> inverse mass ; pad -> evaluate -------------------------------
    def _forward_axis(self, u, out, outer, inner, scalar_product=False):
        self._scalar_product_axis(u, out, outer, inner)
        if not scalar_product:
            if self.has_nonhomogeneous_bcs:
                self._subtract_bc_mass(out, outer, inner)
            self.apply_inverse_mass(out, outer, inner)
            if self.has_nonhomogeneous_bcs:
                self._set_bc_rows(out, outer, inner)

    def _set_boundary_dofs_axis(self, out, outer, inner):
        if self.bcs.num_bcs():
            self._set_bc_rows(out, outer, inner)
