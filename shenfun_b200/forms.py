"""A small Cartesian form language: TestFunction / TrialFunction, Dx / grad / div /
curl, `inner`, `project`, and the cached `Inner` / `Project` objects.

Host-side mirror of the *interface* of shenfun/forms/arguments.py:420-1202
(`Expr`, `TestFunction`, `TrialFunction`), forms/operators.py (`Dx`, `grad`,
`div`, `curl`), forms/inner.py:16-566 (`inner`, `Inner`) and forms/project.py:16-272
(`project`, `Project`) for the cases the hot path needs: constant-coefficient terms on
Cartesian tensor-product spaces.  Bilinear forms give `TPMatrix` objects (1-D
matrices per axis, see matrices.py); linear forms are evaluated on the device as
banded / structured mat-vecs along the non-periodic axis times diagonal Fourier scales.
Curvilinear coordinates and sympy-valued coefficients are out of scope
(SURVEY.md section 2, rows 12 and 18).
"""
import numbers
import numpy as np
import torch

from .arrays import Function, Array
from . import ops

__all__ = ('Expr', 'TestFunction', 'TrialFunction', 'Dx', 'grad', 'div', 'curl', 'inner', 'project',
           'Inner', 'Project')


class Expr:
    """Sum of scaled derivatives of basis functions or of known Functions.

    `terms[c]` is the list of `(scale, comp, derivs, ref)` making up output component `c`:
    `scale * d^{derivs}/dx^{derivs} ref[comp]`, where `ref` is a TestFunction / TrialFunction
    or a `Function` holding expansion coefficients (a term may reference its own Function,
    so sums like `Dx(H[1], 0, 1) - Dx(H[0], 1, 2)` are fine).
    """

    def __init__(self, basis, terms=None, rank=None):
        if terms is None:                       # Expr(Function): arguments.py:452-470
            e = as_expr(basis)
            basis, terms, rank = e.basis, e.terms, e.rank
        self.basis = basis
        self.terms = terms
        self.rank = rank           # 0: scalar expression, 1: vector, 2: tensor

    @property
    def argument(self):
        args = {(t[3].argument if isinstance(t[3], BasisFunction) else 2) for comp in self.terms for t in comp}
        assert len(args) == 1, 'an expression cannot mix test, trial and known functions'
        return args.pop()

    def function_space(self):
        return self.basis.function_space()

    def base(self):
        return self.basis

    @property
    def dimensions(self):
        return self.function_space().dimensions

    def num_components(self):
        return len(self.terms)

    def _scaled(self, a):
        return Expr(self.basis, [[(s*a, c, d, r) for (s, c, d, r) in t] for t in self.terms], self.rank)

    def __mul__(self, a):
        assert isinstance(a, numbers.Number), 'only constant coefficients are on the device path'
        return self._scaled(a)

    __rmul__ = __mul__

    def __truediv__(self, a):
        return self._scaled(1./a)

    def __neg__(self):
        return self._scaled(-1)

    def __add__(self, other):
        other = as_expr(other)
        assert other.rank == self.rank and len(other.terms) == len(self.terms)
        return Expr(self.basis, [a+b for a, b in zip(self.terms, other.terms)], self.rank)

    def __sub__(self, other):
        return self + (-as_expr(other))

    def __getitem__(self, i):
        assert self.rank == 1
        return Expr(self.basis, [self.terms[i]], 0)

    def with_basis(self, new):
        """Same expression with every reference replaced by `new` (arguments.py: `ul._basis = ...`,
        used by the IMEX integrators to turn u into a TrialFunction)."""
        return Expr(new, [[(s, c, d, new) for (s, c, d, r) in t] for t in self.terms], self.rank)


class BasisFunction(Expr):
    argument = 0

    def __init__(self, space):
        self.space = space
        nd = space.dimensions
        ncomp = space.num_components() if space.is_composite_space else 1
        terms = [[(1.0, c, (0,)*nd, self)] for c in range(ncomp)]
        Expr.__init__(self, self, terms, 1 if space.is_composite_space else 0)

    def function_space(self):
        return self.space


class TestFunction(BasisFunction):
    """arguments.py:1095-1147"""
    argument = 0
    __test__ = False


class TrialFunction(BasisFunction):
    """arguments.py:1149-1202"""
    argument = 1


def as_expr(u):
    if isinstance(u, Expr):
        return u
    if isinstance(u, Function):
        space = u.function_space()
        nd = space.dimensions
        ncomp = space.num_components() if space.is_composite_space else 1
        return Expr(u, [[(1.0, c, (0,)*nd, u)] for c in range(ncomp)], 1 if space.is_composite_space else 0)
    raise TypeError(f'cannot interpret {type(u)} as an expression')


def Dx(u, x, k=1):
    """k'th derivative along axis x (forms/operators.py `Dx`)."""
    u = as_expr(u)
    terms = []
    for t in u.terms:
        terms.append([(s, c, tuple(d[i]+(k if i == x else 0) for i in range(len(d))), r) for (s, c, d, r) in t])
    return Expr(u.basis, terms, u.rank)


def grad(u):
    """Gradient: scalar -> vector; vector -> tensor with component (i, j) = d u_i / d x_j
    stored at index i*nd + j."""
    u = as_expr(u)
    nd = u.dimensions
    if u.rank == 0:
        return Expr(u.basis, [Dx(u, j, 1).terms[0] for j in range(nd)], 1)
    assert u.rank == 1, 'grad of scalars and vectors only'
    terms = []
    for i in range(nd):
        ui = Expr(u.basis, [u.terms[i]], 0)
        for j in range(nd):
            terms.append(Dx(ui, j, 1).terms[0])
    return Expr(u.basis, terms, 2)


def div(u):
    """Divergence: vector -> scalar; tensor -> vector (contraction over the last index)."""
    u = as_expr(u)
    nd = u.dimensions
    if u.rank == 1:
        terms = []
        for i in range(nd):
            terms += Dx(Expr(u.basis, [u.terms[i]], 0), i, 1).terms[0]
        return Expr(u.basis, [terms], 0)
    assert u.rank == 2
    out = []
    for i in range(nd):
        terms = []
        for j in range(nd):
            terms += Dx(Expr(u.basis, [u.terms[i*nd+j]], 0), j, 1).terms[0]
        out.append(terms)
    return Expr(u.basis, out, 1)


def curl(u):
    u = as_expr(u)
    assert u.rank == 1
    if u.dimensions == 3:
        w0 = Dx(u[2], 1, 1) - Dx(u[1], 2, 1)
        w1 = Dx(u[0], 2, 1) - Dx(u[2], 0, 1)
        w2 = Dx(u[1], 0, 1) - Dx(u[0], 1, 1)
        return Expr(u.basis, [w0.terms[0], w1.terms[0], w2.terms[0]], 1)
    assert u.dimensions == 2
    w = Dx(u[1], 0, 1) - Dx(u[0], 1, 1)
    return Expr(u.basis, [w.terms[0]], 0)


def _scalar_space(space, comp):
    return space.flatten()[comp] if space.is_composite_space else space


def _bilinear(test, trial):
    """All TPMatrices of sum_c (trial_expr[c], test_expr[c])."""
    from .matrices import TPMatrix
    mats = []
    ncomp = len(test.terms)
    assert len(trial.terms) == ncomp
    for c in range(ncomp):
        for (s0, c0, d0, r0) in test.terms[c]:
            for (s1, c1, d1, r1) in trial.terms[c]:
                tsp = _scalar_space(r0.function_space(), c0)
                usp = _scalar_space(r1.function_space(), c1)
                mats.append(TPMatrix.assemble(tsp, usp, d0, d1, s0*s1, global_index=(c0, c1)))
    return mats


class Inner:
    """Cached linear form (expr(Function), v): the matrices are assembled once, every call is
    a handful of mat-vec launches (forms/inner.py:514-566).  Terms that share the source
    Function component and the test component are fused into one line operator."""

    def __init__(self, v, uh):
        from .la import LinearFormOperator
        assert isinstance(v, TestFunction)
        self.space = v.function_space()
        self.output_array = Function(self.space)
        self.parts = []
        if isinstance(uh, Array):
            self.parts.append(('array', uh, None, None))
            return
        e = as_expr(uh)
        assert e.argument == 2, 'Inner needs an expression of known Functions'
        ncomp = len(v.terms)
        assert len(e.terms) == ncomp, 'test and trial expressions must have the same number of components'
        groups = {}
        for c in range(ncomp):
            for (s0, c0, d0, r0) in v.terms[c]:
                for (s1, c1, d1, r1) in e.terms[c]:
                    from .matrices import TPMatrix
                    tsp = _scalar_space(self.space, c0)
                    usp = _scalar_space(r1.function_space(), c1)
                    M = TPMatrix.assemble(tsp, usp, d0, d1, s0*s1)
                    groups.setdefault((c0, id(r1), c1), [r1, c0, c1, []])[3].append(M)
        for (r1, c0, c1, mats) in groups.values():
            self.parts.append(('op', r1, (c0, c1), LinearFormOperator(mats)))

    def __call__(self):
        out = ops.plain(self.output_array)
        composite = self.space.is_composite_space
        ncomp = out.shape[0] if composite else 1
        written = [False]*ncomp          # the first term of a component overwrites: no zero fill
        for kind, ref, idx, op in self.parts:
            if kind == 'array':
                tmp = self.space.scalar_product(ref, Function(self.space))
                if all(written):
                    out.add_(ops.plain(tmp))
                else:
                    for c in range(ncomp):
                        dst = out[c] if composite else out
                        src = ops.plain(tmp)[c] if composite else ops.plain(tmp)
                        if written[c]:
                            dst.add_(src)
                        else:
                            dst.copy_(src)
                            written[c] = True
                continue
            c0, c1 = idx
            src = ops.plain(ref)
            if ref.function_space().is_composite_space:
                src = src[c1]
            dst = out[c0] if composite else out
            k = c0 if composite else 0
            op.matvec(src, dst, accumulate=written[k])
            written[k] = True
        for c in range(ncomp):
            if not written[c]:
                (out[c] if composite else out).zero_()
        return self.output_array

    def __add__(self, other):
        assert isinstance(other, Inner) and other.space is self.space
        new = Inner.__new__(Inner)
        new.space = self.space
        new.output_array = Function(self.space)
        new.parts = self.parts+other.parts
        return new

    __radd__ = __add__


class Project:
    """Cached projection of an expression onto T (forms/project.py:172-272): right hand side
    through `Inner`, then the mass matrix of T (banded solve along the non-periodic axis).

    A single term `s * d^d u / dx^d` of a known Function u that lives in T itself, differentiated along
    periodic axes only, needs neither: (v, D u) = (B (x) D_F) u and the projection applies B^-1, so the
    result is the spectral array of u times the Fourier factors -- one point-wise pass instead of a
    banded mat-vec and a banded solve (six of the nine velocity gradients of demo/ChannelFlow.py:69-77).
    SHB_PROJECT_SHORTCUT=0 keeps the generic path."""

    def __init__(self, uh, T, output_array=None):
        self.T = T
        self.output_array = Function(T) if output_array is None else output_array
        self._short = self._fourier_shortcut(uh, T)
        self.rhs = Inner(TestFunction(T), uh) if self._short is None else None

    @staticmethod
    def _fourier_shortcut(uh, T):
        import os
        if os.environ.get('SHB_PROJECT_SHORTCUT', '1') == '0' or T.is_composite_space:
            return None
        e = as_expr(uh)
        if len(e.terms) != 1 or len(e.terms[0]) != 1:
            return None
        s, c, d, r = e.terms[0][0]
        if not isinstance(r, Function):
            return None
        usp = _scalar_space(r.function_space(), c)
        if len(usp.bases) != len(T.bases):
            return None
        for ax, (tb, ub) in enumerate(zip(T.bases, usp.bases)):
            if type(tb) is not type(ub) or hash(tb) != hash(ub):
                return None
            if tb.family() != 'fourier' and d[ax] != 0:
                return None
        sl = T.local_slice(True)
        nd = len(T.bases)
        arr = None
        for ax, tb in enumerate(T.bases):
            if tb.family() != 'fourier' or d[ax] == 0:
                continue
            num = np.asarray(tb.spectral_matrix(0, d[ax], tb).scaled_diagonal(0))
            den = np.asarray(tb.spectral_matrix(0, 0, tb).scaled_diagonal(0))
            if num.shape != den.shape or not np.all(den != 0):
                return None
            f = num/den
            if f.shape[0] != tb.spectral_len():
                return None
            shape = [1]*nd
            shape[ax] = sl[ax].stop-sl[ax].start
            part = f[sl[ax]].reshape(shape)
            arr = part if arr is None else arr*part
        if arr is None:
            return None
        arr = np.asarray(arr*s, dtype=complex)
        if np.abs(arr.imag).max() == 0:
            mask, ti = arr.real, False
        elif np.abs(arr.real).max() == 0:
            mask, ti = arr.imag, True
        else:
            return None
        first = min(ax for ax in range(nd) if arr.shape[ax] > 1)
        tshape = tuple(T.shape(True))[first:]
        mask = np.ascontiguousarray(np.broadcast_to(mask.reshape(arr.shape[first:]), tshape), dtype=float)
        return r, c, ops.as_device_tensor(mask), ti

    def __call__(self):
        from .la import mass_solve
        if self._short is not None:
            r, c, mask, ti = self._short
            src = ops.plain(r)
            if r.function_space().is_composite_space:
                src = src[c]
            ops.scale_mask(ops.plain(self.output_array), src, mask, times_i=ti)
            return self.output_array
        rhs = self.rhs()
        flat = self.T.flatten() if self.T.is_composite_space else [self.T]
        if all(all(b.family() == 'fourier' for b in s.bases) for s in flat):
            ops.plain(self.output_array).copy_(ops.plain(rhs))
        else:
            mass_solve(self.T, rhs, self.output_array)
        return self.output_array


def _linear(test, expr, output_array):
    """One-off evaluation of (expr(Function), test)."""
    I = Inner(test, expr)
    res = I()
    if output_array is None:
        return res
    ops.plain(output_array).copy_(ops.plain(res))
    return output_array


def inner(expr0, expr1, output_array=None, assemble=None, kind=None, **kw):
    """Weighted inner product (forms/inner.py:16-511).

    * test x trial            -> list of TPMatrix (a single one is returned bare)
    * test x expr(Function)   -> Function of coefficients (linear form, matvec)
    * test x Array            -> scalar product of the Array (forms/inner.py:361-364)
    * number x Array/Function -> integral over the domain (forms/inner.py:182-199)
    """
    if isinstance(expr0, numbers.Number) or isinstance(expr1, numbers.Number):
        a, arr = (expr0, expr1) if isinstance(expr0, numbers.Number) else (expr1, expr0)
        return a*_integrate(arr)
    if isinstance(expr0, Array) or isinstance(expr1, Array):
        test, arr = (expr1, expr0) if isinstance(expr0, Array) else (expr0, expr1)
        assert isinstance(test, TestFunction)
        space = test.function_space()
        if output_array is None:
            output_array = Function(space)
        return space.scalar_product(arr, output_array)
    e0, e1 = as_expr(expr0), as_expr(expr1)
    if e0.argument > e1.argument:
        e0, e1 = e1, e0
    assert e0.argument == 0, 'one argument must be a TestFunction expression'
    if e1.argument == 1:
        mats = _bilinear(e0, e1)
        return mats[0] if len(mats) == 1 else mats
    return _linear(e0, e1, output_array)


def _integrate(arr):
    """(1, u) = integral of u over the true domain: un-weighted quadrature weights of every
    axis divided by the domain factors (forms/inner.py:182-199 with utilities `dx` :26-61).
    A Function is transformed to physical space first."""
    if isinstance(arr, Function):
        arr = arr.backward()
    space = arr.function_space()
    t = ops.plain(arr)
    sc = space.flatten()[0] if space.is_composite_space else space
    nd = sc.dimensions
    sl = sc.local_slice(False)
    df = 1.0
    for ax, b in enumerate(sc.bases):
        w = b.points_and_weights(b.physical_len(), weighted=False)[1]
        w = np.broadcast_to(w, (b.physical_len(),))[sl[ax]]
        df *= float(b.domain_factor())
        shape = [1]*nd
        shape[ax] = len(w)
        t = t*torch.as_tensor(np.array(w, dtype=float), device=t.device).view(shape)
    val = t.sum()
    val = space.comm.allreduce(val)
    val = val.item()/df
    return val.real if isinstance(val, complex) and abs(val.imag) < 1e-14*max(1.0, abs(val)) else val


def project(uh, T, output_array=None, fill=True, use_to_ortho=True, use_assign=True):
    """Projection onto T (forms/project.py:16-170).

    Array -> forward transform; expression of a Function -> solve B u = (expr, v).
    """
    if output_array is None:
        output_array = Function(T)
    if isinstance(uh, Array):
        return T.forward(uh, output_array)
    return Project(uh, T, output_array=output_array)()
