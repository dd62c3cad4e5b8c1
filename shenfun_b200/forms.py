"""A small Cartesian form language: TestFunction / TrialFunction, Dx / grad / div /
curl, `inner` and `project`.

Host-side mirror of the *interface* of shenfun/forms/arguments.py:420-1202
(`Expr`, `TestFunction`, `TrialFunction`), forms/operators.py (`Dx`, `grad`,
`div`, `curl`), forms/inner.py:16-511 (`inner`) and forms/project.py:16-170
(`project`) for the cases the hot path needs: constant-coefficient terms on
Cartesian tensor-product spaces.  Bilinear forms give `TPMatrix` objects (1-D
matrices per axis, see matrices.py); linear forms are evaluated on the device.
Curvilinear coordinates and sympy-valued coefficients are out of scope
(SURVEY.md section 2, rows 12 and 18).
"""
import numbers
import numpy as np
import torch

from .arrays import Function, Array
from . import ops

__all__ = ('Expr', 'TestFunction', 'TrialFunction', 'Dx', 'grad', 'div', 'curl', 'inner', 'project')


class Expr:
    """Sum of scaled derivatives of the components of one basis.

    `terms[c]` is the list of `(scale, comp, derivs)` making up output component
    `c`: `scale * d^{derivs}/dx^{derivs} basis[comp]`.  `basis` is a
    TestFunction / TrialFunction, or a `Function` holding expansion coefficients.
    """

    def __init__(self, basis, terms, rank):
        self.basis = basis
        self.terms = terms
        self.rank = rank           # 0: scalar expression, 1: vector

    @property
    def argument(self):
        b = self.basis
        if isinstance(b, BasisFunction):
            return b.argument
        return 2               # a Function: known coefficients

    def function_space(self):
        b = self.basis
        return b.function_space() if isinstance(b, (BasisFunction, Function)) else b.space

    @property
    def dimensions(self):
        return self.function_space().dimensions

    def num_components(self):
        return len(self.terms)

    def _scaled(self, a):
        return Expr(self.basis, [[(s*a, c, d) for (s, c, d) in t] for t in self.terms], self.rank)

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
        assert other.basis is self.basis and other.rank == self.rank
        return Expr(self.basis, [a+b for a, b in zip(self.terms, other.terms)], self.rank)

    def __sub__(self, other):
        return self + (-as_expr(other))

    def __getitem__(self, i):
        assert self.rank == 1
        return Expr(self.basis, [self.terms[i]], 0)


class BasisFunction(Expr):
    argument = 0

    def __init__(self, space):
        self.space = space
        nd = space.dimensions
        ncomp = space.num_components() if space.is_composite_space else 1
        terms = [[(1.0, c, (0,)*nd)] for c in range(ncomp)]
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
        return Expr(u, [[(1.0, c, (0,)*nd)] for c in range(ncomp)], 1 if space.is_composite_space else 0)
    raise TypeError(f'cannot interpret {type(u)} as an expression')


def Dx(u, x, k=1):
    """k'th derivative along axis x (forms/operators.py `Dx`)."""
    u = as_expr(u)
    terms = []
    for t in u.terms:
        terms.append([(s, c, tuple(d[i]+(k if i == x else 0) for i in range(len(d)))) for (s, c, d) in t])
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
        for (s0, c0, d0) in test.terms[c]:
            for (s1, c1, d1) in trial.terms[c]:
                tsp = _scalar_space(test.function_space(), c0)
                usp = _scalar_space(trial.function_space(), c1)
                mats.append(TPMatrix.assemble(tsp, usp, d0, d1, s0*s1, global_index=(c0, c1)))
    return mats


def _linear(test, expr, output_array):
    """Evaluate (expr(Function), test) -> coefficients, component by component."""
    from .matrices import TPMatrix
    uf = expr.basis
    tspace = test.function_space()
    if output_array is None:
        output_array = Function(tspace)
    out = ops.plain(output_array)
    uv = ops.plain(uf)
    out.zero_()
    ncomp = len(test.terms)
    assert len(expr.terms) == ncomp, 'test and trial expressions must have the same number of components'
    for c in range(ncomp):
        for (s0, c0, d0) in test.terms[c]:
            dst = out[c0] if tspace.is_composite_space else out
            for (s1, c1, d1) in expr.terms[c]:
                tsp = _scalar_space(tspace, c0)
                usp = _scalar_space(uf.function_space(), c1)
                M = TPMatrix.assemble(tsp, usp, d0, d1, s0*s1)
                src = uv[c1] if uf.function_space().is_composite_space else uv
                M.matvec(src, dst, accumulate=True)
    return output_array


def inner(expr0, expr1, output_array=None, assemble=None, kind=None, **kw):
    """Weighted inner product (forms/inner.py:16-511).

    * test x trial            -> list of TPMatrix (a single one is returned bare)
    * test x expr(Function)   -> Function of coefficients (linear form, matvec)
    * test x Array            -> scalar product of the Array (forms/inner.py:361-364)
    * number x Array          -> integral over the domain (forms/inner.py:268)
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
    """(1, u) with the quadrature weights of every axis (utilities/__init__.py:26-61 `dx`
    up to the weight function; forms/inner.py:268)."""
    space = arr.function_space()
    t = ops.plain(arr)
    nd = space.dimensions
    sl = space.local_slice(False)
    for ax, b in enumerate(space.flatten()[0].bases if space.is_composite_space else space.bases):
        w = b.points_and_weights(b.physical_len())[1]
        w = np.broadcast_to(w, (b.physical_len(),))[sl[ax]]/b.domain_factor() if b.family() != 'fourier' \
            else np.broadcast_to(w, (b.physical_len(),))[sl[ax]]
        shape = [1]*nd
        shape[ax] = len(w)
        t = t*torch.as_tensor(np.ascontiguousarray(w), device=t.device).view(shape)
    val = t.sum()
    val = space.comm.allreduce(val)
    val = val.item()
    return val.real if isinstance(val, complex) and abs(val.imag) < 1e-14*max(1.0, abs(val)) else val


def project(uh, T, output_array=None, fill=True, use_to_ortho=True, use_assign=True):
    """Projection onto T (forms/project.py:16-170).

    Array -> forward transform; expression of a Function -> solve B u = (expr, v).
    """
    if output_array is None:
        output_array = Function(T)
    if isinstance(uh, Array):
        return T.forward(uh, output_array)
    e = as_expr(uh)
    v = TestFunction(T)
    rhs = _linear(v, e, None if False else Function(T))
    flat = T.flatten() if T.is_composite_space else [T]
    if all(s.is_orthogonal() and all(b.family() == 'fourier' for b in s.bases) for s in flat):
        # Fourier mass matrix is the identity (weight 1/L, fourier/bases.py:79-81)
        ops.plain(output_array).copy_(ops.plain(rhs))
        return output_array
    from .matrices import mass_solve
    return mass_solve(T, rhs, output_array)
