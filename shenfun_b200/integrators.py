"""Time integrators on device arrays.

Host-side mirror of shenfun/utilities/integrators.py: `IntegratorBase` :56-118 and
`RK4` :520-600.  The per-stage algebra (dU += L*u_hat, the two axpy updates) is one
fused kernel (`shb_rk4_stage`) instead of five numpy passes.
"""
import types
import numpy as np
import torch

from .arrays import Function
from .forms import TrialFunction, TestFunction, inner, Expr
from .matrices import TPMatrix, get_simplified_tpmatrices
from . import ops

__all__ = ('IntegratorBase', 'RK4', 'PDEIMEXRK', 'IMEXRK011', 'IMEXRK111', 'IMEXRK222', 'IMEXRK443')


class IntegratorBase:
    """integrators.py:56-118"""

    def __init__(self, T, L=None, N=None, update=None, **params):
        _p = {'dt': 0}
        _p.update(params)
        self.params = _p
        self.T = T
        if L is not None:
            self.LinearRHS = types.MethodType(L, self)
        if N is not None:
            self.NonlinearRHS = types.MethodType(N, self)
        if update is not None:
            self.update = types.MethodType(update, self)

    def update(self, u, u_hat, t, tstep, **par):
        pass

    def LinearRHS(self, *args, **kwargs):
        return 0

    def NonlinearRHS(self, *args, **kwargs):
        return 0

    def setup(self, dt):
        pass

    def _linear_operator(self):
        """Diagonal of the linear operator as a real (or, for odd derivatives, complex) device tensor, or None
        (integrators.py:569-585)."""
        ut = TrialFunction(self.T)
        vt = TestFunction(self.T)
        L = self.LinearRHS(ut, **self.params)
        if isinstance(L, Expr):
            L = inner(vt, L)
            L = get_simplified_tpmatrices(L)[0]
        if isinstance(L, list):
            L = L[0]
        if isinstance(L, TPMatrix):
            assert L.is_diagonal(), 'RK4 needs a diagonal linear operator (Fourier spaces)'
            d = L.diagonal_array()
            shape = self.T.flatten()[0].shape(True) if self.T.is_composite_space else self.T.shape(True)
            d = d.expand(shape).contiguous()
            if d.is_complex() and bool((d.imag != 0).any()):
                return d.contiguous()            # complex operator: shb_rk4_stage_cplx
            return d.real.contiguous() if d.is_complex() else d
        if isinstance(L, torch.Tensor):
            return L
        return None


class RK4(IntegratorBase):
    """Classical 4th order Runge-Kutta (integrators.py:520-600)."""

    def __init__(self, T, L=None, N=None, update=None, **params):
        IntegratorBase.__init__(self, T, L=L, N=N, update=update, **params)
        self.U_hat0 = Function(T)
        self.U_hat1 = Function(T)
        self.dU = Function(T)
        self.a = np.array([1./6., 1./3., 1./3., 1./6.])
        self.b = np.array([0.5, 0.5, 1.])

    def setup(self, dt):
        self.params['dt'] = dt

    def solve(self, u, u_hat, dt, trange):
        if abs(self.params['dt']-dt) > 1e-12:
            self.setup(dt)
        t, end_time = trange
        tstep = 0
        L = self._linear_operator()
        uh, U0, U1 = ops.plain(u_hat), ops.plain(self.U_hat0), ops.plain(self.U_hat1)
        while t < end_time-1e-8:
            t += dt
            tstep += 1
            U0.copy_(uh)
            U1.copy_(uh)
            for rk in range(4):
                dU = self.NonlinearRHS(u, u_hat, self.dU, **self.params)
                ops.rk4_stage(ops.plain(dU), uh, U0, U1, L, self.a[rk]*dt,
                              self.b[rk]*dt if rk < 3 else 0.0, rk == 3)
            self.update(u, u_hat, t, tstep, **self.params)
        return u_hat


class PDEIMEXRK:
    """Implicit-explicit Runge-Kutta for du/dt = N + L u (Ascher, Ruuth, Spiteri 1997):
    shenfun/utilities/integrators.py:702-817.  One linear solve per stage (the diagonal of the
    implicit tableau is constant), right hand sides through cached `Inner` mat-vecs."""

    def __init__(self, v, u, L, N, dt, solver=None, name='U-equation', latex=None):
        from .forms import as_expr, Expr
        from . import la
        self.v = v
        self.u = as_expr(u)
        self.u_ = self.u.basis          # the Function that holds the solution
        self.L = L
        self.N = N
        self.dt = dt
        self._solver = solver if solver is not None else la.SolverGeneric1ND
        self.solvers = []
        self.name = name
        self.latex = latex
        T = self.T = v.function_space()
        dev = ops.plain(self.u_).device
        shape = tuple(T.shape(True))
        self.rhs = Function(T)
        self.Krhs = torch.zeros((self.steps(),)+shape, dtype=torch.complex128 if T.spectral_dtype.is_complex
                                else torch.float64, device=dev)
        self.Lrhs = torch.zeros((max(self.steps()-1, 1),)+shape, dtype=self.Krhs.dtype, device=dev)
        self.mask = T._device_mask() if hasattr(T, '_device_mask') else None

    @classmethod
    def steps(cls):
        raise NotImplementedError

    def stages(self):
        raise NotImplementedError

    def assemble(self):
        from .forms import inner, Inner, TrialFunction, Expr
        a = self.stages()[0]
        dt = self.dt
        ul = self.u.with_basis(TrialFunction(self.u_.function_space()))
        mats = inner(self.v, ul - (dt*a[1, 1])*self.L(ul))
        self.solvers.append(self._solver(mats))
        self.linear_rhs = Inner(self.v, self.L(self.u))
        self.u0_rhs = Inner(self.v, self.u)
        if isinstance(self.N, (list, tuple)):
            parts = [Inner(self.v, f) for f in self.N]
            nl = parts[0]
            for p in parts[1:]:
                nl = nl+p
            self.nonlinear_rhs = nl
        else:
            self.nonlinear_rhs = Inner(self.v, self.N)

    def compute_rhs(self, rk=0):
        """rhs = (u0, v) + dt sum_j b[rk+1, j] N_j + dt sum_j a[rk+1, j+1] L_j, masked
        (integrators.py:692-696, 803-812) in ONE pass (`shb_lincomb`); the newest N / L evaluations are
        written to their slots in the same pass when a later stage needs them."""
        a, b = self.stages()[:2]
        nsteps = self.steps()
        n_new = ops.plain(self.nonlinear_rhs())
        if rk == 0:
            self.u0_rhs()        # only at the start of a step
        terms, coefs, saves = [ops.plain(self.u0_rhs.output_array)], [1.0], []
        for j in range(0, rk):
            if b[rk+1, j] != 0:
                terms.append(self.Krhs[j])
                coefs.append(self.dt*b[rk+1, j])
        terms.append(n_new)
        coefs.append(self.dt*b[rk+1, rk])
        if rk+1 < nsteps:
            saves.append((len(terms)-1, self.Krhs[rk]))
        if rk > 0:
            l_new = ops.plain(self.linear_rhs())
            for j in range(0, rk-1):
                if a[rk+1, j+1] != 0:
                    terms.append(self.Lrhs[j])
                    coefs.append(self.dt*a[rk+1, j+1])
            terms.append(l_new)
            coefs.append(self.dt*a[rk+1, rk])
            if rk+1 < nsteps:
                saves.append((len(terms)-1, self.Lrhs[rk-1]))
        ops.lincomb(ops.plain(self.rhs), terms, coefs, self.mask, saves)
        return self.rhs

    def solve_step(self, rk=0):
        return self.solvers[0](self.rhs, self.u_)


class IMEXRK011(PDEIMEXRK):
    @classmethod
    def steps(cls):
        return 1

    def stages(self):
        a = np.array([[0, 0], [0, 0]])
        b = np.array([[0, 0], [1, 0]])
        return a, b, (1, 0)


class IMEXRK111(PDEIMEXRK):
    @classmethod
    def steps(cls):
        return 1

    def stages(self):
        a = np.array([[0, 0], [0, 1]])
        b = np.array([[0, 0], [1, 0]])
        return a, b, (0, 1)


class IMEXRK222(PDEIMEXRK):
    @classmethod
    def steps(cls):
        return 2

    def stages(self):
        gamma = (2-np.sqrt(2))/2
        delta = 1-1/(2*gamma)
        a = np.array([[0, 0, 0], [0, gamma, 0], [0, 1-gamma, gamma]])
        b = np.array([[0, 0, 0], [gamma, 0, 0], [delta, 1-delta, 0]])
        return a, b, (0, gamma, 1)


class IMEXRK443(PDEIMEXRK):
    @classmethod
    def steps(cls):
        return 4

    def stages(self):
        a = np.array([[0, 0, 0, 0, 0], [0, 1/2, 0, 0, 0], [0, 1/6, 1/2, 0, 0], [0, -1/2, 1/2, 1/2, 0],
                      [0, 3/2, -3/2, 1/2, 1/2]])
        b = np.array([[0, 0, 0, 0, 0], [1/2, 0, 0, 0, 0], [11/18, 1/18, 0, 0, 0], [5/6, -5/6, 1/2, 0, 0],
                      [1/4, 7/4, 3/4, -7/4, 0]])
        return a, b, (0, 1/2, 2/3, 1/2, 1)
