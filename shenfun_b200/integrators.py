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

__all__ = ('IntegratorBase', 'RK4')


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
        """Diagonal of the linear operator as a real device tensor, or None
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
