"""Periodic Poisson problem  div grad u = f  on (0, 2 pi)^3 with three Fourier bases -- the
device twin of demo/fourier_poisson3D.py:1-58 (BASELINE config 0).  Manufactured solution
u = cos(4x) + sin(4y) + sin(6z); the (0,0,0) mode is pinned to zero."""
import numpy as np
import sympy as sp

import shenfun_b200 as sf
from shenfun_b200 import la


def solve(N=(14, 15, 16), comm=None):
    """Returns (u_hat, L2 error) exactly as the reference demo computes them (:31-52)."""
    comm = sf.comm if comm is None else comm
    x, y, z = sp.symbols('x,y,z', real=True)
    ue = sp.cos(4*x)+sp.sin(4*y)+sp.sin(6*z)
    fe = ue.diff(x, 2)+ue.diff(y, 2)+ue.diff(z, 2)
    K0 = sf.FunctionSpace(N[0], 'F', dtype='D')
    K1 = sf.FunctionSpace(N[1], 'F', dtype='D')
    K2 = sf.FunctionSpace(N[2], 'F', dtype='d')
    T = sf.TensorProductSpace(comm, (K0, K1, K2))
    u = sf.TrialFunction(T)
    v = sf.TestFunction(T)
    fj = sf.Array(T, buffer=fe)
    f_hat = sf.inner(v, fj, output_array=sf.Function(T))
    A = sf.inner(v, sf.div(sf.grad(u)))
    sol = la.SolverDiagonal(A)
    u_hat = sol(f_hat, sf.Function(T), constraints=((0, 0),))
    uq = u_hat.backward()
    uj = sf.Array(T, buffer=ue)
    error = np.sqrt(sf.inner(1, (uj-uq)**2))
    # evaluation at arbitrary points (demo :54-58)
    point = np.array([[0.1, 0.5], [0.5, 0.6], [0.1, 0.2]])
    p = T.eval(point, u_hat)
    ul = sp.lambdify((x, y, z), ue)
    assert np.allclose(p, ul(*point)), (p, ul(*point))
    return u_hat, float(abs(error))
