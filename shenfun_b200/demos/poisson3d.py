"""3-D Poisson equation, one non-periodic direction: demo/poisson3D.py on the device.

    nabla^2 u = f,   u(+-1, y, z) = 0,  periodic in y and z

Chebyshev or Legendre ShenDirichlet x Fourier x Fourier (BASELINE config 3 at 512^3 with
family 'C').  Same steps as the reference demo (:62-104): right hand side by the weighted
scalar product, left hand side `inner(v, div(grad(u)))`, one banded (Legendre:
`la.SolverGeneric1ND`) or Chebyshev `la.Helmholtz` solve per Fourier wavenumber pair,
backward transform, L2 error against the manufactured solution.
Homogeneous Dirichlet data and the Galerkin test space are used (the demo's Petrov-Galerkin
test space and x,z-dependent boundary data are outside the device hot path).
"""
import numpy as np
import sympy as sp
import torch

import shenfun_b200 as sf
from shenfun_b200 import la, ops


def manufactured():
    x, y, z = sp.symbols('x,y,z', real=True)
    ue = (sp.cos(4*y)+sp.sin(2*z))*(1-x**2)*sp.sin(2*x+0.3)
    fe = ue.diff(x, 2)+ue.diff(y, 2)+ue.diff(z, 2)
    return ue, fe


def build(N, family='C', alpha=0.0):
    """Spaces, matrices and solver for nabla^2 u - alpha u = f."""
    SD = sf.FunctionSpace(N[0], family=family, bc=(0, 0))
    K1 = sf.FunctionSpace(N[1], family='F', dtype='D')
    K2 = sf.FunctionSpace(N[2], family='F', dtype='d')
    T = sf.TensorProductSpace(sf.comm, (SD, K1, K2))
    u = sf.TrialFunction(T)
    v = sf.TestFunction(T)
    expr = sf.div(sf.grad(u)) if alpha == 0 else sf.div(sf.grad(u))-alpha*u
    matrices = sf.inner(v, expr)
    if family.lower().startswith('c'):
        H = la.Helmholtz(matrices)
        constraints = ()
    else:
        H = la.SolverGeneric1ND(matrices)
        constraints = ()
    return T, H, constraints


def main(N=(32, 32, 32), family='C'):
    ue, fe = manufactured()
    T, H, constraints = build(N, family)
    v = sf.TestFunction(T)
    fj = sf.Array(T, buffer=fe)
    f_hat = sf.inner(v, fj, output_array=sf.Function(T))
    u_hat = H(f_hat, sf.Function(T), constraints=constraints)
    uq = u_hat.backward()
    uj = sf.Array(T, buffer=ue)
    err = torch.sqrt(((uj.v-uq.v)**2).mean()).item()
    return err, u_hat


def bench(n=512, family='C', reps=5):
    """Solve-only and full-chain timings at n^3 (random spectral right hand side, seed 1234)."""
    T, H, constraints = build((n, n, n), family, alpha=0.0)
    g = torch.Generator(device='cuda').manual_seed(1234)
    f_hat = sf.Function(T)
    fr = torch.view_as_real(f_hat.v)
    fr.copy_(torch.randn(fr.shape, dtype=torch.float64, device='cuda', generator=g))
    f_hat.v[T.bases[0].dim():] = 0
    f_keep = f_hat.copy()
    u_hat = sf.Function(T)
    u = sf.Array(T)
    u.v.copy_(torch.randn(u.shape, dtype=torch.float64, device='cuda', generator=g))

    def t(fn):
        fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)/reps
    ms_solve = t(lambda: H(f_hat, u_hat, constraints=constraints))
    ms_chain = t(lambda: (T.backward(H(T.scalar_product(u, f_hat), u_hat, constraints=constraints), u)))
    # residual check of the solve: || A u - f || / || f ||
    H(f_keep, u_hat, constraints=constraints)
    c = H.matvec(u_hat, sf.Function(T))
    d = T.bases[0].dim()
    res = (torch.linalg.vector_norm(c.v[:d]-f_keep.v[:d])/torch.linalg.vector_norm(f_keep.v[:d])).item()
    nbytes = 2*np.prod(T.shape(True))*16
    return {'workload': f'{family} ShenDirichlet x Fourier^2 {n}^3 Poisson', 'solve_ms': ms_solve,
            'solves_per_s': 1e3/ms_solve, 'solve_algorithmic_GBs': nbytes/ms_solve/1e6,
            'scalar_product_solve_backward_ms': ms_chain, 'residual': res}
