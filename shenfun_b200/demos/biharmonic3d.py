"""3-D biharmonic equation, one clamped direction: demo/biharmonic3D.py on the device.

    nabla^4 u = f,   u = u' = 0 at x = +-1,  periodic in y and z

Legendre ShenBiharmonic x Fourier x Fourier (BASELINE config 4 at 256^3).  The Legendre
transforms run as fp64 tensor-core contractions; the solve is one pentadiagonal-type banded
system per Fourier wavenumber pair (`la.SolverGeneric1ND`, reference demo :58-70).
"""
import numpy as np
import sympy as sp
import torch

import shenfun_b200 as sf
from shenfun_b200 import la


def manufactured():
    x, y, z = sp.symbols('x,y,z', real=True)
    ue = (sp.sin(4*y)+sp.cos(2*z))*(1-x**2)**2*sp.cos(2*x)
    fe = (ue.diff(x, 4)+ue.diff(y, 4)+ue.diff(z, 4)+2*ue.diff(x, 2, y, 2)+2*ue.diff(x, 2, z, 2)
          + 2*ue.diff(y, 2, z, 2))
    return ue, fe


def build(N, family='L'):
    SB = sf.FunctionSpace(N[0], family=family, bc=(0, 0, 0, 0))
    K1 = sf.FunctionSpace(N[1], family='F', dtype='D')
    K2 = sf.FunctionSpace(N[2], family='F', dtype='d')
    T = sf.TensorProductSpace(sf.comm, (SB, K1, K2))
    u = sf.TrialFunction(T)
    v = sf.TestFunction(T)
    matrices = sf.inner(v, sf.div(sf.grad(sf.div(sf.grad(u)))))
    # Chebyshev: the dedicated solver for the upper-triangular fourth-derivative matrix
    # (chebyshev/la.py:354-558); Legendre: banded per-wavenumber systems
    if family.lower().startswith('c'):
        return T, la.Biharmonic(matrices)
    return T, la.SolverGeneric1ND(matrices)


def main(N=(32, 32, 32), family='L'):
    ue, fe = manufactured()
    T, H = build(N, family)
    v = sf.TestFunction(T)
    fj = sf.Array(T, buffer=fe)
    f_hat = sf.inner(v, fj, output_array=sf.Function(T))
    u_hat = H(f_hat, sf.Function(T))
    uq = u_hat.backward()
    uj = sf.Array(T, buffer=ue)
    err = torch.sqrt(((uj.v-uq.v)**2).mean()).item()
    return err, u_hat


def bench(n=256, family='L', reps=5):
    T, H = build((n, n, n), family)
    u = sf.Array(T)
    u.v.normal_()
    f_hat = sf.Function(T)
    u_hat = sf.Function(T)

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
    b0 = T.bases[0]
    S = tuple(T.shape(True))
    w = torch.randn(S, dtype=torch.complex128, device='cuda')
    w2 = torch.empty_like(w)
    inner_ = S[1]*S[2]
    import os
    from shenfun_b200 import ops
    ms_gemm = t(lambda: b0._backward_axis(w, w2, 1, inner_))
    os.environ['SHB_GEMM_PARITY'] = '0'          # the dense contraction (what cuBLAS computes below)
    try:
        ms_dense = t(lambda: b0._backward_axis(w, w2, 1, inner_))
    finally:
        del os.environ['SHB_GEMM_PARITY']
    flops = 2.0*n*n*inner_*2
    parity = family == 'L' and ops._parity_halves(b0._operators()[1]) is not None
    ms_fwd = t(lambda: T.forward(u, f_hat))
    ms_solve = t(lambda: H(f_hat, u_hat))
    ms_chain = t(lambda: T.backward(H(T.scalar_product(u, f_hat), u_hat), u))
    # denominator for the FP64 tensor-pipe utilisation: a library DGEMM measured on the same device
    a = torch.randn(4096, 4096, dtype=torch.float64, device='cuda')
    bm = torch.randn(4096, 4096, dtype=torch.float64, device='cuda')
    ms_ref = t(lambda: torch.matmul(a, bm))
    dgemm_peak = 2*4096.0**3/ms_ref/1e9
    del a, bm
    # cuBLAS on the stage's own shape: (n x n) operator times n x (n*(n/2+1)*2) real columns
    W = n*(n//2+1)*2
    ar = torch.randn(n, n, dtype=torch.float64, device='cuda')
    xr = torch.randn(n, W, dtype=torch.float64, device='cuda')
    cr = torch.empty(n, W, dtype=torch.float64, device='cuda')
    ms_same = t(lambda: torch.matmul(ar, xr, out=cr))
    same_tflops = 2.0*n*n*W/ms_same/1e9
    del ar, xr, cr
    return {'workload': f'{family} ShenBiharmonic x Fourier^2 {n}^3', 'legendre_stage_ms': ms_gemm,
            'legendre_stage_kernel': 'gemm_axis_parity_kernel (even/odd split: half the dense flops)' if parity else 'gemm_axis_kernel',
            'legendre_stage_dense_equivalent_tflops': flops/ms_gemm/1e9,
            'legendre_stage_executed_tflops': flops*(0.5 if parity else 1.0)/ms_gemm/1e9,
            'legendre_stage_dense_kernel_ms': ms_dense,
            'legendre_stage_tflops': flops/ms_dense/1e9, 'cublas_dgemm_4096_tflops': dgemm_peak,
            'legendre_stage_frac_of_dgemm_peak': flops/ms_dense/1e9/dgemm_peak,
            'cublas_dgemm_same_shape_tflops': same_tflops, 'legendre_stage_frac_of_cublas_same_shape': flops/ms_dense/1e9/same_tflops,
            'forward_ms': ms_fwd, 'solve_ms': ms_solve,
            'scalar_product_solve_backward_ms': ms_chain}
