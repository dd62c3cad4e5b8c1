"""CUDA-event timing of the line solvers at the BASELINE sizes (public API, L2 flushed between calls)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import shenfun_b200 as sf
from shenfun_b200 import la, ops

PEAK = 6546.9
flush = torch.empty(64*1024*1024, dtype=torch.float64, device='cuda')


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts)//2]


def helmholtz(N, fam='C'):
    SD = sf.FunctionSpace(N[0], fam, bc=(0, 0))
    T = sf.TensorProductSpace(sf.comm, (SD, sf.FunctionSpace(N[1], 'F', dtype='D'), sf.FunctionSpace(N[2], 'F', dtype='d')))
    u, v = sf.TrialFunction(T), sf.TestFunction(T)
    H = la.Helmholtz(sf.inner(v, sf.div(sf.grad(u))-2.5*u))
    f = sf.Function(T)
    torch.view_as_real(f.v).normal_()
    f.v[SD.dim():] = 0
    x = sf.Function(T)
    ms = timeit(lambda: H(f, x))
    nbytes = 2*f.v.numel()*16
    c = H.matvec(x, sf.Function(T))
    res = (torch.linalg.vector_norm(c.v[:SD.dim()]-f.v[:SD.dim()])/torch.linalg.vector_norm(f.v[:SD.dim()])).item()
    print(f'Helmholtz {N} lean={getattr(H, "lean", None)}: {ms:.3f} ms  {nbytes/ms/1e6:.0f} GB/s algorithmic  '
          f'frac {nbytes/ms/1e6/PEAK:.2f}  residual {res:.1e}', flush=True)


def banded(shape, kl, ku, per_line):
    n_u = shape[0]
    n = n_u-2*max(kl, 1)
    inner = shape[1]*shape[2]
    nd = kl+1+ku
    rng = np.random.default_rng(0)
    tab = rng.uniform(-1, 1, (nd, n)); tab[kl] += 6
    tab_d = ops.as_device_tensor(tab)
    if per_line:
        fac = torch.empty((nd, 1, n, inner), dtype=torch.float64, device='cuda')
        sc = torch.ones(inner, dtype=torch.float64, device='cuda')
        ops.banded_factor(kl, ku, 2, [tab_d], [sc], fac, 1, inner, n)
    else:
        fac = torch.empty((nd, 1, n, 1), dtype=torch.float64, device='cuda')
        ops.banded_factor(kl, ku, 2, [tab_d], [None], fac, 1, 1, n)
        fac = fac.view(nd, n).contiguous()
    u = torch.randn(shape, dtype=torch.complex128, device='cuda')
    ms = timeit(lambda: ops.banded_solve(kl, ku, 2, u, fac, 1, inner, n, n_u, not per_line))
    nbytes = 2*u.numel()*16
    print(f'banded {shape} kl={kl} ku={ku} per_line={per_line}: {ms:.3f} ms  {nbytes/ms/1e6:.0f} GB/s algorithmic  '
          f'frac {nbytes/ms/1e6/PEAK:.2f}', flush=True)


if __name__ == "__main__":
    helmholtz((512, 512, 512))
    helmholtz((1024, 192, 512))
    banded((256, 256, 129), 2, 2, False)
    banded((256, 256, 129), 2, 2, True)
    banded((512, 512, 257), 1, 1, False)
    banded((512, 512, 257), 1, 1, True)
    banded((1024, 96, 513), 1, 1, True)
    banded((1024, 96, 513), 2, 2, False)
