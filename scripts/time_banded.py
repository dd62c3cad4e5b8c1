"""CUDA-event timing of the banded line solves (shared and per-line factors)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from shenfun_b200 import ops

def run(shape, kl, ku, per_line, reps=10):
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
    flush = torch.empty(64*1024*1024, dtype=torch.float64, device='cuda')
    ts = []
    for it in range(reps+2):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ops.banded_solve(kl, ku, 2, u, fac, 1, inner, n, n_u, not per_line); e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts)//2]
    nbytes = 2*u.numel()*16+(fac.numel()*8 if per_line else 0)
    print(f'{shape} kl={kl} ku={ku} per_line={per_line}: {ms:.3f} ms  {nbytes/ms/1e6:.0f} GB/s')

run((256, 256, 129), 2, 2, False)
run((256, 256, 129), 2, 2, True)
run((512, 512, 257), 1, 1, False)
run((512, 512, 257), 1, 1, True)
run((1024, 96, 513), 1, 1, True)
run((1024, 96, 513), 2, 2, False)
