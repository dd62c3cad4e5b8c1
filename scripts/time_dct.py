"""CUDA-event timing of the Chebyshev axis passes (DCT-II scalar product / DCT-III backward)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import shenfun_b200 as sf
from shenfun_b200 import ops

def run(N, pf, shape_rest):
    SD = sf.FunctionSpace(N, 'C', bc=(0, 0), padding_factor=pf)
    Np = SD.physical_len()
    inner = shape_rest[0]*shape_rest[1]
    uh = torch.randn((N,)+shape_rest, dtype=torch.complex128, device='cuda')
    up = torch.empty((Np,)+shape_rest, dtype=torch.complex128, device='cuda')
    flush = torch.empty(64*1024*1024, dtype=torch.float64, device='cuda')
    for name, fn in (('backward', lambda: SD._backward_axis(uh, up, 1, inner)),
                     ('scalar_product', lambda: SD._scalar_product_axis(up, uh, 1, inner))):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts)//2]
        nbytes = (uh.numel()+up.numel())*16
        print(f'N={N} pf={pf} {name}: {ms:.3f} ms  {nbytes/ms/1e6:.0f} GB/s  frac {nbytes/ms/1e6/6546.9:.2f}')

run(512, 1.0, (512, 257))
run(512, 1.5, (256, 257))
run(1024, 1.5, (96, 257))
