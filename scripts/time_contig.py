"""CUDA-event timing of the contiguous-line passes: R2C / C2R at 512 (pair) and R2C 768 -> 257 modes (3/2 rule)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shenfun_b200 import ops
C = torch.complex128
u = torch.randn(512*512, 512, dtype=torch.float64, device='cuda')
a = torch.randn(512*512, 257, dtype=C, device='cuda')
up = torch.randn(768*768, 768, dtype=torch.float64, device='cuda')
ap = torch.empty(768*768, 257, dtype=C, device='cuda')
flush = torch.empty(64*1024*1024, dtype=torch.float64, device='cuda')
passes = {
    'r2c_512': (lambda: ops.fft_r2c_forward(u, a, 512*512, 1, 512, 512, 1/512), u.numel()*8+a.numel()*16),
    'c2r_512': (lambda: ops.fft_c2r_backward(a, u, 512*512, 1, 512, 512), u.numel()*8+a.numel()*16),
    'r2c_768_trunc': (lambda: ops.fft_r2c_forward(up, ap, 768*768, 1, 768, 512, 1/768), up.numel()*8+ap.numel()*16),
}
for name, (fn, nbytes) in passes.items():
    for _ in range(3):
        fn()
    ts = []
    for _ in range(15):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts)//2]
    print(f'{name}: {ms:.3f} ms  {nbytes/ms/1e6:.0f} GB/s  frac {nbytes/ms/1e6/6546.9:.2f}  PF_MB={os.environ.get("SHB_PF_CONTIG_MB")}')
