"""CUDA-event timing of the three forward axis passes of the 512^3 pair."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shenfun_b200 import ops
C = torch.complex128
u = torch.randn(512, 512, 512, dtype=torch.float64, device='cuda')
a = torch.empty(512, 512, 257, dtype=C, device='cuda')
b = torch.empty(512, 512, 257, dtype=C, device='cuda')
flush = torch.empty(64*1024*1024, dtype=torch.float64, device='cuda')
passes = {
    'r2c_z': (lambda: ops.fft_r2c_forward(u, a, 512*512, 1, 512, 512, 1/512), u.numel()*8+a.numel()*16),
    'c2c_y': (lambda: ops.fft_c2c_forward(a, b, 512, 257, 512, 512, 1/512), 2*a.numel()*16),
    'c2c_x': (lambda: ops.fft_c2c_forward(b, a, 1, 512*257, 512, 512, 1/512), 2*a.numel()*16),
    'c2c_x_bwd': (lambda: ops.fft_c2c_backward(a, b, 1, 512*257, 512, 512), 2*a.numel()*16),
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
    print(f'{name}: {ms:.3f} ms  {nbytes/ms/1e6:.0f} GB/s  frac {nbytes/ms/1e6/6546.9:.2f}')
