import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shenfun_b200 import ops
C = torch.complex128
a = torch.randn(512, 512, 257, dtype=C, device='cuda')
b = torch.empty(768, 512, 257, dtype=C, device='cuda')
c = torch.empty(768, 768, 257, dtype=C, device='cuda')
for it in range(2):
    if it == 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    ops.fft_c2c_backward(a, b, 1, 512*257, 768, 512)
    ops.fft_c2c_backward(b, c, 768, 257, 768, 512)
    ops.fft_c2c_forward(c, b, 768, 257, 768, 512, 1/768)
    ops.fft_c2c_forward(b, a, 1, 512*257, 768, 512, 1/768)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
