import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, shenfun_b200 as sf
from shenfun_b200 import ops
def tl(fn, reps=10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/reps
C = torch.complex128
S = 512*512*257*16; R = 512**3*8
a = torch.randn(512,512,257, dtype=C, device='cuda'); 
u = torch.randn(512,512,512, dtype=torch.float64, device='cuda')
def rep(name, ms, nbytes): print(f'{name:28s} {ms:7.3f} ms  {nbytes/ms/1e6:7.0f} GB/s  {nbytes/ms/1e6/6546.9:5.2f}')
rep('r2c z 512', tl(lambda: ops.fft_r2c_forward(u, a, 512*512, 1, 512, 512, 1/512)), R+S)
rep('c2r z 512', tl(lambda: ops.fft_c2r_backward(a, u, 512*512, 1, 512, 512)), R+S)
rep('c2c y fwd inplace', tl(lambda: ops.fft_c2c_forward(a, a, 512, 257, 512, 512, 1/512)), 2*S)
rep('c2c x fwd inplace', tl(lambda: ops.fft_c2c_forward(a, a, 1, 512*257, 512, 512, 1/512)), 2*S)
rep('c2c x bwd inplace', tl(lambda: ops.fft_c2c_backward(a, a, 1, 512*257, 512, 512)), 2*S)
b = torch.empty(768,512,257, dtype=C, device='cuda')
rep('pad x bwd 512->768', tl(lambda: ops.fft_c2c_backward(a, b, 1, 512*257, 768, 512)), 2.5*S)
rep('trunc x fwd 768->512', tl(lambda: ops.fft_c2c_forward(b, a, 1, 512*257, 768, 512, 1/768)), 2.5*S)
c = torch.empty(768,768,257, dtype=C, device='cuda')
rep('pad y bwd 512->768', tl(lambda: ops.fft_c2c_backward(b, c, 768, 257, 768, 512)), 1.5*2.5*S)
rep('trunc y fwd 768->512', tl(lambda: ops.fft_c2c_forward(c, b, 768, 257, 768, 512, 1/768)), 1.5*2.5*S)
up = torch.empty(768,768,768, dtype=torch.float64, device='cuda')
rep('pad c2r z 257->768', tl(lambda: ops.fft_c2r_backward(c, up, 768*768, 1, 768, 512)), 2.25*S+768**3*8)
rep('trunc r2c z 768->257', tl(lambda: ops.fft_r2c_forward(up, c, 768*768, 1, 768, 512, 1/768)), 2.25*S+768**3*8)
