"""Time the 3/2-rule C2R pass on contiguous lines (768 x 768 lines of 257 modes -> 768 points)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shenfun_b200 import ops
L = int(os.environ.get('LINES', 768*768))
N, n_fft = 512, 768
ah = torch.randn(L, N//2+1, dtype=torch.complex128, device='cuda')
out = torch.empty(L, n_fft, dtype=torch.float64, device='cuda')
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
ts = []
for it in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.fft_c2r_backward(ah, out, L, 1, n_fft, N); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts = sorted(ts[2:])
gb = (L*(N//2+1)*16+L*n_fft*8)/1e9
print(f'c2r 512->768, {L} lines: {ts[len(ts)//2]:.3f} ms  {gb/ts[len(ts)//2]:.2f} TB/s ({gb:.2f} GB)  NO_PAIR3={os.environ.get("SHB_NO_PAIR3")}')
