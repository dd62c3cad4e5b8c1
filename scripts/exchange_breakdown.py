"""Per-step device timing of the slab-distributed 512^3 forward transform (run under torchrun):
r2c pass, producer pass (its stores are the all-to-all), barrier, consumer pass, plus the raw NVLink
rates of this box (copy-engine peer copy and an SM copy kernel storing into peer memory)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import shenfun_b200 as sf
from shenfun_b200 import ops

comm = sf.init_from_env('nccl')
rank, P = comm.Get_rank(), comm.Get_size()
dev = torch.device('cuda', torch.cuda.current_device())
N = (512, 512, 512)
T = sf.TensorProductSpace(comm, [sf.FunctionSpace(N[0], 'F', dtype='D'), sf.FunctionSpace(N[1], 'F', dtype='D'),
                                 sf.FunctionSpace(N[2], 'F', dtype='d')])
u = sf.Array(T); u.v.normal_()
uh = sf.Function(T)
for _ in range(3):
    T.forward(u, uh)


def timeit(fn, reps=20, sync_ranks=True):
    fn(); torch.cuda.synchronize()
    if sync_ranks:
        dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)/reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


g0, n1, n2 = T.shape(False)[0], 512, 257
s1 = T.shape(True)[1]
pitch = 272
w = torch.empty((g0, n1, n2), dtype=torch.complex128, device=dev)
b2, b1, b0 = T.bases[2], T.bases[1], T.bases[0]
res = {}
res['r2c_ms'] = timeit(lambda: b2._forward_axis(u.v, w, g0*n1, 1))
ex = [s for s in T.forward.steps if s.kind == 'exchange'][0].exchange
block = 512*s1*pitch
rbuf, hdl = ex.peer_buffers(block, dev)
chunk = block//P
peers = [int(p)+chunk*16*rank for p in hdl.buffer_ptrs]
cr = n1//P
wf = w.reshape(-1)
prod = lambda: ops.fft_c2c_lines(False, wf, None, peers, g0, pitch, 0, n1*n2, n2, cr*pitch, pitch, 512, 512, 1/512., False, 0, 0, cr, 0, ncols_valid=n2)
res['producer_pitched_ms'] = timeit(prod)
nv_bytes = g0*n1*n2*16*(P-1)/P
res['producer_pitched_nvlink_GBs'] = nv_bytes/res['producer_pitched_ms']/1e6
# natural (un-pitched) layout through the older entry point
rb2, h2 = ex.peer_buffers(block, dev)
pe2 = [int(p)+(512*s1*n2//P)*16*rank for p in h2.buffer_ptrs]
prod_old = lambda: ops.fft_c2c_peer(False, w, pe2, g0, n2, 512, 512, 1/512., False, 0, cr)
res['producer_natural_ms'] = timeit(prod_old)
res['producer_natural_nvlink_GBs'] = nv_bytes/res['producer_natural_ms']/1e6
# same pass with local stores only (HBM-bound reference)
loc = torch.empty((P, g0, cr, n2), dtype=torch.complex128, device=dev)
res['producer_local_ms'] = timeit(lambda: ops.fft_c2c_chunked(False, w, loc, g0, n2, 512, 512, 1/512., False, 0, cr))
res['barrier_ms'] = timeit(lambda: hdl.barrier())
out = torch.empty((512, s1, n2), dtype=torch.complex128, device=dev)
rfl = rbuf.reshape(-1)
cons = lambda: ops.fft_c2c_lines(False, rfl, out.reshape(-1), None, s1, n2, 0, pitch, s1*pitch, n2, s1*n2, 512, 512, 1/512., False, 512//P, (512//P)*s1*pitch, 0, 0)
res['consumer_ms'] = timeit(cons)
res['forward_total_ms'] = timeit(lambda: T.forward(u, uh))
res['backward_total_ms'] = timeit(lambda: T.backward(uh, u))
# ---- backward: producer along axis 0 (lines (j, c)), consumer along axis 1, c2r
vh = torch.randn((512, s1, n2, 2), dtype=torch.float64, device=dev).view(torch.complex128).reshape(512, s1, n2)
blockb = g0*n1*pitch
rbb, hb = ex.peer_buffers(blockb, dev)
chb = blockb//P
peb = [int(p)+chb*16*rank for p in hb.buffer_ptrs]
cr0 = 512//P
prodb = lambda: ops.fft_c2c_lines(True, vh.reshape(-1), None, peb, s1, pitch, 0, n2, s1*n2, pitch, s1*pitch, 512, 512, 1.0, False, 0, 0, cr0, 0, ncols_valid=n2)
res['bwd_producer_pitched_ms'] = timeit(prodb)
res['bwd_producer_nvlink_GBs'] = nv_bytes/res['bwd_producer_pitched_ms']/1e6
locb = torch.empty((512, s1, n2), dtype=torch.complex128, device=dev)
res['bwd_producer_local_ms'] = timeit(lambda: ops.fft_c2c_backward(vh, locb, 1, s1*n2, 512, 512))
w2 = torch.empty((g0, n1, n2), dtype=torch.complex128, device=dev)
cr1 = n1//P
consb = lambda: ops.fft_c2c_lines(True, rbb.reshape(-1), w2.reshape(-1), None, g0, n2, 0, cr1*pitch, pitch, n1*n2, n2, 512, 512, 1.0, False, cr1, g0*cr1*pitch, 0, 0)
res['bwd_consumer_ms'] = timeit(consb)
consb2 = lambda: ops.fft_c2c_lines(True, rbb.reshape(-1), w2.reshape(-1), None, g0, pitch, 0, cr1*pitch, pitch, n1*n2, n2, 512, 512, 1.0, False, cr1, g0*cr1*pitch, 0, 0, ncols_valid=n2)
res['bwd_consumer_tiles_on_pitch_ms'] = timeit(consb2)
natb = torch.randn((P, g0, cr1, n2, 2), dtype=torch.float64, device=dev).view(torch.complex128).reshape(P, g0, cr1, n2)
res['bwd_consumer_natural_layout_ms'] = timeit(lambda: ops.fft_c2c_chunked(True, natb, w2, g0, n2, 512, 512, 1.0, False, cr1, 0))
cons2 = lambda: ops.fft_c2c_lines(False, rfl, out.reshape(-1), None, s1, pitch, 0, pitch, s1*pitch, n2, s1*n2, 512, 512, 1/512., False, 512//P, (512//P)*s1*pitch, 0, 0, ncols_valid=n2)
res['fwd_consumer_tiles_on_pitch_ms'] = timeit(cons2)
res['c2r_ms'] = timeit(lambda: b2._backward_axis(w2, u.v, g0*n1, 1))
# raw link rates: copy engine and SM copy kernel into the next peer's buffer
nxt = (rank+1) % P
peer_t = hdl.get_buffer(nxt, (chunk,), torch.complex128, 0) if hasattr(hdl, 'get_buffer') else None
src = torch.randn(chunk, 2, dtype=torch.float64, device=dev).view(torch.complex128).reshape(-1)
if peer_t is not None:
    res['ce_or_sm_copy_to_peer_GBs'] = chunk*16/timeit(lambda: peer_t.copy_(src))/1e6
    res['copy_bytes'] = chunk*16
if rank == 0:
    print(json.dumps(res, indent=1))
dist.destroy_process_group()
