"""Eager vs CUDA-graph replay of the 512^3 forward+backward pair (run under torchrun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import shenfun_b200 as sf
from shenfun_b200.graphs import GraphedCall

comm = sf.init_from_env()
N = (512, 512, 512)
T = sf.TensorProductSpace(comm, [sf.FunctionSpace(n, 'F', dtype=dt) for n, dt in zip(N, 'DDd')])
u = sf.Array(T); u.v.normal_()
uh = sf.Function(T); ub = sf.Array(T)

def pair():
    T.forward(u, uh)
    T.backward(uh, ub)

def timeit(fn, steps=50):
    for _ in range(5):
        fn()
    comm.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/steps
    return comm.allreduce(ms, op='max')

ms_eager = timeit(pair)
ref = ub.v.clone()
try:
    g = GraphedCall(pair, comm=comm)
    ub.v.zero_()
    g()
    torch.cuda.synchronize()
    err = (ub.v-ref).abs().max().item()
    ms_graph = timeit(g)
    if comm.Get_rank() == 0:
        print(f'ranks {comm.Get_size()}: eager {ms_eager:.3f} ms ({1e3/ms_eager:.0f} pairs/s), graph {ms_graph:.3f} ms '
              f'({1e3/ms_graph:.0f} pairs/s), max diff {err:.1e}')
except Exception as e:  # noqa: BLE001
    if comm.Get_rank() == 0:
        print(f'ranks {comm.Get_size()}: eager {ms_eager:.3f} ms; graph capture failed: {type(e).__name__}: {str(e)[:300]}')
