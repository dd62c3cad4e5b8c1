"""CUDA-graph capture of a fixed sequence of transforms / solves on static arrays.

At 8 GPUs a 512^3 transform pair is ~0.8 ms of device work spread over a dozen launches and
four cross-rank barriers; the Python/ctypes cost of issuing them is then of the same order as
the work.  Capturing the sequence once and replaying it removes that host cost.  Every kernel
of libshenb200 launches on torch's current stream, so it is captured like any torch op; the
peer-store exchange (symmetric-memory barriers) and NCCL collectives are stream-ordered and
capture as well.  Tables (twiddles, operator matrices, work buffers) must exist before capture:
the constructor runs the sequence `warmup` times first.
"""
import torch


class GraphedCall:
    def __init__(self, fn, warmup=3, comm=None):
        self.fn = fn
        for _ in range(max(1, warmup)):
            fn()
        torch.cuda.synchronize()
        if comm is not None and comm.Get_size() > 1:
            comm.barrier()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            fn()
        torch.cuda.synchronize()

    def __call__(self):
        self.graph.replay()
