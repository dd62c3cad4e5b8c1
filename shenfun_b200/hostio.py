"""Streaming host <-> device pipeline for workloads whose inputs and results live in (pinned)
host memory: the H2D copy of step k+1, the device work of step k and the D2H copy of step k-1
run on three streams with double-buffered device arrays, so a sequence of steps is bounded by
the slowest of the three (PCIe is full duplex) instead of their sum.
"""
import torch


class HostPipeline:
    def __init__(self, in_shape, in_dtype, out_shape, out_dtype, fn, depth=2, device=None):
        """fn(d_in, d_out) launches the device work on the current stream."""
        self.fn = fn
        self.depth = depth
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
        self.d_in = [torch.empty(in_shape, dtype=in_dtype, device=dev) for _ in range(depth)]
        self.d_out = [torch.empty(out_shape, dtype=out_dtype, device=dev) for _ in range(depth)]
        self.s_in = torch.cuda.Stream(device=dev)
        self.s_out = torch.cuda.Stream(device=dev)
        ev = lambda: [torch.cuda.Event() for _ in range(depth)]  # noqa: E731
        self.in_ready, self.in_free, self.out_ready, self.out_free = ev(), ev(), ev(), ev()
        main = torch.cuda.current_stream()
        for k in range(depth):
            self.in_free[k].record(main)
            self.out_free[k].record(main)
        self.i = 0

    def step(self, h_in, h_out):
        """Enqueue one step: h_in (pinned) -> device -> fn -> h_out (pinned).  Asynchronous."""
        k = self.i % self.depth
        self.i += 1
        main = torch.cuda.current_stream()
        self.s_in.wait_event(self.in_free[k])
        with torch.cuda.stream(self.s_in):
            self.d_in[k].copy_(h_in, non_blocking=True)
            self.in_ready[k].record(self.s_in)
        main.wait_event(self.in_ready[k])
        main.wait_event(self.out_free[k])
        self.fn(self.d_in[k], self.d_out[k])
        self.in_free[k].record(main)
        self.out_ready[k].record(main)
        self.s_out.wait_event(self.out_ready[k])
        with torch.cuda.stream(self.s_out):
            h_out.copy_(self.d_out[k], non_blocking=True)
            self.out_free[k].record(self.s_out)

    def drain(self):
        """Make the current stream wait for every outstanding copy."""
        main = torch.cuda.current_stream()
        for k in range(self.depth):
            main.wait_event(self.out_free[k])
            main.wait_event(self.in_free[k])
