"""Streaming host <-> device pipeline for workloads whose inputs and results live in (pinned)
host memory: the H2D copy of step k+1, the device work of step k and the D2H copy of step k-1
run on three streams with double-buffered device arrays, so a sequence of steps is bounded by
the slowest of the three (PCIe is full duplex) instead of their sum.
"""
import os

import torch


def bind_host_to_device_numa_node(device=None):
    """Restrict this process to the CPUs of the NUMA node the GPU hangs off, so that pinned host
    buffers allocated afterwards are first-touched there and H2D / D2H copies of the ranks of one box
    do not all cross the socket interconnect.  Returns the node number (None when the topology is not
    exposed in sysfs; nothing is changed then)."""
    try:
        dev = torch.cuda.current_device() if device is None else device
        pr = torch.cuda.get_device_properties(dev)
        bus = f'{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0'
        base = f'/sys/bus/pci/devices/{bus}'
        node = int(open(f'{base}/numa_node').read().strip())
        cpus = open(f'{base}/local_cpulist').read().strip()
        ids = set()
        for part in cpus.split(','):
            if '-' in part:
                a, b = part.split('-')
                ids.update(range(int(a), int(b)+1))
            elif part:
                ids.add(int(part))
        allowed = os.sched_getaffinity(0)
        ids = ids & allowed if ids & allowed else allowed
        if node >= 0 and ids:
            os.sched_setaffinity(0, ids)
            return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        pass
    return None


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
