"""Communicator facade: one process per GPU over torch.distributed.

Mirrors the handful of `mpi4py.MPI.Comm` methods the reference's user code
touches (`comm.Get_size()`, `comm.Get_rank()`, `comm.reduce`, `comm.allreduce`,
`comm.bcast`, `comm.barrier`); the data-path exchange itself is in
`shenfun_b200.transfer`.
"""
import os
import torch
import torch.distributed as dist


class Comm:
    def __init__(self, group=None, self_only=False):
        self._group = group
        self._self_only = self_only

    @property
    def group(self):
        return self._group

    def _active(self):
        return (not self._self_only) and dist.is_available() and dist.is_initialized()

    def Get_size(self):
        return dist.get_world_size(self._group) if self._active() else 1

    def Get_rank(self):
        return dist.get_rank(self._group) if self._active() else 0

    size = property(Get_size)
    rank = property(Get_rank)

    def barrier(self):
        if self._active():
            dist.barrier(self._group)

    Barrier = barrier

    def _device(self):
        if self._active() and dist.get_backend(self._group) == 'nccl':
            return torch.device('cuda', torch.cuda.current_device())
        return torch.device('cpu')

    def allreduce(self, value, op='sum'):
        """Sum (default) / max / min of a python scalar or tensor over ranks."""
        if not self._active() or self.Get_size() == 1:
            return value
        was_tensor = torch.is_tensor(value)
        t = value.clone() if was_tensor else torch.tensor(value, dtype=torch.float64)
        dev = t.device
        t = t.to(self._device())
        rop = {'sum': dist.ReduceOp.SUM, 'max': dist.ReduceOp.MAX, 'min': dist.ReduceOp.MIN}[op]
        dist.all_reduce(t, op=rop, group=self._group)
        return t.to(dev) if was_tensor else t.item()

    def reduce(self, value, op='sum', root=0):
        return self.allreduce(value, op)

    def bcast(self, obj, root=0):
        if not self._active() or self.Get_size() == 1:
            return obj
        lst = [obj]
        dist.broadcast_object_list(lst, src=root, group=self._group)
        return lst[0]

    def allgather_blocks(self, t, axis):
        """Concatenate the ranks' blocks of a tensor along `axis` (uneven extents allowed); every rank
        gets the whole array.  Host-side plumbing (refine, diagnostics), not a data-path collective."""
        if not self._active() or self.Get_size() == 1:
            return t
        P = self.Get_size()
        dev = self._device()
        n = torch.tensor([t.shape[axis]], dtype=torch.int64, device=dev)
        sizes = [torch.zeros_like(n) for _ in range(P)]
        dist.all_gather(sizes, n, group=self._group)
        sizes = [int(x.item()) for x in sizes]
        m = max(sizes)
        shp = list(t.shape)
        shp[axis] = m
        send = torch.zeros(shp, dtype=t.dtype, device=dev)
        send.narrow(axis, 0, t.shape[axis]).copy_(t)
        parts = [torch.empty_like(send) for _ in range(P)]
        dist.all_gather(parts, send, group=self._group)
        return torch.cat([p.narrow(axis, 0, k) for p, k in zip(parts, sizes)], dim=axis).to(t.device)

    def Dup(self):
        return self

    def Free(self):
        pass


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (no-op for 1 process)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29511')
        dist.init_process_group(backend=backend)
    return comm


comm = Comm()
COMM_SELF = Comm(self_only=True)     # single-rank communicator (e.g. the 1-D (0,0)-mode problems)
