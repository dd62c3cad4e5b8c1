"""Global redistribution between slab decompositions (multi-GPU).

Stands in for `mpi4py_fft.pencil.Transfer.forward/backward` (third-party; call
sites shenfun/tensorproductspace.py:89,183,252), which the reference implements
as `MPI_Alltoallw` with sub-array datatypes.  Here: one NCCL all-to-all of
peer-contiguous blocks over NVLink (`torch.distributed.all_to_all_single` is the
plumbing).  The local axis permutation into / out of the peer-contiguous layout
is a pack / unpack pass around the collective.

Layout: the array is distributed along `gather_axis` before the exchange and
along `scatter_axis` after it (and the reverse for the backward direction).
"""
import numpy as np
import torch
import torch.distributed as dist


def _block(N, P, i):
    n, r = divmod(N, P)
    return i*n + min(i, r), n + (1 if i < r else 0)


class SlabExchange:
    def __init__(self, comm, gshape, gather_axis, scatter_axis):
        self.comm = comm
        self.gshape = tuple(gshape)
        self.ga = gather_axis
        self.sa = scatter_axis
        self.P = comm.Get_size()
        self.rank = comm.Get_rank()
        P = self.P
        self.g_blocks = [_block(self.gshape[self.ga], P, i) for i in range(P)]
        self.s_blocks = [_block(self.gshape[self.sa], P, i) for i in range(P)]

    def _shapes(self, lead):
        """Local shapes (with `lead` leading batch dims) before / after the forward exchange."""
        a = list(self.gshape)
        a[self.ga] = self.g_blocks[self.rank][1]
        b = list(self.gshape)
        b[self.sa] = self.s_blocks[self.rank][1]
        return tuple(a), tuple(b)

    def __call__(self, src, dst, ncomp, direction):
        """direction 'forward'/'scalar_product': gather `ga`, scatter `sa`; 'backward': reverse.

        When the scattered axis is the leading one of a chunk the peer blocks are already
        contiguous on that side, so only ONE local strided copy (pack or unpack) is needed:
        the other side of the all-to-all reads from / writes to the array itself."""
        lead = src.dim() - len(self.gshape)
        if direction == 'backward':
            ga, sa, g_blocks, s_blocks = self.sa, self.ga, self.s_blocks, self.g_blocks
        else:
            ga, sa, g_blocks, s_blocks = self.ga, self.sa, self.g_blocks, self.s_blocks
        P, rank = self.P, self.rank
        batch = int(np.prod(src.shape[:lead])) if lead else 1
        # leading batch (vector components) is exchanged component by component when it would
        # break contiguity of the peer blocks
        if batch > 1:
            sv = src.reshape((batch,)+tuple(src.shape[lead:]))
            dv = dst.reshape((batch,)+tuple(dst.shape[lead:]))
            for c in range(batch):
                self._exchange(sv[c], dv[c], ga, sa, g_blocks, s_blocks)
        else:
            self._exchange(src.reshape(src.shape[lead:]), dst.reshape(dst.shape[lead:]), ga, sa, g_blocks, s_blocks)

    def _exchange(self, src, dst, ga, sa, g_blocks, s_blocks):
        P, rank = self.P, self.rank
        send_sizes = []
        recv_sizes = []
        for p in range(P):
            sh = list(src.shape)
            sh[sa] = s_blocks[p][1]
            send_sizes.append(int(np.prod(sh)))
            sh = list(dst.shape)
            sh[ga] = g_blocks[p][1]
            recv_sizes.append(int(np.prod(sh)))
        # send side: chunk p = src[.., s_blocks[p] along sa, ..]; contiguous already when sa == 0
        if sa == 0:
            sendbuf = src.reshape(-1)
        else:
            sendbuf = self._buffer('send', src.numel(), src)
            off = 0
            for p in range(P):
                chunk = src.narrow(sa, s_blocks[p][0], s_blocks[p][1])
                sendbuf[off:off+send_sizes[p]].view(chunk.shape).copy_(chunk)
                off += send_sizes[p]
        # receive side: block from p spans g_blocks[p] along ga; lands in place when ga == 0
        if ga == 0:
            recvbuf = dst.reshape(-1)
        else:
            recvbuf = self._buffer('recv', dst.numel(), dst)
        _all_to_all(recvbuf, sendbuf, recv_sizes, send_sizes, self.comm.group)
        if ga != 0:
            off = 0
            for p in range(P):
                tgt = dst.narrow(ga, g_blocks[p][0], g_blocks[p][1])
                tgt.copy_(recvbuf[off:off+recv_sizes[p]].view(tgt.shape))
                off += recv_sizes[p]

    def _buffer(self, name, numel, like):
        key = (name, like.dtype, like.device)
        if not hasattr(self, '_bufs'):
            self._bufs = {}
        b = self._bufs.get(key)
        if b is None or b.numel() < numel:
            b = torch.empty(numel, dtype=like.dtype, device=like.device)
            self._bufs[key] = b
        return b[:numel]


def _all_to_all(recvbuf, sendbuf, recv_sizes, send_sizes, group):
    complex_in = sendbuf.is_complex()
    if complex_in:
        sendbuf = torch.view_as_real(sendbuf).reshape(-1)
        recv_real = torch.view_as_real(recvbuf).reshape(-1)
        send_sizes = [2*s for s in send_sizes]
        recv_sizes = [2*s for s in recv_sizes]
    else:
        recv_real = recvbuf
    backend = dist.get_backend(group)
    if backend == 'gloo':
        # gloo has no all_to_all: pairwise exchange (CPU tests of the host logic only)
        rank = dist.get_rank(group)
        P = dist.get_world_size(group)
        soff = np.concatenate(([0], np.cumsum(send_sizes)))
        roff = np.concatenate(([0], np.cumsum(recv_sizes)))
        reqs = []
        for p in range(P):
            if p == rank:
                recv_real[roff[p]:roff[p+1]] = sendbuf[soff[p]:soff[p+1]]
                continue
            reqs.append(dist.isend(sendbuf[soff[p]:soff[p+1]].contiguous(), dst=dist.get_global_rank(group, p) if group else p, group=group))
            reqs.append(dist.irecv(recv_real[roff[p]:roff[p+1]], src=dist.get_global_rank(group, p) if group else p, group=group))
        for r in reqs:
            r.wait()
    else:
        dist.all_to_all_single(recv_real, sendbuf, recv_sizes, send_sizes, group=group)
