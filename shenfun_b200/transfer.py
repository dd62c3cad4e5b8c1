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

    def chunk_rows(self, direction, side):
        """Rows per peer chunk along the axis of the kernel on the send / receive side."""
        if direction == 'backward':
            ga, sa = self.sa, self.ga
        else:
            ga, sa = self.ga, self.sa
        ax = sa if side == 'send' else ga
        return self.gshape[ax]//self.P

    # ---- direct peer stores (symmetric memory over NVLink) -----------------------------------
    def peer_available(self):
        """NCCL backend, equal blocks, <= 8 ranks and torch symmetric memory usable."""
        import os
        if os.environ.get('SHB_PEER_EXCHANGE', '1') == '0' or getattr(self, '_peer_broken', False):
            return False
        if not (dist.is_initialized() and dist.get_backend(self.comm.group) == 'nccl'):
            return False
        return self.uniform() and self.P <= 8

    def peer_buffers(self, numel, device):
        """Two alternating symmetric receive buffers of `numel` complex128 each: (local tensor,
        rendezvous handle).  Alternation + one barrier per exchange keeps a rank from overwriting
        a buffer a peer is still reading."""
        import torch.distributed._symmetric_memory as symm
        st = getattr(self, '_peer', None)
        if st is None or st['numel'] < numel:
            bufs = []
            for _ in range(2):
                t = symm.empty(numel, dtype=torch.complex128, device=device)
                hdl = symm.rendezvous(t, self.comm.group if self.comm.group is not None else dist.group.WORLD)
                bufs.append((t, hdl))
            st = {'numel': numel, 'bufs': bufs, 'turn': 0}
            self._peer = st
        st['turn'] ^= 1
        return st['bufs'][st['turn']]

    def uniform(self):
        """Equal blocks on both axes (needed by the fused, kernel-side pack / unpack)."""
        return self.gshape[self.ga] % self.P == 0 and self.gshape[self.sa] % self.P == 0

    def __call__(self, src, dst, ncomp, direction, src_packed=False, dst_packed=False):
        """Redistribute.  direction 'forward'/'scalar_product': un-distribute `ga`, distribute
        `sa`; 'backward': the reverse.

        Peer-chunked ("packed") layout: chunk p holds the block exchanged with rank p in the
        array's own index order.  `src_packed`: the producing kernel already wrote that layout
        (src is the flat send buffer); `dst_packed`: the consuming kernel will read it (dst is
        the flat receive buffer).  Otherwise one strided copy packs / unpacks -- or none at all
        when the blocks are contiguous in the array (leading axis, single component)."""
        nd = len(self.gshape)
        if direction == 'backward':
            ga, sa, g_blocks, s_blocks = self.sa, self.ga, self.s_blocks, self.g_blocks
        else:
            ga, sa, g_blocks, s_blocks = self.ga, self.sa, self.g_blocks, self.s_blocks
        P, rank = self.P, self.rank
        # local shapes before / after (with the leading component batch)
        sh_a = list(self.gshape)
        sh_a[ga] = g_blocks[rank][1]
        sh_b = list(self.gshape)
        sh_b[sa] = s_blocks[rank][1]
        lead = (ncomp,) if ncomp > 1 else ()
        L = len(lead)
        send_sizes, recv_sizes, send_shapes, recv_shapes = [], [], [], []
        for p in range(P):
            s = list(sh_a)
            s[sa] = s_blocks[p][1]
            send_shapes.append(lead+tuple(s))
            send_sizes.append(int(np.prod(send_shapes[-1])))
            r = list(sh_b)
            r[ga] = g_blocks[p][1]
            recv_shapes.append(lead+tuple(r))
            recv_sizes.append(int(np.prod(recv_shapes[-1])))
        # ---- send side
        if src_packed:
            sendbuf = src.reshape(-1)
        elif sa == 0 and ncomp == 1:
            sendbuf = src.reshape(-1)                        # blocks already contiguous
        else:
            srcv = src.reshape(lead+tuple(sh_a))
            sendbuf = self._buffer('send', srcv.numel(), src)
            off = 0
            for p in range(P):
                chunk = srcv.narrow(sa+L, s_blocks[p][0], s_blocks[p][1])
                sendbuf[off:off+send_sizes[p]].view(chunk.shape).copy_(chunk)
                off += send_sizes[p]
        # ---- receive side
        direct = (ga == 0 and ncomp == 1)
        if dst_packed or direct:
            recvbuf = dst.reshape(-1)
        else:
            recvbuf = self._buffer('recv', sum(recv_sizes), dst)
        _all_to_all(recvbuf, sendbuf, recv_sizes, send_sizes, self.comm.group)
        if not (dst_packed or direct):
            dstv = dst.reshape(lead+tuple(sh_b))
            off = 0
            for p in range(P):
                tgt = dstv.narrow(ga+L, g_blocks[p][0], g_blocks[p][1])
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
