"""Tensor-product spaces with fused device transform plans.

Host-side mirror of shenfun/tensorproductspace.py:
  `TensorProductSpace.__init__` / `configure_backwards` :314-541 (planning),
  `ForwardTransform` / `BackwardTransform` / `ScalarTransform` :46-259,
  `get_dealiased` :543-578, `VectorSpace` / `CompositeSpace` :1190-1784,
  `VectorTransform` :1786-1795.

Differences in *how* (not what):
  * a plan is a short list of per-axis kernel launches (plus one all-to-all
    when the space is slab-distributed over several GPUs); pad/truncate,
    normalisation, stencils and mass solves are fused into those launches, and
    un-padded complex axes run in place, so the 512^3 R2C x C2C x C2C transform
    touches each array exactly three times;
  * vector / tensor valued spaces are transformed in ONE batched launch per
    axis (the component index is folded into the `outer` batch extent) instead
    of a Python loop over components;
  * multi-GPU uses the slab decomposition of mpi4py-fft (`slab=True`,
    :393-396): physical space sharded on axis (first-transformed+1) % ndim,
    spectral space on the axis transformed just before it.
"""
import os
import numpy as np
import torch

from .arrays import Function, Array
from .comm import comm as default_comm
from . import ops

__all__ = ('TensorProductSpace', 'VectorSpace', 'TensorSpace', 'CompositeSpace')


def _block(N, P, i):
    """Balanced block decomposition used by mpi4py-fft's Pencil (start, length)."""
    n, r = divmod(N, P)
    start = i*n + min(i, r)
    return start, n + (1 if i < r else 0)


class _Step:
    """One stage of a plan: an axis transform or a global exchange."""
    __slots__ = ('kind', 'axis', 'base', 'in_shape', 'out_shape', 'in_dtype', 'out_dtype',
                 'inplace_ok', 'exchange', 'fuse')

    def __init__(self, kind, axis, base, in_shape, out_shape, in_dtype, out_dtype,
                 inplace_ok=False, exchange=None):
        self.kind = kind
        self.axis = axis
        self.base = base
        self.in_shape = tuple(in_shape)
        self.out_shape = tuple(out_shape)
        self.in_dtype = in_dtype
        self.out_dtype = out_dtype
        self.inplace_ok = inplace_ok
        self.exchange = exchange
        self.fuse = None      # exchange steps: [producer_fused, consumer_fused]


class _WorkPool:
    """Two flat ping-pong byte buffers, grown on demand and reused by every call."""

    def __init__(self):
        self.buf = [None, None]

    def get(self, which, shape, dtype, device):
        nbytes = int(np.prod(shape))*torch.empty((), dtype=dtype).element_size()
        b = self.buf[which]
        if b is None or b.numel() < nbytes or b.device != device:
            b = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)
            self.buf[which] = b
        return b[:nbytes].view(dtype).view(shape)


class _Transform:
    """Callable plan (forward, backward or scalar product) of a space."""

    def __init__(self, space, steps, direction):
        self.space = space
        self.steps = steps
        self.direction = direction
        self._pool = space._pool

    @property
    def input_shape(self):
        return self.steps[0].in_shape

    @property
    def output_shape(self):
        return self.steps[-1].out_shape

    def _run(self, u, out, ncomp, last_override=None, overrides=None):
        """`last_override` / `overrides` {step index: fn(cur, dst, outer, inner)} replace the axis
        kernel of a step (fused C2R-cross; the do-nothing pass of `TensorProductSpace._tensor_bcs`)."""
        from ._lib import ShenB200Error
        self._last_override = last_override
        self._overrides = overrides
        try:
            return self._run_impl(u, out, ncomp)
        except ShenB200Error as e:
            fused = [s for s in self.steps if s.kind == 'exchange' and s.fuse and any(s.fuse)]
            if 'fft_c2c_chunked' not in str(e) or not fused:
                raise
            # no register-resident kernel for this size: pack / unpack around the all-to-all instead
            sp = self.space
            for tr in (sp.forward, sp.backward, sp.scalar_product):
                for s in tr.steps:
                    if s.kind == 'exchange':
                        s.fuse = None
            return self._run_impl(u, out, ncomp)

    def _run_impl(self, u, out, ncomp):
        """u, out: plain contiguous tensors with a leading component batch folded in."""
        steps = self.steps
        n = len(steps)
        lead = (ncomp,) if ncomp > 1 else ()
        # destination of every step: the user output if all later steps run in place
        dest_is_out = [False]*n
        dest_is_out[n-1] = True
        for i in range(n-2, -1, -1):
            s = steps[i+1]
            dest_is_out[i] = dest_is_out[i+1] and s.inplace_ok and s.in_shape == s.out_shape \
                and s.in_dtype == s.out_dtype
        cur = u
        cur_is_user_input = True
        which = 0
        bwd = self.direction == 'backward'
        cur_packed = False
        peer_done = None
        skip_to = 0
        for i, s in enumerate(steps):
            if i < skip_to:
                continue
            # producer pass + exchange + consumer pass cut into column pieces and overlapped
            if i+3 < n and not bwd and s.kind == 'axis' and steps[i+1].kind == 'axis':
                # [local pass] [producer] [exchange] [consumer]: the local pass is pipelined under the producer
                res = self._overlapped_exchange(i+1, cur, out, dest_is_out, which, ncomp, pre=i)
                if res is not None:
                    cur, which = res
                    cur_is_user_input = False
                    cur_packed = False
                    skip_to = i+4
                    continue
            if i+2 < n:
                res = self._overlapped_exchange(i, cur, out, dest_is_out, which, ncomp)
                if res is not None:
                    cur, which = res
                    cur_is_user_input = False
                    cur_packed = False
                    skip_to = i+3
                    continue
            full_out_shape = lead + s.out_shape
            # fused all-to-all: the axis kernel before an exchange writes the send layout, the one
            # after it reads the receive layout (see transfer.SlabExchange / shb_fft_c2c_chunked)
            out_chunk = in_chunk = 0
            nxt = steps[i+1] if i+1 < n else None
            prv = steps[i-1] if i > 0 else None
            if s.kind == 'axis' and nxt is not None and nxt.kind == 'exchange' and nxt.fuse and nxt.fuse[0]:
                out_chunk = nxt.exchange.chunk_rows(self.direction, 'send')
            if s.kind == 'axis' and prv is not None and prv.kind == 'exchange' and prv.fuse and prv.fuse[1]:
                in_chunk = prv.exchange.chunk_rows(self.direction, 'recv')
            fused_io = bool(out_chunk or in_chunk)
            can_inplace = (s.inplace_ok and s.in_shape == s.out_shape and s.in_dtype == s.out_dtype
                           and not cur_is_user_input and not fused_io and s.kind == 'axis')
            if dest_is_out[i] and not out_chunk and not (s.kind == 'exchange' and s.fuse and s.fuse[1]):
                if can_inplace and cur.data_ptr() == out.data_ptr():
                    dst = cur
                else:
                    dst = out
            elif can_inplace:
                dst = cur
            else:
                dst = self._pool.get(which, full_out_shape, s.out_dtype, u.device)
                if dst.data_ptr() == cur.data_ptr():
                    which ^= 1
                    dst = self._pool.get(which, full_out_shape, s.out_dtype, u.device)
                which ^= 1
            if s.kind == 'exchange':
                if peer_done is not None:
                    # the producer already stored every block into the peers' receive buffers
                    rbuf, hdl = peer_done
                    hdl.barrier()
                    dst = rbuf[:int(np.prod(full_out_shape))].view(full_out_shape)
                    peer_done = None
                    self.space.exchange_impl = 'peer-store (axis kernel writes the peers\' receive buffers over NVLink)'
                else:
                    s.exchange(cur, dst, ncomp, self.direction, src_packed=cur_packed,
                               dst_packed=bool(s.fuse and s.fuse[1]))
                    self.space.exchange_impl = ('nccl all_to_all_single, pack/unpack fused into the axis kernels'
                                                if (cur_packed or (s.fuse and s.fuse[1])) else
                                                'nccl all_to_all_single, separate pack/unpack')
                cur_packed = False
            else:
                ax = s.axis
                outer = ncomp*int(np.prod(s.in_shape[:ax], dtype=np.int64))
                inner = int(np.prod(s.in_shape[ax+1:], dtype=np.int64))
                kw = {}
                if fused_io:
                    kw = dict(in_chunk=in_chunk, out_chunk=out_chunk)
                if out_chunk and nxt.fuse[1] and nxt.exchange.peer_available():
                    try:
                        numel = int(np.prod(full_out_shape))
                        rbuf, hdl = nxt.exchange.peer_buffers(numel, u.device)
                        P_, r_ = nxt.exchange.P, nxt.exchange.rank
                        slot = (numel//P_)*16*r_
                        kw['peers'] = [int(p)+slot for p in hdl.buffer_ptrs]
                        peer_done = (rbuf, hdl)
                    except Exception:  # noqa: BLE001  (no P2P / symmetric memory on this box)
                        nxt.exchange._peer_broken = True
                        kw.pop('peers', None)
                        peer_done = None
                if i == n-1 and getattr(self, '_last_override', None) is not None:
                    self._last_override(cur, dst, outer, inner)
                elif getattr(self, '_overrides', None) and i in self._overrides:
                    self._overrides[i](cur, dst, outer, inner)
                elif bwd:
                    s.base._backward_axis(cur, dst, outer, inner, **kw)
                else:
                    s.base._forward_axis(cur, dst, outer, inner,
                                         scalar_product=(self.direction == 'scalar_product'), **kw)
                cur_packed = bool(out_chunk)
            cur = dst
            cur_is_user_input = False
        if cur.data_ptr() != out.data_ptr():
            out.copy_(cur)
        return out

    # ---- slab exchange through pitched peer buffers ---------------------------------------------
    def _overlapped_exchange(self, i, cur, out, dest_is_out, which, ncomp=1, pre=None):
        """Steps i, i+1, i+2 = [C2C pass whose stores ARE the all-to-all] [exchange] [C2C pass that
        reads the receive buffers] of a 3-D space exchanged between axes 0 and 1, through
        `shb_fft_c2c_lines`.  Two things differ from the step-by-step path:

        * the receive buffers have a row pitch rounded up to 16 complex (256 B): with the odd R2C
          extent (257) the rows of the natural layout start 16 B-aligned only, every 256 B store of
          the producer then straddles 128 B lines and the NVLink write rate drops to ~530 GB/s;
        * optionally (SHB_EXCHANGE_CHUNKS = K > 1) the last axis is cut into K column ranges, each a
          set of complete lines for both passes, and barrier + consumer of range k run on a second
          stream while the producer of range k+1 runs.  Measured on 2 B200 (profiles/): every extra
          piece costs ~50 us (one more symmetric-memory barrier, shorter grids) and the two kernels
          time-share the SMs instead of overlapping, so K = 1 is the default.

        Returns (result tensor, work-pool toggle) or None when the pattern / sizes are not covered
        (the caller then runs the steps one by one).  Replaces one `Alltoallw` per re-alignment
        (tensorproductspace.py:89,183,252)."""
        import os
        from ._lib import ShenB200Error
        steps = self.steps
        sp, se, sc = steps[i], steps[i+1], steps[i+2]
        K = int(os.environ.get('SHB_EXCHANGE_CHUNKS', '1'))
        if os.environ.get('SHB_EXCHANGE_LINES', '1') == '0' or getattr(self.space, '_no_overlap', False):
            return None
        # row pieces of the preceding local pass run under the producer: measured on 2 B200 (1.62 -> 1.64 ms per
        # pair): the producer needs every SM to keep enough remote stores in flight, sharing them gains nothing
        KO = int(os.environ.get('SHB_PRE_OVERLAP', '1'))
        if pre is not None and (KO <= 1 or ncomp != 1 or steps[pre].axis != 2 or sp.axis != 1 or K != 1):
            return None
        if not (sp.kind == 'axis' and se.kind == 'exchange' and sc.kind == 'axis' and se.fuse and all(se.fuse)):
            return None
        ex = se.exchange
        if len(sp.in_shape) != 3 or {sp.axis, sc.axis} != {0, 1} or not ex.peer_available():
            return None
        for st in (sp, sc):
            b = st.base
            if not (b.family() == 'fourier' and b.short_name() == 'C2C' and b.physical_len() in (64, 128, 256, 512, 1024)):
                return None
        bwd = self.direction == 'backward'
        P_, rank = ex.P, ex.rank
        n2 = sp.in_shape[2]
        pitch = ((n2+15)//16)*16
        K = max(1, min(K, pitch//16))
        dev = cur.device
        block = int(np.prod(se.out_shape))//n2*pitch          # pitched receive buffer of ONE component
        try:
            rbuf, hdl = ex.peer_buffers(ncomp*block, dev)
        except Exception:  # noqa: BLE001  (no P2P / symmetric memory on this box)
            ex._peer_broken = True
            return None
        chunk = block//P_                                      # one peer's block, one component
        slot = ncomp*chunk*16*rank
        peers = [int(p)+slot for p in hdl.buffer_ptrs]
        lead = (ncomp,) if ncomp > 1 else ()
        # destination of the consumer pass
        if dest_is_out[i+2] and out.data_ptr() != cur.data_ptr():
            dst = out
        else:
            dst = self._pool.get(which, lead+sc.out_shape, sc.out_dtype, dev)
            if dst.data_ptr() == cur.data_ptr():
                which ^= 1
                dst = self._pool.get(which, lead+sc.out_shape, sc.out_dtype, dev)
            which ^= 1
        bp, bc = sp.base, sc.base
        npf, Np = bp.physical_len(), bp.N
        ncf, Nc = bc.physical_len(), bc.N
        if pre is not None:
            # intermediate of the local pass (never the user's output: the producer reads it)
            spre = steps[pre]
            mid = self._pool.get(which, lead+spre.out_shape, spre.out_dtype, dev)
            if mid.data_ptr() in (cur.data_ptr(), dst.data_ptr()):
                which ^= 1
                mid = self._pool.get(which, lead+spre.out_shape, spre.out_dtype, dev)
                if mid.data_ptr() in (cur.data_ptr(), dst.data_ptr()):
                    return None
            which ^= 1
            pre_in, cur = cur, mid
        curf, dstf, rflat = cur.reshape(-1), dst.reshape(-1), rbuf.reshape(-1)
        if bwd:
            # producer: along axis 0 of (n0, s1, n2), lines (j, c) of ONE component per launch;
            # consumer: along axis 1 of (ncomp*g0, n1, n2), lines (a, c)
            s1, g0 = sp.in_shape[1], sc.out_shape[0]
            cr_p, cr_c = npf//P_, Nc//P_
            comp_in, comp_out = Np*s1*n2, 0
            prod = dict(na=s1, in_as=n2, in_rs=s1*n2, out_as=pitch, out_rs=s1*pitch, out_cr=cr_p, per_comp=True,
                        in_comp=Np*s1*n2, peer_comp=cr_p*s1*pitch)
            cons = dict(na=ncomp*g0, in_as=cr_c*pitch, in_rs=pitch, out_as=ncf*n2, out_rs=n2, in_cr=cr_c,
                        in_cs=ncomp*g0*cr_c*pitch, per_comp=False)
            pscale, cscale = 1.0, 1.0
            pdd, cdd = bp.dealias_direct, bc.dealias_direct
        else:
            # producer: along axis 1 of (ncomp*g0, n1, n2), lines (a, c);
            # consumer: along axis 0 of (n0, s1, n2), lines (j, c) of ONE component per launch
            g0, s1 = sp.in_shape[0], sc.out_shape[1]
            cr_p, cr_c = Np//P_, ncf//P_
            prod = dict(na=ncomp*g0, in_as=npf*n2, in_rs=n2, out_as=cr_p*pitch, out_rs=pitch, out_cr=cr_p, per_comp=False)
            cons = dict(na=s1, in_as=pitch, in_rs=s1*pitch, out_as=n2, out_rs=s1*n2, in_cr=cr_c,
                        in_cs=ncomp*cr_c*s1*pitch, per_comp=True, in_comp=cr_c*s1*pitch, out_comp=Nc*s1*n2)
            pscale, cscale = 1./npf, 1./ncf
            pdd = cdd = False

        def producer(c0, cw, cv):
            if prod['per_comp']:
                for q in range(ncomp):
                    pq = [p+16*q*prod['peer_comp'] for p in peers]
                    ops.fft_c2c_lines(bwd, curf[q*prod['in_comp']:], None, pq, prod['na'], cw, c0, prod['in_as'],
                                      prod['in_rs'], prod['out_as'], prod['out_rs'], npf, Np, pscale, pdd, 0, 0,
                                      prod['out_cr'], 0, ncols_valid=cv)
            else:
                ops.fft_c2c_lines(bwd, curf, None, peers, prod['na'], cw, c0, prod['in_as'], prod['in_rs'], prod['out_as'],
                                  prod['out_rs'], npf, Np, pscale, pdd, 0, 0, prod['out_cr'], 0, ncols_valid=cv)

        def consumer(c0, cw, cv):
            if cons['per_comp']:
                for q in range(ncomp):
                    ops.fft_c2c_lines(bwd, rflat[q*cons['in_comp']:], dstf[q*cons['out_comp']:], None, cons['na'], cw, c0,
                                      cons['in_as'], cons['in_rs'], cons['out_as'], cons['out_rs'], ncf, Nc, cscale, cdd,
                                      cons['in_cr'], cons['in_cs'], 0, 0, ncols_valid=cv)
            else:
                ops.fft_c2c_lines(bwd, rflat, dstf, None, cons['na'], cw, c0, cons['in_as'], cons['in_rs'], cons['out_as'],
                                  cons['out_rs'], ncf, Nc, cscale, cdd, cons['in_cr'], cons['in_cs'], 0, 0, ncols_valid=cv)

        # column pieces: boundaries on multiples of 16 columns, the last piece carries the padding
        nt = pitch//16
        bounds = [16*round(k*nt/K) for k in range(K+1)]
        try:
            if pre is not None:
                # forward only: rows [a0, a1) of the leading axis go through the local pass (HBM-bound) on
                # the main stream and then through the producer (NVLink-bound) on a second stream
                st = self.space.__dict__.get('_ovl_streams')
                if st is None:
                    st = self.space._ovl_streams = (torch.cuda.Stream(device=dev), [torch.cuda.Event() for _ in range(16)])
                side, events = st
                main = torch.cuda.current_stream()
                g0_ = spre.in_shape[0]
                KO = max(1, min(KO, g0_))
                rb = [round(k*g0_/KO) for k in range(KO+1)]
                pin, pmid = pre_in.reshape(spre.in_shape), cur.reshape(spre.out_shape)
                inner_rows = int(np.prod(spre.in_shape[1:2]))
                for k in range(KO):
                    a0, a1 = rb[k], rb[k+1]
                    spre.base._forward_axis(pin[a0:a1], pmid[a0:a1], (a1-a0)*inner_rows, 1,
                                            scalar_product=(self.direction == 'scalar_product'))
                    ev = events[k % len(events)]
                    ev.record(main)
                    side.wait_event(ev)
                    with torch.cuda.stream(side):
                        pk = [p+16*a0*prod['out_as'] for p in peers]
                        ops.fft_c2c_lines(bwd, curf[a0*prod['in_as']:], None, pk, a1-a0, pitch, 0, prod['in_as'],
                                          prod['in_rs'], prod['out_as'], prod['out_rs'], npf, Np, pscale, pdd, 0, 0,
                                          prod['out_cr'], 0, ncols_valid=n2)
                done = events[KO % len(events)]
                done.record(side)
                main.wait_event(done)
                hdl.barrier()
                consumer(0, n2, n2)
            elif K == 1:
                tm = getattr(self.space, '_time_producer', None)     # (start, end) CUDA events: bench.py's roofline
                if tm is not None:
                    tm[0].record()
                producer(0, pitch, n2)
                if tm is not None:
                    tm[1].record()
                hdl.barrier()
                consumer(0, n2, n2)
            else:
                st = self.space.__dict__.get('_ovl_streams')
                if st is None:
                    prio = {'cons': -1, 'prod': 0, 'none': 0}[os.environ.get('SHB_OVL_PRIO', 'cons')]
                    st = self.space._ovl_streams = (torch.cuda.Stream(device=dev, priority=prio),
                                                    [torch.cuda.Event() for _ in range(16)])
                side, events = st
                main = torch.cuda.current_stream()
                for k in range(K):
                    c0, cw = bounds[k], bounds[k+1]-bounds[k]
                    cv = min(cw, n2-c0)
                    producer(c0, cw, cv)
                    ev = events[k % len(events)]
                    ev.record(main)
                    side.wait_event(ev)
                    with torch.cuda.stream(side):
                        hdl.barrier(channel=k % 8)
                        consumer(c0, cv, cv)
                done = events[K % len(events)]
                done.record(side)
                main.wait_event(done)
        except ShenB200Error as e:
            if 'fft_c2c_lines' not in str(e):
                raise
            self.space._no_overlap = True
            return None
        self.space.exchange_impl = ('peer-store into 256 B-pitched receive buffers (axis kernel writes the peers\' memory '
                                    'over NVLink)' + (f', {K} overlapped column pieces' if K > 1 else '')
                                    + (f'; forward: preceding local pass pipelined under it in {KO} row pieces' if pre is not None
                                       else ''))
        return dst, which

    def __call__(self, input_array, output_array=None, kind=None, **kw):
        space = self.space
        fo = self.direction != 'backward'
        if output_array is None:
            output_array = Function(space, val=None) if fo else Array(space, val=None)
        u = ops.as_device_tensor(input_array)
        out = ops.plain(output_array)
        ncomp = 1
        in_shape = tuple(u.shape)
        if in_shape != self.input_shape:
            if in_shape[1:] == self.input_shape or in_shape[-len(self.input_shape):] == self.input_shape:
                ncomp = int(np.prod(in_shape[:len(in_shape)-len(self.input_shape)]))
            else:
                raise ValueError(f'{self.direction}: input shape {in_shape} does not match plan '
                                 f'{self.input_shape}')
        if tuple(out.shape[len(out.shape)-len(self.output_shape):]) != self.output_shape:
            raise ValueError(f'{self.direction}: output shape {tuple(out.shape)} does not match plan '
                             f'{self.output_shape}')
        if u.dtype != self.steps[0].in_dtype:
            u = u.to(self.steps[0].in_dtype)
        if not u.is_contiguous():
            u = u.contiguous()
        if not out.is_contiguous():
            raise ValueError('output array must be contiguous')
        self._run(u, out, ncomp)
        return output_array


class TensorProductSpace:
    """Cartesian product of 1-D bases with device transforms.

    Parameters follow shenfun/tensorproductspace.py:262-316.
    """

    def __init__(self, comm, bases, axes=None, dtype=None, slab=True, collapse_fourier=False,
                 backward_from_pencil=False, coordinates=None, modify_spaces_inplace=False, **kw):
        self.comm = default_comm if comm is None else comm
        # Only the slab decomposition of mpi4py-fft (`slab=True`, one exchange; :393-396) exists on the
        # device path: a pencil request (slab=False with > 1 rank, an explicit `Subcomm` or a tuple /
        # list process grid as in demo/poisson3D.py:67) must not silently run something else.
        if isinstance(self.comm, (tuple, list)) or type(self.comm).__name__ == 'Subcomm':
            dims = list(self.comm.dims) if hasattr(self.comm, 'dims') else list(self.comm)
            if sum(1 for d in dims if d != 1) > 1:
                raise NotImplementedError('pencil decompositions (process grid %r) are not built: the device '
                                          'path is slab-distributed over <= 8 GPUs' % (dims,))
            self.comm = getattr(self.comm, 'comm', default_comm) if not isinstance(self.comm, (tuple, list)) \
                else default_comm
        if not slab and self.comm.Get_size() > 1 and len(bases) > 2:
            raise NotImplementedError('slab=False (pencil decomposition) is not built on the device path; '
                                      'use slab=True (one all-to-all per transform)')
        if collapse_fourier and self.comm.Get_size() > 1:
            raise NotImplementedError('collapse_fourier=True is not built on the device path')
        if not modify_spaces_inplace:
            bases = [b.get_unplanned() for b in bases]
        self.bases = tuple(bases)
        nd = len(self.bases)
        for i, b in enumerate(self.bases):
            b.axis = i
            b.tensorproductspace = self
        if axes is None:
            axes = tuple(range(nd))
        flat = []
        for a in axes:
            if isinstance(a, (tuple, list)):
                flat.extend(int(x) % nd for x in a)
            else:
                flat.append(int(a) % nd)
        assert sorted(flat) == list(range(nd)), 'axes must be a permutation'
        self.axes = tuple(flat)
        self._padded_space = {}
        self._pool = _WorkPool()
        self.coors = None
        self._fuse_exchange = bool(kw.pop('fuse_exchange', True))
        self.exchange_impl = 'none (single rank)'     # set by the plans: which exchange implementation ran

        first = self.bases[self.axes[-1]]
        # tensorproductspace.py:365-385: real input needs an R2C (or a real
        # polynomial base) transformed first
        if dtype is None:
            dtype = first.dtype
        dtype = np.dtype(dtype)
        self.physical_dtype = torch.float64 if dtype.kind == 'f' else torch.complex128
        for ax in self.axes[:-1]:
            b = self.bases[ax]
            if b.family() == 'fourier' and b.short_name() == 'R2C' and nd > 1:
                raise ValueError('an R2C base must be the first transformed axis (last in axes)')
        if first.family() == 'fourier' and first.short_name() == 'C2C':
            self.physical_dtype = torch.complex128

        P = self.comm.Get_size()
        rank = self.comm.Get_rank()
        self._P, self._rank = P, rank
        fwd_order = list(reversed(self.axes))
        self._dist_phys = None
        self._dist_spec = None
        if P > 1:
            assert nd >= 2, 'cannot distribute a 1-D space'
            dp = (fwd_order[0]+1) % nd
            i = fwd_order.index(dp)
            assert i >= 1
            self._dist_phys = dp
            self._dist_spec = fwd_order[i-1]

        # global shapes
        self._gshape_phys = tuple(b.physical_len() for b in self.bases)
        self._gshape_spec = tuple(b.spectral_len() for b in self.bases)
        self._build_plans()

    # ---- planning -------------------------------------------------------------------
    def _local(self, gshape, dist_axis):
        shape = list(gshape)
        start = [0]*len(shape)
        if dist_axis is not None:
            start[dist_axis], shape[dist_axis] = _block(gshape[dist_axis], self._P, self._rank)
        return tuple(shape), tuple(start)

    def _build_plans(self):
        from .transfer import SlabExchange
        nd = len(self.bases)
        fwd_order = list(reversed(self.axes))
        # forward: physical -> spectral
        gshape = list(self._gshape_phys)
        dist = self._dist_phys
        dt = self.physical_dtype
        fsteps = []
        for i, ax in enumerate(fwd_order):
            b = self.bases[ax]
            if dist is not None and ax == dist:
                new_dist = fwd_order[i-1]
                in_shape, _ = self._local(gshape, dist)
                out_shape, _ = self._local(gshape, new_dist)
                ex = SlabExchange(self.comm, tuple(gshape), gather_axis=dist, scatter_axis=new_dist)
                st = _Step('exchange', None, None, in_shape, out_shape, dt, dt, exchange=ex)
                # fuse the pack into the kernel of the axis transformed just before (new_dist) and the
                # unpack into the kernel of the axis transformed next (ax), when both can address
                # the peer-chunked layout and the blocks are equal
                prod = fsteps[-1].base if fsteps and fsteps[-1].kind == 'axis' else None
                can = ex.uniform() and dt == torch.complex128 and self._fuse_exchange
                st.fuse = [bool(can and prod is not None and getattr(prod, 'supports_chunked', False)
                                and fsteps[-1].axis == new_dist),
                           bool(can and getattr(b, 'supports_chunked', False))]
                fsteps.append(st)
                dist = new_dist
            in_shape, _ = self._local(gshape, dist)
            gshape[ax] = b.spectral_len()
            out_shape, _ = self._local(gshape, dist)
            odt = b.spectral_dtype_for(dt)
            fsteps.append(_Step('axis', ax, b, in_shape, out_shape, dt, odt, inplace_ok=b.inplace_ok()))
            dt = odt
        self.spectral_dtype = dt
        assert dist == self._dist_spec
        # backward = mirror image
        bsteps = []
        for s in reversed(fsteps):
            b_ = _Step(s.kind, s.axis, s.base, s.out_shape, s.in_shape, s.out_dtype, s.in_dtype,
                       inplace_ok=s.inplace_ok, exchange=s.exchange)
            if s.fuse:
                b_.fuse = [s.fuse[1], s.fuse[0]]      # producer / consumer swap roles
            bsteps.append(b_)
        self.forward = _Transform(self, fsteps, 'forward')
        self.backward = _Transform(self, bsteps, 'backward')
        self.scalar_product = _Transform(self, fsteps, 'scalar_product')

    # ---- inhomogeneous boundary conditions in a tensor-product space ----------------------------
    def get_nonhomogeneous_axes(self):
        """tensorproductspace.py:1085-1095"""
        return [i for i, b in enumerate(self.bases) if getattr(b, 'has_nonhomogeneous_bcs', False)]

    def _tensor_bcs(self, base):
        """Boundary values of `base` as the plan sees them (`BoundaryValues.set_tensor_bcs`,
        tensorproductspace.py:1909-2032): the boundary data, a function of the other coordinates, is
        stored in the rows of the boundary dofs of a physical array and taken through the forward
        plan with a do-nothing pass in place of this base's own transform.  Returns
        (partial, final): `partial[b]` [outer][inner] are the values at the moment this base is
        transformed (axes transformed earlier already spectral: what `_add_mass_rhs` needs, :2091-2106),
        `final[b]` the fully transformed values (what `set_boundary_dofs(final=True)` writes, :2108-2128)."""
        ax = base.axis
        cache = self.__dict__.setdefault('_tbc_cache', {})
        if ax in cache:
            return cache[ax]
        if len(self.get_nonhomogeneous_axes()) > 1:
            raise NotImplementedError('inhomogeneous boundary conditions along more than one axis')
        conds = base.bcs.conditions()
        nbc = len(conds)
        nd = self.dimensions
        fw = self.forward
        idx = [i for i, s in enumerate(fw.steps) if s.kind == 'axis' and s.axis == ax][0]
        b = torch.zeros(self.shape(False), dtype=self.physical_dtype, device=ops.device())
        # the boundary rows live at the tail of the (padded) physical axis; that axis is never the
        # distributed one when its own transform runs, but it can be in physical space
        sl_phys = self.local_slice(False)
        n_phys = base.physical_len()
        mesh = None
        for j, (_, _, val) in enumerate(conds):
            row = n_phys-nbc+j
            if not (sl_phys[ax].start <= row < sl_phys[ax].stop):
                continue
            sel = [slice(None)]*nd
            sel[ax] = row-sl_phys[ax].start
            if isinstance(val, (int, float, complex, np.number)):
                b[tuple(sel)] = val
            else:
                import sympy as sp
                e = sp.sympify(val)
                syms = sorted(e.free_symbols, key=str)
                if mesh is None:
                    mesh = self.local_mesh(bcast=True)
                args = []
                for sym in syms:
                    k = 'xyzrs'.index(str(sym))
                    assert k != ax and k < nd, 'boundary data may depend on the other coordinates only'
                    take = [slice(None)]*nd
                    take[ax] = 0
                    args.append(np.asarray(mesh[k])[tuple(take)])
                data = sp.lambdify(syms, e, 'numpy')(*args) if syms else complex(e)
                b[tuple(sel)] = ops.as_device_tensor(np.ascontiguousarray(np.broadcast_to(
                    data, tuple(n for i, n in enumerate(self.shape(False)) if i != ax)))).to(b.dtype)
        dim = base.dim()
        grabbed = {}

        def do_nothing(cur, dst, outer, inner):
            v = cur.view(outer, n_phys, inner)
            o = dst.view(outer, base.N, inner)
            o.zero_()
            o[:, dim:dim+nbc, :] = v[:, n_phys-nbc:, :]
            grabbed['partial'] = o[:, dim:dim+nbc, :].transpose(0, 1).contiguous()   # [nbc][outer][inner]
        out = torch.empty(self.shape(True), dtype=self.spectral_dtype, device=b.device)
        fw._run(b, out, 1, overrides={idx: do_nothing})
        fw._overrides = None
        sel = [slice(None)]*nd
        sel[ax] = slice(dim, dim+nbc)
        final = out[tuple(sel)].movedim(ax, 0).contiguous()                          # [nbc] + other axes
        cache[ax] = (grabbed['partial'], final)
        return cache[ax]

    def _set_boundary_dofs(self, u):
        """Function.set_boundary_dofs (forms/arguments.py:1525-1539; final values, :2108-2128)."""
        t = ops.plain(u)
        for base in self.bases:
            nbc = base.bcs.num_bcs() if hasattr(base, 'bcs') else 0
            if not nbc:
                continue
            ax, dim = base.axis, base.dim()
            sel = [slice(None)]*self.dimensions
            sel[ax] = slice(dim, dim+nbc)
            lead = t.dim()-self.dimensions
            tgt = t[(slice(None),)*lead+tuple(sel)]
            if getattr(base, 'has_nonhomogeneous_bcs', False):
                tgt.copy_(self._tensor_bcs(base)[1].movedim(0, ax))
            else:
                tgt.zero_()
        return u

    def backward_cross(self, u_hat, b, out):
        """out = backward(u_hat) x b for three-component fields: the last (complex-to-real) axis
        pass is fused with the cross product (`shb_fft_c2r_cross`), falling back to a separate
        cross-product pass when no fused kernel covers the size.  u_hat: (3,)+spectral shape,
        b / out: (3,)+physical shape (out may not alias b)."""
        from ._lib import ShenB200Error
        uh, bt, ot = ops.as_device_tensor(u_hat), ops.plain(b), ops.plain(out)
        last = self.backward.steps[-1]
        base = last.base
        ok = (last.kind == 'axis' and base.family() == 'fourier' and base.short_name() == 'R2C'
              and last.axis == self.dimensions-1 and not getattr(self, '_no_cross_fusion', False))
        if ok:
            def fused(cur, dst, outer, inner):
                assert inner == 1 and outer % 3 == 0
                ops.fft_c2r_cross(cur, bt, dst, outer//3, base.physical_len(), base.N, 1.0, base.dealias_direct)
            try:
                self.backward._run(uh if uh.is_contiguous() else uh.contiguous(), ot, 3, last_override=fused)
                return out
            except ShenB200Error as e:
                if 'fft_c2r_cross' not in str(e):
                    raise
                self._no_cross_fusion = True
            finally:
                self.backward._last_override = None
        tmp = torch.empty_like(ot)
        self.backward._run(uh if uh.is_contiguous() else uh.contiguous(), tmp, 3)
        ops.cross3d(tmp, bt, ot)
        return out

    def curl_fusion_available(self):
        """True when `backward_cross_hat(u_hat, out, curl_of=K)` can form i k x u_hat inside the first padded
        pass (`shb_fft_c2c_backward_curl`): one rank, Fourier^3 with 3/2-rule padding, the leading axis
        transformed first and served by the residue-split padding kernel.  Opt-in (SHB_NS_CURL_FUSION=1):
        measured on B200 at 512^3 it saves the 6.5 GB per stage of the stored vorticity spectrum but the
        two-loads-per-element first pass gives the time back (122.8 vs 121.0 ms per step)."""
        st = self.backward.steps
        if self.comm.Get_size() != 1 or self.dimensions != 3 or len(st) != 3 or any(s.kind != 'axis' for s in st):
            return False
        s0, b0 = st[0], st[0].base
        return (s0.axis == 0 and b0.family() == 'fourier' and b0.short_name() == 'C2C' and not b0.dealias_direct
                and 2*b0.physical_len() == 3*b0.N and b0.N//2 in (64, 128, 256, 512)
                and os.environ.get('SHB_NS_CURL_FUSION') == '1')

    def backward_cross_hat(self, ab_hat, out, curl_of=None):
        """out = backward(a_hat) x backward(b_hat) with both spectral fields stacked in one
        (6,)+spectral array `ab_hat` = [a_hat; b_hat]: the earlier axis passes run once over the six
        components, the last pass transforms all six lines of a spatial line in one kernel and
        writes only the cross product (`shb_fft_c2r_cross`, spectral `b`).  Falls back to
        `backward` + `backward_cross` when no fused kernel covers the size.

        `curl_of=(k0, k1, k2)` (flat wavenumber tables; requires `curl_fusion_available()`): `ab_hat` is the
        (3,)+spectral field a_hat alone and b_hat = i k x a_hat is formed inside the first pass."""
        from ._lib import ShenB200Error
        ab, ot = ops.as_device_tensor(ab_hat), ops.plain(out)
        assert ab.shape[0] == (6 if curl_of is None else 3) and ab.is_contiguous()
        last = self.backward.steps[-1]
        base = last.base
        ok = (last.kind == 'axis' and base.family() == 'fourier' and base.short_name() == 'R2C'
              and last.axis == self.dimensions-1 and not getattr(self, '_no_cross_fusion', False))
        overrides = None
        if curl_of is not None:
            assert ok and self.curl_fusion_available(), 'curl fusion is not available for this space'
            b0 = self.backward.steps[0].base
            k0, k1, k2 = curl_of

            def first(cur, dst, outer, inner):
                assert outer == 6 and inner % k2.numel() == 0
                ops.fft_c2c_backward_curl(cur, dst, inner, k2.numel(), b0.physical_len(), b0.N, k0, k1, k2)
            overrides = {0: first}
        if ok:
            def fused(cur, dst, outer, inner):
                assert inner == 1 and outer % 6 == 0
                half = cur.view(2, -1)
                ops.fft_c2r_cross(half[0], half[1], dst, outer//6, base.physical_len(), base.N, 1.0,
                                  base.dealias_direct)
            try:
                self.backward._run(ab, ot, 6, last_override=fused, overrides=overrides)
                return out
            except ShenB200Error as e:
                if 'fft_c2r_cross' not in str(e) or curl_of is not None:
                    raise
                self._no_cross_fusion = True
            finally:
                self.backward._last_override = None
                self.backward._overrides = None
        bp = torch.empty_like(ot)
        self.backward._run(ab[3:], bp, 3)
        return self.backward_cross(ab[:3], bp, out)

    def eval(self, points, coefficients, output_array=None, method=1):
        """Function values at arbitrary points from the expansion coefficients
        (tensorproductspace.py:668-854).  points: (D, npoints).  Every axis contributes a basis
        matrix E_a[p, k] evaluated on the host; the coefficients are contracted axis by axis on the
        device (not a hot path: O(npoints * N^D)).  For a real-to-complex axis the conjugate half
        of the spectrum is added as in the reference (:718-737): modes 1 .. N/2-1 (N even) or
        1 .. (N-1)/2 (N odd).  Ranks hold their local block; the result is summed over ranks."""
        pts = np.atleast_2d(np.asarray(points, dtype=float))
        assert pts.shape[0] == self.dimensions, 'points must have shape (D, npoints)'
        npts = pts.shape[1]
        uh = ops.as_device_tensor(coefficients)
        sl = self.local_slice(True)
        dev = uh.device
        letters = 'abcd'[:self.dimensions]
        mats, r2c_axis, conj_mask = [], None, None
        for ax, base in enumerate(self.bases):
            E = np.asarray(base.evaluate_basis_all(pts[ax]))[:, sl[ax]]
            if base.family() == 'fourier' and base.short_name() == 'R2C':
                r2c_axis = ax
                k = np.arange(base.spectral_len())[sl[ax]]
                last = base.N//2 if base.N % 2 == 0 else base.N//2+1      # first index NOT conjugated
                conj_mask = ((k >= 1) & (k < last)).astype(float)
            mats.append(torch.as_tensor(np.ascontiguousarray(E), device=dev).to(torch.complex128))
        u = uh.to(torch.complex128)
        expr = ','.join('p'+c for c in letters)+','+letters+'->p'
        val = torch.einsum(expr, *mats, u)
        if r2c_axis is not None:
            m2 = list(mats)
            m2[r2c_axis] = mats[r2c_axis]*torch.as_tensor(conj_mask, device=dev)
            val = val+torch.conj(torch.einsum(expr, *m2, u))
        if self.comm.Get_size() > 1:
            import torch.distributed as dist
            buf = torch.view_as_real(val.contiguous())
            dist.all_reduce(buf, group=self.comm.group)
            val = torch.view_as_complex(buf)
        real_out = not self.physical_dtype.is_complex
        res = (val.real if real_out else val).cpu().numpy()
        if output_array is not None:
            output_array[...] = res
            return output_array
        return res

    # ---- bookkeeping API (mpi4py_fft.PFFT + tensorproductspace.py:580-1180) ------------
    @property
    def dimensions(self):
        return len(self.bases)

    @property
    def rank(self):
        return 0

    @property
    def is_composite_space(self):
        return 0

    @property
    def is_padded(self):
        return any(b.is_padded for b in self.bases)

    def num_components(self):
        return 1

    def flatten(self):
        return [self]

    def __len__(self):
        return len(self.bases)

    def __getitem__(self, i):
        return self.bases[i]

    def __iter__(self):
        return iter(self.bases)

    def global_shape(self, forward_output=False):
        return self._gshape_spec if forward_output else self._gshape_phys

    def shape(self, forward_output=True):
        """Local shape of this rank's block."""
        if forward_output:
            return self._local(self._gshape_spec, self._dist_spec)[0]
        return self._local(self._gshape_phys, self._dist_phys)[0]

    def local_shape(self, forward_output=True):
        return self.shape(forward_output)

    def local_slice(self, forward_output=True):
        if forward_output:
            shape, start = self._local(self._gshape_spec, self._dist_spec)
        else:
            shape, start = self._local(self._gshape_phys, self._dist_phys)
        return tuple(slice(s, s+n) for s, n in zip(start, shape))

    def dims(self):
        return tuple(b.dim() for b in self.bases)

    def dim(self):
        return int(np.prod(self.dims()))

    def size(self, forward_output=False):
        return int(np.prod(self.shape(forward_output)))

    def slice(self):
        return tuple(b.slice() for b in self.bases)

    @property
    def domains(self):
        return [tuple(b.domain) for b in self.bases]

    @property
    def volume(self):
        """tensorproductspace.py:1030-1036"""
        vol = 1.0
        for d in self.domains:
            vol *= float(d[1]-d[0])
        return vol

    def get_nonperiodic_axes(self):
        return [i for i, b in enumerate(self.bases) if b.family() != 'fourier']

    def get_nondiagonal_axes(self):
        return self.get_nonperiodic_axes()

    def is_orthogonal(self):
        return all(getattr(b, 'is_orthogonal', True) for b in self.bases)

    def destroy(self):
        self._pool.buf = [None, None]
        for sp in self._padded_space.values():
            sp.destroy()

    # ---- meshes and wavenumbers --------------------------------------------------------
    def mesh(self, map_true_domain=True):
        return [b.mesh(bcast=True, map_true_domain=map_true_domain) for b in self.bases]

    def local_mesh(self, bcast=False, map_true_domain=True, device=False):
        """tensorproductspace.py:950-983.  Host numpy arrays by default; device=True
        gives float64 CUDA tensors (broadcastable) for on-device initialisation."""
        sl = self.local_slice(False)
        X = [x[tuple(s if i == ax else slice(None) for i in range(self.dimensions))]
             for ax, (x, s) in enumerate(zip(self.mesh(map_true_domain), sl))]
        if bcast:
            X = [np.broadcast_to(x, self.shape(False)) for x in X]
        if device:
            X = [ops.as_device_tensor(np.ascontiguousarray(x)) for x in X]
        return X

    def wavenumbers(self, scaled=False, eliminate_highest_freq=False):
        return [b.wavenumbers(bcast=True, scaled=scaled, eliminate_highest_freq=eliminate_highest_freq)
                for b in self.bases]

    def local_wavenumbers(self, broadcast=False, scaled=False, eliminate_highest_freq=False, device=False):
        """tensorproductspace.py:1018-1052"""
        sl = self.local_slice(True)
        K = self.wavenumbers(scaled=scaled, eliminate_highest_freq=eliminate_highest_freq)
        K = [k[tuple(s if i == ax else slice(None) for i in range(self.dimensions))]
             for ax, (k, s) in enumerate(zip(K, sl))]
        if broadcast:
            K = [np.broadcast_to(k, self.shape(True)) for k in K]
        if device:
            K = [ops.as_device_tensor(np.ascontiguousarray(k, dtype=float)) for k in K]
        return K

    def get_mask_nyquist(self):
        """tensorproductspace.py:1063-1076: product of the per-axis Nyquist masks (local)."""
        sl = self.local_slice(True)
        mask = None
        for ax, b in enumerate(self.bases):
            f = b.get_mask_nyquist(bcast=True)
            if f is None:
                continue
            f = f[tuple(sl[ax] if i == ax else slice(None) for i in range(self.dimensions))]
            mask = f if mask is None else mask*f
        return mask

    def mask_nyquist(self, u_hat, mask=None):
        """Zero Nyquist coefficients in place (utilities/__init__.py:441-444 `apply_mask`)."""
        if mask is None:
            mask = self._device_mask()
        if mask is not None:
            t = ops.plain(u_hat)
            if not isinstance(mask, torch.Tensor):
                mask = ops.as_device_tensor(np.ascontiguousarray(np.broadcast_to(mask, self.shape(True)), dtype=float))
            if t.is_complex() and t.is_contiguous() and mask.dtype == torch.float64 and mask.is_contiguous() \
                    and t.numel() % mask.numel() == 0 and tuple(t.shape[t.dim()-mask.dim():]) == tuple(mask.shape):
                ops.apply_mask(t, mask)               # shb_apply_mask (applymask.pyx:15-169)
            else:
                t.mul_(mask)
        return u_hat

    def _device_mask(self):
        if not hasattr(self, '_mask_dev'):
            m = self.get_mask_nyquist()
            self._mask_dev = None if m is None else ops.as_device_tensor(
                np.ascontiguousarray(np.broadcast_to(m, self.shape(True)), dtype=float))
        return self._mask_dev

    # ---- derived spaces ----------------------------------------------------------------
    def get_dealiased(self, padding_factor=1.5, dealias_direct=False):
        """tensorproductspace.py:543-578: padded twin with the same spectral layout."""
        if isinstance(padding_factor, (int, float)):
            padding_factor = (padding_factor,)*len(self.bases)
        padding_factor = tuple(float(p) for p in padding_factor)
        key = (padding_factor, bool(dealias_direct))
        if key in self._padded_space:
            return self._padded_space[key]
        if all(abs(p-1) < 1e-8 for p in padding_factor) and not dealias_direct:
            return self
        bases = [b.get_dealiased(padding_factor=p, dealias_direct=dealias_direct)
                 for b, p in zip(self.bases, padding_factor)]
        sp = TensorProductSpace(self.comm, bases, axes=self.axes, modify_spaces_inplace=True)
        assert sp.shape(True) == self.shape(True)
        self._padded_space[key] = sp
        return sp

    def get_unplanned(self, **kw):
        return TensorProductSpace(self.comm, [b.get_unplanned(**kw) for b in self.bases], axes=self.axes)

    def get_refined(self, N):
        return TensorProductSpace(self.comm, [b.get_refined(n) for b, n in zip(self.bases, N)], axes=self.axes)

    def get_orthogonal(self):
        return TensorProductSpace(self.comm, [b.get_orthogonal() for b in self.bases], axes=self.axes)


class CompositeSpace:
    """Product of tensor-product spaces (tensorproductspace.py:1190-1560).

    Components must share local shapes so that the component index can be folded
    into the batch extent of every axis kernel.
    """

    def __init__(self, spaces):
        self.spaces = list(spaces)
        s0 = self.flatten()[0]
        for s in self.flatten():
            if s.shape(True) != s0.shape(True) or s.shape(False) != s0.shape(False):
                raise NotImplementedError('components with different shapes')
        self.comm = s0.comm
        self._padded_space = {}
        self.forward = _CompositeTransform(self, 'forward')
        self.backward = _CompositeTransform(self, 'backward')
        self.scalar_product = _CompositeTransform(self, 'scalar_product')

    @property
    def physical_dtype(self):
        return self.flatten()[0].physical_dtype

    @property
    def spectral_dtype(self):
        return self.flatten()[0].spectral_dtype

    @property
    def dimensions(self):
        return self.flatten()[0].dimensions

    @property
    def rank(self):
        return 1

    @property
    def is_composite_space(self):
        return 1

    @property
    def tensor_rank(self):
        return 1

    def num_components(self):
        return len(self.flatten())

    def flatten(self):
        out = []
        for s in self.spaces:
            out.extend(s.flatten())
        return out

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)

    def __iter__(self):
        return iter(self.spaces)

    def _lead(self):
        return (len(self.spaces),)

    def shape(self, forward_output=True):
        return self._lead() + tuple(self.spaces[0].shape(forward_output))

    def local_shape(self, forward_output=True):
        return self.shape(forward_output)

    def global_shape(self, forward_output=False):
        return self._lead() + tuple(self.spaces[0].global_shape(forward_output))

    def local_slice(self, forward_output=True):
        return tuple(slice(0, n) for n in self._lead()) + tuple(self.flatten()[0].local_slice(forward_output))

    def local_mesh(self, *a, **k):
        return self.flatten()[0].local_mesh(*a, **k)

    def mesh(self, *a, **k):
        return self.flatten()[0].mesh(*a, **k)

    def local_wavenumbers(self, *a, **k):
        return self.flatten()[0].local_wavenumbers(*a, **k)

    def get_mask_nyquist(self):
        return self.flatten()[0].get_mask_nyquist()

    def mask_nyquist(self, u_hat, mask=None):
        return self.flatten()[0].mask_nyquist(u_hat, mask)

    def get_dealiased(self, padding_factor=1.5, dealias_direct=False):
        key = (str(padding_factor), bool(dealias_direct))
        if key not in self._padded_space:
            cache = {}
            new = []
            for s in self.spaces:
                if id(s) not in cache:
                    cache[id(s)] = s.get_dealiased(padding_factor, dealias_direct)
                new.append(cache[id(s)])
            self._padded_space[key] = type(self)(new) if type(self) is CompositeSpace else type(self)(new[0])
        return self._padded_space[key]

    def get_orthogonal(self):
        return type(self)([s.get_orthogonal() for s in self.spaces]) if type(self) is CompositeSpace \
            else type(self)(self.spaces[0].get_orthogonal())

    def destroy(self):
        for s in self.flatten():
            s.destroy()


class _CompositeTransform:
    """Batched transform of all components (replaces the per-component loop of
    `VectorTransform.__call__`, tensorproductspace.py:1786-1795)."""

    def __init__(self, space, direction):
        self.space = space
        self.direction = direction

    def __call__(self, input_array, output_array=None, kind=None, **kw):
        space = self.space
        fo = self.direction != 'backward'
        if output_array is None:
            output_array = Function(space, val=None) if fo else Array(space, val=None)
        flat = space.flatten()
        same = all(s is flat[0] for s in flat)
        if same:
            getattr(flat[0], self.direction)(input_array, output_array, kind=kind)
        else:
            u = ops.plain(input_array) if isinstance(input_array, torch.Tensor) else input_array
            out = ops.plain(output_array)
            ncomp = len(flat)
            u = u.reshape((ncomp,) + tuple(u.shape[-space.dimensions:]))
            out_v = out.view((ncomp,) + tuple(out.shape[-space.dimensions:]))
            for i, s in enumerate(flat):
                getattr(s, self.direction)(u[i], out_v[i], kind=kind)
        return output_array


class VectorSpace(CompositeSpace):
    """`ndim` copies of one scalar space (tensorproductspace.py:1563-1650)."""

    def __init__(self, space):
        if isinstance(space, (list, tuple)):
            spaces = list(space)
        else:
            spaces = [space]*space.dimensions
        CompositeSpace.__init__(self, spaces)

    def get_dealiased(self, padding_factor=1.5, dealias_direct=False):
        key = (str(padding_factor), bool(dealias_direct))
        if key not in self._padded_space:
            if all(s is self.spaces[0] for s in self.spaces):
                self._padded_space[key] = VectorSpace(self.spaces[0].get_dealiased(padding_factor, dealias_direct))
            else:
                self._padded_space[key] = VectorSpace([s.get_dealiased(padding_factor, dealias_direct)
                                                       for s in self.spaces])
        return self._padded_space[key]


class TensorSpace(CompositeSpace):
    """ndim x ndim copies of one scalar space (tensorproductspace.py:1652-1784)."""

    def __init__(self, space):
        self._scalar = space
        nd = space.dimensions
        CompositeSpace.__init__(self, [space]*(nd*nd))

    @property
    def rank(self):
        return 2

    @property
    def tensor_rank(self):
        return 2

    def _lead(self):
        nd = self._scalar.dimensions
        return (nd, nd)

    def get_dealiased(self, padding_factor=1.5, dealias_direct=False):
        key = (str(padding_factor), bool(dealias_direct))
        if key not in self._padded_space:
            self._padded_space[key] = TensorSpace(self._scalar.get_dealiased(padding_factor, dealias_direct))
        return self._padded_space[key]
