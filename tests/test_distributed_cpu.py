"""Host-side multi-rank logic on CPU: two gloo processes check the slab bookkeeping of
TensorProductSpace (local shapes / slices as mpi4py-fft's slab decomposition) and the
SlabExchange redistribution against a numpy model of the global array."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import shenfun_b200 as sf
        from shenfun_b200.transfer import SlabExchange
        comm = sf.Comm()
        assert comm.Get_size() == world and comm.Get_rank() == rank
        # --- bookkeeping of a slab-distributed Fourier space (uneven split on purpose)
        N = (10, 9, 8)
        T = sf.TensorProductSpace(comm, [sf.FunctionSpace(N[0], 'F', dtype='D'), sf.FunctionSpace(N[1], 'F', dtype='D'),
                                         sf.FunctionSpace(N[2], 'F', dtype='d')])
        assert T.global_shape(False) == (10, 9, 8) and T.global_shape(True) == (10, 9, 5)
        n0 = [5, 5][rank]
        n1 = [5, 4][rank]
        assert T.shape(False) == (n0, 9, 8), T.shape(False)
        assert T.shape(True) == (10, n1, 5), T.shape(True)
        assert T.local_slice(False)[0] == slice(rank*5, rank*5+5)
        assert T.local_slice(True)[1] == (slice(0, 5) if rank == 0 else slice(5, 9))
        Tp = T.get_dealiased((1.5, 1.5, 1.5))
        assert Tp.shape(True) == T.shape(True)
        assert Tp.global_shape(False) == (15, 13, 12)
        # --- exchange against a numpy model of the global array
        rng = np.random.default_rng(5)
        G = rng.standard_normal((10, 9, 5))+1j*rng.standard_normal((10, 9, 5))
        ex = SlabExchange(comm, (10, 9, 5), gather_axis=0, scatter_axis=1)
        src = torch.from_numpy(np.ascontiguousarray(G[T.local_slice(False)[0]]))
        dst = torch.empty((10, n1, 5), dtype=torch.complex128)
        ex(src, dst, 1, 'forward')
        assert np.array_equal(dst.numpy(), G[:, T.local_slice(True)[1]])
        back = torch.empty_like(src)
        ex(dst, back, 1, 'backward')
        assert np.array_equal(back.numpy(), src.numpy())
        # vector valued (leading component axis)
        GV = rng.standard_normal((3, 10, 9, 5))
        srcv = torch.from_numpy(np.ascontiguousarray(GV[:, T.local_slice(False)[0]]))
        dstv = torch.empty((3, 10, n1, 5), dtype=torch.float64)
        ex(srcv, dstv, 3, 'forward')
        assert np.array_equal(dstv.numpy(), GV[:, :, T.local_slice(True)[1]])
        # --- the pencil facade of the mpi4py-fft surface (shenfun_b200.pencil) over the same exchange
        from shenfun_b200.pencil import Subcomm, Pencil, DistArray
        sc = Subcomm(comm, [0, 1, 1])
        assert sc.dims == (2, 1, 1)
        pA = Pencil(sc, (10, 9, 5), axis=2)
        assert pA.subshape == (5, 9, 5) and pA.substart == (rank*5, 0, 0)
        pB = pA.pencil(0)                                  # whole along 0, the ranks move to axis 2? no: to the
        assert pB.subcomm.dims == (1, 1, 2)                # axis that was aligned (mpi4py-fft swaps the two)
        pC = Pencil(Subcomm(comm, [1, 0, 1]), (10, 9, 5), axis=0)
        tr = Pencil(sc, (10, 9, 5), axis=1).transfer(pC, np.complex128)
        a = torch.from_numpy(np.ascontiguousarray(G[rank*5:rank*5+5]))
        b = torch.empty(pC.subshape, dtype=torch.complex128)
        tr.forward(a, b)
        assert np.array_equal(b.numpy(), G[:, pC.substart[1]:pC.substart[1]+pC.subshape[1]])
        a2 = torch.empty_like(a)
        tr.backward(b, a2)
        assert np.array_equal(a2.numpy(), a.numpy())
        da = DistArray((10, 9, 5), subcomm=sc, dtype=np.complex128, alignment=1, buffer=a.clone(), device='cpu')
        assert da.alignment == 1 and da.commsizes == [2, 1, 1] and da.local_slice()[0] == slice(rank*5, rank*5+5)
        db = da.redistribute(0)
        assert db.alignment == 0 and np.array_equal(db.v.numpy(), G[:, db.local_slice()[1]])
        try:
            Subcomm(comm, [1, 1, 1])
            raise AssertionError('grid of one process on two ranks must be rejected')
        except ValueError:
            pass
        try:
            sf.TensorProductSpace(comm, [sf.FunctionSpace(8, 'F', dtype='D'), sf.FunctionSpace(8, 'F', dtype='D'),
                                         sf.FunctionSpace(8, 'F', dtype='d')], slab=False)
            raise AssertionError('pencil request must raise')
        except NotImplementedError:
            pass
        # --- checkpoint: two ranks write their blocks of one global array; each reads its own back,
        # rank 0 also checks the whole file against the numpy model
        import shutil
        import tempfile
        base = os.path.join(tempfile.gettempdir(), f'shb_chk_{port}')
        U = sf.Function(T, buffer=np.ascontiguousarray(G[T.local_slice(True)]))
        chk = sf.Checkpoint(base, checkevery=1, data={'0': {'U': [U]}})
        chk.update(0.25, 3)
        V = sf.Function(T)
        sf.Checkpoint(base).read(V, 'U')
        assert np.array_equal(V.get(), G[T.local_slice(True)])
        if rank == 0:
            assert np.array_equal(np.load(os.path.join(base+'.chk', 'U', '0.npy')), G)
        comm.barrier()
        if rank == 0:
            shutil.rmtree(base+'.chk', ignore_errors=True)
            shutil.rmtree(base+'.old.chk', ignore_errors=True)
        # refine on a distributed space: every rank ends with its block of the globally refined array
        from shenfun_b200.arrays import _refine_axis
        for Nn in ((16, 12, 12), (8, 6, 6)):
            R = U.refine(Nn)
            Tn = R.function_space()
            gt = torch.from_numpy(G.copy())
            for ax, (b, nn) in enumerate(zip(T.bases, Nn)):
                gt = _refine_axis(b, gt, ax, nn)
            assert tuple(gt.shape) == tuple(Tn.global_shape(True))
            assert np.array_equal(R.get(), gt.numpy()[Tn.local_slice(True)])
        # reductions through the communicator facade
        assert comm.allreduce(float(rank+1)) == 3.0
        assert comm.bcast({'a': rank}, root=0) == {'a': 0}
        results[rank] = 'ok'
    except Exception as e:  # noqa: BLE001
        import traceback
        results[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_two_rank_slab_logic_gloo():
    world = 2
    port = 29600+os.getpid() % 300
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    for r in range(world):
        assert results.get(r) == 'ok', results.get(r)
