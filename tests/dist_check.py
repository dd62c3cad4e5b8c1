"""Run under torchrun (one rank per GPU): slab-distributed transforms and the NS step must
reproduce the single-GPU results on identical global inputs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import shenfun_b200 as sf  # noqa: E402
from oracle import tensorproduct as otp  # noqa: E402


def rel(a, b):
    return np.linalg.norm((a-b).ravel())/np.linalg.norm(b.ravel())


def main():
    comm = sf.init_from_env('nccl')
    rank, world = comm.Get_rank(), comm.Get_size()
    worst = 0.0
    # the last three cases put BASELINE line lengths on both exchanged axes: the 512-point producer is
    # the 16-column peer-store kernel (`fft_strided_kernel<Sched<512,...>, ., 16, true>`), the 768-point
    # ones the 3/2-padded passes of the NS step, 1024 x 256 the channel-flow planes
    cases = [((32, 24, 20), 1.0), ((32, 24, 20), 1.5), ((30, 27, 16), 1.0), ((64, 64, 64), 1.5),
             ((512, 512, 16), 1.0), ((512, 512, 16), 1.5), ((1024, 256, 16), 1.0)]
    for N, pf in cases:
        bases = [sf.FunctionSpace(N[0], 'F', dtype='D'), sf.FunctionSpace(N[1], 'F', dtype='D'),
                 sf.FunctionSpace(N[2], 'F', dtype='d')]
        T = sf.TensorProductSpace(comm, bases)
        Tp = T.get_dealiased((pf,)*3)
        d = [('C2C', N[0], int(pf*N[0])), ('C2C', N[1], int(pf*N[1])), ('R2C', N[2], int(pf*N[2]))]
        rng = np.random.default_rng(17)
        ug = rng.standard_normal(Tp.global_shape(False))
        ref_hat = otp.forward(ug, d)
        u = sf.Array(Tp, buffer=np.ascontiguousarray(ug[Tp.local_slice(False)]))
        uh = Tp.forward(u)
        e1 = rel(uh.get(), ref_hat[T.local_slice(True)])
        vg = rng.standard_normal(T.global_shape(True))+1j*rng.standard_normal(T.global_shape(True))
        ref_v = otp.backward(vg, d)
        v = Tp.backward(sf.Function(T, buffer=np.ascontiguousarray(vg[T.local_slice(True)])))
        e2 = rel(v.get(), ref_v[Tp.local_slice(False)])
        # batched vector transform
        TV = sf.VectorSpace(Tp)
        Ug = rng.standard_normal((3,)+tuple(Tp.global_shape(False)))
        Uh = TV.forward(sf.Array(TV, buffer=np.ascontiguousarray(Ug[(slice(None),)+Tp.local_slice(False)])))
        e3 = rel(Uh.get(), otp.forward(Ug, d, lead=1)[(slice(None),)+T.local_slice(True)])
        worst = max(worst, e1, e2, e3)
        if max(e1, e2, e3) > 1e-12:
            print(f'[rank {rank}] N={N} pf={pf}: forward {e1:.2e} backward {e2:.2e} vector {e3:.2e}', flush=True)
        if rank == 0:
            print(f'N={N} pf={pf}: forward {e1:.2e} backward {e2:.2e} vector {e3:.2e}', flush=True)
    # Chebyshev x Fourier^2 Poisson solve distributed == oracle-free check: solve o matvec
    from shenfun_b200 import la
    SD = sf.FunctionSpace(32, 'C', bc=(0, 0))
    T = sf.TensorProductSpace(comm, (SD, sf.FunctionSpace(24, 'F', dtype='D'), sf.FunctionSpace(20, 'F', dtype='d')))
    uu, vv = sf.TrialFunction(T), sf.TestFunction(T)
    H = la.Helmholtz(sf.inner(vv, sf.div(sf.grad(uu))-1.5*uu))
    f = sf.Function(T)
    torch.manual_seed(3+rank)
    torch.view_as_real(f.v).normal_()
    f.v[SD.dim():] = 0
    x = H(f, sf.Function(T))
    c = H.matvec(x, sf.Function(T))
    e4 = (torch.linalg.vector_norm(c.v[:SD.dim()]-f.v[:SD.dim()])/torch.linalg.vector_norm(f.v[:SD.dim()])).item()
    # NS: distributed step == single-GPU golden
    from shenfun_b200.demos import navier_stokes as ns
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ns_16.npz'))
    s = ns.FusedNS((16, 16, 16), dt=0.01, padding_factor=1.5, comm=comm).initialize()
    for _ in range(5):
        s.step()
    # global relative error (a rank may own only wavenumbers the Taylor-Green field never excites)
    loc = g['U_hat_pad15_5'][(slice(None),)+s.T.local_slice(True)]
    en = torch.tensor([np.linalg.norm((s.U_hat.get()-loc).ravel())**2, np.linalg.norm(loc.ravel())**2],
                      dtype=torch.float64, device='cuda')
    dist.all_reduce(en)
    e5 = float(torch.sqrt(en[0]/en[1]))
    # channel flow (MKM), 3 IMEXRK222 steps == reference golden (global relative error)
    from shenfun_b200.demos.channelflow import MKM
    gm = np.load(os.path.join(ROOT, 'tests', 'golden', 'mkm_32_16_16.npz'))
    c = MKM(N=(32, 16, 16), Re=180., dt=0.001, conv=0, padding_factor=(1.5, 1.5, 1.5), timestepper='IMEXRK222',
            comm=comm)
    t, ts = c.initialize()
    c.solve(t=t, tstep=ts, end_time=0.003)
    loc = gm['u3_conv0'][(slice(None),)+c.TD.local_slice(True)]
    en = torch.tensor([np.linalg.norm((c.u_.get()-loc).ravel())**2, np.linalg.norm(loc.ravel())**2],
                      dtype=torch.float64, device='cuda')
    dist.all_reduce(en)
    e6 = float(torch.sqrt(en[0]/en[1]))
    if rank == 0:
        print(f'MKM 3 steps vs reference {e6:.2e}', flush=True)
    worst = max(worst, e4, e5, e6)
    if max(e4, e5) > 1e-12:
        print(f'[rank {rank}] helmholtz {e4:.2e} NS {e5:.2e}', flush=True)
    w = torch.tensor([worst], dtype=torch.float64, device='cuda')
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f'helmholtz solve/matvec {e4:.2e}  NS 5 steps {e5:.2e}  WORST {w.item():.3e} world={world}', flush=True)
    dist.destroy_process_group()
    sys.exit(0 if w.item() < 1e-12 else 1)


if __name__ == '__main__':
    main()
