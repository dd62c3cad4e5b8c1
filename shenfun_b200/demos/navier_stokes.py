"""Triply periodic Navier-Stokes (Taylor-Green), demo/NavierStokes.py on the device.

Two drivers with identical numerics:
  * `solve_reference_style` -- the demo's own structure (form language, `project`,
    `inner`, `A.solve`, generic `RK4`), every array on the GPU;
  * `FusedNS` -- the same algebra with the spectral point-wise work of one RK stage
    folded into two kernels (`shb_curl_hat`, `shb_ns_rk_stage`) and the 3/2-rule
    transforms batched over the three components.
"""
import numpy as np
import torch

import shenfun_b200 as sf
from shenfun_b200 import ops
from shenfun_b200.forms import (TrialFunction, TestFunction, inner, project, grad, div, curl)
from shenfun_b200.matrices import get_simplified_tpmatrices
from shenfun_b200.integrators import RK4
from shenfun_b200.utilities import cross


def make_spaces(N, comm=None):
    comm = sf.comm if comm is None else comm
    V0 = sf.FunctionSpace(N[0], 'F', dtype='D')
    V1 = sf.FunctionSpace(N[1], 'F', dtype='D')
    V2 = sf.FunctionSpace(N[2], 'F', dtype='d')
    T = sf.TensorProductSpace(comm, (V0, V1, V2))
    return T, sf.VectorSpace(T)


def taylor_green(TV, U):
    """demo/NavierStokes.py:50-52"""
    X = TV.local_mesh(True, device=True)
    U[0] = torch.sin(X[0])*torch.cos(X[1])*torch.cos(X[2])
    U[1] = -torch.cos(X[0])*torch.sin(X[1])*torch.cos(X[2])
    U[2] = 0
    return U


def solve_reference_style(N=(32, 32, 32), nu=0.000625, dt=0.01, end_time=0.1, padding_factor=1.0):
    """demo/NavierStokes.py:13-61 with RK4; returns (U_hat, U, kinetic energy)."""
    T, TV = make_spaces(N)
    pad = abs(padding_factor-1) > 1e-8
    TVp = TV.get_dealiased((padding_factor,)*3) if pad else TV
    u = TrialFunction(T)
    v = TestFunction(T)
    U = sf.Array(TV)
    Up = sf.Array(TVp)
    U_hat = sf.Function(TV)
    P_hat = sf.Function(T)
    curl_hat = sf.Function(TV)
    W = sf.Array(TVp)
    curl_ = sf.Array(TVp)
    A = inner(grad(u), grad(v))
    A = get_simplified_tpmatrices(A)[0]
    vT = TestFunction(TV)

    def LinearRHS(self, u, **params):
        return nu*div(grad(u))

    def NonlinearRHS(self, U, U_hat, dU, **params):
        project(curl(U_hat), TV, output_array=curl_hat)
        TVp.backward(curl_hat, curl_)
        TVp.backward(U_hat, Up)
        cross(W, Up, curl_)
        TVp.forward(W, dU)
        A.solve(inner(div(dU), v), P_hat)
        dU += inner(grad(P_hat), vT)
        return dU

    taylor_green(TV, U)
    U_hat = TV.forward(U, U_hat)
    integ = RK4(TV, L=LinearRHS, N=NonlinearRHS)
    U_hat = integ.solve(U, U_hat, dt, (0, end_time))
    U = U_hat.backward(U)
    k = 0.5*(U.v*U.v).sum().item()/np.prod(np.array(N))
    k = T.comm.allreduce(k)
    return U_hat, U, k


class FusedNS:
    """Same time stepping with fused spectral kernels and batched padded transforms."""

    def __init__(self, N=(32, 32, 32), nu=0.000625, dt=0.01, padding_factor=1.5, comm=None, pressure=False):
        self.pressure = pressure     # also return P_hat every stage (the demo keeps it; the update does not need it)
        self.N, self.nu, self.dt = tuple(N), nu, dt
        self.T, self.TV = make_spaces(N, comm)
        pad = abs(padding_factor-1) > 1e-8
        self.TVp = self.TV.get_dealiased((padding_factor,)*3) if pad else self.TV
        self.Tp = self.TVp.flatten()[0]
        T = self.T
        self.K = [k.reshape(-1).contiguous() for k in T.local_wavenumbers(scaled=True, device=True)]
        self.U = sf.Array(self.TV)
        # [U_hat; curl_hat] stacked: the nonlinear term transforms both in one batched backward
        self._uw_hat = torch.zeros((6,)+tuple(self.TV.shape(True))[1:], dtype=torch.complex128,
                                   device=self.K[0].device)
        self.U_hat = sf.Function(self.TV, buffer=self._uw_hat[:3])
        self.U_hat0 = sf.Function(self.TV)
        self.U_hat1 = sf.Function(self.TV)
        self.dU = sf.Function(self.TV)
        self.curl_hat = sf.Function(self.TV, buffer=self._uw_hat[3:])
        self.P_hat = sf.Function(T)
        self.Up = sf.Array(self.TVp)
        self.a = [1./6., 1./3., 1./3., 1./6.]
        self.b = [0.5, 0.5, 1.0, 0.0]
        self.t = 0.0
        self.tstep = 0
        self._curl_ready = False
        # opt-in (SHB_NS_CURL_FUSION=1; one rank, 3/2 rule, N/2 in {64 .. 512}): i k x u_hat is formed inside the
        # first padded pass and the vorticity spectrum is never stored.  Not faster on B200 (see
        # TensorProductSpace.curl_fusion_available), so the default keeps the stored spectrum.
        self._curl_fused = pad and hasattr(self.Tp, 'curl_fusion_available') and self.Tp.curl_fusion_available()

    def initialize(self):
        taylor_green(self.TV, self.U)
        self.TV.forward(self.U, self.U_hat)
        self._curl_ready = False
        return self

    def step(self):
        dt, nu = self.dt, self.nu
        uh = self.U_hat.v
        U0, U1, dU = self.U_hat0.v, self.U_hat1.v, self.dU.v
        U0.copy_(uh)
        U1.copy_(uh)
        k0, k1, k2 = self.K
        for rk in range(4):
            if self._curl_fused:
                self.Tp.backward_cross_hat(uh, self.Up, curl_of=(k0, k1, k2))
                self.TVp.forward(self.Up, self.dU)
                ops.ns_rk_stage(dU, uh, U0, U1, self.P_hat.v if self.pressure else None, k0, k1, k2, nu,
                                self.a[rk]*dt, self.b[rk]*dt, rk == 3, keep_dU=False, curl_out=None)
                continue
            if not self._curl_ready:
                ops.curl_hat(uh, self.curl_hat.v, k0, k1, k2)
            # W = backward(U_hat) x backward(curl_hat): one batched backward over the six spectral
            # components, the last axis pass fused with the cross product (no physical U / curl)
            self.Tp.backward_cross_hat(self._uw_hat, self.Up)
            self.TVp.forward(self.Up, self.dU)
            ops.ns_rk_stage(dU, uh, U0, U1, self.P_hat.v if self.pressure else None, k0, k1, k2, nu,
                            self.a[rk]*dt, self.b[rk]*dt, rk == 3, keep_dU=False, curl_out=self.curl_hat.v)
            self._curl_ready = True      # the stage kernel left i k x u_hat of the updated velocity
        self.t += dt
        self.tstep += 1

    def energy(self):
        U = self.U_hat.backward(self.U)
        k = 0.5*(U.v*U.v).sum().item()/np.prod(np.array(self.N))
        return self.T.comm.allreduce(k)


def bench(n=512, steps=3, warmup=1, padding_factor=1.5, comm=None):
    """Taylor-Green NS steps/s at n^3 with 3/2-rule dealiasing, RK4 (BASELINE config 2)."""
    ns = FusedNS((n, n, n), nu=0.000625, dt=0.001, padding_factor=padding_factor, comm=comm).initialize()
    for _ in range(warmup):
        ns.step()
    if comm is not None:
        comm.barrier()
    torch.cuda.synchronize()
    l0 = ops.LAUNCHES
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ns.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/steps
    if comm is not None and comm.Get_size() > 1:
        ms = comm.allreduce(ms, op='max')
    k = ns.energy()
    out = {'metric': 'Taylor-Green NS timesteps/s', 'value': 1e3/ms, 'unit': 'steps/s', 'ms_per_step': ms,
           'workload': f'{n}^3 triply periodic, 3/2-rule padding ({int(n*padding_factor)}^3 physical), RK4, '
                       f'nu=0.000625 dt=0.001', 'kinetic_energy': k,
           'gpu_launches_per_step': (ops.LAUNCHES-l0)//steps}
    del ns
    torch.cuda.empty_cache()
    return out
