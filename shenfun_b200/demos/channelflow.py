"""Turbulent channel flow (Kim-Moin-Moser formulation): demo/ChannelFlow.py `KMM` and
demo/MKM.py `MKM` on the device (BASELINE config 5).

Wall-normal direction x: Chebyshev (ShenBiharmonic for u, ShenDirichlet for g, v, w; orthogonal
for the curl components), streamwise y and spanwise z: Fourier.  Fourth-order equation for the
wall-normal velocity, second-order equation for the wall-normal vorticity g, continuity for
v, w; the (0, 0) Fourier mode of v, w is a 1-D problem.  Time stepping: IMEXRK
(shenfun/utilities/integrators.py:702-892), linear solves: chebyshev.la.Biharmonic / Helmholtz.
Same structure and operator calls as the reference classes (ChannelFlow.py:5-330,
MKM.py:6-118); file output, checkpointing and plotting are left out.
"""
import numpy as np
import torch

import shenfun_b200 as sf
from shenfun_b200 import la, ops
from shenfun_b200.forms import (TestFunction, Dx, div, grad, curl, Project, Expr, inner)
from shenfun_b200 import integrators
from shenfun_b200.utilities import cross


class KMM:
    def __init__(self, N=(32, 32, 32), domain=((-1, 1), (0, 2*np.pi), (0, np.pi)), nu=0.01, dt=0.1, conv=0,
                 dpdy=1, family='C', padding_factor=(1, 1.5, 1.5), moderror=100, timestepper='IMEXRK222',
                 comm=None):
        comm = sf.comm if comm is None else comm
        self.comm = comm
        self.N, self.nu, self.dt, self.conv = N, nu, dt, conv
        self.moderror = moderror
        self.padding_factor = padding_factor
        self.dpdy = dpdy
        self.PDE = PDE = getattr(integrators, timestepper)
        FS = sf.FunctionSpace
        self.B0 = FS(N[0], family, bc=(0, 0, 0, 0), domain=domain[0])
        self.D0 = FS(N[0], family, bc=(0, 0), domain=domain[0])
        self.C0 = FS(N[0], family, domain=domain[0])
        self.F1 = FS(N[1], 'F', dtype='D', domain=domain[1])
        self.F2 = FS(N[2], 'F', dtype='d', domain=domain[2])
        self.D00 = FS(N[0], family, bc=(0, 0), domain=domain[0])
        TPS = sf.TensorProductSpace
        self.TB = TPS(comm, (self.B0, self.F1, self.F2))
        self.TD = TPS(comm, (self.D0, self.F1, self.F2))
        self.TC = TPS(comm, (self.C0, self.F1, self.F2))
        self.BD = sf.VectorSpace([self.TB, self.TD, self.TD])
        self.CD = sf.VectorSpace(self.TD)
        self.CC = sf.VectorSpace([self.TD, self.TC, self.TC])
        self.TDp = self.TD.get_dealiased(padding_factor)
        # 1-D spaces of the (0, 0) Fourier mode live on every rank (tiny), used by the owner only
        self.T00 = TPS(sf.COMM_SELF, (self.D00,))
        self.C00 = TPS(sf.COMM_SELF, (self.D00.get_orthogonal(),))

        self.u_ = sf.Function(self.BD)
        self.H_ = sf.Function(self.CD)
        self.curl = sf.Function(self.CC)
        self.g_ = self.curl[0]
        self.ub = sf.Array(self.BD)
        self.v00 = sf.Function(self.T00)
        self.w00 = sf.Function(self.T00)
        self.mask = self.TB._device_mask()
        self.X = self.TD.local_mesh(bcast=True)
        self.K = self.TD.local_wavenumbers(scaled=True, device=True)
        sl = self.TD.local_slice(True)
        self.owns_00 = (sl[1].start == 0 and sl[2].start == 0)

        self.dudx = Project(Dx(self.u_[0], 0, 1), self.TD)
        if conv == 0:
            self.dudy = Project(Dx(self.u_[0], 1, 1), self.TB)
            self.dudz = Project(Dx(self.u_[0], 2, 1), self.TB)
            self.dvdx = Project(Dx(self.u_[1], 0, 1), self.TC)
            self.dvdy = Project(Dx(self.u_[1], 1, 1), self.TD)
            self.dvdz = Project(Dx(self.u_[1], 2, 1), self.TD)
            self.dwdx = Project(Dx(self.u_[2], 0, 1), self.TC)
            self.dwdy = Project(Dx(self.u_[2], 1, 1), self.TD)
            self.dwdz = Project(Dx(self.u_[2], 2, 1), self.TD)
        self.curly = Project(curl(self.u_)[1], self.TC, output_array=self.curl[1])
        self.curlz = Project(curl(self.u_)[2], self.TC, output_array=self.curl[2])
        self.divu = Project(div(self.u_), self.TC)

        v = TestFunction(self.TB)
        h = TestFunction(self.TD)
        cheb = self.B0.family() == 'chebyshev'
        sol1 = la.Biharmonic if cheb else la.SolverGeneric1ND
        sol2 = la.Helmholtz if cheb else la.SolverGeneric1ND
        H = self.H_
        self.pdes = {
            'u': PDE(v, div(grad(self.u_[0])), lambda f: self.nu*div(grad(f)),
                     Dx(Dx(H[1], 0, 1), 1, 1)+Dx(Dx(H[2], 0, 1), 2, 1)-Dx(H[0], 1, 2)-Dx(H[0], 2, 2),
                     dt=self.dt, solver=sol1),
            'g': PDE(h, self.g_, lambda f: self.nu*div(grad(f)), Dx(H[1], 2, 1)-Dx(H[2], 1, 1),
                     dt=self.dt, solver=sol2),
        }
        K2 = self.K[1]*self.K[1]+self.K[2]*self.K[2]
        K2s = torch.where(K2 == 0, torch.ones_like(K2), K2)
        self.K_over_K2 = [self.K[1]/K2s, self.K[2]/K2s]
        if self.owns_00:
            v0 = TestFunction(self.T00)
            self.h1 = sf.Function(self.T00)
            self.h2 = sf.Function(self.T00)
            source = sf.Array(self.C00, val=-self.dpdy)
            self.pdes1d = {
                'v0': PDE(v0, self.v00, lambda f: self.nu*div(grad(f)), [-Expr(self.h1), source], dt=self.dt,
                          solver=sol2),
                'w0': PDE(v0, self.w00, lambda f: self.nu*div(grad(f)), -Expr(self.h2), dt=self.dt, solver=sol2),
            }

    # -- nonlinear term -----------------------------------------------------------------------
    def convection(self):
        H = self.H_
        pf = self.padding_factor
        self.up = self.u_.backward(padding_factor=pf)
        up = self.up.v
        if self.conv == 0:
            g = []
            for name in ('dudx', 'dudy', 'dudz', 'dvdx', 'dvdy', 'dvdz', 'dwdx', 'dwdy', 'dwdz'):
                g.append(ops.plain(getattr(self, name)().backward(padding_factor=pf).v).contiguous())
            # u . grad u for the three components in one pass over the twelve physical arrays
            if getattr(self, '_conv_work', None) is None or self._conv_work.shape != up.shape:
                self._conv_work = torch.empty_like(ops.plain(up))
            cb = ops.dot_grad3d(ops.plain(up), g, self._conv_work)
            for i in range(3):
                self.TDp.forward(cb[i], H[i])
        else:
            self.curly()
            self.curlz()
            cp = self.curl.backward(padding_factor=pf)
            cb = torch.empty_like(up)
            cross(cb, cp.v, up)
            for i in range(3):
                self.TDp.forward(cb[i], H[i])
        ops.apply_mask(H.v, self.mask)

    def compute_vw(self, rk):
        u = self.u_.v
        if self.owns_00:
            self.v00.v.copy_(u[1, :, 0, 0].real)
            self.w00.v.copy_(u[2, :, 0, 0].real)
            self.h1.v.copy_(self.H_.v[1, :, 0, 0].real)
            self.h2.v.copy_(self.H_.v[2, :, 0, 0].real)
        f = self.dudx().v
        g = self.g_.v
        ops.kmm_vw(ops.plain(f), ops.plain(g), self.K_over_K2[0], self.K_over_K2[1], u[1], u[2])
        if self.owns_00:
            self.pdes1d['v0'].compute_rhs(rk)
            u[1, :, 0, 0] = self.pdes1d['v0'].solve_step(rk).v.to(u.dtype)
            self.pdes1d['w0'].compute_rhs(rk)
            u[2, :, 0, 0] = self.pdes1d['w0'].solve_step(rk).v.to(u.dtype)
        return u

    def energy_and_divergence(self):
        ub = self.u_.backward(self.ub)
        e = [inner(1, sf.Array(self.TD, buffer=(ub.v[i]*ub.v[i]).contiguous())) for i in range(3)]
        divu = self.divu().backward()
        e3 = np.sqrt(inner(1, sf.Array(self.TC, buffer=(divu.v*divu.v).contiguous())))
        return e, e3

    def update(self, t, tstep):
        pass

    def prepare_step(self, rk):
        self.convection()

    def assemble(self):
        for pde in self.pdes.values():
            pde.assemble()
        if self.owns_00:
            for pde in self.pdes1d.values():
                pde.assemble()
        self._assembled = True

    def solve(self, t=0, tstep=0, end_time=1000):
        if not getattr(self, '_assembled', False):
            self.assemble()
        while t < end_time-1e-8:
            for rk in range(self.PDE.steps()):
                self.prepare_step(rk)
                for eq in self.pdes.values():
                    eq.compute_rhs(rk)
                for eq in self.pdes.values():
                    eq.solve_step(rk)
                self.compute_vw(rk)
            t += self.dt
            tstep += 1
            self.update(t, tstep)
        return t, tstep


class MKM(KMM):
    """Re_tau = 180 channel (demo/MKM.py:6-118)."""

    def __init__(self, N=(32, 32, 32), domain=((-1, 1), (0, 2*np.pi), (0, np.pi)), Re=180., dt=0.1, conv=0,
                 family='C', padding_factor=(1, 1.5, 1.5), moderror=100, timestepper='IMEXRK222', comm=None):
        self.utau = 1
        KMM.__init__(self, N=N, domain=domain, nu=1./Re, dt=dt, conv=conv, family=family,
                     padding_factor=padding_factor, moderror=moderror, timestepper=timestepper,
                     dpdy=-self.utau**2, comm=comm)
        self.Re = Re
        self.flux = np.array([618.97])

    def initialize(self):
        X = self.X
        Y = np.where(X[0] < 0, 1+X[0], 1-X[0])
        utau = self.nu*self.Re
        Um = 46.9091*utau
        Xplus = Y*self.Re
        Yplus = X[1]*self.Re
        Zplus = X[2]*self.Re
        duplus = Um*0.2/utau
        alfaplus = self.F1.domain[1]/200.
        betaplus = self.F2.domain[1]/100.
        sigma = 0.00055
        epsilon = Um/200.
        U = sf.Array(self.BD)
        U[1] = Um*(Y-0.5*Y**2)
        dd = utau*duplus/2.0*Xplus/40.*np.exp(-sigma*Xplus**2+0.5)*np.cos(betaplus*Zplus)
        U[1] += ops.as_device_tensor(np.ascontiguousarray(np.broadcast_to(dd, U.shape[1:])))
        U[2] += ops.as_device_tensor(np.ascontiguousarray(np.broadcast_to(
            epsilon*np.sin(alfaplus*Yplus)*Xplus*np.exp(-sigma*Xplus**2), U.shape[1:])))
        u_ = U.forward(self.u_)
        ops.apply_mask(u_.v, self.mask)
        U = u_.backward(U)
        u_ = U.forward(self.u_)
        self.g_.v.copy_(1j*self.K[1]*u_.v[2]-1j*self.K[2]*u_.v[1])
        return 0, 0

    def update(self, t, tstep):
        # dynamically adjust the flux (MKM.py:104-113)
        ub1 = self.u_[1].backward(self.ub[1])
        beta = inner(1, ub1)
        q = self.flux[0]-beta
        V = self.TD.volume
        if self.owns_00:
            u = self.u_.v
            u[1, :, 0, 0] *= (1+q/V/u[1, 0, 0, 0])
