"""Oracle: one RK4 time step of the triply periodic Navier-Stokes demo, in numpy.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/ and by the
`cpu_baseline` leg of bench.py, never by the product path.

Restates demo/NavierStokes.py:31-44 (`LinearRHS`, `NonlinearRHS`) and the RK4 loop of
shenfun/utilities/integrators.py:587-599 on top of oracle.tensorproduct (scipy.fft / pocketfft):
  curl_hat = i k x U_hat                       (project(curl(U_hat), TV): Fourier mass = identity)
  W        = backward_pad(U_hat) x backward_pad(curl_hat)   (3/2-rule padded physical space)
  dU       = forward_pad(W)
  P_hat    = (i k . dU) / |k|^2 with the (0,0,0) mode 0     (A = inner(grad u, grad v) = |k|^2, :29-30,41)
  dU      += -i k P_hat ... i.e. inner(grad(P_hat), v) = i k P_hat           (:42)
  dU      += nu * (-|k|^2) U_hat                                            (:31-32, integrators.py:593)
Pinned against tests/golden/ns_16.npz (the reference's own run) by tests/test_oracle_fourier.py.
"""
import numpy as np
from . import tensorproduct as otp


class NavierStokes:
    def __init__(self, N, nu=0.000625, dt=0.01, padding_factor=1.5):
        self.N = tuple(N)
        self.nu, self.dt = nu, dt
        pf = padding_factor
        self.d = [('C2C', N[0], N[0]), ('C2C', N[1], N[1]), ('R2C', N[2], N[2])]
        self.dp = [('C2C', N[0], int(np.floor(pf*N[0]))), ('C2C', N[1], int(np.floor(pf*N[1]))),
                   ('R2C', N[2], int(np.floor(pf*N[2])))]
        k0 = np.fft.fftfreq(N[0], 1./N[0])[:, None, None]
        k1 = np.fft.fftfreq(N[1], 1./N[1])[None, :, None]
        k2 = np.fft.rfftfreq(N[2], 1./N[2])[None, None, :]
        self.K = [k0, k1, k2]
        self.K2 = k0**2+k1**2+k2**2
        with np.errstate(divide='ignore'):
            self.iK2 = np.where(self.K2 == 0, 0.0, 1.0/np.where(self.K2 == 0, 1.0, self.K2))
        self.a = [1./6., 1./3., 1./3., 1./6.]
        self.b = [0.5, 0.5, 1.0]

    def initialize(self):
        """Taylor-Green field (demo/NavierStokes.py:50-53)."""
        N = self.N
        X = np.meshgrid(*[2*np.pi*np.arange(n)/n for n in N], indexing='ij')
        U = np.zeros((3,)+N)
        U[0] = np.sin(X[0])*np.cos(X[1])*np.cos(X[2])
        U[1] = -np.cos(X[0])*np.sin(X[1])*np.cos(X[2])
        self.U_hat = otp.forward(U, self.d, lead=1)
        return self

    def nonlinear(self, U_hat):
        K = self.K
        curl_hat = np.empty_like(U_hat)
        curl_hat[0] = 1j*(K[1]*U_hat[2]-K[2]*U_hat[1])
        curl_hat[1] = 1j*(K[2]*U_hat[0]-K[0]*U_hat[2])
        curl_hat[2] = 1j*(K[0]*U_hat[1]-K[1]*U_hat[0])
        curl_ = otp.backward(curl_hat, self.dp, lead=1)
        U = otp.backward(U_hat, self.dp, lead=1)
        W = np.cross(U, curl_, axis=0)
        dU = otp.forward(W, self.dp, lead=1)
        # pressure: A P = inner(div(dU), v) with A = inner(grad u, grad v) = |k|^2
        P_hat = 1j*(K[0]*dU[0]+K[1]*dU[1]+K[2]*dU[2])*self.iK2
        for i in range(3):
            dU[i] += 1j*K[i]*P_hat
        return dU

    def step(self):
        """integrators.py:587-599"""
        dt, nu = self.dt, self.nu
        U0 = self.U_hat.copy()
        U1 = self.U_hat.copy()
        u = self.U_hat
        for rk in range(4):
            dU = self.nonlinear(u)
            dU += -nu*self.K2*u
            if rk < 3:
                u = U0+self.b[rk]*dt*dU
            U1 += self.a[rk]*dt*dU
        self.U_hat = U1
        return self.U_hat
