"""Generate golden vectors by running the REFERENCE's own Python sources.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference imports on top of oracle/refshim (serial stand-ins for the absent
mpi4py / mpi4py-fft, restated on scipy.fft) with
config['optimization']['mode'] = 'numba' (the reference's own numba kernels).
Outputs: tests/golden/*.npz (inputs + reference outputs, fp64), small enough to
commit.  The GPU box never reads /root/reference; tests compare the device
path and the numpy oracle against these files.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'refshim'))
sys.path.insert(0, '/root/reference')
os.chdir(HERE)

import numpy as np  # noqa: E402
import shenfun as sf  # noqa: E402

sf.config['optimization']['mode'] = 'numba'
rng = np.random.default_rng(20261017)


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **kw)
    print('wrote', name, {k: getattr(v, 'shape', v) for k, v in kw.items()})


def fourier_nd():
    """TensorProductSpace forward/backward, plain and 3/2-padded, incl. odd sizes."""
    cases = {
        'f3d_14_15_16': ((14, 15, 16), 'DDd'),
        'f3d_16_16_16': ((16, 16, 16), 'DDd'),
        'f3d_12_9_10': ((12, 9, 10), 'DDd'),
        'f2d_16_12': ((16, 12), 'Dd'),
        'f2d_c_10_12': ((10, 12), 'DD'),
    }
    for name, (N, dts) in cases.items():
        bases = [sf.FunctionSpace(n, 'F', dtype=dt) for n, dt in zip(N, dts)]
        T = sf.TensorProductSpace(sf.comm, bases)
        u = rng.standard_normal(T.shape(False))
        if dts[-1] == 'D':
            u = u + 1j*rng.standard_normal(T.shape(False))
        ua = sf.Array(T)
        ua[:] = u
        uh = np.array(T.forward(ua, sf.Function(T)))
        # random spectral data for backward (not necessarily Hermitian-consistent:
        # exercises the c2r conventions for DC/Nyquist imaginary parts)
        vh = rng.standard_normal(T.shape(True)) + 1j*rng.standard_normal(T.shape(True))
        vf = sf.Function(T)
        vf[:] = vh
        vb = np.array(T.backward(vf, sf.Array(T)))
        out = dict(N=np.array(N), dtypes=dts, u=u, u_hat=uh, v_hat=vh, v=vb)
        for pf in (1.5, 2.0):
            Tp = T.get_dealiased((pf,)*len(N))
            vp = np.array(Tp.backward(vf, sf.Array(Tp)))
            w = rng.standard_normal(Tp.shape(False))
            if dts[-1] == 'D':
                w = w + 1j*rng.standard_normal(Tp.shape(False))
            wa = sf.Array(Tp)
            wa[:] = w
            wh = np.array(Tp.forward(wa, sf.Function(Tp)))
            tag = str(pf).replace('.', '')
            out.update({f'v_pad{tag}': vp, f'w_pad{tag}': w, f'w_hat_pad{tag}': wh})
        save(name, **out)


def navier_stokes():
    """demo/NavierStokes.py (RK4) at 16^3: U_hat after 1 and 5 steps, plain and 3/2-padded,
    and the 32^3 kinetic energy check of the demo (k = 0.124953117517 at t=0.1)."""
    from shenfun import (FunctionSpace, TensorProductSpace, VectorSpace, TrialFunction, TestFunction,
                         Array, Function, inner, grad, div, curl, project, get_simplified_tpmatrices, RK4, comm)

    def run(N, nsteps, pf, nu=0.000625, dt=0.01):
        V0 = FunctionSpace(N[0], 'F', dtype='D')
        V1 = FunctionSpace(N[1], 'F', dtype='D')
        V2 = FunctionSpace(N[2], 'F', dtype='d')
        T = TensorProductSpace(comm, (V0, V1, V2))
        TV = VectorSpace(T)
        TVp = TV.get_dealiased((pf,)*3) if pf > 1 else TV
        u = TrialFunction(T)
        v = TestFunction(T)
        U = Array(TV)
        Up = Array(TVp)
        U_hat = Function(TV)
        P_hat = Function(T)
        curl_hat = Function(TV)
        W = Array(TVp)
        curl_ = Array(TVp)
        A = inner(grad(u), grad(v))
        A = get_simplified_tpmatrices(A)[0]

        def LinearRHS(self, u, **params):
            return nu*div(grad(u))

        def NonlinearRHS(self, U, U_hat, dU, **params):
            nonlocal curl_hat, curl_, P_hat, W, Up
            curl_hat = project(curl(U_hat), TV, output_array=curl_hat)
            curl_ = TVp.backward(curl_hat, curl_)
            Up = TVp.backward(U_hat, Up)
            W[:] = np.cross(Up, curl_, axis=0)
            dU = TVp.forward(W, dU)
            P_hat = A.solve(inner(div(dU), v), P_hat)
            dU += inner(grad(P_hat), TestFunction(TV))
            return dU

        X = T.local_mesh(True)
        U[0] = np.sin(X[0])*np.cos(X[1])*np.cos(X[2])
        U[1] = -np.cos(X[0])*np.sin(X[1])*np.cos(X[2])
        U[2] = 0
        U_hat = TV.forward(U, U_hat)
        integ = RK4(TV, L=LinearRHS, N=NonlinearRHS)
        U_hat = integ.solve(U, U_hat, dt, (0, nsteps*dt))
        U = U_hat.backward(U)
        k = 0.5*np.sum(U*U)/np.prod(np.array(N))
        return np.array(U_hat), float(k)

    out = {}
    N = (16, 16, 16)
    for pf, tag in ((1.0, 'plain'), (1.5, 'pad15')):
        for ns in (1, 5):
            uh, k = run(N, ns, pf)
            out[f'U_hat_{tag}_{ns}'] = uh
            out[f'k_{tag}_{ns}'] = k
    uh, k = run((32, 32, 32), 10, 1.0)
    assert round(k-0.124953117517, 7) == 0, k
    out['k32_t01'] = k
    save('ns_16', **out)


if __name__ == '__main__':
    which = sys.argv[1:] or None
    fns = {k: v for k, v in globals().items() if callable(v) and getattr(v, '__module__', '') == '__main__'
           and k not in ('save',)}
    for k, f in fns.items():
        if which is None or k in which:
            f()
