"""Generate golden vectors by running the REFERENCE's own Python sources.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference imports on top of oracle/refshim (serial stand-ins for the absent
mpi4py / mpi4py-fft, restated on scipy.fft; the reference's own fastgl.cpp compiled
into oracle/_ref by oracle/build_ref.sh) with config['optimization']['mode'] = 'numba'
(the reference's own numba kernels).
Outputs: tests/golden/*.npz (inputs + reference outputs, fp64), small enough to
commit.  The GPU box never reads /root/reference; tests compare the device
path and the numpy oracle against these files.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
os.chdir(HERE)

import numpy as np  # noqa: E402
import refshim  # noqa: E402

sf = refshim.activate()       # reference's shenfun on the serial shims, numba kernels
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


BCS = {
    'ortho': None,
    'dirichlet': (0, 0),
    'neumann': {'left': {'N': 0}, 'right': {'N': 0}},
    'biharmonic': (0, 0, 0, 0),
    'dn': {'left': {'D': 0}, 'right': {'N': 0}},
    'inhom': {'left': {'D': 1.5}, 'right': {'D': -0.5}},
}


def poly_1d():
    """1-D Chebyshev / Legendre bases: matrices, scalar product, forward, backward (plain, padded)."""
    for fam in ('C', 'L'):
        out = {}
        for name, bc in BCS.items():
            for N in (16, 15):
                S = sf.FunctionSpace(N, fam, bc=bc) if bc is not None else sf.FunctionSpace(N, fam)
                tag = f'{name}_{N}'
                u = rng.standard_normal(N)
                out[f'u_{tag}'] = u
                out[f'sp_{tag}'] = np.array(S.scalar_product(u.copy()))
                uh = np.array(S.forward(u.copy()))
                out[f'fw_{tag}'] = uh
                vh = np.zeros(N)
                vh[:S.dim()] = rng.standard_normal(S.dim())
                if name == 'inhom':
                    vh[S.dim():] = uh[S.dim():]
                out[f'vh_{tag}'] = vh
                out[f'bw_{tag}'] = np.array(S.backward(vh.copy()))
                Sp = S.get_dealiased(1.5)
                out[f'bwpad_{tag}'] = np.array(Sp.backward(vh.copy()))
                w = rng.standard_normal(Sp.shape(False))
                out[f'w_{tag}'] = w
                out[f'fwpad_{tag}'] = np.array(Sp.forward(w.copy()))
                if N == 16:
                    K = S.stencil_matrix() if bc is not None else None
                    if K is not None:
                        out[f'K_{tag}'] = K.diags().toarray()
                    for (a, b) in ((0, 0), (0, 1), (0, 2), (1, 1), (0, 4), (2, 2)):
                        try:
                            M = sf.SpectralMatrix((S, a), (S, b))
                            out[f'M{a}{b}_{tag}'] = M.diags().toarray()
                        except Exception as e:  # noqa: BLE001
                            print('skip matrix', fam, name, a, b, repr(e)[:80])
        save('poly1d_'+fam, **out)


def poly_3d():
    """Chebyshev/Legendre x Fourier^2 tensor spaces, two axis orders, plain and 3/2-padded."""
    out = {}
    for fam in ('C', 'L'):
        for name in ('dirichlet', 'biharmonic', 'neumann'):
            bc = BCS[name]
            for layout in ('first', 'middle'):
                SD = sf.FunctionSpace(14, fam, bc=bc)
                K1 = sf.FunctionSpace(12, 'F', dtype='D')
                K2 = sf.FunctionSpace(10, 'F', dtype='d')
                if layout == 'first':
                    T = sf.TensorProductSpace(sf.comm, (SD, K1, K2))
                else:
                    T = sf.TensorProductSpace(sf.comm, (K1, SD, K2), axes=(1, 0, 2))
                tag = f'{fam}_{name}_{layout}'
                u = rng.standard_normal(T.shape(False))
                ua = sf.Array(T)
                ua[:] = u
                out[f'u_{tag}'] = u
                out[f'fw_{tag}'] = np.array(T.forward(ua, sf.Function(T)))
                out[f'sp_{tag}'] = np.array(T.scalar_product(ua, sf.Function(T)))
                vh = sf.Function(T)
                vh[:] = rng.standard_normal(T.shape(True))+1j*rng.standard_normal(T.shape(True))
                sl = [slice(None)]*3
                ax = 0 if layout == 'first' else 1
                sl[ax] = slice(SD.dim(), None)
                vh[tuple(sl)] = 0
                out[f'vh_{tag}'] = np.array(vh)
                out[f'bw_{tag}'] = np.array(T.backward(vh, sf.Array(T)))
                Tp = T.get_dealiased((1.5, 1.5, 1.5))
                out[f'bwpad_{tag}'] = np.array(Tp.backward(vh, sf.Array(Tp)))
                w = rng.standard_normal(Tp.shape(False))
                wa = sf.Array(Tp)
                wa[:] = w
                out[f'w_{tag}'] = w
                out[f'fwpad_{tag}'] = np.array(Tp.forward(wa, sf.Function(Tp)))
    save('poly3d', **out)


def solvers():
    """la.TDMA/PDMA/... along axes; SolverGeneric1ND Poisson / biharmonic; Chebyshev Helmholtz."""
    from shenfun import la, inner, div, grad, TestFunction, TrialFunction, Dx
    from shenfun.chebyshev.la import Helmholtz
    out = {}
    # --- named banded solvers on 1-D matrices along the axes of a 3-D array (tests/test_la.py:20-49)
    N = 16
    cases = {
        'TDMA': (sf.FunctionSpace(N, 'C', bc=(0, 0)), (0, 0), la.TDMA),
        'PDMA': (sf.FunctionSpace(N, 'C', bc=(0, 0, 0, 0)), (0, 0), la.PDMA),
        'TDMA_L': (sf.FunctionSpace(N, 'L', bc=(0, 0)), (0, 0), la.TDMA),
        'PDMA_L': (sf.FunctionSpace(N, 'L', bc=(0, 0, 0, 0)), (0, 0), la.PDMA),
    }
    for name, (S, (a, b), cls) in cases.items():
        M = sf.SpectralMatrix((S, a), (S, b))
        out[f'mat_{name}'] = M.diags().toarray()
        for axis in (0, 1, 2):
            shape = [5, 6, 7]
            shape[axis] = N
            f = rng.standard_normal(shape)+1j*rng.standard_normal(shape)
            sol = cls(M) if cls is not None else la.Solver(M)
            u = np.array(sol(f.copy(), axis=axis))
            out[f'f_{name}_{axis}'] = f
            out[f'u_{name}_{axis}'] = u
    # --- Poisson with SolverGeneric1ND, Legendre Dirichlet x F x F (demo/poisson3D.py structure)
    for fam, tag in (('L', 'leg'), ('C', 'cheb')):
        SD = sf.FunctionSpace(14, fam, bc=(0, 0))
        K1 = sf.FunctionSpace(12, 'F', dtype='D')
        K2 = sf.FunctionSpace(10, 'F', dtype='d')
        T = sf.TensorProductSpace(sf.comm, (SD, K1, K2))
        u = TrialFunction(T)
        v = TestFunction(T)
        f_hat = sf.Function(T)
        f_hat[:] = rng.standard_normal(T.shape(True))+1j*rng.standard_normal(T.shape(True))
        f_hat[SD.dim():] = 0
        out[f'poisson_{tag}_f'] = np.array(f_hat)
        if fam == 'L':
            mats = inner(v, div(grad(u)))
            H = la.SolverGeneric1ND(mats)
            u_hat = H(f_hat.copy(), sf.Function(T), constraints=((0, 0),))
            out[f'poisson_{tag}_u'] = np.array(u_hat)
            # Helmholtz-like without gauge: (v, div grad u) - 2.5 (v, u)
            mats2 = inner(v, div(grad(u))-2.5*u)
            H2 = la.SolverGeneric1ND(mats2)
            out[f'helm_{tag}_u'] = np.array(H2(f_hat.copy(), sf.Function(T)))
        else:
            mats = inner(v, div(grad(u))-2.5*u)
            H = Helmholtz(*mats)
            u_hat = H(f_hat.copy(), sf.Function(T))
            out[f'helm_{tag}_u'] = np.array(u_hat)
            c = sf.Function(T)
            out[f'helm_{tag}_matvec'] = np.array(H.matvec(u_hat, c))
    # --- biharmonic, Legendre x F x F (demo/biharmonic3D.py structure)
    SB = sf.FunctionSpace(16, 'L', bc=(0, 0, 0, 0))
    K1 = sf.FunctionSpace(12, 'F', dtype='D')
    K2 = sf.FunctionSpace(10, 'F', dtype='d')
    T = sf.TensorProductSpace(sf.comm, (K1, K2.get_unplanned() if False else sf.FunctionSpace(10, 'F', dtype='D'), SB)) if False else \
        sf.TensorProductSpace(sf.comm, (SB, K1, K2))
    u = TrialFunction(T)
    v = TestFunction(T)
    f_hat = sf.Function(T)
    f_hat[:] = rng.standard_normal(T.shape(True))+1j*rng.standard_normal(T.shape(True))
    f_hat[SB.dim():] = 0
    mats = inner(v, div(grad(div(grad(u)))))
    H = la.SolverGeneric1ND(mats)
    out['biharm_f'] = np.array(f_hat)
    out['biharm_u'] = np.array(H(f_hat.copy(), sf.Function(T)))
    # --- Chebyshev biharmonic solver (chebyshev/la.py:354-558, tests/test_matrices.py:625-658)
    from shenfun.chebyshev.la import Biharmonic
    SB = sf.FunctionSpace(16, 'C', bc=(0, 0, 0, 0))
    K1 = sf.FunctionSpace(12, 'F', dtype='D')
    K2 = sf.FunctionSpace(10, 'F', dtype='d')
    T = sf.TensorProductSpace(sf.comm, (SB, K1, K2))
    u = TrialFunction(T)
    v = TestFunction(T)
    mats = inner(v, div(grad(div(grad(u)))))
    H = Biharmonic(*mats)
    f_hat = sf.Function(T)
    f_hat[:] = rng.standard_normal(T.shape(True))+1j*rng.standard_normal(T.shape(True))
    f_hat[SB.dim():] = 0
    out['cheb_biharm_f'] = np.array(f_hat)
    u_hat = H(f_hat.copy(), sf.Function(T))
    out['cheb_biharm_u'] = np.array(u_hat)
    out['cheb_biharm_matvec'] = np.array(H.matvec(u_hat, sf.Function(T)))
    save('solvers', **out)


def channel_flow():
    """demo/MKM.py (KMM channel flow, Chebyshev x Fourier^2, IMEXRK222, 3/2 padding): velocity
    coefficients after 1 and 3 time steps for both convection forms.  File output and
    checkpointing (HDF5, absent here) are stubbed out; everything else is the reference's code."""
    import shenfun

    class _NoFile:
        def __init__(self, *a, **k):
            pass

        def update(self, *a, **k):
            pass

        def write(self, *a, **k):
            pass

    shenfun.ShenfunFile = _NoFile
    shenfun.Checkpoint = _NoFile
    sys.path.insert(0, '/root/reference/demo')
    import types
    if 'matplotlib' not in sys.modules:          # plotting is never reached (modplot < 0)
        mpl = types.ModuleType('matplotlib')
        mpl.pyplot = types.ModuleType('matplotlib.pyplot')
        sys.modules['matplotlib'] = mpl
        sys.modules['matplotlib.pyplot'] = mpl.pyplot
    from MKM import MKM
    out = {}
    for conv in (0, 1):
        N = (32, 16, 16)
        c = MKM(N=N, Re=180., dt=0.001, conv=conv, modplot=-1, modsave=10**8, moderror=-1, filename='golden_mkm',
                family='C', checkpoint=10**8, padding_factor=(1.5, 1.5, 1.5), timestepper='IMEXRK222')
        t, tstep = c.initialize(from_checkpoint=False)
        out[f'u0_conv{conv}'] = np.array(c.u_)
        out[f'g0_conv{conv}'] = np.array(c.g_)
        c.solve(t=t, tstep=tstep, end_time=0.001)
        out[f'u1_conv{conv}'] = np.array(c.u_)
        out[f'H1_conv{conv}'] = np.array(c.H_)
        c.solve(t=0.001, tstep=1, end_time=0.003)
        out[f'u3_conv{conv}'] = np.array(c.u_)
        out[f'g3_conv{conv}'] = np.array(c.g_)
    save('mkm_32_16_16', **out)


if __name__ == '__main__':
    which = sys.argv[1:] or None
    fns = {k: v for k, v in globals().items() if callable(v) and getattr(v, '__module__', '') == '__main__'
           and k not in ('save',)}
    for k, f in fns.items():
        if which is None or k in which:
            f()
