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


def poly_1d_gl():
    """Chebyshev bases on the Gauss-Lobatto points (quad='GL': DCT-I in the reference,
    chebyshev/bases.py:331-338, 349-353): scalar product, forward, backward, plain and padded,
    plus the mass and stiffness matrices."""
    out = {}
    for name, bc in BCS.items():
        for N in (16, 15):
            S = sf.FunctionSpace(N, 'C', bc=bc, quad='GL') if bc is not None else sf.FunctionSpace(N, 'C', quad='GL')
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
                for (a, b) in ((0, 0), (0, 2)):
                    try:
                        M = sf.SpectralMatrix((S, a), (S, b))
                        out[f'M{a}{b}_{tag}'] = M.diags().toarray()
                    except Exception as e:  # noqa: BLE001
                        print('skip matrix GL', name, a, b, repr(e)[:80])
    save('poly1d_CGL', **out)


def poly_1d_lgl():
    """Legendre bases on the Gauss-Lobatto points (quad='GL', legendre/bases.py:128-129, 157-164):
    scalar product, forward, backward, plain and padded, mass matrix."""
    out = {}
    for name, bc in BCS.items():
        for N in (16, 15):
            S = sf.FunctionSpace(N, 'L', bc=bc, quad='GL') if bc is not None else sf.FunctionSpace(N, 'L', quad='GL')
            tag = f'{name}_{N}'
            u = rng.standard_normal(N)
            out[f'u_{tag}'] = u
            out[f'x_{tag}'] = S.points_and_weights()[0]
            out[f'wq_{tag}'] = S.points_and_weights()[1]
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
                try:
                    out[f'M00_{tag}'] = sf.SpectralMatrix((S, 0), (S, 0)).diags().toarray()
                except Exception as e:  # noqa: BLE001
                    print('skip matrix LGL', name, repr(e)[:80])
    save('poly1d_LGL', **out)


def jacobi_1d():
    """Orthogonal Jacobi spaces (jacobi/bases.py:150-260, quad='JG'): points, weights, norms, scalar product,
    forward, backward, plain and padded."""
    out = {}
    for (al, be) in ((0, 0), (1, 1), (0.5, 0.5), (2, 1), (0, 1.5)):
        for N in (16, 15):
            S = sf.FunctionSpace(N, 'J', alpha=al, beta=be)
            tag = f'{al}_{be}_{N}'
            u = rng.standard_normal(N)
            out[f'u_{tag}'] = u
            out[f'x_{tag}'], out[f'wq_{tag}'] = S.points_and_weights()
            out[f'h_{tag}'] = np.array([float(S.L2_norm_sq(i)) for i in range(N)])
            out[f'sp_{tag}'] = np.array(S.scalar_product(u.copy()))
            out[f'fw_{tag}'] = np.array(S.forward(u.copy()))
            vh = rng.standard_normal(N)
            out[f'vh_{tag}'] = vh
            out[f'bw_{tag}'] = np.array(S.backward(vh.copy()))
            Sp = S.get_dealiased(1.5)
            out[f'bwpad_{tag}'] = np.array(Sp.backward(vh.copy()))
            w = rng.standard_normal(Sp.shape(False))
            out[f'w_{tag}'] = w
            out[f'fwpad_{tag}'] = np.array(Sp.forward(w.copy()))
    save('jacobi1d', **out)


def chebyshevu_1d():
    """Orthogonal Chebyshev spaces of the second kind (chebyshevu/bases.py:60-190), quad 'GU' and 'GC'."""
    out = {}
    for quad in ('GU', 'GC'):
        for N in (16, 15):
            S = sf.FunctionSpace(N, 'U', quad=quad)
            tag = f'{quad}_{N}'
            u = rng.standard_normal(N)
            out[f'u_{tag}'] = u
            out[f'x_{tag}'], out[f'wq_{tag}'] = S.points_and_weights()
            out[f'sp_{tag}'] = np.array(S.scalar_product(u.copy()))
            out[f'fw_{tag}'] = np.array(S.forward(u.copy()))
            vh = rng.standard_normal(N)
            out[f'vh_{tag}'] = vh
            out[f'bw_{tag}'] = np.array(S.backward(vh.copy()))
            # (the reference's padded transforms of this family fail: chebyshevu/bases.py:252 scales the padded
            # output with a factor of the un-padded length)
    save('chebyshevu1d', **out)


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


def poly_3d_inhom():
    """Inhomogeneous boundary values in a tensor-product space (BoundaryValues.set_tensor_bcs,
    tensorproductspace.py:1909-2032): constants and functions of the other coordinates; the
    polynomial axis first and in the middle; forward, backward of the result, set_boundary_dofs."""
    import sympy as sp
    x, y, z = sp.symbols('x,y,z', real=True)
    out = {}
    for fam in ('C', 'L'):
        for layout in ('first', 'middle'):
            other = (y, z) if layout == 'first' else (x, z)
            kinds = {
                'const': {'left': {'D': 1.5}, 'right': {'D': -0.5}},
                'func': {'left': {'D': sp.cos(2*other[0])*sp.sin(other[1])}, 'right': {'D': 0.25+sp.sin(other[0])}},
                'neumann': {'left': {'N': 0.75}, 'right': {'N': sp.cos(other[1])}},
            }
            for kind, bc in kinds.items():
                SD = sf.FunctionSpace(14, fam, bc=bc)
                K1 = sf.FunctionSpace(12, 'F', dtype='D')
                K2 = sf.FunctionSpace(10, 'F', dtype='d')
                if layout == 'first':
                    T = sf.TensorProductSpace(sf.comm, (SD, K1, K2))
                else:
                    T = sf.TensorProductSpace(sf.comm, (K1, SD, K2), axes=(1, 0, 2))
                tag = f'{fam}_{kind}_{layout}'
                u = rng.standard_normal(T.shape(False))
                ua = sf.Array(T)
                ua[:] = u
                out[f'u_{tag}'] = u
                uh = T.forward(ua, sf.Function(T))
                out[f'fw_{tag}'] = np.array(uh)
                out[f'bw_{tag}'] = np.array(T.backward(uh, sf.Array(T)))
                f0 = sf.Function(T)
                f0.set_boundary_dofs()
                out[f'bdofs_{tag}'] = np.array(f0)
    save('poly3d_inhom', **out)


def refine_assign():
    """Function.refine / Function.assign (forms/arguments.py:1693-1827): spectral padding and
    truncation to another resolution, Fourier^3 (even and odd sizes) and Chebyshev x Fourier^2."""
    out = {}
    cases = {'f3d': ([('F', 12, 'D'), ('F', 10, 'D'), ('F', 8, 'd')], [(18, 16, 12), (8, 6, 6), (12, 15, 9)]),
             'c3d': ([('C', 12, None), ('F', 8, 'D'), ('F', 6, 'd')], [(18, 12, 10), (12, 6, 4), (16, 8, 6)])}
    for tag, (spec, targets) in cases.items():
        bases = [sf.FunctionSpace(n, f, bc=(0, 0)) if f == 'C' else sf.FunctionSpace(n, f, dtype=dt) for f, n, dt in spec]
        T = sf.TensorProductSpace(sf.comm, bases)
        u = rng.standard_normal(T.shape(False))
        ua = sf.Array(T)
        ua[:] = u
        uh = T.forward(ua, sf.Function(T))
        out[f'uh_{tag}'] = np.array(uh)
        for N in targets:
            r = uh.refine(N)
            out[f"ref_{tag}_{'x'.join(map(str, N))}"] = np.array(r)
    save('refine', **out)


def legendre_chebyshev_conversion():
    """legendre/dlt.py: leg2cheb (:353-421, also transposed), cheb2leg (:423-474) along each axis of real
    and complex arrays, even and odd lengths; DLT backward / forward / scalar product values through the
    Legendre space (the reference's default kinds; the FFT-based DLT agrees with them to rounding)."""
    from shenfun.legendre import dlt
    out = {}
    for N in (16, 33):
        for axis in (0, 1):
            shape = [5, 6]
            shape[axis] = N
            x = rng.standard_normal(shape)+1j*rng.standard_normal(shape)
            out[f'x_{N}_{axis}'] = x
            out[f'l2c_{N}_{axis}'] = dlt.leg2cheb(x.copy(), axis=axis)
            out[f'l2cT_{N}_{axis}'] = dlt.leg2cheb(x.copy(), axis=axis, transpose=True)
            out[f'c2l_{N}_{axis}'] = dlt.cheb2leg(x.copy(), axis=axis)
        xr = rng.standard_normal((N, 4))
        out[f'xr_{N}'] = xr
        out[f'l2c_r_{N}'] = dlt.leg2cheb(xr.copy(), axis=0)
        out[f'c2l_r_{N}'] = dlt.cheb2leg(xr.copy(), axis=0)
        L = sf.FunctionSpace(N, 'L')
        uh = sf.Function(L)
        uh[:] = rng.standard_normal(N)
        out[f'dlt_uh_{N}'] = np.array(uh)
        ub = np.array(L.backward(uh.copy()))
        out[f'dlt_bwd_{N}'] = ub
        ua = sf.Array(L)
        ua[:] = ub
        out[f'dlt_sp_{N}'] = np.array(L.scalar_product(ua, sf.Function(L)))
    save('dlt', **out)


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
    T = sf.TensorProductSpace(sf.comm, (SB, K1, K2))
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


def plugin_kernels():
    """The kernels the reference resolves BY NAME (optimization/__init__.py:47-84), called
    directly from its numba module with the reference's positional signatures: banded LU and
    solves, SolverGeneric1ND_solve_data, Helmholtz (Dirichlet and Neumann), Poisson_Solve_ADD,
    the Chebyshev biharmonic trio, the matvecs and the recursive Legendre/Jacobi transforms."""
    from shenfun.optimization import numba as ref
    from shenfun import la
    out = {}
    N = 18
    offs = {'TDMA': (-2, 0, 2), 'TDMA_O': (-1, 0, 1), 'PDMA': (-4, -2, 0, 2, 4), 'FDMA': (-2, 0, 2, 4),
            'TwoDMA': (0, 2), 'ThreeDMA': (0, 2, 4), 'HeptaDMA': (-4, -2, 0, 2, 4, 6, 8), 'DiagMA': (0,)}

    def rand_data(kind, lead=()):
        o = offs[kind]
        d = rng.uniform(-1, 1, lead+(len(o), N))
        d[..., o.index(0), :] += 6.0          # diagonally dominant: LU without pivoting is stable
        return d

    for kind in offs:
        data = rand_data(kind)
        out[f'{kind}_data'] = data.copy()
        lu = data.copy()
        if hasattr(ref, f'{kind}_LU'):
            getattr(ref, f'{kind}_LU')(lu)
        out[f'{kind}_lu'] = lu.copy()
        for axis in (0, 1, 2):
            shape = [4, 5, 6]
            shape[axis] = N
            u = rng.standard_normal(shape)+1j*rng.standard_normal(shape)
            out[f'{kind}_b{axis}'] = u.copy()
            getattr(ref, f'{kind}_Solve')(u, lu, axis)
            out[f'{kind}_u{axis}'] = u
        ur = rng.standard_normal(N+2)          # real 1-D line, longer than the matrix
        out[f'{kind}_b1d'] = ur.copy()
        getattr(ref, f'{kind}_inner_solve')(ur, lu)
        out[f'{kind}_u1d'] = ur
    # SolverGeneric1ND_solve_data: per-line factors data[i, j] (la.py:1300-1319)
    for kind in ('TDMA', 'PDMA', 'FDMA'):
        for naxes in (0, 1, 2):
            lines = (3, 4)
            data = rand_data(kind, lines)
            for i in range(lines[0]):
                for j in range(lines[1]):
                    getattr(ref, f'{kind}_LU')(data[i, j])
            shape = list(lines)
            shape.insert(naxes, N)
            u = rng.standard_normal(shape)+1j*rng.standard_normal(shape)
            out[f'g1nd_{kind}_{naxes}_data'] = data.copy()
            out[f'g1nd_{kind}_{naxes}_b'] = u.copy()
            ref.SolverGeneric1ND_solve_data(u, data, getattr(ref, f'{kind}_inner_solve'), naxes, True)
            out[f'g1nd_{kind}_{naxes}_u'] = u
    # Chebyshev Helmholtz, Dirichlet and Neumann (chebyshev/la.py:131-351)
    Nc = 20
    for tag, bc in (('D', (0, 0)), ('N', {'left': {'N': 0}, 'right': {'N': 0}})):
        S = sf.FunctionSpace(Nc, 'C', bc=bc)
        A = sf.inner(sf.Dx(sf.TrialFunction(S), 0, 2), sf.TestFunction(S))
        B = sf.inner(sf.TrialFunction(S), sf.TestFunction(S))
        B[2] = np.broadcast_to(B[2], A[2].shape)         # as the call site does, chebyshev/la.py:254-255
        B[-2] = np.broadcast_to(B[-2], A[2].shape)
        for k in (0, 2, 4):
            out[f'helm{tag}_A{k}'] = np.broadcast_to(np.asarray(A[k], dtype=float), (A.shape[0]-k,)).copy()
        for k in (-2, 0, 2):
            out[f'helm{tag}_B{k}'] = np.broadcast_to(np.asarray(B[k], dtype=float), (B.shape[0]-abs(k),)).copy()
        neumann = tag == 'N'
        for axis in (0, 1, 2):
            shape = [3, 4, 5]
            shape[axis] = Nc
            sshape = list(shape)
            sshape[axis] = 1
            alfa = rng.uniform(0.5, 1.5, sshape)
            beta = -rng.uniform(0.5, 30.0, sshape)
            if neumann:
                # the singular constant mode (la.pyx:733-736).  Put it on the LAST line: the numba
                # LU_Helmholtz_1D pins it by writing A_0[0] = 1/A_scale into the array shared by all
                # lines (numba/helmholtz.py:74-76), which would leak into every later line; the Cython
                # kernel works on per-call copies (la.pyx:726-736) and is the behaviour mirrored.
                beta.flat[-1] = 0.0
            d0, d1, d2, L = (np.zeros(shape) for _ in range(4))
            ref.LU_Helmholtz(A, B, alfa, beta, neumann, d0, d1, d2, L, axis)
            b = rng.standard_normal(shape)+1j*rng.standard_normal(shape)
            sl = [slice(None)]*3
            sl[axis] = slice(Nc-2, None)
            b[tuple(sl)] = 0
            u = np.zeros_like(b)
            ref.Solve_Helmholtz(b.copy(), u, neumann, d0, d1, d2, L, axis)
            for nm, arr in (('alfa', alfa), ('beta', beta), ('d0', d0), ('d1', d1), ('d2', d2), ('L', L), ('b', b), ('u', u)):
                out[f'helm{tag}_{axis}_{nm}'] = arr
            if not neumann:
                c = np.zeros_like(b)
                ref.Helmholtz_matvec(u, c, alfa, beta, A, B, axis)
                out[f'helm{tag}_{axis}_matvec'] = c
    # Poisson_Solve_ADD (chebyshev/la.py:13-30)
    S = sf.FunctionSpace(Nc, 'C', bc=(0, 0))
    A = sf.inner(sf.Dx(sf.TrialFunction(S), 0, 2), sf.TestFunction(S))
    out['add_A0'] = np.asarray(A[0], dtype=float).copy()
    out['add_A2'] = np.asarray(A[2], dtype=float).copy()
    out['add_scale'] = np.array(float(A.scale))
    for axis in (0, 1, 2):
        shape = [3, 4, 5]
        shape[axis] = Nc
        b = rng.standard_normal(shape)+1j*rng.standard_normal(shape)
        u = np.zeros_like(b)
        u = ref.Poisson_Solve_ADD(A, b.copy(), u, axis)
        out[f'add_{axis}_b'] = b
        out[f'add_{axis}_u'] = u
    # Chebyshev biharmonic (chebyshev/la.py:354-558)
    Nb = 24
    SBs = sf.FunctionSpace(Nb, 'C', bc=(0, 0, 0, 0))
    u_, v_ = sf.TrialFunction(SBs), sf.TestFunction(SBs)
    Sm = sf.inner(sf.Dx(u_, 0, 4), v_)
    Am = sf.inner(sf.Dx(u_, 0, 2), v_)
    Bm = sf.inner(u_, v_)
    dg = {'sii': Sm[0], 'siu': Sm[2], 'siuu': Sm[4], 'ail': Am[-2], 'aii': Am[0], 'aiu': Am[2],
          'bill': Bm[-4], 'bil': Bm[-2], 'bii': Bm[0], 'biu': Bm[2], 'biuu': Bm[4]}
    names = list(dg)
    for k in names:
        out[f'bih_{k}'] = np.asarray(dg[k], dtype=float).copy()
    M = Nb//2-2
    for axis in (0, 1, 2):
        shape = [3, 4, 5]
        shape[axis] = Nb
        sshape = list(shape)
        sshape[axis] = 1
        fshape = list(shape)
        fshape[axis] = M
        fshape = [2]+fshape
        a0 = 1.3
        alfa = -rng.uniform(0.5, 20.0, sshape)
        beta = rng.uniform(0.5, 200.0, sshape)
        u0, u1, u2, l0, l1, ak, bk = (np.zeros(fshape) for _ in range(7))
        ref.LU_Biharmonic(a0, alfa, beta, *[dg[k] for k in names], u0, u1, u2, l0, l1, axis)
        ref.Biharmonic_factor_pr(ak, bk, l0, l1, axis)
        b = rng.standard_normal(shape)+1j*rng.standard_normal(shape)
        sl = [slice(None)]*3
        sl[axis] = slice(Nb-4, None)
        b[tuple(sl)] = 0
        u = np.zeros_like(b)
        ref.Biharmonic_Solve(b.copy(), u, u0, u1, u2, l0, l1, ak, bk, a0, axis)
        c = np.zeros_like(b)
        ref.Biharmonic_matvec(u, c, a0, alfa, beta, *[dg[k] for k in names], axis)
        for nm, arr in (('alfa', alfa), ('beta', beta), ('u0', u0), ('u1', u1), ('u2', u2), ('l0', l0), ('l1', l1),
                        ('ak', ak), ('bk', bk), ('b', b), ('u', u), ('matvec', c)):
            out[f'bih_{axis}_{nm}'] = arr
        out['bih_a0'] = np.array(a0)
    # recursive transforms: Legendre and a Jacobi family (spectralbase.py:448-491, 516-558)
    for tag, S in (('leg', sf.FunctionSpace(20, 'L')), ('jac', sf.FunctionSpace(20, 'J', alpha=0.5, beta=-0.25)),
                   ('chebu', sf.FunctionSpace(20, 'U'))):
        x, w = S.points_and_weights()
        a = np.array(S.get_recursion_matrix(S.N+3, S.N+3).diags('dia').data)
        out[f'tr_{tag}_x'] = x
        out[f'tr_{tag}_w'] = w
        out[f'tr_{tag}_a'] = a
        for axis in (0, 1, 2):
            shape = [3, 4, 5]
            shape[axis] = S.N
            for dt in ('real', 'cplx'):
                uin = rng.standard_normal(shape)
                if dt == 'cplx':
                    uin = uin+1j*rng.standard_normal(shape)
                sp = np.zeros_like(uin)
                ref.transforms.scalar_product(uin, sp, x, w, axis, a)
                ev = np.zeros_like(uin)
                ref.transforms.evaluate_expansion_all(sp, ev, x, axis, a)
                out[f'tr_{tag}_{axis}_{dt}_in'] = uin
                out[f'tr_{tag}_{axis}_{dt}_sp'] = sp
                out[f'tr_{tag}_{axis}_{dt}_ev'] = ev
    save('plugin', **out)


def fourier_poisson():
    """demo/fourier_poisson3D.py:19-52 (BASELINE config 0) at its own size and at 64^3: the
    right-hand side coefficients, the solution coefficients and the L2 error."""
    from sympy import symbols, cos, sin
    from shenfun import inner, div, grad, TestFunction, TrialFunction, FunctionSpace, TensorProductSpace, \
        Array, Function, la
    out = {}
    x, y, z = symbols("x,y,z", real=True)
    ue = cos(4*x)+sin(4*y)+sin(6*z)
    fe = ue.diff(x, 2)+ue.diff(y, 2)+ue.diff(z, 2)
    for N in ((14, 15, 16), (64, 64, 64)):
        K0 = FunctionSpace(N[0], 'F', dtype='D')
        K1 = FunctionSpace(N[1], 'F', dtype='D')
        K2 = FunctionSpace(N[2], 'F', dtype='d')
        T = TensorProductSpace(sf.comm, (K0, K1, K2))
        u = TrialFunction(T)
        v = TestFunction(T)
        fj = Array(T, buffer=fe)
        f_hat = inner(v, fj, output_array=Function(T))
        A = inner(v, div(grad(u)))
        sol = la.SolverDiagonal(A)
        u_hat = sol(f_hat, Function(T), constraints=((0, 0),))
        uq = u_hat.backward()
        uj = Array(T, buffer=ue)
        error = np.sqrt(inner(1, (uj-uq)**2))
        assert abs(error) < 1e-6
        tag = 'x'.join(str(n) for n in N)
        fh, uh = np.array(f_hat), np.array(u_hat)
        if N[0] > 16:
            # keep the fixture small: the modes that are pure rounding noise (< 1e-14 of the largest
            # one) are stored as zero; the 1e-12 relative L2 gate is unaffected
            fh[np.abs(fh) < 1e-14*np.abs(fh).max()] = 0
            uh[np.abs(uh) < 1e-14*np.abs(uh).max()] = 0
        out[f'f_hat_{tag}'] = fh
        out[f'u_hat_{tag}'] = uh
        out[f'error_{tag}'] = np.array(abs(error))
    save('fourier_poisson3d', **out)


if __name__ == '__main__':
    which = sys.argv[1:] or None
    fns = {k: v for k, v in globals().items() if callable(v) and getattr(v, '__module__', '') == '__main__'
           and k not in ('save',)}
    for k, f in fns.items():
        if which is None or k in which:
            f()
