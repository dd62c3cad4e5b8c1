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


if __name__ == '__main__':
    which = sys.argv[1:] or None
    fns = {k: v for k, v in globals().items() if callable(v) and getattr(v, '__module__', '') == '__main__'
           and k not in ('save',)}
    for k, f in fns.items():
        if which is None or k in which:
            f()
