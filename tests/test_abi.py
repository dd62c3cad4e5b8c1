"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol declared in include/shenb200.h (no compute calls without a GPU)."""
import ctypes

from shenfun_b200 import _lib


def test_library_loads_and_exports_every_declared_symbol(lib):
    names = _lib.exported_symbols()
    assert len(names) >= 7
    for name in names:
        assert hasattr(lib, name), f'{name} declared in include/shenb200.h but not exported'
    # every symbol bound by the Python host layer is declared in the header
    for name in lib._signatures:
        assert name in names


def test_version_and_error_string(lib):
    assert lib.shb_version() >= 100
    assert isinstance(lib.shb_last_error(), bytes)


def test_bad_arguments_return_error_codes_not_crashes(lib):
    rc = lib.shb_fft_c2c_forward(None, None, 1, 1, 8, 8, 1.0, None)
    assert rc == -1
    assert b'null' in lib.shb_last_error()
    buf = ctypes.create_string_buffer(16)
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.shb_fft_c2c_forward(p, p, 1, 1, 8, 16, 1.0, None)   # N > n_fft
    assert rc == -1


def test_product_path_fails_loudly_without_gpu():
    import pytest
    import torch
    import numpy as np
    import shenfun_b200 as sf
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    T = sf.TensorProductSpace(sf.comm, [sf.FunctionSpace(8, 'F', dtype='D'), sf.FunctionSpace(8, 'F', dtype='d')])
    with pytest.raises(sf.ShenB200Error):
        T.forward(np.zeros((8, 8)))


def test_plan_descriptor_validation_and_sizes_host_only():
    """`shb_plan_create` is host logic (no device call): sizes of the arrays and of the workspace,
    and the reference's constraints on the axis order (tensorproductspace.py:365-366, 382-385)."""
    import ctypes
    from shenfun_b200 import _lib
    lib = _lib.load()

    def make(shape, kinds, padded, axes, dd=0):
        d = _lib.PlanDesc()
        d.ndim = len(shape)
        for i in range(len(shape)):
            d.shape[i], d.padded[i], d.kind[i], d.axes[i] = shape[i], padded[i], kinds[i], axes[i]
        d.dealias_direct = dd
        h = ctypes.c_void_p()
        return lib.shb_plan_create(ctypes.byref(d), ctypes.byref(h)), h

    rc, h = make((512, 512, 512), (0, 0, 1), (768, 768, 768), (0, 1, 2))
    assert rc == 0
    assert lib.shb_plan_array_elems(h, 0) == 768**3
    assert lib.shb_plan_array_elems(h, 1) == 512*512*257
    # largest intermediate: after the r2c pass (768, 768, 257) complex; two ping-pong buffers, 3 components
    assert lib.shb_plan_workspace_bytes(h, 3) == 2*3*768*768*257*16
    assert lib.shb_plan_destroy(h) == 0
    assert make((8, 8), (1, 0), (8, 8), (0, 1))[0] == -1           # R2C axis not transformed first
    assert b'R2C' in lib.shb_last_error()
    assert make((8, 8), (0, 0), (8, 8), (1, 1))[0] == -1           # not a permutation
    assert make((8, 8), (0, 1), (12, 12), (0, 1), dd=1)[0] == -1   # dealias_direct with padding
    assert make((8, 8), (0, 1), (6, 8), (0, 1))[0] == -1           # padded < shape
