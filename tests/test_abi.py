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
