"""ctypes binding of libshenb200.so (the C ABI declared in include/shenb200.h).

There is no CPU fallback: if the shared library is missing, or a kernel is
launched without a B200-class device, the call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libshenb200.so')

_lib = None


class ShenB200Error(RuntimeError):
    pass


class PlanDesc(ctypes.Structure):
    """`shb_plan_desc` of include/shenb200.h."""
    _fields_ = [('ndim', ctypes.c_int), ('shape', ctypes.c_int64*4), ('padded', ctypes.c_int64*4),
                ('kind', ctypes.c_int*4), ('axes', ctypes.c_int*4), ('dealias_direct', ctypes.c_int)]


def _declare(lib):
    c = ctypes
    vp, i64, i32, f64 = c.c_void_p, c.c_int64, c.c_int, c.c_double
    lib.shb_last_error.restype = c.c_char_p
    lib.shb_last_error.argtypes = []
    lib.shb_version.restype = i32
    lib.shb_device_ok.restype = i32
    sig = {
        'shb_fft_c2c_forward': [vp, vp, i64, i64, i32, i32, f64, vp],
        'shb_fft_c2c_backward': [vp, vp, i64, i64, i32, i32, f64, i32, vp],
        'shb_fft_r2c_forward': [vp, vp, i64, i64, i32, i32, f64, vp],
        'shb_fft_c2r_backward': [vp, vp, i64, i64, i32, i32, f64, i32, vp],
        'shb_fft_c2c_lines': [i32, vp, vp, vp, i32, i64, i64, i64, i64, i64, i64, i64, i64, i32, i32, f64, i32, i32, i64, i32,
                              i64, vp],
        'shb_fft_c2c_chunked': [i32, vp, vp, i64, i64, i32, i32, f64, i32, i32, i32, vp],
        'shb_fft_c2c_peer': [i32, vp, vp, i32, i64, i64, i32, i32, f64, i32, i32, i32, vp],
        'shb_fft_c2r_cross': [vp, vp, vp, i64, i32, i32, f64, i32, vp],
        'shb_cross3d': [vp, vp, vp, i64, i32, vp],
        'shb_curl_hat': [vp, vp, vp, vp, vp, i32, i32, i32, vp],
        'shb_ns_rk_stage': [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f64, f64, f64, i32, vp],
        'shb_rk4_stage': [vp, vp, vp, vp, vp, i64, i32, f64, f64, i32, vp],
        'shb_rk4_stage_cplx': [vp, vp, vp, vp, vp, i64, i32, f64, f64, i32, vp],
        'shb_apply_mask': [vp, vp, i64, i32, vp],
        'shb_scale_mask': [vp, vp, vp, i64, i64, i32, vp],
        'shb_fft_c2c_backward_curl': [vp, vp, i64, i32, i32, i32, vp, vp, vp, f64, vp],
        'shb_kmm_vw': [vp, vp, vp, vp, vp, vp, i64, i64, vp],
        'shb_lincomb': [vp, i32, vp, vp, vp, i64, i64, i32, i32, vp, vp, vp],
        'shb_banded_factor': [i32, i32, i32, i32, vp, vp, vp, i64, i64, i32, i64, vp],
        'shb_banded_solve': [i32, i32, i32, vp, vp, i64, i64, i32, i32, i32, i32, vp],
        'shb_banded_matvec': [vp, vp, i32, i32, vp, vp, vp, i64, i64, i32, i32, i32, i32, i32, i32, vp],
        'shb_tail_matvec': [vp, vp, vp, vp, vp, i32, i64, i64, i32, i32, i32, i32, i32, i32, vp],
        'shb_dot_grad3d': [vp, vp, vp, i64, vp],
        'shb_helmholtz_factor': [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i32, i32, i32, vp],
        'shb_helmholtz_solve': [vp, vp, vp, vp, vp, vp, i64, i64, i32, i32, i32, i32, i32, vp],
        'shb_helmholtz_factor_lean': [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i32, vp],
        'shb_helmholtz_solve_lean': [vp, vp, vp, vp, vp, vp, vp, i64, i64, i32, i32, i32, vp],
        'shb_helmholtz_lean_table_rows': [i32],
        'shb_cheb_scalar_product': [vp, vp, i64, i64, i32, i32, f64, i32, i32, i32, vp, vp, vp],
        'shb_helmholtz_matvec': [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i32, i32, i32, vp],
        'shb_biharmonic_factor': [vp, vp, vp, vp, vp, i64, i64, i32, vp],
        'shb_biharmonic_solve': [vp, vp, vp, i64, i64, i32, i32, i32, vp],
        'shb_biharmonic_matvec': [vp, vp, vp, vp, vp, vp, i64, i64, i32, i32, i32, vp],
        'shb_gemm_axis': [vp, vp, vp, i32, i32, i64, i64, i32, vp],
        'shb_gemm_axis_parity': [vp, vp, vp, i32, i32, i32, i32, i64, i64, i32, vp],
        'shb_cheb_backward': [vp, vp, i64, i64, i32, i32, i32, i32, i32, vp, vp, i32, vp, vp],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.restype = i32
        fn.argtypes = argtypes
    # N-D plan API
    lib.shb_plan_create.restype = i32
    lib.shb_plan_create.argtypes = [c.POINTER(PlanDesc), c.POINTER(vp)]
    lib.shb_plan_destroy.restype = i32
    lib.shb_plan_destroy.argtypes = [vp]
    lib.shb_plan_workspace_bytes.restype = i64
    lib.shb_plan_workspace_bytes.argtypes = [vp, i32]
    lib.shb_plan_array_elems.restype = i64
    lib.shb_plan_array_elems.argtypes = [vp, i32]
    for name in ('shb_plan_forward', 'shb_plan_backward'):
        fn = getattr(lib, name)
        fn.restype = i32
        fn.argtypes = [vp, vp, vp, vp, i32, vp]
    return sig


def load():
    """Load (once) and return the ctypes handle; raise if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ShenB200Error(
                f'{LIB_PATH} not found: build it with ./build.sh or '
                '`python -c "import __graft_entry__ as g; g.build()"` '
                '(there is no CPU fallback)')
        lib = ctypes.CDLL(LIB_PATH)
        lib._signatures = _declare(lib)
        _lib = lib
    return _lib


SHB_ERR_UNSUPPORTED = -3        # include/shenb200.h: shb_status


def check(rc):
    if rc != 0:
        msg = load().shb_last_error().decode()
        raise ShenB200Error(f'libshenb200 error {rc}: {msg}')


def exported_symbols():
    """Names declared by include/shenb200.h (used by the CPU-side ABI test)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), 'include', 'shenb200.h')
    txt = open(hdr).read()
    return sorted(set(re.findall(r'\b(shb_[a-z0-9_]+)\s*\(', txt)))


def stream_ptr():
    """cudaStream_t of torch's current stream as an integer for the C ABI."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
