"""ctypes binding of the C ABI declared in include/season_nerf_b200.h.

There is no CPU fallback: if the shared library is missing and cannot be built (nvcc absent) every
entry point raises.  The library is built in-tree (season_nerf_b200/lib/) by season_nerf_b200/build.py.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libseason_nerf_b200.so")

F32, BF16, F64 = 0, 1, 2

_p, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> argtypes; every function returns int (0 = ok) unless listed in _RESTYPES
SIGNATURES = {
    "snb_version": [],
    "snb_num_sms": [],
    "snb_launch_count": [],
    "snb_error_string": [_i],
    "snb_sample_rays": [_p, _p, _p, _i, _i, _i, _p, _p, _p],
    "snb_camera_rays": [C.POINTER(C.c_double), _p, _p, _ll, _i, _i, C.c_double, C.c_double, C.POINTER(C.c_double), _p, _p, _p,
                        _p, _p],
    "snb_solar_rays": [C.POINTER(C.c_double), C.POINTER(C.c_double), _p, _p, _p, _i, _p, _p, _p, _p, _p],
    "snb_supervised_sample": [_p, _p, _p, _i, _i, _f, _ll, _p, _p],
    "snb_solar_tops": [_p, _ll, C.POINTER(C.c_double), _i, _p, _p],
    "snb_composite_fwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p],
    "snb_composite_bwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "snb_heads_composite_fwd": [_p, _i, _p, _i, _p, _i, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p],
    "snb_heads_composite_bwd": [_p, _i, _p, _i, _p, _i, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "snb_solar_loss_fwd": [_p, _i, _p, _i, _p, _i, _i, _p, _p, _p],
    "snb_solar_loss_bwd": [_p, _i, _p, _i, _p, _i, _i, _p, _p, _p, _p],
    "snb_render_composite_raw": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p],
    "snb_year_sweep_raw": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p],
    "snb_march_transmittance": [_p, _p, _ll, _i, _p, _p],
    "snb_cli_composite": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p],
    "snb_cli_classic_shadow": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p],
    "snb_year_sweep": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p],
    "snb_pe_encode": [_p, _i, _ll, _i, _i, _p, _i, _i, _i, _i, _p],
    "snb_gemm": [_p, _i, _i, _p, _i, _i, _p, _i, _p, _f, _i, _ll, _i, _i, _i, _i, _p],
    "snb_gemm_stats": [_p, _i, _p, _i, _p, _i, _p, _f, _ll, _i, _i, _p, _p],
    "snb_gemm_stats_xf": [_p, _i, _p, _p, _p, _i, _p, _i, _p, _f, _ll, _i, _i, _p, _p, _i, _p],
    "snb_gemm_sine_fwd": [_p, _i, _p, _i, _p, _i, _p, _i, _p, _f, _ll, _i, _i, _p],
    "snb_loss_tail_fwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _f, _i, _p, _p, _p],
    "snb_loss_tail_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _f, _i, _p, _p, _p, _p, _p, _p, _p, _p],
    "snb_stage_weights": [_p, _p, _p, _p, _p, _p, _i, _p],
    "snb_gemm_sine_bwd": [_p, _i, _p, _i, _p, _i, _p, _i, _p, _p, _p, _p, _f, _ll, _i, _i, _p, _p],
    "snb_bn_bwd_apply": [_p, _i, _p, _i, _p, _p, _p, _p, _p, _f, _p, _i, _ll, _i, _i, _p],
    "snb_bn_finalize": [_p, _p, _i, _ll, _i, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p],
    "snb_col_stats": [_p, _i, _i, _ll, _i, _p, _p, _p],
    "snb_sine_fwd": [_p, _i, _p, _p, _p, _i, _ll, _i, _i, _p],
    "snb_sine_bwd_reduce": [_p, _i, _p, _i, _p, _p, _p, _p, _ll, _i, _i, _p, _p, _p],
    "snb_sine_bwd_apply": [_p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _ll, _i, _i, _p],
    "snb_convert": [_p, _i, _i, _p, _i, _i, _ll, _i, _p],
    "snb_fused_eval2": [_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, _p, _ll, _i, _p, _p, _p, _p, _p, _p],
}
_RESTYPES = {"snb_launch_count": _ll, "snb_error_string": C.c_char_p}

_lib = None


class SeasonNerfCudaError(RuntimeError):
    pass


def load():
    """Load (building first if needed and possible) the sm_100a library.  Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    # build() is a no-op when lib/build.stamp equals the digest of the current sources: a library left over from older
    # sources (stale in-tree .so next to new ctypes signatures = undefined behaviour) is rebuilt here, loudly or not at all
    from . import build as _build
    try:
        _build.build()
    except Exception as e:  # no silent fallback: the product path needs the CUDA library
        raise SeasonNerfCudaError(
            "season_nerf_b200: CUDA library %s is missing or stale and could not be built (%s). "
            "Run `python -m season_nerf_b200.build`; there is no CPU fallback." % (LIB_PATH, e))
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch: fail loudly
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, _i)
    _lib = lib
    return lib


def check(code):
    if code != 0:
        msg = load().snb_error_string(code)
        raise SeasonNerfCudaError("season_nerf_b200 CUDA call failed (%d): %s" % (code, msg.decode() if msg else "?"))


def launch_count():
    return int(load().snb_launch_count())
