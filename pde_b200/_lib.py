"""ctypes binding of libheston_b200.so (include/heston_b200.h).

The shared library is the product; this module only declares its C ABI to Python.  There is
no fallback: if the library is missing it is built with nvcc, and if that fails, or no CUDA
device is present when a compute entry point is called, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# PDE_B200_LIB points at an alternative build of the same library (kernel-shape experiments)
LIB_PATH = os.environ.get("PDE_B200_LIB") or os.path.join(HERE, "csrc", "libheston_b200.so")

HB_OK, HB_ERR_INVALID_ARGUMENT, HB_ERR_INVALID_PARAMETER, HB_ERR_CUDA, HB_ERR_STATE = range(5)
HB_MODE_REFGRID, HB_MODE_FFT = 0, 1
HB_NEQ_WIDTH = 22
HB_SABR_CPP, HB_SABR_PY = 0, 1

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p

# every symbol include/heston_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "hb_version": (C.c_int, []),
    "hb_last_error": (C.c_char_p, []),
    "hb_device_count": (C.c_int, []),
    "hb_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(_vp)]),
    "hb_plan_destroy": (C.c_int, [_vp]),
    "hb_surface_set": (C.c_int, [_vp, C.c_int, _dp, _dp, _u8p, _dp, C.c_double, C.c_double, C.c_double]),
    "hb_set_bounds": (C.c_int, [_vp, _dp, _dp]),
    "hb_plan_set_truncation": (C.c_int, [_vp, C.c_double]),
    "hb_plan_log_cut": (C.c_double, [_vp]),
    "hb_plan_n_options": (C.c_int, [_vp]),
    "hb_plan_n_maturities": (C.c_int, [_vp]),
    "hb_price": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "hb_objective": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "hb_normal_eq": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "hb_jacobian": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "hb_implied_vol": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "hb_implied_vol_host": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "hb_greeks": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "hb_greeks_host": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "hb_cf": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, C.c_int, C.c_double, C.c_double, C.c_double,
                        _vp, _vp]),
    "hb_fft_batch": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "hb_sync": (C.c_int, [_vp]),
    "hb_price_host": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "hb_objective_host": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "hb_normal_eq_host": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "hb_jacobian_host": (C.c_int, [_vp, _dp, C.c_int, _dp, _dp]),
    "hb_model_validate": (C.c_int, [_dp]),
    "hb_model_cf": (C.c_int, [_dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _dp,
                              C.c_int]),
    "hb_model_price_options": (C.c_int, [_dp, C.c_int, _dp, C.c_int, _dp, C.c_double, C.c_double, C.c_double, C.c_int,
                                         _dp, C.c_int]),
    "hb_sabr_vols": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "hb_sabr_objective": (C.c_int, [C.c_double, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int,
                                    _vp, _vp]),
    "hb_sabr_vols_host": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, _dp, _dp, C.c_int, _dp]),
    "hb_measure_fp64_peak": (C.c_int, [C.c_int, C.c_double, _dp]),
    "hb_launch_count": (C.c_uint64, []),
    "hb_plan_profile": (C.c_int, [_vp, C.c_int]),
    "hb_plan_profile_read": (C.c_int, [_vp, _dp, C.POINTER(C.c_longlong)]),
}

_lib = None


def load() -> C.CDLL:
    """Load (building first if needed) the CUDA library.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.environ.get("PDE_B200_LIB"):
        from .csrc.build import build

        build()  # no-op when the library is newer than every source and header
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class HestonB200Error(RuntimeError):
    pass


def check(rc: int) -> None:
    """Map hb_status to the exceptions the reference's pybind11 layer raises
    (std::invalid_argument -> ValueError; everything else RuntimeError)."""
    if rc == HB_OK:
        return
    msg = load().hb_last_error().decode()
    if rc in (HB_ERR_INVALID_ARGUMENT, HB_ERR_INVALID_PARAMETER):
        raise ValueError(msg)
    raise HestonB200Error(msg)
