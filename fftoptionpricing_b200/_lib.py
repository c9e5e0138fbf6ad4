"""ctypes binding of libfftop_b200.so (the C ABI declared in include/fftop_b200.h).

The shared library is built in-tree (``fftoptionpricing_b200/csrc/Makefile`` ->
``fftoptionpricing_b200/libfftop_b200.so``).  There is no fallback of any kind: if the library is
missing, or no CUDA device is present, loading / ``Engine()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FFTOP_B200_LIB: another build of the same library (A/B runs of kernel variants under profiles/)
LIB_PATH = os.environ.get("FFTOP_B200_LIB") or os.path.join(_HERE, "libfftop_b200.so")

_dp = C.POINTER(C.c_double)


class fop_model_t(C.Structure):
    _fields_ = [("base", C.c_int), ("time_change", C.c_int)]


# every symbol include/fftop_b200.h declares (tests/test_capi_symbols.py checks the list
# against the header)
_SIGNATURES = {
    "fop_abi_version": (C.c_int, []),
    "fop_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "fop_destroy": (C.c_int, [C.c_void_p]),
    "fop_last_error": (C.c_char_p, [C.c_void_p]),
    "fop_status_string": (C.c_char_p, [C.c_int]),
    "fop_get_stream": (C.c_void_p, [C.c_void_p]),
    "fop_synchronize": (C.c_int, [C.c_void_p]),
    "fop_launch_count": (C.c_longlong, [C.c_void_p]),
    "fop_last_cos_contraction": (C.c_int, [C.c_void_p]),
    "fop_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "fop_profile_read": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, _dp, C.POINTER(C.c_longlong),
                                   C.c_int, C.POINTER(C.c_int)]),
    "fop_model_num_params": (C.c_int, [fop_model_t]),
    "fop_carr_madan_strikes": (C.c_int, [C.c_int, C.c_double, C.c_double, _dp]),
    "fop_fang_oost_strikes": (C.c_int, [C.c_double, C.c_double, C.c_double, C.c_int, _dp]),
    "fop_strike_underlying": (C.c_int, [C.c_double, C.c_double, C.c_double, C.c_int, _dp]),
    "fop_carr_madan": (C.c_int, [C.c_void_p, fop_model_t, C.c_void_p, C.c_int, C.c_int,
                                 C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                 C.c_int, C.c_void_p]),
    "fop_cos": (C.c_int, [C.c_void_p, fop_model_t, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                          C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_int,
                          C.c_void_p]),
    "fop_cos_greeks": (C.c_int, [C.c_void_p, fop_model_t, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                 C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_int,
                                 C.POINTER(C.c_int), C.c_int, C.c_void_p]),
    "fop_fsts": (C.c_int, [C.c_void_p, fop_model_t, C.c_void_p, C.c_int, C.c_int, C.c_double,
                           C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                           C.c_void_p]),
    "fop_objfn_cf": (C.c_int, [C.c_void_p, fop_model_t, C.c_void_p, C.c_int, C.c_void_p,
                               C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]),
    "fop_calibrate_cf": (C.c_int, [C.c_void_p, fop_model_t, _dp, _dp, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_double, C.c_double, C.c_int, C.c_int, C.c_double,
                                   C.c_ulonglong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_int)]),
    "fop_calibrate_cos": (C.c_int, [C.c_void_p, fop_model_t, _dp, _dp, C.c_void_p, C.c_int, C.c_double,
                                    C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_int, C.c_double, C.c_ulonglong, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.POINTER(C.c_int)]),
    "fop_cos_loss": (C.c_int, [C.c_void_p, fop_model_t, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                               C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "fop_fo_estimate": (C.c_int, [_dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                  C.c_double, C.c_int, _dp, C.c_int, _dp]),
    "fop_comm_unique_id": (C.c_int, [C.c_char_p]),
    "fop_comm_init": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
    "fop_gather_losses": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "fop_gather_losses_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "fop_comm_wait": (C.c_int, [C.c_void_p]),
    "fop_comm_destroy": (C.c_int, [C.c_void_p]),
    "fop_fft": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "fop_integrate_fft": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                    C.c_double, C.c_int, C.c_void_p]),
    "fop_cos_u_grid": (C.c_int, [_dp, C.c_int, C.c_double, C.c_int, _dp]),
    "fop_cos_cf_samples": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                     C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fop_carr_madan_u_grid": (C.c_int, [C.c_int, C.c_double, C.c_double, _dp]),
    "fop_carr_madan_cf_samples": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double,
                                            C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                            C.c_void_p]),
    "fop_fsts_u_grid": (C.c_int, [C.c_int, C.c_double, _dp]),
    "fop_fsts_samples": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                   C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                   C.c_int, C.c_void_p]),
    "fop_measure_fp64_peak": (C.c_int, [C.c_void_p, _dp]),
}
# not part of the public header: tuning probe used by profiles/ scripts
_PRIVATE = {"fop_microbench_fp64": (C.c_int, [C.c_void_p, _dp]),
            "fop_microbench_int8": (C.c_int, [C.c_void_p, _dp])}

_lib = None


def load_library() -> C.CDLL:
    """Loads libfftop_b200.so and binds every declared symbol.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            f"`make -C fftoptionpricing_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in {**_SIGNATURES, **_PRIVATE}.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def declared_symbols():
    return sorted(_SIGNATURES)
