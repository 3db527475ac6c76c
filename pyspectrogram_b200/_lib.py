"""ctypes binding of libpsgb200.so (the C ABI in include/psg_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpsgb200.so")

PSG_WINDOW_KAISER = 0
PSG_WINDOW_BOXCAR = 1

PSG_IQ_C64 = 0
PSG_IQ_CI16 = 1
PSG_IQ_CI8 = 2

PSG_OK = 0
PSG_ERR_ARG = -1
PSG_ERR_UNSUPPORTED = -2
PSG_ERR_CUDA = -3
PSG_ERR_NODEVICE = -4
PSG_ERR_NOMEM = -5


class PsgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpsgb200 error {code}: {msg}")
        self.code = code


class PsgUnsupported(PsgError, NotImplementedError):
    pass


class PsgArgumentError(PsgError, ValueError):
    pass


# every symbol include/psg_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_SIGS = {
    "psg_version": (C.c_int, []),
    "psg_last_error": (C.c_char_p, []),
    "psg_device_count": (C.c_int, []),
    "psg_plan_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_double, C.c_int]),
    "psg_plan_destroy": (C.c_int, [_P]),
    "psg_plan_nfft": (C.c_int, [_P]),
    "psg_plan_window": (C.c_int, [_P, _P]),
    "psg_sti_run": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int, _P, C.c_int, C.c_int, C.c_int64,
                              C.c_float, C.c_float, _P, _P, _P]),
    "psg_sti_run_typed": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int64, C.c_int, _P, C.c_int, C.c_int, C.c_int64,
                                    C.c_float, C.c_float, _P, _P, _P]),
    "psg_sti_run_checked": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, _P, C.c_int, C.c_int,
                                      C.c_int64, C.c_float, C.c_float, _P, _P, _P, _P]),
    "psg_sti_host_typed": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, _P, C.c_int, C.c_int,
                                     C.c_int64, C.c_float, C.c_float, _P, _P, _P, _P]),
    "psg_median_time": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P]),
    "psg_minmax_time": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P, _P]),
    "psg_gather_bins": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, C.c_int, C.c_float, C.c_float, _P, _P]),
    "psg_sti_host": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int, _P, C.c_int, C.c_int,
                               C.c_int64, C.c_float, C.c_float, _P, _P, _P, _P]),
    "psg_debug_set_force_generic": (C.c_int, [C.c_int]),
    "psg_debug_set_variant": (C.c_int, [C.c_char_p]),
    "psg_debug_set_split_scratch": (C.c_int, [C.c_int64]),
    "psg_debug_set_mode_r_multi": (C.c_int, [C.c_int]),
    "psg_debug_set_host_chunk": (C.c_int, [C.c_int64]),
    "psg_debug_set_items_per_slot": (C.c_int, [C.c_int]),
    "psg_variant_count": (C.c_int, []),
    "psg_variant_name": (C.c_char_p, [C.c_int]),
    "psg_variant_logn": (C.c_int, [C.c_int]),
    "psg_window_table": (C.c_int, [C.c_int, C.c_int, C.c_double, _P, C.POINTER(C.c_double)]),
    "psg_launch_count": (C.c_int64, []),
    "psg_plan_variant": (C.c_char_p, [_P]),
}

_lib = None


def load():
    """Load (once) and return the ctypes library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m pyspectrogram_b200.build` "
            "(there is no CPU fallback for the PSD/STI path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the build is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGS)


def check(rc):
    if rc >= 0:
        return rc
    msg = load().psg_last_error().decode("utf-8", "replace")
    if rc == PSG_ERR_UNSUPPORTED:
        raise PsgUnsupported(rc, msg)
    if rc == PSG_ERR_ARG:
        raise PsgArgumentError(rc, msg)
    raise PsgError(rc, msg)
