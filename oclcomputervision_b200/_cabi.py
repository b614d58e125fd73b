"""ctypes binding of libraisr_b200.so (C-ABI declared in include/raisr_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C oclcomputervision_b200/csrc``.
There is no Python or CPU fallback: if the shared object is missing or no CUDA device is usable,
every call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_longlong, c_size_t, c_ubyte, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libraisr_b200.so")

RAISR_HOST, RAISR_DEVICE = 0, 1
E_ARG, E_CUDA, E_STATE, E_NOMEM, E_UNSUPPORTED = -1, -2, -3, -4, -5

# name -> (restype, argtypes); must list every symbol of include/raisr_b200.h
SIGNATURES = {
    "raisr_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int]),
    "raisr_destroy": (None, [c_void_p]),
    "raisr_set_filters": (c_int, [c_void_p, c_int, c_void_p, c_size_t]),
    "raisr_get_effective_filters": (c_int, [c_void_p, c_int, c_void_p, c_size_t, POINTER(c_int), POINTER(c_float)]),
    "raisr_pack_taps_b24": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "raisr_set_quantizers": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int]),
    "raisr_set_stream": (c_int, [c_void_p, c_void_p]),
    "raisr_set_option": (c_int, [c_void_p, c_char_p, c_longlong]),
    "raisr_upsample_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_void_p, c_int, c_int,
                                  c_size_t, c_int, c_int, c_int, POINTER(c_float)]),
    "raisr_upsample_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_void_p, c_int, c_int,
                                   c_size_t, c_int, c_int, c_int, POINTER(c_float)]),
    "raisr_upsample_bgra_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_void_p, c_int, c_int,
                                       c_size_t, c_int, c_int, c_int, POINTER(c_float)]),
    "raisr_upsample_bgra_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_void_p, c_int, c_int,
                                        c_size_t, c_int, c_int, c_int, POINTER(c_float)]),
    "raisr_bilinear_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_void_p, c_int, c_int,
                                  c_size_t, c_int, c_int, c_int, POINTER(c_float)]),
    "raisr_resize_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_int, c_void_p, c_int, c_int, c_size_t,
                                c_int, c_int, c_int, POINTER(c_float)]),
    "ocv_hist_grid_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_void_p, c_int, POINTER(c_float)]),
    "ocv_histeq_global_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_void_p, c_size_t, c_void_p, c_int, POINTER(c_float)]),
    "ocv_histeq_local_block_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_void_p, c_size_t, c_void_p, c_int, c_int,
                                          c_int, c_int, c_int, POINTER(c_float)]),
    "raisr_debug_hash": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_int, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_int]),
    "raisr_debug_hash_bgra": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_int]),
    "raisr_upsample_band_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_int, c_int, c_void_p,
                                       c_size_t, c_int, c_int, c_int]),
    "raisr_band_src_rows": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    "raisr_ipc_export": (c_int, [c_void_p, POINTER(c_ubyte)]),
    "raisr_ipc_open": (c_int, [POINTER(c_ubyte), POINTER(c_void_p)]),
    "raisr_ipc_close": (c_int, [c_void_p]),
    "raisr_p2p_copy2d": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t]),
    "raisr_flag_set": (c_int, [c_void_p, c_void_p, ctypes.c_uint]),
    "raisr_flag_wait": (c_int, [c_void_p, c_void_p, ctypes.c_uint, c_int]),
    "raisr_copy2d": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_int]),
    "raisr_timer_mark": (c_int, [c_void_p, c_int]),
    "raisr_timer_elapsed_ms": (c_int, [c_void_p, c_int, c_int, POINTER(c_float)]),
    "raisr_dev_alloc": (c_int, [c_void_p, POINTER(c_void_p), c_size_t]),
    "raisr_dev_free": (c_int, [c_void_p, c_void_p]),
    "raisr_host_alloc": (c_int, [POINTER(c_void_p), c_size_t]),
    "raisr_host_register": (c_int, [c_void_p, c_size_t]),
    "raisr_host_unregister": (c_int, [c_void_p]),
    "raisr_host_free": (c_int, [c_void_p]),
    "raisr_sync": (c_int, [c_void_p]),
    "raisr_launch_count": (c_longlong, [c_void_p]),
    "raisr_device_info": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), c_char_p, c_int]),
    "raisr_last_kernel_ms": (c_int, [c_void_p, POINTER(c_float), POINTER(c_float)]),
    "raisr_measure_ffma_tflops": (c_int, [c_void_p, POINTER(c_float)]),
    "raisr_last_error": (c_char_p, []),
    "raisr_version": (c_char_p, []),
}

_lib = None


class RaisrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libraisr_b200: %s (code %d)" % (msg, code))
        self.code = code


def load() -> ctypes.CDLL:
    """Load the shared object and bind every declared symbol.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libraisr_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C oclcomputervision_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().raisr_last_error()
        raise RaisrError(rc, msg.decode() if msg else "unknown error")
