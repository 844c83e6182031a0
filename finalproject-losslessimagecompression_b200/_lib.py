"""ctypes binding of libflic_b200.so -- the C ABI declared in include/flic_b200.h.

There is no fallback: if the library has not been built (python __graft_entry__.py, or
python finalproject-losslessimagecompression_b200/build.py) every product entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# FLIC_B200_LIB selects another build of the same library (kernel experiments); default: in-tree
LIB_PATH = os.environ.get("FLIC_B200_LIB") or os.path.join(HERE, "libflic_b200.so")

# status bits (include/flic_b200.h)
ST_ZERO_SCALE = 1
ST_OUT_OF_WINDOW = 2
ST_UNDERRUN = 4
ST_NONFINITE = 8
ST_BAD_END_STATE = 16
ST_NO_SYMBOL = 32
ST_TOO_LONG = 64

E_ARG, E_CAPACITY, E_NOMEM, E_STATUS = -1, -2, -3, -4

_vp = C.c_void_p
_i64 = C.c_int64
_u64 = C.c_uint64

# name -> (restype, argtypes); mirrors include/flic_b200.h one to one
SIGNATURES = {
    "flic_abi_version": (C.c_int, []),
    "flic_last_error": (C.c_char_p, []),
    "flic_kernel_launches": (_i64, []),
    "flic_last_coder_kernel": (C.c_char_p, [C.c_int]),
    "flic_set_decode_kernel": (C.c_int, [C.c_int]),
    "flic_decode_cluster_capacity": (_i64, [C.c_int]),
    "flic_cdf_tables": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "flic_debug_expf": (C.c_int, [_vp, _vp, _i64, _vp]),
    "flic_debug_part1": (C.c_int, [_vp, _vp, _i64, _vp]),
    "flic_debug_div_check": (C.c_int, [_i64, _u64, C.c_int, _vp, _vp]),
    "flic_debug_push_check": (C.c_int, [_i64, _u64, _vp, _vp]),
    "flic_dlogistic_log_prob": (C.c_int, [_vp, _vp, _vp, _i64, _i64, C.c_int, C.c_float, _vp, _vp, _vp]),
    "flic_dlogistic_sample": (C.c_int, [_vp, _vp, _vp, _i64, C.c_int, _vp, _vp]),
    "flic_encode_workspace_bytes": (_i64, [_i64, _i64]),
    "flic_rans_encode": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp]),
    "flic_rans_decode": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, C.c_int, _vp]),
    "flic_rans_decode_resume": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "flic_gather_words": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "flic_couple_add_round": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, C.c_int, C.c_int, _vp]),
    "flic_u8_to_grid": (C.c_int, [_vp, _vp, _i64, _vp]),
    "flic_grid_to_u8": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "flic_permute_channels": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "flic_squeeze": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, C.c_int, C.c_int, _vp]),
    "flic_codec_create": (C.c_int, [C.c_int, _i64, _i64, C.POINTER(_vp)]),
    "flic_codec_destroy": (None, [_vp]),
    "flic_host_alloc": (C.c_int, [C.POINTER(_vp), _i64]),
    "flic_host_free": (None, [_vp]),
    "flic_codec_encode": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, C.POINTER(_i64)]),
    "flic_codec_decode": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "flic_codec_probe_copies": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    "flic_rans_encode_single": (C.c_int, [_vp, _u64, _i64, _vp, _vp, _vp, _vp, C.POINTER(_i64), C.POINTER(_u64),
                                          C.POINTER(C.c_int32)]),
    "flic_rans_decode_single": (C.c_int, [_vp, _u64, _vp, _i64, _i64, _vp, _vp, _vp, C.POINTER(_u64),
                                          C.POINTER(C.c_int32)]),
}

_lib = None


class FlicError(RuntimeError):
    """A C-ABI call failed (CUDA error or bad argument)."""


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FlicError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root. "
                "There is no CPU fallback for this path.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the header and the library drifted apart
            fn.restype = res
            fn.argtypes = args
        if L.flic_abi_version() != 2:
            raise FlicError("libflic_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    msg = lib().flic_last_error().decode("utf-8", "replace")
    raise FlicError(f"{what or 'flic call'} failed (code {rc}): {msg}")


def kernel_launches() -> int:
    return int(lib().flic_kernel_launches())


def status_message(bits: int) -> str:
    names = [(ST_ZERO_SCALE, "scale == 0"), (ST_OUT_OF_WINDOW, "symbol outside the 2048-bin window / off the 1/256 grid"),
             (ST_UNDERRUN, "word buffer under-run / output too small"), (ST_NONFINITE, "non-finite scale or |mean| > 16384"),
             (ST_BAD_END_STATE, "decoder did not end at 1<<32"), (ST_NO_SYMBOL, "no symbol matches (corrupt stream)"),
             (ST_TOO_LONG, "a single stream of 2^32 words or more")]
    return "; ".join(n for b, n in names if bits & b) or "ok"


def raise_for_status(bits: int) -> None:
    """Map a stream status word to the exception the reference raises where it has one."""
    if not bits:
        return
    if bits & ST_ZERO_SCALE:
        raise ZeroDivisionError("float division")  # rans/rans.cpp:1435-1437
    raise ValueError("rANS stream invalid: " + status_message(bits))
