"""ctypes/numpy face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package never does (tests/test_no_oracle_in_product.py
greps for it).

Two checkers live here:
  * the C restatement of rans/rans.pyx (oracle/rans_oracle.c -> oracle/liboracle.so);
  * the reference's own rans.pyx re-cythonised into oracle/_ref/ (`ref_rans()`), when built.
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import sys

import numpy as np

from . import build as _build

_lib = None

ERR_NAMES = {0: "ok", 1: "ZeroDivisionError(float division)", 2: "ZeroDivisionError(integer division)",
             3: "OverflowError(negative to unsigned)", 4: "buffer under-run"}


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = _build.ORACLE_SO
        if not os.path.exists(path) or os.path.exists(os.path.join(_build.HERE, "rans_oracle.c")):
            path = _build.build_oracle()
        L = C.CDLL(path)
        f32p, u32p, u64p, i64p, i32p = (C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_int64), C.POINTER(C.c_int32))
        L.flic_oracle_expf_restated.restype = C.c_float
        L.flic_oracle_expf_restated.argtypes = [C.c_float]
        L.flic_oracle_expf_restated_fma.restype = C.c_float
        L.flic_oracle_expf_restated_fma.argtypes = [C.c_float]
        L.flic_oracle_expf_host.restype = C.c_float
        L.flic_oracle_expf_host.argtypes = [C.c_float]
        L.flic_oracle_expf_sweep.restype = C.c_uint64
        L.flic_oracle_expf_sweep.argtypes = [C.c_uint32, C.c_uint32, C.c_int, u32p, C.c_uint32]
        L.flic_oracle_expf_compare.restype = None
        L.flic_oracle_expf_compare.argtypes = [C.c_uint32, C.c_int64, f32p, u64p, u64p, u32p, C.c_uint32]
        L.flic_oracle_part1_compare.restype = None
        L.flic_oracle_part1_compare.argtypes = [C.c_uint32, C.c_int64, i32p, u64p, u32p, C.c_uint32]
        L.flic_oracle_cdf.restype = C.c_int
        L.flic_oracle_cdf.argtypes = [C.c_float] * 4
        L.flic_oracle_lower.restype = C.c_int
        L.flic_oracle_lower.argtypes = [C.c_float]
        L.flic_oracle_tables.restype = C.c_int
        L.flic_oracle_tables.argtypes = [f32p, f32p, f32p, C.c_int64, i32p, u64p, u64p]
        L.flic_oracle_encode.restype = C.c_int
        L.flic_oracle_encode.argtypes = [C.c_uint64, C.c_int64, f32p, f32p, f32p, u32p, i64p, u64p]
        L.flic_oracle_decode.restype = C.c_int
        L.flic_oracle_decode.argtypes = [C.c_uint64, u32p, C.c_int64, C.c_int64, f32p, f32p, f32p, u64p, i64p]
        L.flic_oracle_encode_streams.restype = C.c_int
        L.flic_oracle_encode_streams.argtypes = [f32p, f32p, f32p, i64p, C.c_int64, u32p, i64p, u64p, i32p, C.c_int]
        L.flic_oracle_decode_streams.restype = C.c_int
        L.flic_oracle_decode_streams.argtypes = [f32p, f32p, i64p, C.c_int64, u32p, i64p, u64p, i32p, f32p, C.c_int]
        _lib = L
    return _lib


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _raise(rc: int):
    if rc in (1, 2):
        raise ZeroDivisionError(ERR_NAMES[rc])
    if rc == 3:
        raise OverflowError(ERR_NAMES[rc])
    if rc:
        raise IndexError(ERR_NAMES.get(rc, str(rc)))


def tables(x, mean, scale):
    """(lower_i int32, start uint64, freq uint64) per symbol -- rans/rans.pyx:49-56."""
    x, mean, scale = _f32(x), _f32(mean), _f32(scale)
    n = x.size
    lo = np.empty(n, np.int32)
    st = np.empty(n, np.uint64)
    fr = np.empty(n, np.uint64)
    _raise(lib().flic_oracle_tables(_p(x, C.c_float), _p(mean, C.c_float), _p(scale, C.c_float), n,
                                    _p(lo, C.c_int32), _p(st, C.c_uint64), _p(fr, C.c_uint64)))
    return lo, st, fr


def encode(state: int, n: int, x, mean, scale):
    """Same contract as the reference's rans.encode (rans/rans.pyx:37) on numpy arrays:
    returns (final_state:int, words: np.uint32[...] in emission order)."""
    x, mean, scale = _f32(x), _f32(mean), _f32(scale)
    buf = np.empty(max(int(n), 1), np.uint32)
    nw = C.c_int64(0)
    st = C.c_uint64(0)
    _raise(lib().flic_oracle_encode(C.c_uint64(state), int(n), _p(x, C.c_float), _p(mean, C.c_float),
                                    _p(scale, C.c_float), _p(buf, C.c_uint32), C.byref(nw), C.byref(st)))
    return int(st.value), buf[: nw.value].copy()


def decode(state: int, buffer_, n: int, mean_, scale_):
    """Same contract as rans.decode (rans/rans.pyx:69): buffer_/mean_/scale_ REVERSED by the
    caller; returns (end_state:int, message np.float32[n] reversed)."""
    buf = np.ascontiguousarray(np.asarray(buffer_, dtype=np.uint32))
    mean, scale = _f32(mean_), _f32(scale_)
    msg = np.empty(max(int(n), 1), np.float32)
    st = C.c_uint64(0)
    used = C.c_int64(0)
    _raise(lib().flic_oracle_decode(C.c_uint64(state), _p(buf, C.c_uint32), buf.size, int(n),
                                    _p(mean, C.c_float), _p(scale, C.c_float), _p(msg, C.c_float),
                                    C.byref(st), C.byref(used)))
    return int(st.value), msg[: int(n)].copy()


def encode_streams(x, mean, scale, offsets, n_threads: int = 1):
    """Code every stream [offsets[s], offsets[s+1]) from state 1<<32.
    Returns (packed words uint32, word_offsets int64[n+1], final_states uint64[n], status int32[n])."""
    x, mean, scale = _f32(x), _f32(mean), _f32(scale)
    off = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64))
    ns = off.size - 1
    words = np.zeros(max(x.size, 1), np.uint32)
    counts = np.zeros(max(ns, 1), np.int64)
    states = np.zeros(max(ns, 1), np.uint64)
    status = np.zeros(max(ns, 1), np.int32)
    lib().flic_oracle_encode_streams(_p(x, C.c_float), _p(mean, C.c_float), _p(scale, C.c_float),
                                     _p(off, C.c_int64), ns, _p(words, C.c_uint32), _p(counts, C.c_int64),
                                     _p(states, C.c_uint64), _p(status, C.c_int32), int(n_threads))
    counts = counts[:ns]
    woff = np.zeros(ns + 1, np.int64)
    np.cumsum(counts, out=woff[1:])
    packed = np.empty(int(woff[-1]), np.uint32)
    for s in range(ns):
        packed[woff[s]: woff[s + 1]] = words[off[s]: off[s] + counts[s]]
    return packed, woff, states[:ns].copy(), status[:ns].copy()


def decode_streams(packed, word_offsets, states, mean, scale, offsets, n_threads: int = 1):
    """Inverse of encode_streams.  Returns (x float32 in forward order, end_states, status)."""
    mean, scale = _f32(mean), _f32(scale)
    off = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64))
    woff = np.asarray(word_offsets, dtype=np.int64)
    packed = np.asarray(packed, dtype=np.uint32)
    ns = off.size - 1
    words = np.zeros(max(mean.size, 1), np.uint32)
    counts = np.ascontiguousarray(np.diff(woff)) if ns else np.zeros(1, np.int64)
    for s in range(ns):
        words[off[s]: off[s] + counts[s]] = packed[woff[s]: woff[s + 1]]
    st = np.ascontiguousarray(np.asarray(states, dtype=np.uint64)).copy()
    status = np.zeros(max(ns, 1), np.int32)
    out = np.zeros(max(mean.size, 1), np.float32)
    lib().flic_oracle_decode_streams(_p(mean, C.c_float), _p(scale, C.c_float), _p(off, C.c_int64), ns,
                                     _p(words, C.c_uint32), _p(counts, C.c_int64), _p(st, C.c_uint64),
                                     _p(status, C.c_int32), _p(out, C.c_float), int(n_threads))
    return out[: mean.size], st[:ns], status[:ns]


def expf_restated(x: float, fma: bool = False) -> float:
    L = lib()
    return float((L.flic_oracle_expf_restated_fma if fma else L.flic_oracle_expf_restated)(C.c_float(x)))


def expf_host(x: float) -> float:
    return float(lib().flic_oracle_expf_host(C.c_float(x)))


def expf_sweep(lo_bits: int, hi_bits: int, fma: bool = False, max_bad: int = 16):
    bad = np.zeros(max_bad, np.uint32)
    n = lib().flic_oracle_expf_sweep(lo_bits, hi_bits, int(fma), _p(bad, C.c_uint32), max_bad)
    return int(n), bad[: min(int(n), max_bad)].copy()


def expf_compare(lo_bits: int, got: np.ndarray, max_bad: int = 16):
    """got[i] = candidate expf of the float with bit pattern lo_bits+i.
    Returns (mismatches vs host libm, mismatches vs the restatement, first bad bit patterns)."""
    got = np.ascontiguousarray(got, dtype=np.float32)
    bh, br = C.c_uint64(0), C.c_uint64(0)
    bad = np.zeros(max_bad, np.uint32)
    lib().flic_oracle_expf_compare(lo_bits, got.size, _p(got, C.c_float), C.byref(bh), C.byref(br),
                                   _p(bad, C.c_uint32), max_bad)
    return int(bh.value), int(br.value), bad[: min(int(bh.value), max_bad)].copy()


def part1_compare(lo_bits: int, got: np.ndarray, max_bad: int = 16):
    """got[i] = candidate part1 for the float argument with bit pattern lo_bits+i (rans.pyx:25-26,34).
    Returns (mismatches vs the reference arithmetic with the host libm, first bad bit patterns)."""
    got = np.ascontiguousarray(got, dtype=np.int32)
    b = C.c_uint64(0)
    bad = np.zeros(max_bad, np.uint32)
    lib().flic_oracle_part1_compare(lo_bits, got.size, _p(got, C.c_int32), C.byref(b), _p(bad, C.c_uint32), max_bad)
    return int(b.value), bad[: min(int(b.value), max_bad)].copy()


# ---- the reference's own Cython module, rebuilt (oracle/_ref) -------------------------------
_ref_mod = None


def ref_rans():
    """The reference's rans module (encode/decode over Python lists), or None if not built."""
    global _ref_mod
    if _ref_mod is None:
        path = _build.build_ref()
        if path is None or not os.path.exists(path):
            return None
        spec = importlib.util.spec_from_file_location("rans", path)
        mod = importlib.util.module_from_spec(spec)
        saved = sys.modules.get("rans")
        try:
            spec.loader.exec_module(mod)
        finally:
            if saved is not None:
                sys.modules["rans"] = saved
        _ref_mod = mod
    return _ref_mod


# ---- elementwise pieces of the flow (numpy restatements) -------------------------------------

def round_nbits(x: np.ndarray, nbits: int = 8) -> np.ndarray:
    """roundlib.py:18-38: Round(x) = rint(x * 2^nbits) / 2^nbits in float32, ties to even."""
    bins = np.float32(2 ** nbits)
    return (np.rint(np.asarray(x, np.float32) * bins) / bins).astype(np.float32)


def couple_forward(xb: np.ndarray, t: np.ndarray, nbits: int = 8) -> np.ndarray:
    """couplelib.py:49-51: zb = xb + Round(dense(xa)); t is dense(xa)."""
    return (np.asarray(xb, np.float32) + round_nbits(t, nbits)).astype(np.float32)


def couple_backward(zb: np.ndarray, t: np.ndarray, nbits: int = 8) -> np.ndarray:
    """couplelib.py:58-59: xb = zb - Round(dense(za))."""
    return (np.asarray(zb, np.float32) - round_nbits(t, nbits)).astype(np.float32)


def quantise_input_u8(img_u8: np.ndarray) -> np.ndarray:
    """trainer.py:61,72: ToTensor (k/255, float32) then Round(nbits=8)."""
    v = np.asarray(img_u8, np.float32) / np.float32(255.0)
    return round_nbits(v, 8)
