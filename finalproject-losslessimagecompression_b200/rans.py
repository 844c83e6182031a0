"""The reference's `rans` module, on the GPU.

Drop-in surface (reference: rans/rans.pyx, call sites trainer.py:315,317, coder.py:25,36,
rans/test.py:16,22):

    encode(state, n, x_, mean_, scale_) -> (state, buffer)
    decode(state, buffer_, n, mean_, scale_) -> (state, message)

Same argument names, order and meaning, Python lists in and out, decode's inputs REVERSED by
the caller exactly as with the reference.  The work is done by libflic_b200.so on cuda:<current
device>; there is no CPU path -- if the library or a GPU is missing these functions raise.

Tensor surface (what the flow's compress/decompress and the benchmark use):

    cdf_tables(x, mean, scale) -> (start, freq)
    encode_streams(x, mean, scale, stream_offsets, ...) -> EncodedStreams
    decode_streams(enc, mean, scale, stream_offsets, ...) -> x

Differences from the reference, all of them places where it misbehaves:
scale == 0 raises ZeroDivisionError as there; a symbol outside its 2048-bin window or off the
1/256 grid raises ValueError instead of silently producing an undecodable stream; a truncated
buffer raises ValueError instead of reading out of bounds.
"""
from __future__ import annotations

import array as _array
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

INITIAL_STATE = 1 << 32  # trainer.py:310


# --------------------------------------------------------------------------------------------
# tensor API (device pointers, current torch stream, no synchronisation)
# --------------------------------------------------------------------------------------------

def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (this path has no CPU implementation)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous().view(-1)


def _offsets(stream_offsets, n_symbols: int, device: torch.device, validate: bool = True) -> torch.Tensor:
    """int64[n_streams+1] on `device`; None means one stream over everything.

    The kernels index the symbol arrays with these numbers, so a partition that does not start at
    0, runs backwards or ends beyond the arrays reads and writes out of bounds.  Offsets that
    arrive on the host are checked for free; CUDA tensors cost one synchronising reduction, which
    `validate=False` skips for callers that build the partition themselves (the flow does)."""
    if stream_offsets is None:
        return torch.tensor([0, n_symbols], dtype=torch.int64, device=device)
    off = torch.as_tensor(stream_offsets, dtype=torch.int64)
    if off.dim() != 1 or off.numel() < 1:
        raise ValueError("stream_offsets must be 1-D with at least one entry")
    if validate:
        bad = (off[0] != 0) | (off[-1] > n_symbols)
        if off.numel() > 1:
            bad = bad | (off[1:] < off[:-1]).any()
        if bool(bad):
            raise ValueError("stream_offsets must start at 0, be non-decreasing and end within the symbol arrays")
    return off.to(device).contiguous()


def uniform_offsets(n_streams: int, stream_len: int, device=None) -> torch.Tensor:
    """Offsets of n_streams equal-length streams (e.g. one per image of a level)."""
    return torch.arange(0, (n_streams + 1) * stream_len, stream_len, dtype=torch.int64, device=device) \
        if stream_len > 0 else torch.zeros(n_streams + 1, dtype=torch.int64, device=device)


def cdf_tables(x: torch.Tensor, mean: torch.Tensor, scale: torch.Tensor):
    """Per-symbol (start, freq) of encode pass 1 (rans/rans.pyx:49-56), as int64-free uint32 bit
    patterns in int32 tensors... returned as torch.int64 for convenience of comparison.

    Returns (start, freq, status_word) where status_word is a 1-element int32 CUDA tensor."""
    xv, mv, sv = _f32c(x, "x"), _f32c(mean, "mean"), _f32c(scale, "scale")
    n = xv.numel()
    if mv.numel() != n or sv.numel() != n:
        raise ValueError("x, mean, scale must have the same number of elements")
    dev = xv.device
    start = torch.empty(n, dtype=torch.int32, device=dev)
    freq = torch.empty(n, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().flic_cdf_tables(xv.data_ptr(), mv.data_ptr(), sv.data_ptr(), n, start.data_ptr(),
                                              freq.data_ptr(), status.data_ptr(), _stream_ptr(dev)), "flic_cdf_tables")
    return start, freq, status


@dataclass
class EncodedStreams:
    """Result of encode_streams: for stream s the reference's (state, buffer) pair is
    (final_states[s], words[word_offsets[s]:word_offsets[s+1]]) -- uint64 / uint32 bit patterns
    held in int64 / int32 CUDA tensors."""
    words: torch.Tensor         # int32 [capacity >= total words]; only [:n_words] is meaningful
    word_offsets: torch.Tensor  # int64 [n_streams + 1]
    final_states: torch.Tensor  # int64 [n_streams]
    status: torch.Tensor        # int32 [n_streams]
    n_symbols: int

    @property
    def n_streams(self) -> int:
        return self.final_states.numel()

    def n_words(self) -> int:
        """Total words (synchronises)."""
        return int(self.word_offsets[-1].item())

    def bits(self) -> int:
        """Cost the way the reference accounts for it: 64 per stream + 32 per word (trainer.py:326-327)."""
        return 64 * self.n_streams + 32 * self.n_words()

    def check(self) -> "EncodedStreams":
        """Raise if any stream is invalid (synchronises)."""
        bits = int(torch.bitwise_or(self.status, 0).max().item()) if self.status.numel() else 0
        if bits:
            allbits = 0
            for b in (1, 2, 4, 8, 16, 32, 64):
                if bool((self.status & b).any().item()):
                    allbits |= b
            _lib.raise_for_status(allbits)
        return self

    def record_stream(self, stream) -> "EncodedStreams":
        """The tensors were produced on a side CUDA stream and will be used on `stream`."""
        for t in (self.words, self.word_offsets, self.final_states, self.status):
            t.record_stream(stream)
        return self

    def trimmed(self) -> "EncodedStreams":
        n = self.n_words()
        return EncodedStreams(self.words[:n].clone(), self.word_offsets, self.final_states, self.status, self.n_symbols)


class Workspace:
    """Reusable device scratch for encode_streams (worst-case word scratch, counts, scan temporaries,
    and the packed output).  Grows on demand; one per device and stream of use."""

    def __init__(self):
        self.buf = None
        self.packed = None

    def get(self, n_symbols: int, n_streams: int, device):
        need = int(_lib.lib().flic_encode_workspace_bytes(n_symbols, n_streams))
        if self.buf is None or self.buf.numel() < need or self.buf.device != device:
            self.buf = torch.empty(need, dtype=torch.uint8, device=device)
        if self.packed is None or self.packed.numel() < max(n_symbols, 1) or self.packed.device != device:
            self.packed = torch.empty(max(n_symbols, 1), dtype=torch.int32, device=device)
        return self.buf, self.packed


_default_ws: dict = {}


def encode_streams(x, mean, scale, stream_offsets=None, init_states=None, workspace: Workspace | None = None,
                   own_output: bool = True, validate: bool = True) -> EncodedStreams:
    """rANS-encode every stream [stream_offsets[s], stream_offsets[s+1]) of the flat symbol arrays.

    x, mean, scale: float32 CUDA tensors of equal numel (any shape; flattened row-major, which is
    the reference's `.reshape(-1)` order, trainer.py:311-313).  scale is exp(logscale).
    Asynchronous on the current stream.  own_output=False returns views into the workspace
    (valid until its next use) and skips one copy.
    """
    xv, mv, sv = _f32c(x, "x"), _f32c(mean, "mean"), _f32c(scale, "scale")
    n = xv.numel()
    if mv.numel() != n or sv.numel() != n:
        raise ValueError("x, mean, scale must have the same number of elements")
    dev = xv.device
    off = _offsets(stream_offsets, n, dev, validate)
    ns = off.numel() - 1
    ws = workspace or _default_ws.setdefault((dev.index, _stream_ptr(dev)), Workspace())
    buf, packed = ws.get(n, ns, dev)
    word_offsets = torch.empty(ns + 1, dtype=torch.int64, device=dev)
    states = torch.empty(ns, dtype=torch.int64, device=dev)
    status = torch.empty(ns, dtype=torch.int32, device=dev)
    init_ptr = 0
    if init_states is not None:
        init_states = torch.as_tensor(init_states).to(device=dev, dtype=torch.int64).contiguous()
        if init_states.numel() != ns:
            raise ValueError("init_states must have one entry per stream")
        init_ptr = init_states.data_ptr()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().flic_rans_encode(
            xv.data_ptr(), mv.data_ptr(), sv.data_ptr(), off.data_ptr(), ns, n, init_ptr, buf.data_ptr(),
            buf.numel(), packed.data_ptr(), packed.numel(), word_offsets.data_ptr(), states.data_ptr(),
            status.data_ptr(), _stream_ptr(dev)), "flic_rans_encode")
    words = packed
    enc = EncodedStreams(words, word_offsets, states, status, n)
    return enc.trimmed() if own_output else enc


def decode_streams(enc: EncodedStreams, mean, scale, stream_offsets=None, check_end: bool = True,
                   out: torch.Tensor | None = None, validate: bool = True, states: torch.Tensor | None = None,
                   words_left: torch.Tensor | None = None, return_words_left: bool = False):
    """Inverse of encode_streams.  Returns (x float32[n_symbols], end_states int64[n_streams],
    status int32[n_streams]); x is in forward order.  Asynchronous on the current stream.

    Continuation (coder.py:29-38 chains one state through the levels): `states` replaces
    enc.final_states as the states to start from, `words_left` (int64 per stream) says how many
    of each stream's words are still unread, and with return_words_left the unread counts after
    this call are returned as a fourth value.  A chained stream is decoded level by level, last
    level first, feeding each call's (end_states, words_left) into the next."""
    mv, sv = _f32c(mean, "mean"), _f32c(scale, "scale")
    n = mv.numel()
    if sv.numel() != n:
        raise ValueError("mean and scale must have the same number of elements")
    dev = mv.device
    off = _offsets(stream_offsets, n, dev, validate)
    ns = off.numel() - 1
    if enc.final_states.numel() != ns or enc.word_offsets.numel() != ns + 1:
        raise ValueError("encoded stream count does not match stream_offsets")
    x_out = out if out is not None else torch.empty(n, dtype=torch.float32, device=dev)
    if x_out.numel() != n or x_out.dtype != torch.float32 or not x_out.is_contiguous():
        raise ValueError("out must be a contiguous float32 tensor of n_symbols elements")
    end_states = torch.empty(ns, dtype=torch.int64, device=dev)
    status = torch.empty(ns, dtype=torch.int32, device=dev)
    words = enc.words.to(dev)
    word_offsets = enc.word_offsets.to(dev)
    if validate and ns > 0 and int(word_offsets[-1].item()) > words.numel():
        raise ValueError("word_offsets end beyond the word array")
    st_in = enc.final_states if states is None else states
    st_in = st_in.to(device=dev, dtype=torch.int64).contiguous()
    if st_in.numel() != ns:
        raise ValueError("states must have one entry per stream")
    left_in_ptr = 0
    if words_left is not None:
        words_left = words_left.to(device=dev, dtype=torch.int64).contiguous()
        if words_left.numel() != ns:
            raise ValueError("words_left must have one entry per stream")
        left_in_ptr = words_left.data_ptr()
    left_out = torch.empty(ns, dtype=torch.int64, device=dev) if return_words_left else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().flic_rans_decode_resume(
            words.data_ptr(), word_offsets.data_ptr(), st_in.data_ptr(), left_in_ptr,
            mv.data_ptr(), sv.data_ptr(), off.data_ptr(), ns, x_out.data_ptr(), end_states.data_ptr(),
            left_out.data_ptr() if left_out is not None else 0, status.data_ptr(), int(bool(check_end)),
            _stream_ptr(dev)), "flic_rans_decode_resume")
    if return_words_left:
        return x_out, end_states, status, left_out
    return x_out, end_states, status


def chain_levels(levels: list) -> EncodedStreams:
    """Concatenate, stream by stream, the words of EncodedStreams that were encoded one after the
    other with the state carried from each into the next (init_states = previous final_states):
    the result is what the reference's chained coder.Encode (coder.py:18-27) produces per stream --
    the last level's state and every level's words in emission order.  Asynchronous; no host sync."""
    if not levels:
        raise ValueError("no levels")
    dev = levels[0].words.device
    ns = levels[0].n_streams
    counts = [e.word_offsets[1:] - e.word_offsets[:-1] for e in levels]
    total = counts[0].clone()
    for c in counts[1:]:
        if c.numel() != ns:
            raise ValueError("levels must have the same streams")
        total = total + c
    word_offsets = torch.zeros(ns + 1, dtype=torch.int64, device=dev)
    torch.cumsum(total, 0, out=word_offsets[1:])
    capacity = sum(max(e.n_symbols, 0) for e in levels)
    if any(e.n_symbols < 0 for e in levels):       # streams that came out of a container: exact sizes are known
        capacity = sum(int(e.word_offsets[-1].item()) for e in levels)
    words = torch.empty(max(capacity, 1), dtype=torch.int32, device=dev)
    status = levels[0].status.clone()
    for e in levels[1:]:
        status |= e.status
    start = word_offsets[:-1].clone()
    with torch.cuda.device(dev):
        for e, c in zip(levels, counts):
            _lib.check(_lib.lib().flic_gather_words(e.words.data_ptr(), e.word_offsets.data_ptr(), start.data_ptr(), ns,
                                                    words.data_ptr(), words.numel(), status.data_ptr(), _stream_ptr(dev)),
                       "flic_gather_words")
            start = start + c
    return EncodedStreams(words, word_offsets, levels[-1].final_states, status, sum(max(e.n_symbols, 0) for e in levels))


def check_status(status: torch.Tensor) -> None:
    """Raise for any non-zero stream status (synchronises)."""
    if status.numel() == 0:
        return
    if bool((status != 0).any().item()):
        allbits = 0
        for b in (1, 2, 4, 8, 16, 32, 64):
            if bool((status & b).any().item()):
                allbits |= b
        _lib.raise_for_status(allbits)


# --------------------------------------------------------------------------------------------
# host codec (C-ABI host entry points: copies inside) -- what a non-torch caller binds
# --------------------------------------------------------------------------------------------

class HostCodec:
    """Owns a flic_codec (device workspace + streams) for host-buffer encode/decode."""

    def __init__(self, max_symbols: int, max_streams: int = 1, device: int | None = None):
        if not torch.cuda.is_available():
            raise _lib.FlicError("no CUDA device: this path has no CPU implementation")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.max_symbols, self.max_streams = int(max_symbols), int(max_streams)
        h = C.c_void_p()
        _lib.check(_lib.lib().flic_codec_create(self.device, self.max_symbols, self.max_streams, C.byref(h)),
                   "flic_codec_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().flic_codec_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _np(a, dtype):
        return np.ascontiguousarray(np.asarray(a, dtype=dtype))

    def encode(self, x, mean, scale, stream_offsets):
        """numpy in / numpy out: (words uint32, word_offsets int64, states uint64, status int32)."""
        x, mean, scale = self._np(x, np.float32), self._np(mean, np.float32), self._np(scale, np.float32)
        off = self._np(stream_offsets, np.int64)
        ns = off.size - 1
        words = np.empty(max(x.size, 1), np.uint32)
        woff = np.zeros(ns + 1, np.int64)
        states = np.zeros(max(ns, 1), np.uint64)
        status = np.zeros(max(ns, 1), np.int32)
        nw = C.c_int64(0)
        _lib.check(_lib.lib().flic_codec_encode(self._h, x.ctypes.data, mean.ctypes.data, scale.ctypes.data,
                                                off.ctypes.data, ns, words.ctypes.data, words.size, woff.ctypes.data,
                                                states.ctypes.data, status.ctypes.data, C.byref(nw)),
                   "flic_codec_encode")
        return words[: nw.value], woff, states[:ns], status[:ns]

    def decode(self, words, word_offsets, states, mean, scale, stream_offsets):
        """numpy in / numpy out: (x float32 forward order, end_states uint64, status int32)."""
        words = self._np(words, np.uint32)
        woff, off = self._np(word_offsets, np.int64), self._np(stream_offsets, np.int64)
        states = self._np(states, np.uint64)
        mean, scale = self._np(mean, np.float32), self._np(scale, np.float32)
        ns = off.size - 1
        x = np.empty(max(mean.size, 1), np.float32)
        end = np.zeros(max(ns, 1), np.uint64)
        status = np.zeros(max(ns, 1), np.int32)
        _lib.check(_lib.lib().flic_codec_decode(self._h, words.ctypes.data, woff.ctypes.data, states.ctypes.data,
                                                mean.ctypes.data, scale.ctypes.data, off.ctypes.data, ns,
                                                x.ctypes.data, end.ctypes.data, status.ctypes.data),
                   "flic_codec_decode")
        return x[: mean.size], end[:ns], status[:ns]

    def encode_single(self, state: int, n: int, x, mean, scale):
        x, mean, scale = self._np(x, np.float32), self._np(mean, np.float32), self._np(scale, np.float32)
        buf = np.empty(max(n, 1), np.uint32)
        nw, st, status = C.c_int64(0), C.c_uint64(0), C.c_int32(0)
        rc = _lib.lib().flic_rans_encode_single(self._h, C.c_uint64(state), n, x.ctypes.data, mean.ctypes.data,
                                                scale.ctypes.data, buf.ctypes.data, C.byref(nw), C.byref(st),
                                                C.byref(status))
        if rc == _lib.E_STATUS:
            _lib.raise_for_status(status.value)
        _lib.check(rc, "flic_rans_encode_single")
        return int(st.value), buf[: nw.value]

    def decode_single(self, state: int, buffer_rev, n: int, mean_rev, scale_rev):
        buf = self._np(buffer_rev, np.uint32)
        mean, scale = self._np(mean_rev, np.float32), self._np(scale_rev, np.float32)
        msg = np.empty(max(n, 1), np.float32)
        st, status = C.c_uint64(0), C.c_int32(0)
        rc = _lib.lib().flic_rans_decode_single(self._h, C.c_uint64(state), buf.ctypes.data, buf.size, n,
                                                mean.ctypes.data, scale.ctypes.data, msg.ctypes.data, C.byref(st),
                                                C.byref(status))
        if rc == _lib.E_STATUS:
            _lib.raise_for_status(status.value)
        _lib.check(rc, "flic_rans_decode_single")
        return int(st.value), msg[:n]


_compat_codec: HostCodec | None = None


def _codec_for(n: int) -> HostCodec:
    global _compat_codec
    c = _compat_codec
    if c is None or c.max_symbols < n or c.device != torch.cuda.current_device():
        if c is not None:
            c.close()
        _compat_codec = c = HostCodec(max(int(n * 1.25), 1 << 16), 1)
    return c


# --------------------------------------------------------------------------------------------
# the reference's two functions
# --------------------------------------------------------------------------------------------

def _f32(v, n):
    """The first n entries of a Python list as float32, the way the reference converts them
    (rans/rans.cpp:2366-2452: PyFloat_AsDouble per item, then a C cast to float; a non-number is a
    TypeError).  array.array does exactly that, twice as fast as numpy's list path."""
    return np.frombuffer(_array.array("f", v if len(v) == n else v[:n]), dtype=np.float32)


def _u32(v):
    # rans/rans.cpp:2561-2640: __Pyx_PyInt_As_unsigned_int per item (OverflowError above 2^32 - 1)
    return np.frombuffer(_array.array("I", v), dtype=np.uint32) if v else np.empty(0, np.uint32)


def _list_arg(v, name):
    # rans/rans.cpp:1585-1587,1991-1993: exact `list` or None, anything else is a TypeError
    if v is None:
        return []
    if type(v) is not list:
        raise TypeError(f"Argument '{name}' has incorrect type (expected list, got {type(v).__name__})")
    return v


def encode(state, n, x_, mean_, scale_):
    """rans.encode (rans/rans.pyx:37-67): returns (state, buffer) with buffer a list of uint32 words
    in emission order.  Uses the first n entries of the lists."""
    x_, mean_, scale_ = _list_arg(x_, "x_"), _list_arg(mean_, "mean_"), _list_arg(scale_, "scale_")
    state, n = int(state), int(n)
    if not 0 <= state < (1 << 64):
        raise OverflowError("can't convert to unsigned long long")  # rans/rans.cpp:1571
    if n <= 0:
        return state, []
    if min(len(x_), len(mean_), len(scale_)) < n:
        raise IndexError("n exceeds the length of the symbol lists")
    x, mean, scale = _f32(x_, n), _f32(mean_, n), _f32(scale_, n)
    st, buf = _codec_for(n).encode_single(state, n, x, mean, scale)
    return st, buf.tolist()


def decode(state, buffer_, n, mean_, scale_):
    """rans.decode (rans/rans.pyx:69-110): buffer_, mean_, scale_ REVERSED by the caller
    (trainer.py:317); returns (state, message) with message reversed, floats s/256."""
    buffer_, mean_, scale_ = _list_arg(buffer_, "buffer_"), _list_arg(mean_, "mean_"), _list_arg(scale_, "scale_")
    state, n = int(state), int(n)
    if not 0 <= state < (1 << 64):
        raise OverflowError("can't convert to unsigned long long")  # rans/rans.cpp:1977
    if n <= 0:
        return state, []
    if min(len(mean_), len(scale_)) < n:
        raise IndexError("n exceeds the length of the parameter lists")
    buf, mean, scale = _u32(buffer_), _f32(mean_, n), _f32(scale_, n)
    st, msg = _codec_for(max(n, buf.size)).decode_single(state, buf, n, mean, scale)
    return st, msg.astype(np.float64).tolist()
