"""Wire format of a compressed batch (the reference only *accounts* for the
cost, 64 bits per stream + 32 per word, trainer.py:326-327, and never serialises).

Layout (little endian):
  header   magic 'FLIC' | u16 version | u16 n_levels | u32 n_images | u32 C | u32 H | u32 W |
           u32 codec_batch | u32 streams_per_segment | u64 model_tag | u32 flags (version 2)
  flags bit 0 ("chained"): every image is ONE stream whose state runs through all its levels, as
  coder.Encode chains it (coder.py:18-27), and a chunk has one section instead of one per level.
  then for every chunk (ceil(n_images / codec_batch) of them) and every level, one section:
           u32 n_streams | u64 final_state[n_streams] | u32 n_words[n_streams] | u32 words[sum]
A stream's payload (final_state, words in emission order) is byte-identical to what the
reference's encode() returns for the same slice, so any section can be checked against, or
decoded by, the reference coder.  Overhead relative to the reference's accounting: 32 bits per
stream for the word count and 44 bytes per file.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np
import torch

from .rans import EncodedStreams

MAGIC = b"FLIC"
VERSION = 2          # version 2 adds the flags word; version 1 containers are still read
_HEADER = struct.Struct("<4sHHIIIIIIQ")
_FLAG_CHAINED = 1


@dataclass
class CompressedBatch:
    n_images: int
    shape: tuple            # (C, H, W) of one image
    n_levels: int
    codec_batch: int        # images per network pass; the decoder must use the same (determinism)
    streams_per_segment: int  # streams per (image, level); 0 = one stream per level per chunk
    model_tag: int = 0
    sections: list = field(default_factory=list)  # [chunk][level] -> EncodedStreams ([chunk][0] when chained)
    chained: bool = False
    _trimmed: int = 0       # chunks whose word arrays have been cut to size

    def finalize(self, upto: int | None = None) -> "CompressedBatch":
        """Cut the word arrays of chunks [.., upto] (default: all) down to the words they hold.
        compress() enqueues everything without waiting for the GPU and leaves worst-case arrays
        behind; this is where the word counts are read on the host (one sync per array)."""
        last = len(self.sections) - 1 if upto is None else min(upto, len(self.sections) - 1)
        while self._trimmed <= last:
            for e in self.sections[self._trimmed]:
                n = e.n_words()
                if e.words.numel() != n:
                    e.words = e.words[:n].clone()
            self._trimmed += 1
        return self

    def n_streams(self) -> int:
        return sum(e.n_streams for ch in self.sections for e in ch)

    def n_words(self) -> int:
        return sum(e.n_words() for ch in self.sections for e in ch)

    def reference_bits(self) -> int:
        """Cost by the reference's accounting (trainer.py:326-327)."""
        return 64 * self.n_streams() + 32 * self.n_words()

    def bits_per_dim(self) -> float:
        C, H, W = self.shape
        return self.reference_bits() / (self.n_images * C * H * W)

    def to_bytes(self) -> bytes:
        C, H, W = self.shape
        self.finalize()
        out = [_HEADER.pack(MAGIC, VERSION, self.n_levels, self.n_images, C, H, W, self.codec_batch,
                            self.streams_per_segment, self.model_tag),
               struct.pack("<I", _FLAG_CHAINED if self.chained else 0)]
        for chunk in self.sections:
            for e in chunk:
                woff = e.word_offsets.cpu().numpy()
                nw = int(woff[-1]) if woff.size else 0
                out.append(struct.pack("<I", e.n_streams))
                out.append(e.final_states.cpu().numpy().view(np.uint64).astype("<u8").tobytes())
                out.append(np.diff(woff).astype("<u4").tobytes())
                out.append(e.words[:nw].cpu().numpy().view(np.uint32).astype("<u4").tobytes())
        return b"".join(out)

    @staticmethod
    def from_bytes(blob: bytes, device="cuda") -> "CompressedBatch":
        if len(blob) < _HEADER.size:
            raise ValueError("truncated container")
        magic, ver, n_levels, n_images, C, H, W, codec_batch, sps, tag = _HEADER.unpack_from(blob, 0)
        if magic != MAGIC or ver not in (1, 2):
            raise ValueError("not a FLIC container")
        pos = _HEADER.size
        flags = 0
        if ver >= 2:
            if pos + 4 > len(blob):
                raise ValueError("truncated container")
            (flags,) = struct.unpack_from("<I", blob, pos)
            pos += 4
        cb = CompressedBatch(n_images, (C, H, W), n_levels, codec_batch, sps, tag, chained=bool(flags & _FLAG_CHAINED))
        n_chunks = (n_images + codec_batch - 1) // codec_batch if codec_batch else 0
        mv = memoryview(blob)
        for _ in range(n_chunks):
            chunk = []
            for _ in range(1 if cb.chained else n_levels):
                if pos + 4 > len(blob):
                    raise ValueError("truncated container")
                (ns,) = struct.unpack_from("<I", blob, pos)
                pos += 4
                if pos + 12 * ns > len(blob):
                    raise ValueError("truncated container")
                states = np.frombuffer(mv, "<u8", ns, pos).astype(np.uint64)
                pos += 8 * ns
                counts = np.frombuffer(mv, "<u4", ns, pos).astype(np.int64)
                pos += 4 * ns
                woff = np.zeros(ns + 1, np.int64)
                np.cumsum(counts, out=woff[1:])
                nw = int(woff[-1])
                if pos + 4 * nw > len(blob):
                    raise ValueError("truncated container")
                words = np.frombuffer(mv, "<u4", nw, pos).astype(np.uint32)
                pos += 4 * nw
                chunk.append(EncodedStreams(
                    torch.from_numpy(words.view(np.int32).copy()).to(device),
                    torch.from_numpy(woff).to(device),
                    torch.from_numpy(states.view(np.int64).copy()).to(device),
                    torch.zeros(ns, dtype=torch.int32, device=device), -1))
            cb.sections.append(chunk)
        if pos != len(blob):
            raise ValueError("trailing bytes in container")
        cb._trimmed = len(cb.sections)
        return cb
