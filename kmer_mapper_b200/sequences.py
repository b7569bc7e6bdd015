"""Ragged byte sequences: the hand-off format between the reader and the kernels.

The reference gets ``chunk.sequence`` from bionumpy as an EncodedRaggedArray (flat bytes + row
lengths; command_line_interface.py:71,110).  Here it is a pair of flat arrays, which is also what
the C ABI takes: ``bases uint8[B]`` (ASCII, reads back to back, no separators) and
``offsets int64[R+1]``.
"""
from __future__ import annotations

import numpy as np


class RaggedSequence:
    def __init__(self, bases, offsets):
        self.bases = bases
        self.offsets = offsets

    def __len__(self):  # number of reads, like len(chunk_sequence) at command_line_interface.py:38
        return len(self.offsets) - 1

    @property
    def lengths(self):
        return np.diff(np.asarray(self.offsets))

    def ravel(self):
        return self.bases

    def __getitem__(self, i):
        return self.bases[int(self.offsets[i]):int(self.offsets[i + 1])]

    @classmethod
    def from_strings(cls, reads) -> "RaggedSequence":
        bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
        bases = np.frombuffer(b"".join(bs), dtype=np.uint8).copy()
        offsets = np.zeros(len(bs) + 1, dtype=np.int64)
        np.cumsum([len(b) for b in bs], out=offsets[1:])
        return cls(bases, offsets)


def as_ragged(chunk_sequence) -> RaggedSequence:
    """Accepts a RaggedSequence, a ``(bases, offsets)`` pair, one read (str/bytes/1-d uint8 array), a
    list of reads, or a bionumpy-style ragged array exposing ``.ravel()`` and row lengths."""
    if isinstance(chunk_sequence, RaggedSequence):
        return chunk_sequence
    if isinstance(chunk_sequence, tuple) and len(chunk_sequence) == 2:
        return RaggedSequence(chunk_sequence[0], chunk_sequence[1])
    if isinstance(chunk_sequence, (str, bytes, bytearray)):
        return RaggedSequence.from_strings([chunk_sequence])
    if isinstance(chunk_sequence, (list,)):
        return RaggedSequence.from_strings(chunk_sequence)
    if isinstance(chunk_sequence, np.ndarray) and chunk_sequence.ndim == 1:
        b = np.ascontiguousarray(chunk_sequence).view(np.uint8) if chunk_sequence.dtype.itemsize == 1 else None
        if b is None:
            raise ValueError("a single read must be a 1-d byte array")
        return RaggedSequence(b, np.array([0, b.shape[0]], dtype=np.int64))
    # bionumpy EncodedRaggedArray / npstructures RaggedArray: flat data + row lengths
    if hasattr(chunk_sequence, "ravel") and hasattr(chunk_sequence, "shape"):
        flat = chunk_sequence.ravel()
        flat = flat.raw() if hasattr(flat, "raw") else flat
        lengths = getattr(chunk_sequence, "lengths", None)
        if lengths is None:
            shape = chunk_sequence.shape
            lengths = shape[-1] if isinstance(shape, tuple) and not np.isscalar(shape[-1]) else None
        if lengths is None:
            raise TypeError("cannot find the row lengths of %r" % type(chunk_sequence))
        offsets = np.zeros(len(lengths) + 1, dtype=np.int64)
        np.cumsum(np.asarray(lengths), out=offsets[1:])
        return RaggedSequence(np.ascontiguousarray(np.asarray(flat)).view(np.uint8), offsets)
    raise TypeError("unsupported chunk sequence type %r" % type(chunk_sequence))
