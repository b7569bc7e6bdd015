"""Chunked FASTA/FASTQ(.gz) reader: file bytes -> flat bases + read offsets in pinned host memory.

Replaces what the reference gets from bionumpy at command_line_interface.py:102-103,109-111
(``bnp.open(path).read_chunks(min_chunk_size)`` -> ``chunk.sequence``): read at least
``min_chunk_size`` bytes, cut at the last complete record, carry the remainder, strip
headers / '+' lines / qualities / newlines, and hand over a ragged byte array.  Formats by suffix
(Readme.md:11, util.py:78-98): .fa/.fasta/.fna (multi-line allowed), .fq/.fastq (4-line records),
each optionally .gz.  PARITY UNPINNED against bionumpy (absent dependency); tests cross-check
against an independent pure-Python parser.

Parsing is vectorised numpy over the raw byte buffer (newline scan + range masks); the output of
each chunk lands in one of two alternating pinned buffers so that the asynchronous host-to-device
copy of chunk i (copy stream, C ABI) overlaps the parsing of chunk i+1.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os

import numpy as np

from . import _lib
from .sequences import RaggedSequence

FASTA_SUFFIXES = (".fa", ".fasta", ".fna", ".fsa")
FASTQ_SUFFIXES = (".fq", ".fastq")


class ReadChunk:
    """What the mapping loop consumes: ``chunk.sequence`` (command_line_interface.py:71,110)."""

    def __init__(self, sequence: RaggedSequence):
        self.sequence = sequence

    def __len__(self):
        return len(self.sequence)


class _PinnedPool:
    """Two alternating pinned host buffers (grow on demand); pageable numpy memory when no GPU
    runtime is available (parsing itself never needs a GPU)."""

    def __init__(self, pinned=True):
        self._bufs = [None, None]
        self._ptrs = [None, None]
        self._i = 0
        self._pinned = bool(pinned) and _lib.device_count() > 0

    def take(self, n_bytes: int) -> np.ndarray:
        i = self._i
        self._i ^= 1
        buf = self._bufs[i]
        if buf is None or buf.shape[0] < n_bytes:
            cap = max(int(n_bytes * 1.25) + 4096, 1 << 16)
            self._release(i)
            if self._pinned:
                p = C.c_void_p()
                _lib.check(_lib.lib().kmb_host_alloc(C.byref(p), cap))
                self._ptrs[i] = p
                buf = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(cap,))
            else:
                buf = np.empty(cap, dtype=np.uint8)
            self._bufs[i] = buf
        return buf

    def _release(self, i):
        self._bufs[i] = None
        if self._ptrs[i] is not None:
            _lib.lib().kmb_host_free(self._ptrs[i])
            self._ptrs[i] = None

    def close(self):
        self._release(0)
        self._release(1)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _format_of(path: str) -> str:
    name = str(path).lower()
    if name.endswith(".gz"):
        name = name[:-3]
    if name.endswith(FASTA_SUFFIXES):
        return "fasta"
    if name.endswith(FASTQ_SUFFIXES):
        return "fastq"
    raise RuntimeError("Unsupported file suffix in %s (expected .fa/.fasta/.fq/.fastq, optionally .gz)" % path)


def _line_bounds(buf: np.ndarray):
    """Start and end (exclusive, without '\\n' and a trailing '\\r') of every line in buf."""
    nl = np.flatnonzero(buf == 10)
    starts = np.empty(nl.shape[0] + 1, dtype=np.int64)
    starts[0] = 0
    starts[1:] = nl + 1
    ends = np.empty_like(starts)
    ends[:-1] = nl
    ends[-1] = buf.shape[0]
    if starts[-1] >= buf.shape[0]:  # buffer ends with a newline: no trailing partial line
        starts, ends = starts[:-1], ends[:-1]
    cr = (ends > starts) & (buf[np.maximum(ends - 1, 0)] == 13)
    ends = ends - cr
    return starts, ends


def _gather_ranges(buf, starts, ends, out):
    """Concatenate buf[starts[i]:ends[i]] into out; returns the number of bytes written."""
    lens = ends - starts
    total = int(lens.sum())
    if total == 0:
        return 0
    delta = np.zeros(buf.shape[0] + 1, dtype=np.int8)
    nz = lens > 0
    np.add.at(delta, starts[nz], 1)
    np.add.at(delta, ends[nz], -1)
    mask = np.cumsum(delta[:-1], dtype=np.int8).view(np.bool_)
    np.compress(mask, buf, out=out[:total])
    return total


def parse_fastq(buf: np.ndarray, pool: _PinnedPool, final: bool):
    """Returns (RaggedSequence, n_bytes_consumed).  Records are 4 lines; an incomplete trailing record
    is left for the next chunk unless ``final``."""
    starts, ends = _line_bounds(buf)
    n_lines = starts.shape[0]
    complete_lines = n_lines if (final or (buf.shape[0] and buf[-1] == 10)) else n_lines - 1
    n_rec = complete_lines // 4
    if final and complete_lines % 4 == 3:
        n_rec += 0  # a record without its quality line is malformed; drop it like a truncated file
    if n_rec == 0:
        return RaggedSequence(np.zeros(0, np.uint8), np.zeros(1, np.int64)), 0
    if not np.all(buf[starts[0:4 * n_rec:4]] == ord("@")):
        bad = int(np.flatnonzero(buf[starts[0:4 * n_rec:4]] != ord("@"))[0])
        raise ValueError("FASTQ record %d does not start with '@'" % bad)
    s = starts[1:4 * n_rec:4]
    e = ends[1:4 * n_rec:4]
    consumed = int(starts[4 * n_rec]) if 4 * n_rec < n_lines else buf.shape[0]
    offsets = np.zeros(n_rec + 1, dtype=np.int64)
    np.cumsum(e - s, out=offsets[1:])
    out = pool.take(int(offsets[-1]) + 16)
    n = _gather_ranges(buf, s, e, out)
    return RaggedSequence(out[:n], offsets), consumed


def parse_fasta(buf: np.ndarray, pool: _PinnedPool, final: bool):
    """Multi-line FASTA.  A record is complete once the next header (or end of file) is seen."""
    starts, ends = _line_bounds(buf)
    n_lines = starts.shape[0]
    if n_lines == 0:
        return RaggedSequence(np.zeros(0, np.uint8), np.zeros(1, np.int64)), 0
    is_header = buf[np.minimum(starts, buf.shape[0] - 1)] == ord(">")
    is_header &= ends > starts - 1
    hdr = np.flatnonzero(is_header)
    if hdr.size == 0:
        if final:
            raise ValueError("FASTA data without a '>' header line")
        return RaggedSequence(np.zeros(0, np.uint8), np.zeros(1, np.int64)), 0
    if final:
        n_rec = hdr.size
        last_line = n_lines
        consumed = buf.shape[0]
    else:
        n_rec = hdr.size - 1            # the last record may continue in the next chunk
        last_line = int(hdr[-1])
        consumed = int(starts[last_line])
    if n_rec <= 0:
        return RaggedSequence(np.zeros(0, np.uint8), np.zeros(1, np.int64)), 0
    first_line = int(hdr[0])
    seq_lines = np.flatnonzero(~is_header[first_line:last_line]) + first_line
    s, e = starts[seq_lines], ends[seq_lines]
    # bases per record: sum of its sequence-line lengths
    rec_of_line = np.searchsorted(hdr[:n_rec + (0 if final else 1)], seq_lines, side="right") - 1
    per_rec = np.bincount(rec_of_line, weights=(e - s), minlength=n_rec).astype(np.int64)[:n_rec]
    offsets = np.zeros(n_rec + 1, dtype=np.int64)
    np.cumsum(per_rec, out=offsets[1:])
    out = pool.take(int(offsets[-1]) + 16)
    n = _gather_ranges(buf, s, e, out)
    return RaggedSequence(out[:n], offsets), consumed


class ReadFile:
    """``bnp.open(path)`` replacement: ``.read_chunks(min_chunk_size)`` yields ReadChunk objects."""

    def __init__(self, path, pinned=True):
        self.path = str(path)
        self.format = _format_of(self.path)
        self._pool = _PinnedPool(pinned)

    def _open(self):
        if self.path.lower().endswith(".gz"):
            return gzip.open(self.path, "rb")  # handles multi-member archives
        return open(self.path, "rb", buffering=0)

    def read_chunks(self, min_chunk_size=5_000_000):
        parse = parse_fastq if self.format == "fastq" else parse_fasta
        carry = b""
        with self._open() as f:
            while True:
                block = f.read(int(min_chunk_size))
                final = len(block) == 0
                data = carry + block if carry else block
                if not data:
                    break
                buf = np.frombuffer(data, dtype=np.uint8)
                seq, consumed = parse(buf, self._pool, final)
                carry = bytes(data[consumed:]) if consumed < len(data) else b""
                if len(seq):
                    yield ReadChunk(seq)
                if final:
                    break

    def close(self):
        self._pool.close()


def open_reads(path, pinned=True) -> ReadFile:
    return ReadFile(path, pinned=pinned)
