"""Chunked FASTA/FASTQ(.gz) reader: file bytes -> flat bases + read offsets in pinned host memory.

Replaces what the reference gets from bionumpy at command_line_interface.py:102-103,109-111
(``bnp.open(path).read_chunks(min_chunk_size)`` -> ``chunk.sequence``): read at least
``min_chunk_size`` bytes, cut at the last complete record, carry the remainder, strip
headers / '+' lines / qualities / newlines, and hand over a ragged byte array.  Formats by suffix
(Readme.md:11, util.py:78-98): .fa/.fasta/.fna (multi-line allowed), .fq/.fastq (4-line records),
each optionally .gz.  PARITY UNPINNED against bionumpy (absent dependency); tests cross-check
against an independent pure-Python parser.

Three overlapped stages: a background thread reads (and inflates) the next block of text while the
native multi-threaded parser (``kmb_parse_reads``, csrc/kmb_reader.cpp) turns the current one into
bases + offsets written straight into one of two alternating pinned buffers, while the previous
chunk's asynchronous host-to-device copy and kernels run (copy stream, C ABI).
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import queue
import threading

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
            cap = (max(int(n_bytes * 1.25) + 4096, 1 << 16) + 63) & ~63
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


class ReadFile:
    """``bnp.open(path)`` replacement: ``.read_chunks(min_chunk_size)`` yields ReadChunk objects."""

    def __init__(self, path, pinned=True, n_threads=None):
        self.path = str(path)
        self.format = _format_of(self.path)
        self._bases_pool = _PinnedPool(pinned)
        self._offsets_pool = _PinnedPool(pinned)
        self.n_threads = int(n_threads or min(os.cpu_count() or 1, 16))

    def _open(self):
        if self.path.lower().endswith(".gz"):
            return gzip.open(self.path, "rb")  # handles multi-member archives
        return open(self.path, "rb", buffering=0)

    def _blocks(self, block_bytes):
        """Background thread: read (and inflate) the next block while the current one is parsed and mapped."""
        q = queue.Queue(maxsize=2)

        def produce():
            try:
                with self._open() as f:
                    while True:
                        block = f.read(block_bytes)
                        q.put(block)
                        if not block:
                            return
            except BaseException as e:  # surfaced in the consumer
                q.put(e)

        t = threading.Thread(target=produce, daemon=True)
        t.start()
        while True:
            item = q.get()
            if isinstance(item, BaseException):
                raise item
            yield item
            if not item:
                return

    def _parse(self, text: np.ndarray, final: bool):
        """One call of the native parser; returns (RaggedSequence, bytes consumed)."""
        lib = _lib.lib()
        n_text = int(text.shape[0])
        fmt = 1 if self.format == "fastq" else 0
        n_reads, n_bases, consumed = C.c_uint64(), C.c_uint64(), C.c_uint64()
        rc = lib.kmb_parse_reads(text.ctypes.data, n_text, fmt, int(final), self.n_threads, None, 0, None, 0,
                                 C.byref(n_reads), C.byref(n_bases), C.byref(consumed))
        if rc != _lib.KMB_OK:
            raise ValueError("%s: malformed %s record" % (self.path, self.format.upper()))
        if n_reads.value == 0:
            return RaggedSequence(np.zeros(0, np.uint8), np.zeros(1, np.int64)), int(consumed.value)
        bases = self._bases_pool.take(n_bases.value + 16)
        offsets = self._offsets_pool.take(8 * (n_reads.value + 1)).view(np.int64)
        rc = lib.kmb_parse_reads(text.ctypes.data, n_text, fmt, int(final), self.n_threads, bases.ctypes.data,
                                 bases.shape[0], offsets.ctypes.data, offsets.shape[0], C.byref(n_reads),
                                 C.byref(n_bases), C.byref(consumed))
        _lib.check(rc)
        return RaggedSequence(bases[:n_bases.value], offsets[:n_reads.value + 1]), int(consumed.value)

    def read_chunks(self, min_chunk_size=5_000_000):
        carry = b""
        seen_data = False
        for block in self._blocks(int(min_chunk_size)):
            final = len(block) == 0
            data = carry + block if carry else block
            if not data:
                break
            seen_data = True
            text = np.frombuffer(data, dtype=np.uint8)
            seq, consumed = self._parse(text, final)
            if final and consumed < len(data) and len(seq) == 0 and self.format == "fasta":
                raise ValueError("%s: FASTA data without a '>' header line" % self.path)
            carry = bytes(data[consumed:]) if consumed < len(data) else b""
            if len(seq):
                yield ReadChunk(seq)
            if final:
                break
        del seen_data

    def close(self):
        self._bases_pool.close()
        self._offsets_pool.close()


def open_reads(path, pinned=True, n_threads=None) -> ReadFile:
    return ReadFile(path, pinned=pinned, n_threads=n_threads)
