"""Chunked FASTA/FASTQ(.gz) reader: file bytes -> flat bases + read offsets in pinned host memory.

Replaces what the reference gets from bionumpy at command_line_interface.py:102-103,109-111
(``bnp.open(path).read_chunks(min_chunk_size)`` -> ``chunk.sequence``): read at least
``min_chunk_size`` bytes, cut at the last complete record, carry the remainder, strip
headers / '+' lines / qualities / newlines, and hand over a ragged byte array.  Formats by suffix
(Readme.md:11, util.py:78-98): .fa/.fasta/.fna (multi-line allowed), .fq/.fastq (4-line records),
each optionally .gz.  PARITY UNPINNED against bionumpy (absent dependency); tests cross-check
against an independent pure-Python parser.

Plain files are parsed in place from a read-only mapping of the page cache by the native multi-threaded
parser (``kmb_parse_reads``, csrc/kmb_reader.cpp), bases + offsets written straight into one of two
alternating pinned buffers; .gz files are inflated member-parallel by the native library
(``kmb_gunzip_members``) in a background thread, a few blocks ahead of the parser.  The previous chunk's
asynchronous host-to-device copy and kernels run meanwhile (copy stream, C ABI).
"""
from __future__ import annotations

import ctypes as C
import mmap
import os
import queue
import threading

import numpy as np

from . import _lib
from .sequences import RaggedSequence

FASTA_SUFFIXES = (".fa", ".fasta", ".fna", ".fsa")
FASTQ_SUFFIXES = (".fq", ".fastq")


class ReadChunk:
    """What the mapping loop consumes: ``chunk.sequence`` (command_line_interface.py:71,110).

    LIFETIME: ``sequence`` is a view into one of the reader's rotating pinned buffers (N_CHUNK_BUFFERS of them): it
    stays valid while the next N_CHUNK_BUFFERS - 1 chunks are produced and is overwritten after that.  The mapping
    loop consumes a chunk before asking for the next, so it never notices; anything that keeps chunks around
    (``list(read_chunks(...))``, a prefetch queue) must take ``chunk.copy()``."""

    def __init__(self, sequence: RaggedSequence):
        self.sequence = sequence

    def __len__(self):
        return len(self.sequence)

    def copy(self) -> "ReadChunk":
        """An independent chunk in ordinary (pageable) memory."""
        return ReadChunk(RaggedSequence(np.array(self.sequence.bases, copy=True), np.array(self.sequence.offsets, copy=True)))


N_CHUNK_BUFFERS = 4


class TextChunk:
    """A whole-record window of a reads file's raw text: address + length (valid until the next chunk is asked for).
    No numpy view of the file mapping is handed out, so the mapping can be closed even while a chunk object is alive."""

    def __init__(self, ptr: int, n: int, keep=None, fd=None, offset=0):
        self.ptr, self.n, self._keep = int(ptr), int(n), keep
        self.fd, self.offset = fd, int(offset)     # a window of an open plain file: read with pread, never mapped

    def __len__(self):
        return self.n

    def tobytes(self) -> bytes:
        if self.fd is not None:
            return os.pread(self.fd, self.n, self.offset)
        return C.string_at(self.ptr, self.n)


class _PinnedPool:
    """N_CHUNK_BUFFERS rotating pinned host buffers (grow on demand); pageable numpy memory when no GPU
    runtime is available (parsing itself never needs a GPU)."""

    def __init__(self, pinned=True, n=N_CHUNK_BUFFERS):
        self._bufs = [None] * n
        self._ptrs = [None] * n
        self._i = 0
        self._pinned = bool(pinned) and _lib.device_count() > 0

    def take(self, n_bytes: int) -> np.ndarray:
        i = self._i
        self._i = (self._i + 1) % len(self._bufs)
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

    def untake(self):
        """Give back the buffer of the last take() (nothing was handed to a consumer)."""
        self._i = (self._i - 1) % len(self._bufs)

    def _release(self, i):
        self._bufs[i] = None
        if self._ptrs[i] is not None:
            _lib.lib().kmb_host_free(self._ptrs[i])
            self._ptrs[i] = None

    def close(self):
        for i in range(len(self._bufs)):
            self._release(i)

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


class ParallelGzip:
    """Multi-member gzip (bgzip/BGZF files, `cat a.gz b.gz`, the synthetic FASTQ.gz of the benchmark configs)
    inflated member-parallel by the native library (``kmb_gunzip_members``, csrc/kmb_gunzip.cpp: speculative member
    starts, worker pool, chain walk).  A member that inflates to more than ``max_member_bytes`` (a plain
    single-member .gz) makes the reader fall back to sequential streaming from that point, through the native
    single-core decoder (``kmb_gzstream_*``, csrc/kmb_inflate.cpp: ~2x zlib, CRC-32 checked).
    """

    HEADROOM = 1 << 20   # free bytes in front of every block: the consumer puts its carried-over partial record there
    # Capacity of one inflate call.  The blocks handed out end where the next member no longer fits, so their
    # boundaries are a function of this number and the file alone -- NOT of the thread count: the ranks of a
    # multi-GPU job take every world_size-th block, and ranks with different CPU shares must cut the stream at
    # the same places (a capacity that grew with n_threads made two ranks with 2 and 12 threads drop or double reads).
    BATCH_BYTES = 256 << 20
    MAX_MEMBER_BYTES = 64 << 20   # a member that inflates to more than this is a plain single-member .gz: stream it

    def __init__(self, path, n_threads, max_member_bytes=None):
        self.path = path
        self.n_threads = max(1, int(n_threads))
        self.max_member_bytes = int(self.MAX_MEMBER_BYTES if max_member_bytes is None else max_member_bytes)

    def arrays(self, block_bytes, n_buffers=4, start=0):
        """Yields (buffer, n): the next n bytes of text are buffer[HEADROOM : HEADROOM + n], at least block_bytes of
        them except at the end of the file.  Buffers rotate: one stays valid while the next n_buffers - 1 are made.
        ``start``: offset of the member to begin with."""
        size = os.path.getsize(self.path)
        if size == 0:
            return
        lib = _lib.lib()
        cap = max(int(block_bytes) + self.max_member_bytes, self.BATCH_BYTES)   # rank-independent, see BATCH_BYTES
        bufs = [None] * n_buffers
        turn = 0

        def next_buffer():
            nonlocal turn
            if bufs[turn] is None:
                bufs[turn] = np.empty(self.HEADROOM + cap, dtype=np.uint8)
            buf = bufs[turn]
            turn = (turn + 1) % n_buffers
            return buf

        fallback_from = None
        with open(self.path, "rb") as f, mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            whole = np.frombuffer(mm, dtype=np.uint8)
            base_ptr = whole.ctypes.data
            try:
                pos = int(start)
                while pos < size:
                    buf = next_buffer()
                    consumed, produced, flag = C.c_uint64(), C.c_uint64(), C.c_int()
                    rc = lib.kmb_gunzip_members(base_ptr + pos, size - pos, self.n_threads, buf.ctypes.data + self.HEADROOM,
                                                cap, self.max_member_bytes, C.byref(consumed), C.byref(produced), C.byref(flag))
                    if rc == _lib.KMB_ERR_BAD_ARG:
                        raise OSError("%s: not a gzip member at offset %d" % (self.path, pos))
                    _lib.check(rc)
                    pos += consumed.value
                    if produced.value:
                        yield buf, int(produced.value)
                    if flag.value == 1:
                        fallback_from = pos
                        break
                    if flag.value == 2:
                        if bytes(mm[pos:pos + 4096]).strip(b"\0") or bytes(mm[pos:]).strip(b"\0"):
                            raise OSError("%s: corrupt gzip member at offset %d" % (self.path, pos))
                        break            # zero padding after the last member
                    if consumed.value == 0 and flag.value == 0:
                        raise OSError("%s: gzip member at offset %d does not fit the reader's buffer" % (self.path, pos))
                if fallback_from is not None:
                    # One long deflate stream (plain `gzip file`): the native single-core decoder (kmb_inflate.cpp),
                    # block after block; the format may refer back 32 KB, so the tail of each block is copied in
                    # front of the next one (into the headroom; the parser's carry-over lands there later).
                    stream = C.c_void_p()
                    _lib.check(lib.kmb_gzstream_open(base_ptr + fallback_from, size - fallback_from, self.n_threads,
                                                     C.byref(stream)))
                    try:
                        window = np.zeros(0, dtype=np.uint8)
                        finished = C.c_int(0)
                        while not finished.value:
                            buf = next_buffer()
                            buf[self.HEADROOM - window.shape[0]:self.HEADROOM] = window
                            produced = C.c_uint64()
                            rc = lib.kmb_gzstream_read(stream, buf.ctypes.data + self.HEADROOM, cap, window.shape[0],
                                                       C.byref(produced), C.byref(finished))
                            if rc != _lib.KMB_OK:
                                raise OSError("%s: %s" % (self.path, lib.kmb_gzstream_error(stream).decode()))
                            n = int(produced.value)
                            if n:
                                have = self.HEADROOM + n
                                window = buf[max(self.HEADROOM - window.shape[0], have - 32768):have].copy()
                                yield buf, n
                    finally:
                        lib.kmb_gzstream_close(stream)
            finally:
                del whole

    def blocks(self, block_bytes):
        """Yields decompressed text in file order as bytes, then b''."""
        for buf, n in self.arrays(block_bytes):
            yield buf[self.HEADROOM:self.HEADROOM + n].tobytes()
        yield b""


class ReadFile:
    """``bnp.open(path)`` replacement: ``.read_chunks(min_chunk_size)`` yields ReadChunk objects."""

    def __init__(self, path, pinned=True, n_threads=None):
        self.path = str(path)
        self.format = _format_of(self.path)
        self._bases_pool = _PinnedPool(pinned)
        self._offsets_pool = _PinnedPool(pinned)
        self.n_threads = int(n_threads or min(len(os.sched_getaffinity(0)) or 1, 32))
        self._bases_per_byte = 0.0   # densest window seen so far: sizes the next window's output buffers
        self._reads_per_byte = 0.0

    def _gz_arrays(self, block_bytes, start=0):
        """Background thread: inflate the next blocks while the current one is parsed and mapped.  Yields
        (buffer, n) like ParallelGzip.arrays, then None."""
        q = queue.Queue(maxsize=1)

        def produce():
            try:
                for item in ParallelGzip(self.path, self.n_threads).arrays(block_bytes, n_buffers=4, start=start):
                    q.put(item)          # 1 queued + 1 being parsed + 1 being filled < 4 buffers
                q.put(None)
            except BaseException as e:  # surfaced in the consumer
                q.put(e)

        t = threading.Thread(target=produce, daemon=True)
        t.start()
        while True:
            item = q.get()
            if isinstance(item, BaseException):
                raise item
            yield item
            if item is None:
                return

    def _parse(self, text_ptr: int, n_text: int, final: bool, count_only: bool = False):
        """The native parser on one window of text (address + length: no view of the file mapping is created, so
        the mapping can be closed even while an exception's traceback is alive); returns (RaggedSequence, bytes
        consumed).  The output buffers
        are sized from the previous window (bases and reads per byte of text, plus slack), so the usual window
        costs one call; a window that needs more answers KMB_ERR_NOMEM with the exact sizes and is parsed again."""
        lib = _lib.lib()
        fmt = 1 if self.format == "fastq" else 0
        n_reads, n_bases, consumed = C.c_uint64(), C.c_uint64(), C.c_uint64()
        if count_only:      # another rank's chunk: only where it ends matters
            rc = lib.kmb_parse_reads(text_ptr, n_text, fmt, int(final), self.n_threads, None, 0, None, 0,
                                     C.byref(n_reads), C.byref(n_bases), C.byref(consumed))
            if rc == _lib.KMB_ERR_BAD_ARG:
                raise ValueError("%s: malformed %s record" % (self.path, self.format.upper()))
            _lib.check(rc)
            return RaggedSequence(np.zeros(0, np.uint8), np.zeros(1, np.int64)), int(consumed.value)
        want_bases = int(n_text * self._bases_per_byte * 1.05) + 4096
        want_reads = int(n_text * self._reads_per_byte * 1.05) + 64
        for attempt in range(2):
            bases = self._bases_pool.take(want_bases + 16)
            offsets = self._offsets_pool.take(8 * (want_reads + 1)).view(np.int64)
            rc = lib.kmb_parse_reads(text_ptr, n_text, fmt, int(final), self.n_threads, bases.ctypes.data,
                                     bases.shape[0], offsets.ctypes.data, offsets.shape[0], C.byref(n_reads),
                                     C.byref(n_bases), C.byref(consumed))
            if rc == _lib.KMB_ERR_NOMEM and attempt == 0:
                self._bases_pool.untake()
                self._offsets_pool.untake()
                want_bases, want_reads = n_bases.value, n_reads.value
                continue
            break
        if rc == _lib.KMB_ERR_BAD_ARG:
            raise ValueError("%s: malformed %s record" % (self.path, self.format.upper()))
        _lib.check(rc)
        if consumed.value:
            self._bases_per_byte = max(self._bases_per_byte, n_bases.value / consumed.value)
            self._reads_per_byte = max(self._reads_per_byte, n_reads.value / consumed.value)
        if n_reads.value == 0:
            self._bases_pool.untake()
            self._offsets_pool.untake()
            return RaggedSequence(np.zeros(0, np.uint8), np.zeros(1, np.int64)), int(consumed.value)
        return RaggedSequence(bases[:n_bases.value], offsets[:n_reads.value + 1]), int(consumed.value)

    def _shard_start(self, base_ptr, file_size, cut):
        """First record start at or after byte ``cut`` (the file size when there is none)."""
        if cut <= 0:
            return 0
        if cut >= file_size:
            return file_size
        off = C.c_uint64()
        fmt = 1 if self.format == "fastq" else 0
        _lib.check(_lib.lib().kmb_find_record_start(base_ptr + cut - 1, file_size - (cut - 1), fmt, C.byref(off)))
        return cut - 1 + off.value

    def _windows_of_file(self, min_chunk_size, rank=0, world_size=1):
        """Plain (uncompressed) file: the parser reads the page cache through a read-only mapping -- no copy of the
        text into Python objects, no carry-over buffer: the next window simply starts where the last complete
        record ended.  With world_size > 1 this rank takes the records that START inside its 1/world_size of the
        file's bytes, so the ranks of a multi-GPU job parse disjoint parts.  Yields RaggedSequence objects."""
        file_size = os.path.getsize(self.path)
        if file_size == 0:
            return
        with open(self.path, "rb") as f, mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            if hasattr(mm, "madvise"):
                mm.madvise(mmap.MADV_SEQUENTIAL)
            whole = np.frombuffer(mm, dtype=np.uint8)
            base_ptr = whole.ctypes.data
            try:
                pos = self._shard_start(base_ptr, file_size, file_size * rank // world_size)
                size = self._shard_start(base_ptr, file_size, file_size * (rank + 1) // world_size)
                while pos < size:
                    want = int(min_chunk_size)
                    while True:
                        end = min(size, pos + want)
                        final = end == size
                        if not final and hasattr(mm, "madvise"):   # start reading the next window from disk now
                            a = end - end % mmap.PAGESIZE
                            mm.madvise(mmap.MADV_WILLNEED, a, min(want, size - a))
                        seq, consumed = self._parse(base_ptr + pos, end - pos, final)
                        if consumed or final:
                            break
                        want *= 2                                   # one record longer than the window
                    if final and consumed < end - pos and len(seq) == 0 and self.format == "fasta":
                        raise ValueError("%s: FASTA data without a '>' header line" % self.path)
                    pos += consumed if not final else end - pos
                    yield seq
            finally:
                del whole

    # ---- raw text for the device-side parser (kmb_mapper_map_text) ------------------------------------------------
    def _next_record_start(self, ptr, n, at):
        """First record start in text[ptr, ptr+n) at or after byte ``at`` (n when there is none)."""
        if at <= 0:
            return 0
        if at >= n:
            return n
        off = C.c_uint64()
        fmt = 1 if self.format == "fastq" else 0
        _lib.check(_lib.lib().kmb_find_record_start(ptr + at - 1, n - (at - 1), fmt, C.byref(off)))
        return at - 1 + off.value

    def _last_record_start(self, ptr, n):
        """Last record start in text[ptr, ptr+n) (0 when the only one is the beginning): where a block of inflated text
        is cut so that what is handed on holds whole records; the rest is carried over to the next block."""
        window = 1 << 16
        while True:
            lo = max(0, n - window)
            best, p = None, lo
            while True:
                q = self._next_record_start(ptr, n, p + 1) if p or lo else self._next_record_start(ptr, n, 1)
                if q >= n:
                    break
                best, p = q, q
            if best is not None or lo == 0:
                return best or 0
            window *= 8

    def text_chunks(self, min_chunk_size=64 << 20, rank=0, world_size=1, gz_start=0):
        """Whole-record windows of the file's TEXT (TextChunk: address + length, valid until the next one is asked for),
        for the device-side parser.  Sharding as in read_chunks: a contiguous byte range of a plain file per rank, every
        world_size-th block of a .gz.  ``gz_start``: begin at the gzip member at this offset and drop the text in front
        of the first record start after the first newline (what ``Mapper.map_gz`` left to the host decoders)."""
        if not self.path.lower().endswith(".gz"):
            file_size = os.path.getsize(self.path)
            if file_size == 0:
                return
            # The file is never mapped: the library preads every window straight into pinned staging (no page faults,
            # no copy through Python); only a few KB around each nominal cut are read here to find the record start.
            fd = os.open(self.path, os.O_RDONLY)
            try:
                def record_start_at(cut):
                    """First record start at or after byte ``cut`` of the file."""
                    if cut <= 0:
                        return 0
                    window = 1 << 16
                    while cut < file_size:
                        buf = np.frombuffer(os.pread(fd, window, cut - 1), dtype=np.uint8)
                        got = self._next_record_start(buf.ctypes.data, int(buf.shape[0]), 1)
                        if got < buf.shape[0] or cut - 1 + buf.shape[0] >= file_size:
                            return min(file_size, cut - 1 + got)
                        window *= 8          # a record longer than the window
                    return file_size

                pos = record_start_at(file_size * rank // world_size)
                size = record_start_at(file_size * (rank + 1) // world_size)
                while pos < size:
                    end = min(size, pos + int(min_chunk_size))
                    if end < size:
                        end = min(size, record_start_at(end))
                    yield TextChunk(0, end - pos, fd=fd, offset=pos)
                    pos = end
            finally:
                os.close(fd)
            return
        head = ParallelGzip.HEADROOM
        carry = np.zeros(0, dtype=np.uint8)
        skip_partial = gz_start > 0
        for i, item in enumerate(self._gz_arrays(int(min_chunk_size), start=gz_start)):
            final = item is None
            if final:
                if carry.shape[0] and i % world_size == rank:
                    yield TextChunk(carry.ctypes.data, carry.shape[0], carry)
                break
            buf, n = item
            if skip_partial:
                # what map_gz's last batch has already mapped: everything in front of the first record start after the
                # first newline (the same rule on the same bytes: kmb_find_record_start from the member's first byte)
                skip_partial = False
                first = self._next_record_start(buf.ctypes.data + head, n, 1)
                if first >= n and n >= (1 << 20):
                    raise ValueError("%s: a record longer than 1 MB where the host decoders take over" % self.path)
                buf = buf[first:]
                n -= first
                if n == 0:
                    continue
            if carry.shape[0] <= head:
                start = head - carry.shape[0]
                buf[start:head] = carry
                text = buf[start:head + n]
            else:
                text = np.concatenate([carry, buf[head:head + n]])
            cut = self._last_record_start(text.ctypes.data, int(text.shape[0]))
            carry = text[cut:].copy()
            if cut and i % world_size == rank:
                yield TextChunk(text.ctypes.data, cut, text)

    def read_chunks(self, min_chunk_size=5_000_000, rank=0, world_size=1):
        """Chunks of at least ``min_chunk_size`` bytes of the file, cut at record boundaries.  ``rank`` /
        ``world_size``: only this rank's share -- a contiguous byte range of a plain file, every world_size-th chunk
        of a .gz (which has to be inflated from the start by everyone)."""
        if not self.path.lower().endswith(".gz"):
            for seq in self._windows_of_file(min_chunk_size, rank, world_size):
                if len(seq):
                    yield ReadChunk(seq)
            return
        head = ParallelGzip.HEADROOM
        carry = np.zeros(0, dtype=np.uint8)
        for i, item in enumerate(self._gz_arrays(int(min_chunk_size))):
            final = item is None
            if final:
                if not carry.shape[0]:
                    break
                text = carry
            else:
                buf, n = item
                if carry.shape[0] <= head:        # the carried-over partial record goes right in front of the new text
                    start = head - carry.shape[0]
                    buf[start:head] = carry
                    text = buf[start:head + n]
                else:                             # a record longer than the headroom (a chromosome-sized FASTA entry)
                    text = np.concatenate([carry, buf[head:head + n]])
            seq, consumed = self._parse(text.ctypes.data, int(text.shape[0]), final, count_only=i % world_size != rank)
            if final and consumed < text.shape[0] and len(seq) == 0 and self.format == "fasta":
                raise ValueError("%s: FASTA data without a '>' header line" % self.path)
            carry = text[consumed:].copy()
            if len(seq):
                yield ReadChunk(seq)
            if final:
                break

    def close(self):
        self._bases_pool.close()
        self._offsets_pool.close()


def open_reads(path, pinned=True, n_threads=None) -> ReadFile:
    return ReadFile(path, pinned=pinned, n_threads=n_threads)
