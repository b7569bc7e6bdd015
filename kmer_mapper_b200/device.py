"""Python handles over the C ABI: a device-resident index and a mapper (count buffer + streams).

These are the objects the reference-shaped functions in mapper.py / gpu_counter.py /
command_line_interface.py are built from.  Host arrays are numpy, device arrays are torch CUDA
tensors (torch is used only to own device memory and streams).
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import FLAG_NO_N_TO_A, FLAG_REVCOMP, InvalidBaseError, KmbError, as_buffer, check, lib

DEFAULT_MAX_FREQUENCY = 1000  # mapper.pyx:19


def current_device() -> int:
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


def _order_after_torch(*buffers, same_stream=None):
    """The library runs on its own non-blocking streams.  When an input lives in a torch CUDA tensor, the
    torch work that produced it must have finished before our kernels read it -- unless the mapper was
    put on torch's current stream (set_stream), where stream order already guarantees it."""
    for b in buffers:
        if _lib.is_torch_tensor(b) and b.is_cuda:
            import torch
            cur = torch.cuda.current_stream(b.device)
            if same_stream is not None and cur.cuda_stream == same_stream and same_stream != 0:
                return
            cur.synchronize()
            return


def _index_attr(index, name):
    try:
        return getattr(index, name)
    except AttributeError:
        raise AttributeError("index object has no attribute %r (mapper.pyx:22-29 reads _hashes_to_index, _n_kmers, "
                             "_nodes, _kmers, _frequencies, _modulo)" % name)


class DeviceIndex:
    """The six arrays mapper.pyx:22-29 reads, copied to one GPU and re-laid out there
    (bucket directory + packed entries + L2-resident occupancy filter)."""

    def __init__(self, hashes_to_index, n_kmers, nodes, kmers, frequencies, modulo, device=None):
        _lib.require_device()
        self.device = current_device() if device is None else int(device)
        _order_after_torch(hashes_to_index, n_kmers, nodes, kmers, frequencies)
        keep = []
        ptrs = []
        for arr, dt, nm in ((hashes_to_index, np.int32, "hashes_to_index"), (n_kmers, np.int32, "n_kmers"),
                            (nodes, np.int32, "nodes"), (kmers, np.uint64, "kmers"),
                            (frequencies, np.uint16, "frequencies")):
            k, p, n = as_buffer(arr, dt, nm)
            keep.append(k)
            ptrs.append((p, n))
        modulo = int(modulo)
        if ptrs[0][1] != modulo or ptrs[1][1] != modulo:
            raise ValueError("hashes_to_index and n_kmers must have `modulo`=%d elements (got %d, %d)"
                             % (modulo, ptrs[0][1], ptrs[1][1]))
        n_entries = ptrs[2][1]
        if ptrs[3][1] != n_entries or ptrs[4][1] != n_entries:
            raise ValueError("nodes, kmers and frequencies must have the same length")
        h = C.c_void_p()
        check(lib().kmb_index_create(self.device, ptrs[0][0], ptrs[1][0], modulo, ptrs[2][0], ptrs[3][0], ptrs[4][0],
                                     n_entries, C.byref(h)))
        self._h = h
        self._finalizer = weakref.finalize(self, lib().kmb_index_destroy, h)
        mx, ne, mo, db = C.c_int64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib().kmb_index_info(h, C.byref(mx), C.byref(ne), C.byref(mo), C.byref(db)))
        self._max_node_id, self.n_entries, self.modulo, self.device_bytes = mx.value, ne.value, mo.value, db.value
        fb = C.c_uint64()
        check(lib().kmb_index_filter_bytes(h, C.byref(fb)))
        self.filter_bytes = fb.value
        nm, no, nl = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib().kmb_index_layout(h, C.byref(nm), C.byref(no), C.byref(nl)))
        self.n_main_lines, self.n_overflow_lines, self.n_live_entries = nm.value, no.value, nl.value

    @classmethod
    def from_index(cls, index, device=None) -> "DeviceIndex":
        """From any duck-typed index object (mapper.pyx:22-29).  The device copy is cached on the
        object so that per-chunk calls (command_line_interface.py:51) do not re-upload it."""
        if isinstance(index, DeviceIndex):
            return index
        device = current_device() if device is None else int(device)
        cache = getattr(index, "_kmb_device_index", None)
        arrays = tuple(_index_attr(index, a) for a in ("_hashes_to_index", "_n_kmers", "_nodes", "_kmers", "_frequencies"))
        modulo = int(_index_attr(index, "_modulo"))
        # The cache entry keeps references to the arrays it was built from and is valid only while the index still
        # holds those very objects (`is`): an id() alone could be reused by a new array after the old one was freed.
        # In-place edits of an array cannot be seen this way: call DeviceIndex.invalidate(index) after such an edit.
        if cache is not None and cache[0] == (device, modulo) and len(cache[1]) == len(arrays) and \
                all(a is b for a, b in zip(cache[1], arrays)):
            return cache[2]
        di = cls(*arrays, modulo=modulo, device=device)
        try:
            index._kmb_device_index = ((device, modulo), arrays, di)
        except Exception:
            pass
        return di

    @staticmethod
    def invalidate(index) -> None:
        """Drop the device copy cached on an index object (after editing its arrays in place)."""
        if hasattr(index, "_kmb_device_index"):
            try:
                del index._kmb_device_index
            except Exception:
                pass

    def max_node_id(self) -> int:
        """KmerIndex.max_node_id() = nodes.max() (command_line_interface.py:51,79,117)."""
        return max(self._max_node_id, 0)

    def in_graph_index(self, kmers):
        """mapper.pyx:81-130: uint8[n], 1 iff the bucket of kmers[i] holds its key."""
        keep, p, n = as_buffer(kmers, np.uint64, "kmers")
        _order_after_torch(kmers)
        if _lib.is_torch_tensor(kmers):
            import torch
            out = torch.empty(n, dtype=torch.uint8, device=kmers.device)
            check(lib().kmb_in_graph_index(self._h, p, n, out.data_ptr()))
            return out
        out = np.zeros(n, dtype=np.uint8)
        check(lib().kmb_in_graph_index(self._h, p, n, out.ctypes.data))
        return out

    def close(self):
        self._finalizer()


class Mapper:
    """node_counts uint32[n_counts] on the device + streams + staging (kmb_mapper).  Cumulative
    across calls like the reference's additive map-reduce (command_line_interface.py:124-130)."""

    def __init__(self, index: DeviceIndex, n_counts=None, max_index_lookup_frequency=DEFAULT_MAX_FREQUENCY,
                 counts_tensor=None):
        self.index = index
        self.n_counts = index.max_node_id() + 1 if n_counts is None else int(n_counts)
        self._counts_tensor = counts_tensor
        cptr = None
        if counts_tensor is not None:
            _order_after_torch(counts_tensor)
            keep, cptr, n = as_buffer(counts_tensor, np.uint32, "counts_tensor")
            if n != self.n_counts:
                raise ValueError("counts_tensor has %d elements, n_counts is %d" % (n, self.n_counts))
        h = C.c_void_p()
        check(lib().kmb_mapper_create(index._h, self.n_counts, cptr, int(max_index_lookup_frequency), C.byref(h)))
        self._h = h
        self._stream = None
        self._finalizer = weakref.finalize(self, lib().kmb_mapper_destroy, h)

    def set_stream(self, cuda_stream=None):
        """Run on a caller stream (an int cudaStream_t or a torch.cuda.Stream); None = the mapper's own."""
        s = getattr(cuda_stream, "cuda_stream", cuda_stream)
        check(lib().kmb_mapper_set_stream(self._h, s))
        self._stream = s or None

    def map_kmers(self, kmers, revcomp=False, k=31):
        keep, p, n = as_buffer(kmers, np.uint64, "kmers")
        _order_after_torch(kmers, same_stream=self._stream)
        check(lib().kmb_mapper_map_kmers(self._h, p, n, FLAG_REVCOMP if revcomp else 0, int(k)))

    def map_reads(self, bases, offsets, k, revcomp=False, n_to_a=True):
        kb, pb, nb = as_buffer(bases, np.uint8, "bases")
        ko, po, no = as_buffer(offsets, np.int64, "offsets")
        if no < 1:
            raise ValueError("offsets must have n_reads+1 elements")
        flags = (FLAG_REVCOMP if revcomp else 0) | (0 if n_to_a else FLAG_NO_N_TO_A)
        _order_after_torch(bases, offsets, same_stream=self._stream)
        check(lib().kmb_mapper_map_reads(self._h, pb, nb, po, no - 1, int(k), flags))

    def map_text(self, text, fmt, k, revcomp=False, n_to_a=True):
        """One chunk of raw FASTA (``fmt`` 0 / "fasta") or FASTQ (1 / "fastq") text holding whole records: parsed on
        the device and mapped (``kmb_mapper_map_text``).  ``text``: a uint8 numpy array (any host memory), a torch
        uint8 CUDA tensor, bytes, or a reader.TextChunk."""
        if isinstance(text, (bytes, bytearray, memoryview)):
            text = np.frombuffer(text, dtype=np.uint8)
        fmt = {"fasta": 0, "fastq": 1}.get(fmt, fmt)
        flags = (FLAG_REVCOMP if revcomp else 0) | (0 if n_to_a else FLAG_NO_N_TO_A)
        if getattr(text, "fd", None) is not None:            # reader.TextChunk of a plain file: (fd, offset, n)
            check(lib().kmb_mapper_map_text_fd(self._h, text.fd, text.offset, text.n, int(fmt), int(k), flags))
            return
        if hasattr(text, "ptr") and hasattr(text, "n"):      # reader.TextChunk in memory
            kt, pt, nt = text, text.ptr, text.n
        else:
            kt, pt, nt = as_buffer(text, np.uint8, "text")
        _order_after_torch(text, same_stream=self._stream)
        check(lib().kmb_mapper_map_text(self._h, pt, nt, int(fmt), int(k), flags))

    def map_gz(self, gz, fmt, k, revcomp=False, n_to_a=True, shard_index=0, shard_count=1) -> int:
        """A multi-member .gz of FASTA / FASTQ text, from its first byte: inflated, parsed and mapped on the device
        (``kmb_mapper_map_gz``).  ``gz``: a uint8 numpy array (e.g. over a read-only file mapping) or bytes.  Returns the
        offset from which the host decoders have to continue (``len(gz)`` when the device did everything, 0 when it
        did nothing: a plain single-member .gz), skipping the first partial record there unless the offset is 0."""
        if isinstance(gz, (bytes, bytearray, memoryview)):
            gz = np.frombuffer(gz, dtype=np.uint8)
        fmt = {"fasta": 0, "fastq": 1}.get(fmt, fmt)
        flags = (FLAG_REVCOMP if revcomp else 0) | (0 if n_to_a else FLAG_NO_N_TO_A)
        kg, pg, ng = as_buffer(gz, np.uint8, "gz")
        resume = C.c_uint64(0)
        check(lib().kmb_mapper_map_gz(self._h, pg, ng, int(fmt), int(k), flags, int(shard_index), int(shard_count), C.byref(resume)))
        return int(resume.value)

    def flush(self):
        """Queue the slot-counter -> node-count pass on the mapper's stream (no host wait)."""
        check(lib().kmb_mapper_flush(self._h))

    def sync(self):
        rc = lib().kmb_mapper_sync(self._h)
        if rc == _lib.KMB_ERR_INVALID_BASE:
            off = C.c_int64(-1)
            lib().kmb_mapper_bad_offset(self._h, C.byref(off))
            raise InvalidBaseError(off.value, lib().kmb_last_error().decode())
        check(rc)

    def counts(self, out=None) -> np.ndarray:
        """uint32[n_counts] on the host (mapper.pyx:37,72); ``out`` may be a caller buffer (e.g. pinned)."""
        self.sync()
        if out is None:
            out = np.zeros(self.n_counts, dtype=np.uint32)
        elif out.dtype != np.uint32 or out.shape != (self.n_counts,) or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous uint32[%d] array" % self.n_counts)
        check(lib().kmb_mapper_read_counts(self._h, out.ctypes.data, self.n_counts))
        return out

    def counts_device_ptr(self) -> int:
        p, n = C.c_void_p(), C.c_uint64()
        check(lib().kmb_mapper_counts_device(self._h, C.byref(p), C.byref(n)))
        return p.value

    def lookup_counts(self, keys) -> np.ndarray:
        """counts[node of the first entry whose key == keys[i]], 0 when absent (Counter.__getitem__,
        gpu_counter.py:33)."""
        keep, p, n = as_buffer(keys, np.uint64, "keys")
        out = np.zeros(n, dtype=np.uint32)
        check(lib().kmb_mapper_lookup_counts(self._h, p, n, out.ctypes.data))
        return out

    def write_counts(self, values):
        """Replace the counts (``counter._values = ...``, command_line_interface.py:136)."""
        keep, p, n = as_buffer(values, np.uint32, "values")
        _order_after_torch(values, same_stream=self._stream)
        check(lib().kmb_mapper_write_counts(self._h, p, n))

    def reset(self):
        check(lib().kmb_mapper_reset(self._h))

    def stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        check(lib().kmb_mapper_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def candidates(self) -> int:
        """Look-ups that passed the filter and fetched an index sector since the last reset."""
        v = C.c_uint64()
        check(lib().kmb_mapper_candidates(self._h, C.byref(v)))
        return v.value

    def kernel_time(self):
        """(total ms, n kernels) of the mapping kernels since the last call (option time_kernels=1)."""
        ms, n = C.c_double(), C.c_uint64()
        check(lib().kmb_mapper_kernel_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def apply_time(self):
        """(total ms, n launches) of the apply passes (hit log -> node counts) since the last call."""
        ms, n = C.c_double(), C.c_uint64()
        check(lib().kmb_mapper_apply_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def close(self):
        self._finalizer()


class _BorrowedMapper:
    """Context manager behind ``borrowed_mapper``."""

    def __init__(self, di, n_counts, cutoff):
        self.di, self.key = di, (int(n_counts), int(cutoff))
        self.entry = None
        self.mapper = None

    def __enter__(self) -> Mapper:
        import threading
        cache = self.di.__dict__.setdefault("_mapper_cache", {})
        entry = cache.get(self.key)
        if entry is None:
            entry = cache[self.key] = [threading.Lock(), None]
        if entry[0].acquire(blocking=False):
            self.entry = entry
            if entry[1] is None:
                entry[1] = Mapper(self.di, self.key[0], self.key[1])
            else:
                entry[1].reset()
            self.mapper = entry[1]
        else:                      # another thread is using the cached one: a private mapper for this call
            self.mapper = Mapper(self.di, self.key[0], self.key[1])
        return self.mapper

    def __exit__(self, exc_type, exc, tb):
        if self.entry is not None:
            if exc_type is not None:        # do not keep a mapper whose state is unknown
                try:
                    self.entry[1].close()
                finally:
                    self.entry[1] = None
            self.entry[0].release()
        else:
            self.mapper.close()
        return False


def borrowed_mapper(di: DeviceIndex, n_counts: int, max_index_lookup_frequency: int = DEFAULT_MAX_FREQUENCY):
    """A zeroed Mapper for one drop-in call (``map_kmers_to_graph_index``, ``map_cpu``).  The reference calls these
    once per 2.5 MB chunk (command_line_interface.py:51) and gets a fresh count array each time; creating and
    destroying a device mapper per call would cost a cudaMalloc + memset of the whole count array (320 MB at human
    scale), two streams and a pinned block each time.  One mapper per (index, n_counts, cut-off) is kept on the
    DeviceIndex and reset -- asynchronously -- instead."""
    return _BorrowedMapper(di, n_counts, max_index_lookup_frequency)


__all__ = ["DeviceIndex", "Mapper", "KmbError", "InvalidBaseError", "DEFAULT_MAX_FREQUENCY", "borrowed_mapper"]
