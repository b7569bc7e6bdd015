"""Drop-in for the reference's Cython module kmer_mapper/mapper.pyx (same names, arguments, dtypes,
error behaviour), computed by the CUDA kernels behind include/kmer_mapper_b200.h."""
from __future__ import annotations

import logging
import time

import numpy as np

from . import _lib
from .device import DEFAULT_MAX_FREQUENCY, DeviceIndex, borrowed_mapper


def _check_kmers(kmers):
    """The reference's signature is ``np.uint64_t[::1] kmers`` (mapper.pyx:19): anything that is not a
    C-contiguous 1-d uint64 buffer raises ValueError at the call."""
    if _lib.is_torch_tensor(kmers):
        return kmers
    a = np.asarray(kmers) if not isinstance(kmers, np.ndarray) else kmers
    if a.dtype != np.uint64:
        raise ValueError("Buffer dtype mismatch, expected 'uint64_t' but got %r" % a.dtype.name)
    if a.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % a.ndim)
    if not a.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous")
    return a


def map_kmers_to_graph_index(index, max_node_id, kmers, max_index_lookup_frequency=DEFAULT_MAX_FREQUENCY):
    """mapper.pyx:19-72.  For each query k-mer: bucket ``kmer % modulo``; every entry of the bucket with
    an equal key (no break) and ``frequency <= max_index_lookup_frequency`` adds 1 to its node.
    Returns a fresh ``uint32[max_node_id+1]`` (mapper.pyx:37).

    ``index`` is duck-typed exactly like the reference: any object with ``_hashes_to_index, _n_kmers,
    _nodes`` (int32), ``_kmers`` (uint64), ``_frequencies`` (uint16) and ``_modulo``.  Unlike the
    reference (boundscheck off, mapper.pyx:15-18) an index whose nodes exceed ``max_node_id`` or whose
    buckets leave the arrays raises instead of corrupting memory.
    """
    t = time.perf_counter()
    kmers = _check_kmers(kmers)
    di = DeviceIndex.from_index(index)
    with borrowed_mapper(di, int(max_node_id) + 1, int(max_index_lookup_frequency)) as m:
        m.map_kmers(kmers)
        out = m.counts()           # a fresh host array per call (mapper.pyx:37)
    logging.debug("Time spent looking up hashes: %.3f", time.perf_counter() - t)  # mapper.pyx:71
    return out


def in_graph_index(index, kmers, max_index_lookup_frequency=DEFAULT_MAX_FREQUENCY):
    """mapper.pyx:81-130: uint8[len(kmers)], 1 where the k-mer is a key of the index.  The frequency
    argument is accepted and ignored, like the reference (:112-127)."""
    kmers = _check_kmers(kmers)
    return DeviceIndex.from_index(index).in_graph_index(kmers)


def in_graph_index_no_memory_maps(index, kmers, max_index_lookup_frequency=DEFAULT_MAX_FREQUENCY):
    """mapper.pyx:137-190: same result as in_graph_index (the reference keeps a second copy that takes
    np.ndarray buffers "so that ray-stuff works", :138)."""
    return in_graph_index(index, kmers, max_index_lookup_frequency)
