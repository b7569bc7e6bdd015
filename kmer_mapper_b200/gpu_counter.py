"""Drop-in for the reference's kmer_mapper/gpu_counter.py.

The reference's GpuCounter wraps the third-party ``cucounter`` hash table (gpu_counter.py:14-16).
Here the table is the same device index the mapper uses, built over the unique keys with
"node" i = key i, so ``count`` is the mapping kernel and ``counter[keys]`` is its lookup kernel.
"""
from __future__ import annotations

import logging

import numpy as np

from .device import DeviceIndex, Mapper
from .kmer_index import KmerIndex


def _auto_capacity(n_keys: int) -> int:
    """Bucket count when the caller passes 0 (``--gpu-hash-map-size`` default, command_line_interface.py:178):
    about four buckets per key, odd."""
    return max(4 * int(n_keys), 1021) | 1


class _DeviceCounter:
    """The three operations gpu_counter.py uses from cucounter.Counter: construct over unique keys,
    ``count(kmers, count_revcomps, k)`` (cumulative), ``counter[keys]``."""

    def __init__(self, unique_kmers, capacity=0):
        keys = np.ascontiguousarray(unique_kmers, dtype=np.uint64)
        modulo = int(capacity) if capacity and capacity > 0 else _auto_capacity(keys.shape[0])
        modulo = min(modulo, 2 ** 32 - 1)
        idx = KmerIndex.from_flat_kmers(hashes=keys, nodes=np.arange(keys.shape[0], dtype=np.int64), modulo=modulo,
                                        frequencies=np.ones(keys.shape[0], dtype=np.uint16))
        idx.convert_to_int32()
        self._index = DeviceIndex.from_index(idx)
        self._mapper = Mapper(self._index, n_counts=max(keys.shape[0], 1), max_index_lookup_frequency=65535)

    def count(self, kmers, count_revcomps=False, k=31):
        self._mapper.map_kmers(kmers, revcomp=bool(count_revcomps), k=k)

    def __getitem__(self, keys):
        return self._mapper.lookup_counts(np.ascontiguousarray(keys, dtype=np.uint64))


class GpuCounter:
    def __init__(self, unique_kmers, kmers, nodes, k):
        self.unique_kmers = unique_kmers
        self.kmers = kmers
        self.nodes = nodes
        self.counter = None
        self.k = k

    def initialize_cuda(self, modulo):
        """gpu_counter.py:13-16; ``modulo`` is the table capacity hint, 0 = choose."""
        logging.info("N unique kmers: %d" % len(self.unique_kmers))
        self.counter = _DeviceCounter(self.unique_kmers, modulo)

    @classmethod
    def from_kmers_and_nodes(cls, kmers, nodes, k) -> "GpuCounter":
        unique_kmers = np.unique(kmers)
        return cls(unique_kmers, kmers, nodes, k)

    def count(self, kmers, count_revcomps=False):
        """gpu_counter.py:23-24; cumulative across calls."""
        self.counter.count(kmers, count_revcomps, self.k)

    def get_node_counts(self, min_nodes=0):
        """gpu_counter.py:26-37: per-entry counts scattered onto nodes with
        ``np.bincount(nodes, counts, minlength=min_nodes)`` -- float64, unfiltered, exactly like the
        reference method (the mapping CLI does not use this route, see command_line_interface.map_gpu)."""
        counts = np.zeros(len(self.kmers), dtype=np.uint32)
        chunk_size = 10_000_000
        start = 0
        kmers = np.ascontiguousarray(self.kmers, dtype=np.uint64)
        for chunk in np.array_split(kmers, max(1, len(kmers) // chunk_size)):
            logging.debug("Querying chunk %d-%d" % (start, start + len(chunk)))
            counts[start:start + len(chunk)] = self.counter[chunk]
            start += len(chunk)
        logging.info("Doing bincount")
        return np.bincount(self.nodes, counts, minlength=min_nodes)
