"""Drop-in for the reference's ``kmer_mapper/gpu_counter.py`` (class ``GpuCounter``, gpu_counter.py:5-37).

The reference wraps the third-party ``cucounter`` hash table: count occurrences of a fixed set of unique keys,
then scatter the per-key counts onto nodes.  Here the "hash table" is the same device index the mapper uses,
built over the unique keys with node ``i`` standing for key ``i``; counting is the mapping kernel and the per-key
read-back is its lookup kernel.  Method names, arguments and results follow the reference; the bodies do not.
"""
from __future__ import annotations

import logging

import numpy as np

from .device import DeviceIndex, Mapper
from .kmer_index import KmerIndex

_QUERY_SLICE = 10_000_000  # the reference reads the per-key counts back in slices of this many keys (gpu_counter.py:29)


def _table_size(n_keys: int, requested: int) -> int:
    """Bucket count of the key table.  ``requested`` is the reference's hash-map size hint
    (``--gpu-hash-map-size``, 0 = choose: command_line_interface.py:178); the default gives ~4 buckets per key."""
    size = int(requested) if requested and requested > 0 else max(4 * int(n_keys), 1021) | 1
    return min(size, 2 ** 32 - 1)


class _KeyCounter:
    """What ``cucounter.counter.Counter`` is to the reference (gpu_counter.py:16,24,33): built over unique keys,
    ``count(kmers, count_revcomps, k)`` accumulates, ``counter[keys]`` reads the counts of the given keys."""

    def __init__(self, unique_keys, table_size=0):
        keys = np.ascontiguousarray(unique_keys, dtype=np.uint64)
        n = keys.shape[0]
        table = KmerIndex.from_flat_kmers(hashes=keys, nodes=np.arange(n, dtype=np.int64),
                                          modulo=_table_size(n, table_size), frequencies=np.ones(n, dtype=np.uint16))
        table.convert_to_int32()
        self._index = DeviceIndex.from_index(table)
        # every key has frequency 1, so no cut-off ever applies: the counter counts all occurrences
        self._mapper = Mapper(self._index, n_counts=max(n, 1), max_index_lookup_frequency=65535)

        self._n_keys = n

    @property
    def n_keys(self) -> int:
        return self._n_keys

    def count(self, kmers, count_revcomps=False, k=31):
        self._mapper.map_kmers(kmers, revcomp=bool(count_revcomps), k=k)

    def count_reads(self, bases, offsets, k, count_revcomps=False):
        """The same from raw reads (N -> A, 2-bit, rolling k-mers fused into the counting kernel): no hash array."""
        self._mapper.map_reads(bases, offsets, k, revcomp=bool(count_revcomps), n_to_a=True)

    def __getitem__(self, keys):
        return self._mapper.lookup_counts(np.ascontiguousarray(keys, dtype=np.uint64))

    # ``counter._values``: the per-unique-key counts, in key order (command_line_interface.py:48,119,136)
    @property
    def _values(self):
        return self._mapper.counts()[:self._n_keys]

    @_values.setter
    def _values(self, values):
        v = np.zeros(self._mapper.n_counts, dtype=np.uint32)
        v[:self._n_keys] = np.asarray(values, dtype=np.uint32)
        self._mapper.write_counts(v)


class GpuCounter:
    """Same constructor, attributes and methods as the reference class."""

    def __init__(self, unique_kmers, kmers, nodes, k):
        self.k = k
        self.nodes = nodes
        self.kmers = kmers
        self.unique_kmers = unique_kmers
        self.counter = None  # created by initialize_cuda

    @classmethod
    def from_kmers_and_nodes(cls, kmers, nodes, k) -> "GpuCounter":
        return cls(np.unique(kmers), kmers, nodes, k)

    def initialize_cuda(self, modulo):
        logging.info("Building the device key table over %d unique k-mers" % len(self.unique_kmers))
        self.counter = _KeyCounter(self.unique_kmers, modulo)

    def count(self, kmers, count_revcomps=False):
        """Accumulates over calls, like the reference's counter."""
        if self.counter is None:
            raise RuntimeError("GpuCounter.initialize_cuda() has not been called")
        self.counter.count(kmers, count_revcomps, self.k)

    def get_node_counts(self, min_nodes=0):
        """Per-key counts of every index entry, summed per node: float64 of length
        ``max(min_nodes, nodes.max() + 1)`` and without a frequency cut-off, exactly what the reference method
        returns (gpu_counter.py:37).  The mapping CLI does not take this route: ``map_gpu`` returns the CPU
        route's uint32 counts with the cut-off applied."""
        entry_keys = np.ascontiguousarray(self.kmers, dtype=np.uint64)
        per_entry = np.empty(entry_keys.shape[0], dtype=np.uint32)
        for lo in range(0, entry_keys.shape[0], _QUERY_SLICE):
            hi = min(lo + _QUERY_SLICE, entry_keys.shape[0])
            per_entry[lo:hi] = self.counter[entry_keys[lo:hi]]
        return np.bincount(np.asarray(self.nodes), weights=per_entry, minlength=min_nodes)
