"""``CounterKmerIndex``: the alternative index type of the reference's CPU route
(command_line_interface.py:46-49, 118-119, 133-138; loaded by util.py:63-66).

In the reference this class comes from graph_kmer_index and its ``counter`` from npstructures (both absent third-party
packages: PARITY UNPINNED -- shape and semantics below are restated from the reference's call sites).  The route counts
occurrences per UNIQUE k-mer instead of per node (``kmer_index.counter.count(hashes)``; the per-chunk result is the
counter's value array, ``counter._values``), sums those arrays over the chunks, and turns them into node counts once at
the end (``kmer_index.get_node_counts()`` = per-entry count of the entry's k-mer, summed per node).  It is the same
kernels as the node route minus the node scatter: here the counter is a device index over the unique keys in which
"node i" stands for key i, so counting is the fused mapping kernel and the values array is its count buffer.
"""
from __future__ import annotations

import os

import numpy as np

from .gpu_counter import _KeyCounter

NPZ_KEYS = ("kmers", "nodes")


class CounterKmerIndex:
    def __init__(self, kmers, nodes, counter=None):
        self.kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        self.nodes = np.ascontiguousarray(nodes, dtype=np.int64)
        if self.kmers.shape != self.nodes.shape:
            raise ValueError("CounterKmerIndex: kmers and nodes must have the same length")
        self._counter = counter          # built on first use: needs the GPU

    @property
    def counter(self):
        if self._counter is None:
            self._counter = _KeyCounter(np.unique(self.kmers))
        return self._counter

    @classmethod
    def from_kmer_index(cls, kmer_index) -> "CounterKmerIndex":
        return cls(kmer_index._kmers, kmer_index._nodes)

    def max_node_id(self) -> int:
        return int(self.nodes.max()) if self.nodes.size else 0

    def count_kmers(self, kmers):
        self.counter.count(np.ascontiguousarray(kmers, dtype=np.uint64))

    def get_node_counts(self, min_nodes=0):
        """``np.bincount(nodes, counter[kmers], minlength)``: every entry contributes the count of its k-mer to its
        node; float64 and without a frequency cut-off, like the GPU counter's method (gpu_counter.py:37)."""
        per_entry = self.counter[self.kmers]
        return np.bincount(self.nodes, weights=per_entry, minlength=int(min_nodes))

    # ---- file format of THIS implementation (the reference pickles the object with shared_memory_wrapper.to_file)
    def to_file(self, file_name):
        np.savez(file_name, kmers=self.kmers, nodes=self.nodes, counter_index=np.int64(1))

    @classmethod
    def from_file(cls, file_name) -> "CounterKmerIndex":
        path = file_name
        if not os.path.exists(path) and os.path.exists(str(file_name) + ".npz"):
            path = str(file_name) + ".npz"
        data = np.load(path, allow_pickle=False)
        missing = [k for k in NPZ_KEYS + ("counter_index",) if k not in data.files]
        if missing:
            raise KeyError("%s is not a CounterKmerIndex archive: missing key(s) %s" % (path, missing))
        return cls(data["kmers"], data["nodes"])
