"""The index data model of the mapping path.

The reference takes this from the third-party package graph_kmer_index (setup.py:26, absent from
the reference tree): a ``KmerIndex`` whose six attributes ``_hashes_to_index, _n_kmers, _nodes,
_kmers, _frequencies, _modulo`` are everything the hot loop reads (mapper.pyx:22-29).  This class
carries exactly those, with the construction rule the reference's test uses
(tests/test_mapping.py:34-38: ``KmerIndex.from_flat_kmers(flat, modulo=21)`` then
``convert_to_int32()``) and an ``np.savez`` archive with graph_kmer_index's key names
(PARITY UNPINNED: the key names are recalled, not verifiable here; the loader fails loudly on a
missing key).  Building an index is a one-off host-side step outside the mapped path.
"""
from __future__ import annotations

import os

import numpy as np

DEFAULT_MODULO = 452930477  # graph_kmer_index's default (prime)
NPZ_KEYS = ("hashes_to_index", "n_kmers", "nodes", "ref_offsets", "kmers", "modulo", "frequencies",
            "allele_frequencies")
REQUIRED_KEYS = ("hashes_to_index", "n_kmers", "nodes", "kmers", "modulo")


class KmerIndex:
    def __init__(self, hashes_to_index, n_kmers, nodes, ref_offsets, kmers, modulo=DEFAULT_MODULO,
                 frequencies=None, allele_frequencies=None):
        self._hashes_to_index = hashes_to_index
        self._n_kmers = n_kmers
        self._nodes = nodes
        self._ref_offsets = ref_offsets
        self._kmers = kmers
        self._modulo = int(modulo)
        if frequencies is None:
            frequencies = np.ones(len(kmers), dtype=np.uint16)
        self._frequencies = frequencies
        self._allele_frequencies = allele_frequencies

    # ---- what the mapping path calls (util.py:42-43,61-62; command_line_interface.py:51,79,117) ----
    def convert_to_int32(self):
        self._hashes_to_index = np.ascontiguousarray(self._hashes_to_index, dtype=np.int32)
        self._n_kmers = np.ascontiguousarray(self._n_kmers, dtype=np.int32)
        self._nodes = np.ascontiguousarray(self._nodes, dtype=np.int32)
        self._kmers = np.ascontiguousarray(self._kmers, dtype=np.uint64)
        self._frequencies = np.ascontiguousarray(self._frequencies, dtype=np.uint16)

    def remove_ref_offsets(self):
        self._ref_offsets = None

    def max_node_id(self) -> int:
        return int(np.max(self._nodes)) if len(self._nodes) else 0

    def validate(self):
        """The reference runs with boundscheck off (mapper.pyx:15-18); check once on load instead."""
        m = self._modulo
        if len(self._hashes_to_index) != m or len(self._n_kmers) != m:
            raise ValueError("index: hashes_to_index/n_kmers must have modulo=%d elements" % m)
        n = len(self._kmers)
        if len(self._nodes) != n or len(self._frequencies) != n:
            raise ValueError("index: nodes/kmers/frequencies length mismatch")

    # ---- construction (tests/test_mapping.py:34-38) ----
    @classmethod
    def from_flat_kmers(cls, flat_kmers=None, modulo=DEFAULT_MODULO, hashes=None, nodes=None, ref_offsets=None,
                        frequencies=None) -> "KmerIndex":
        """Sort entries by ``kmer % modulo``; ``hashes_to_index[h]`` = first entry of bucket h,
        ``n_kmers[h]`` = bucket size, ``frequencies[l]`` = number of entries sharing ``kmers[l]``."""
        if flat_kmers is not None:
            hashes = getattr(flat_kmers, "_hashes", None) if hashes is None else hashes
            nodes = getattr(flat_kmers, "_nodes", None) if nodes is None else nodes
            ref_offsets = getattr(flat_kmers, "_ref_offsets", None) if ref_offsets is None else ref_offsets
        kmers = np.asarray(hashes, dtype=np.uint64)
        nodes = np.asarray(nodes, dtype=np.int64)
        modulo = int(modulo)
        h = (kmers % np.uint64(modulo)).astype(np.int64)
        order = np.argsort(h, kind="stable")
        kmers, nodes, h = kmers[order], nodes[order], h[order]
        n_kmers = np.bincount(h, minlength=modulo).astype(np.int64)
        hashes_to_index = np.cumsum(n_kmers) - n_kmers
        if frequencies is None:
            _, inv, cnt = np.unique(kmers, return_inverse=True, return_counts=True)
            frequencies = np.minimum(cnt[inv], 65535) if kmers.size else np.zeros(0, np.int64)
        else:
            frequencies = np.asarray(frequencies)[order]
        if ref_offsets is not None:
            ref_offsets = np.asarray(ref_offsets)[order]
        return cls(hashes_to_index, n_kmers, nodes, ref_offsets, kmers, modulo, np.asarray(frequencies, dtype=np.uint16))

    # ---- .npz ----
    @classmethod
    def from_file(cls, file_name) -> "KmerIndex":
        path = file_name
        if not os.path.exists(path) and os.path.exists(str(file_name) + ".npz"):
            path = str(file_name) + ".npz"
        # allow_pickle stays off: an archive is user input (-i), and unpickling runs arbitrary code.  The only
        # pickled members graph_kmer_index writes are None-valued optional entries (ref_offsets, allele_frequencies
        # saved as object arrays): those fail to load without pickle and are treated as absent.
        data = np.load(path, allow_pickle=False)
        missing = [k for k in REQUIRED_KEYS if k not in data.files]
        if missing:
            raise KeyError("%s is not a KmerIndex archive: missing key(s) %s (has %s)" % (path, missing, data.files))

        def opt(key):
            if key not in data.files:
                return None
            try:
                return data[key]
            except ValueError:      # an object array (a pickled None): absent
                return None

        idx = cls(data["hashes_to_index"], data["n_kmers"], data["nodes"], opt("ref_offsets"), data["kmers"],
                  int(data["modulo"]), opt("frequencies"), opt("allele_frequencies"))
        idx.validate()
        return idx

    def to_file(self, file_name):
        arrays = dict(hashes_to_index=self._hashes_to_index, n_kmers=self._n_kmers, nodes=self._nodes,
                      kmers=self._kmers, modulo=np.int64(self._modulo), frequencies=self._frequencies)
        if self._ref_offsets is not None:
            arrays["ref_offsets"] = self._ref_offsets
        if self._allele_frequencies is not None:
            arrays["allele_frequencies"] = self._allele_frequencies
        np.savez(file_name, **arrays)

    def get(self, kmer):
        """(nodes, ref_offsets, frequencies) of the entries whose key equals ``kmer`` (host-side helper
        used by tests/test_mapping.py:38 of the reference), or None."""
        kmer = int(kmer)
        h = kmer % self._modulo
        s = int(self._hashes_to_index[h])
        e = s + int(self._n_kmers[h])
        sel = np.flatnonzero(np.asarray(self._kmers[s:e], dtype=np.uint64) == np.uint64(kmer)) + s
        if sel.size == 0:
            return None
        ro = None if self._ref_offsets is None else np.asarray(self._ref_offsets)[sel]
        return np.asarray(self._nodes)[sel], ro, np.asarray(self._frequencies)[sel]
