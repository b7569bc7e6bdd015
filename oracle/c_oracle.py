"""ctypes binding of oracle/kmer_oracle.c (TEST INFRASTRUCTURE ONLY, see that file's header)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build_ref import C_OUT, build_c_oracle

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(C_OUT, "libkmer_oracle.so")
        if not os.path.exists(path):
            build_c_oracle(verbose=False)
        _lib = C.CDLL(path)
        _lib.ko_kmer_hashes.restype = C.c_int64
        _lib.ko_map_reads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _idx(index):
    return (np.ascontiguousarray(index._hashes_to_index, dtype=np.int32),
            np.ascontiguousarray(index._n_kmers, dtype=np.int32),
            np.ascontiguousarray(index._nodes, dtype=np.int32),
            np.ascontiguousarray(index._kmers, dtype=np.uint64),
            np.ascontiguousarray(index._frequencies, dtype=np.uint16))


def kmer_hashes(bases, offsets, k, n_to_a=True):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    bad = C.c_int64(-1)
    n = lib().ko_kmer_hashes(_p(bases), _p(offsets), C.c_int64(len(offsets) - 1), C.c_int(k),
                             C.c_int(int(n_to_a)), None, C.byref(bad))
    if n < 0:
        raise ValueError("invalid base at flat offset %d" % bad.value)
    out = np.empty(n, dtype=np.uint64)
    lib().ko_kmer_hashes(_p(bases), _p(offsets), C.c_int64(len(offsets) - 1), C.c_int(k),
                         C.c_int(int(n_to_a)), _p(out), C.byref(bad))
    return out


def map_kmers_to_graph_index(index, max_node_id, kmers, max_index_lookup_frequency=1000):
    h2i, nk, nodes, ikm, freq = _idx(index)
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
    counts = np.zeros(max_node_id + 1, dtype=np.uint32)
    lib().ko_map_kmers(_p(h2i), _p(nk), _p(nodes), _p(ikm), _p(freq), C.c_uint64(int(index._modulo)),
                       _p(kmers), C.c_int64(kmers.shape[0]), C.c_int(max_index_lookup_frequency), _p(counts))
    return counts


def in_graph_index(index, kmers):
    h2i, nk, nodes, ikm, freq = _idx(index)
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
    out = np.zeros(kmers.shape[0], dtype=np.uint8)
    lib().ko_in_graph_index(_p(h2i), _p(nk), _p(ikm), C.c_uint64(int(index._modulo)),
                            _p(kmers), C.c_int64(kmers.shape[0]), _p(out))
    return out


def map_reads(index, max_node_id, bases, offsets, k, max_index_lookup_frequency=1000, n_threads=1,
              counts=None):
    """Returns (counts uint32[max_node_id+1], n_kmers_mapped)."""
    h2i, nk, nodes, ikm, freq = _idx(index)
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    if counts is None:
        counts = np.zeros(max_node_id + 1, dtype=np.uint32)
    n_mapped = C.c_int64(0)
    bad = C.c_int64(-1)
    rc = lib().ko_map_reads(_p(h2i), _p(nk), _p(nodes), _p(ikm), _p(freq), C.c_uint64(int(index._modulo)),
                            _p(bases), _p(offsets), C.c_int64(len(offsets) - 1), C.c_int(k),
                            C.c_int(max_index_lookup_frequency), C.c_int(n_threads), _p(counts),
                            C.byref(n_mapped), C.byref(bad))
    if rc != 0:
        raise ValueError("invalid base at flat offset %d" % bad.value)
    return counts, n_mapped.value
