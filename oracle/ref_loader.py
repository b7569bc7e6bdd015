"""Import the compiled, unmodified reference ``kmer_mapper.mapper`` from ``oracle/_ref``.

TEST INFRASTRUCTURE ONLY.  The two names mapper.pyx imports at module load (mapper.pyx:8-9) and
never uses inside its three functions are satisfied with stub modules, because graph_kmer_index
and kmer_mapper.util's own imports (bionumpy, isal, shared_memory_wrapper) are absent here.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

from .build_ref import REF_OUT, ref_so_path

_cached = None


def load_reference_mapper():
    """Returns the reference's compiled module (attributes map_kmers_to_graph_index,
    in_graph_index, in_graph_index_no_memory_maps) or None when it has not been built."""
    global _cached
    if _cached is not None:
        return _cached
    if not os.path.exists(ref_so_path()):
        return None
    saved = {name: sys.modules.get(name) for name in
             ("graph_kmer_index", "graph_kmer_index.shared_mem", "kmer_mapper", "kmer_mapper.util",
              "kmer_mapper.mapper")}
    gki = types.ModuleType("graph_kmer_index")
    shm = types.ModuleType("graph_kmer_index.shared_mem")
    shm.to_shared_memory = None
    shm.SingleSharedArray = None
    gki.shared_mem = shm
    util = types.ModuleType("kmer_mapper.util")
    util.log_memory_usage_now = lambda *a, **k: None
    sys.modules["graph_kmer_index"] = gki
    sys.modules["graph_kmer_index.shared_mem"] = shm
    sys.modules.pop("kmer_mapper", None)
    sys.modules["kmer_mapper.util"] = util
    sys.path.insert(0, REF_OUT)
    try:
        mod = importlib.import_module("kmer_mapper.mapper")
    finally:
        sys.path.remove(REF_OUT)
        for name in ("graph_kmer_index", "graph_kmer_index.shared_mem", "kmer_mapper", "kmer_mapper.util"):
            if saved[name] is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = saved[name]
    _cached = mod
    return mod
