"""Drop-in for the mapping-path functions of the reference's kmer_mapper/util.py."""
from __future__ import annotations

import ctypes as C
import logging
import resource
import sys

import numpy as np

from . import _lib
from ._lib import FLAG_NO_N_TO_A, InvalidBaseError, as_buffer, lib
from .device import current_device
from .kmer_index import KmerIndex
from .sequences import as_ragged


def log_memory_usage_now(logplace=""):
    """util.py:33-35."""
    memory = int(resource.getrusage(resource.RUSAGE_SELF).ru_maxrss) / 1000000
    logging.info("Memory usage (%s): %.4f GB" % (logplace, memory))


def _looks_like_index(obj) -> bool:
    return all(hasattr(obj, a) for a in ("_hashes_to_index", "_n_kmers", "_nodes", "_kmers", "_modulo"))


def _looks_like_counter_index(obj) -> bool:
    return all(hasattr(obj, a) for a in ("counter", "get_node_counts", "kmers", "nodes"))


def _get_kmer_index_from_args(args):
    """util.py:38-68: ``args.kmer_index`` may be a loaded index object or a path (-i); -b bundles,
    MinimalKmerIndex and CounterKmerIndex files belong to graph_kmer_index / shared_memory_wrapper,
    which are outside the mapped path -- they fail loudly here."""
    kmer_index = getattr(args, "kmer_index", None)
    if kmer_index is not None and not isinstance(kmer_index, (str, bytes)) and _looks_like_counter_index(kmer_index):
        return kmer_index
    if kmer_index is not None and not isinstance(kmer_index, (str, bytes)) and _looks_like_index(kmer_index):
        if hasattr(kmer_index, "convert_to_int32"):
            kmer_index.convert_to_int32()
        if hasattr(kmer_index, "remove_ref_offsets"):
            kmer_index.remove_ref_offsets()
        return kmer_index
    if kmer_index is None:
        if getattr(args, "index_bundle", None) is None:
            logging.error("Either a kmer index (-i) or an index bundle (-b) needs to be specified")
            sys.exit(1)
        raise NotImplementedError("index bundles (-b) are graph_kmer_index IndexBundle pickles; only plain KmerIndex "
                                  ".npz files (-i) are supported on this path")
    if "minimal" in str(kmer_index):
        raise NotImplementedError("MinimalKmerIndex files are not supported; pass a plain KmerIndex .npz")
    try:
        index = KmerIndex.from_file(kmer_index)
        index.convert_to_int32()
        index.remove_ref_offsets()  # not needed, will save us some memory
    except KeyError:
        # util.py:63-66: anything that is not a KmerIndex archive is tried as a CounterKmerIndex
        from .counter_index import CounterKmerIndex
        index = CounterKmerIndex.from_file(kmer_index)
        logging.info("Kmer index is counter index")
    return index


def get_kmer_hashes_from_chunk_sequence(chunk_sequence, kmer_size, n_to_a=False, device=None):
    """util.py:71-75: every in-read window of every read as ``uint64``, read-major then position order,
    ``hash = sum_j code(read[p+j]) * 4**j`` with A,C,G,T = 0,1,2,3 (case-insensitive).  Like the
    reference function it has no N policy of its own (the CPU route replaces N by A *before* calling
    it, command_line_interface.py:41-42; pass ``n_to_a=True`` for that); any other byte raises.

    Host arrays in -> numpy array out; torch CUDA tensors in -> torch CUDA tensor out.
    """
    seq = as_ragged(chunk_sequence)
    _lib.require_device()
    kb, pb, nb = as_buffer(seq.bases, np.uint8, "bases")
    ko, po, no = as_buffer(seq.offsets, np.int64, "offsets")
    on_device = _lib.is_torch_tensor(seq.bases)
    if on_device:
        device = seq.bases.device.index
    device = current_device() if device is None else int(device)
    flags = 0 if n_to_a else FLAG_NO_N_TO_A
    k = int(kmer_size)
    # upper bound on the number of windows: sum_r max(0, L_r - k + 1)
    if on_device:
        import torch
        lengths = seq.offsets[1:] - seq.offsets[:-1]
        n_max = int(torch.clamp(lengths - (k - 1), min=0).sum().item())
        out = torch.empty(n_max, dtype=torch.uint64, device=seq.bases.device)
        out_ptr = out.data_ptr()
    else:
        lengths = np.diff(np.asarray(seq.offsets))
        n_max = int(np.maximum(lengths - (k - 1), 0).sum())
        out = np.empty(n_max, dtype=np.uint64)
        out_ptr = out.ctypes.data
    if on_device:
        torch.cuda.current_stream(seq.bases.device).synchronize()
    n_out, bad = C.c_uint64(0), C.c_int64(-1)
    rc = lib().kmb_hash_reads(device, pb, nb, po, no - 1, k, flags, out_ptr, n_max, C.byref(n_out), C.byref(bad))
    if rc == _lib.KMB_ERR_INVALID_BASE:
        raise InvalidBaseError(bad.value, lib().kmb_last_error().decode())
    _lib.check(rc)
    assert n_out.value == n_max, (n_out.value, n_max)
    logging.debug("N hashes: %d" % n_max)
    return out
