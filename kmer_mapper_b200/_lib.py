"""ctypes binding of the C ABI declared in include/kmer_mapper_b200.h.

The product path has no CPU fallback: if the shared library is missing, or there is no CUDA
device, every compute call raises (``KmbError``) instead of computing something on the host.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._build import LIB_PATH

KMB_OK = 0
KMB_ERR_BAD_ARG = -1
KMB_ERR_INVALID_BASE = -2
KMB_ERR_CUDA = -3
KMB_ERR_BAD_INDEX = -4
KMB_ERR_NOMEM = -5
KMB_ERR_NCCL = -6
COMM_ID_BYTES = 128

FLAG_REVCOMP = 1
FLAG_NO_N_TO_A = 2


class KmbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("kmer_mapper_b200 error %d: %s" % (code, message))
        self.code = code


class InvalidBaseError(ValueError):
    """A byte outside ACGTacgt (and N under the N->A policy) in the reads; bionumpy raises an
    EncodingError at util.py:72-73 in the same situation."""

    def __init__(self, offset: int, message: str):
        super().__init__(message)
        self.offset = offset


_u8p, _u16p, _u32p, _u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64))
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/kmer_mapper_b200.h declares
SIGNATURES = {
    "kmb_last_error": (C.c_char_p, []),
    "kmb_version": (C.c_char_p, []),
    "kmb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "kmb_index_create": (C.c_int, [C.c_int, _vp, _vp, C.c_uint64, _vp, _vp, _vp, C.c_uint64, C.POINTER(_vp)]),
    "kmb_index_destroy": (C.c_int, [_vp]),
    "kmb_index_info": (C.c_int, [_vp, C.POINTER(C.c_int64), _u64p, _u64p, _u64p]),
    "kmb_index_filter_bytes": (C.c_int, [_vp, _u64p]),
    "kmb_index_layout": (C.c_int, [_vp, _u64p, _u64p, _u64p]),
    "kmb_mapper_create": (C.c_int, [_vp, C.c_uint64, _vp, C.c_int, C.POINTER(_vp)]),
    "kmb_mapper_destroy": (C.c_int, [_vp]),
    "kmb_mapper_set_stream": (C.c_int, [_vp, _vp]),
    "kmb_mapper_map_kmers": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint32, C.c_int]),
    "kmb_mapper_map_reads": (C.c_int, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, C.c_int, C.c_uint32]),
    "kmb_mapper_map_text": (C.c_int, [_vp, _vp, C.c_uint64, C.c_int, C.c_int, C.c_uint32]),
    "kmb_mapper_map_text_fd": (C.c_int, [_vp, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint32]),
    "kmb_text_parsed": (C.c_int, [_u64p, _u64p]),
    "kmb_mapper_map_gz": (C.c_int, [_vp, _vp, C.c_uint64, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, _u64p]),
    "kmb_gz_device_stats": (C.c_int, [_u64p, _u64p, _u64p]),
    "kmb_parse_text_device": (C.c_int, [C.c_int, _vp, C.c_uint64, C.c_int, _vp, C.c_uint64, _vp, C.c_uint64, _u64p, _u64p]),
    "kmb_mapper_flush": (C.c_int, [_vp]),
    "kmb_mapper_sync": (C.c_int, [_vp]),
    "kmb_mapper_bad_offset": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "kmb_mapper_read_counts": (C.c_int, [_vp, _vp, C.c_uint64]),
    "kmb_mapper_write_counts": (C.c_int, [_vp, _vp, C.c_uint64]),
    "kmb_mapper_reset": (C.c_int, [_vp]),
    "kmb_mapper_counts_device": (C.c_int, [_vp, C.POINTER(_vp), _u64p]),
    "kmb_mapper_stats": (C.c_int, [_vp, _u64p, _u64p]),
    "kmb_comm_unique_id": (C.c_int, [_vp]),
    "kmb_comm_init_rank": (C.c_int, [C.c_int, C.c_int, C.c_int, _vp, C.POINTER(_vp)]),
    "kmb_comm_destroy": (C.c_int, [_vp]),
    "kmb_mapper_allreduce": (C.c_int, [_vp, _vp]),
    "kmb_in_graph_index": (C.c_int, [_vp, _vp, C.c_uint64, _vp]),
    "kmb_hash_reads": (C.c_int, [C.c_int, _vp, C.c_uint64, _vp, C.c_uint64, C.c_int, C.c_uint32, _vp, C.c_uint64,
                                 _u64p, C.POINTER(C.c_int64)]),
    "kmb_mapper_lookup_counts": (C.c_int, [_vp, _vp, C.c_uint64, _vp]),
    "kmb_codec_actg_from_bytes": (C.c_int, [C.c_int, _vp, C.c_uint64, _vp]),
    "kmb_codec_simple_from_bytes": (C.c_int, [C.c_int, _vp, C.c_uint64, _vp]),
    "kmb_codec_to_bytes": (C.c_int, [C.c_int, _vp, C.c_uint64, _vp]),
    "kmb_codec_complement": (C.c_int, [C.c_int, _vp, C.c_uint64, _vp]),
    "kmb_codec_twobit_swap": (C.c_int, [C.c_int, _vp, C.c_uint64, C.c_int, _vp]),
    "kmb_gunzip_members": (C.c_int, [_vp, C.c_uint64, C.c_int, _vp, C.c_uint64, C.c_uint64, _u64p, _u64p, C.POINTER(C.c_int)]),
    "kmb_gzstream_open": (C.c_int, [_vp, C.c_uint64, C.c_int, C.POINTER(_vp)]),
    "kmb_gzstream_read": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint64, _u64p, C.POINTER(C.c_int)]),
    "kmb_gzstream_error": (C.c_char_p, [_vp]),
    "kmb_gzstream_close": (C.c_int, [_vp]),
    "kmb_find_record_start": (C.c_int, [_vp, C.c_uint64, C.c_int, _u64p]),
    "kmb_host_read_bandwidth": (C.c_int, [_vp, C.c_uint64, C.c_int, C.POINTER(C.c_double)]),
    "kmb_pack_bases": (C.c_int, [_vp, C.c_uint64, C.c_uint32, C.c_int, _vp, C.c_uint64, C.POINTER(C.c_int64)]),
    "kmb_parse_reads": (C.c_int, [_vp, C.c_uint64, C.c_int, C.c_int, C.c_int, _vp, C.c_uint64, _vp, C.c_uint64,
                                  _u64p, _u64p, _u64p]),
    "kmb_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_size_t]),
    "kmb_host_free": (C.c_int, [_vp]),
    "kmb_bench_gather": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(C.c_float)]),
    "kmb_mapper_candidates": (C.c_int, [_vp, _u64p]),
    "kmb_mapper_kernel_time": (C.c_int, [_vp, C.POINTER(C.c_double), _u64p]),
    "kmb_mapper_apply_time": (C.c_int, [_vp, C.POINTER(C.c_double), _u64p]),
    "kmb_set_option": (C.c_int, [C.c_char_p, C.c_int64]),
    "kmb_get_option": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64)]),
    "kmb_launch_count": (C.c_int, [_u64p]),
}

_lib = None


def lib() -> C.CDLL:
    """The loaded CUDA library.  Raises if it has not been built: there is no other code path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KmbError(KMB_ERR_CUDA,
                           "%s is missing: build it with `python -m kmer_mapper_b200._build` "
                           "(nvcc, sm_100a). kmer_mapper_b200 has no CPU fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc == KMB_OK:
        return
    msg = lib().kmb_last_error().decode("utf-8", "replace")
    raise KmbError(rc, msg)


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().kmb_device_count(C.byref(n))
    return n.value if rc == KMB_OK else 0


def require_device() -> None:
    if device_count() < 1:
        raise KmbError(KMB_ERR_CUDA, "no CUDA device: kmer_mapper_b200 computes only on the GPU (no CPU fallback)")


def set_option(name: str, value: int) -> None:
    check(lib().kmb_set_option(name.encode(), int(value)))


def get_option(name: str) -> int:
    v = C.c_int64(0)
    check(lib().kmb_get_option(name.encode(), C.byref(v)))
    return v.value


def launch_count() -> int:
    v = C.c_uint64(0)
    check(lib().kmb_launch_count(C.byref(v)))
    return v.value


# ---- buffers: numpy arrays (host) or torch CUDA tensors (device) -----------------------------------

def is_torch_tensor(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch" and hasattr(x, "data_ptr")


def as_buffer(x, np_dtype, name="array"):
    """Returns (keepalive, pointer, n_elements).  numpy input is made C-contiguous with the exact
    dtype the reference's memoryview casts demand (mapper.pyx:19,22-29); a torch tensor must already
    be contiguous with the matching dtype (no hidden device copies)."""
    np_dtype = np.dtype(np_dtype)
    if is_torch_tensor(x):
        import torch
        want = {"uint8": torch.uint8, "int32": torch.int32, "int64": torch.int64, "uint16": torch.uint16,
                "uint32": torch.uint32, "uint64": torch.uint64}[np_dtype.name]
        ok = x.dtype == want or (x.element_size() == np_dtype.itemsize and not x.dtype.is_floating_point)
        if not ok:
            raise ValueError("%s: tensor dtype %s does not match %s" % (name, x.dtype, np_dtype))
        if not x.is_contiguous():
            raise ValueError("%s: tensor must be contiguous" % name)
        return x, x.data_ptr(), x.numel()
    a = np.asarray(x)
    if a.dtype != np_dtype:
        raise ValueError("%s: Buffer dtype mismatch, expected %s but got %s" % (name, np_dtype, a.dtype))
    if a.ndim != 1:
        raise ValueError("%s: Buffer has wrong number of dimensions (expected 1, got %d)" % (name, a.ndim))
    if not a.flags.c_contiguous:
        raise ValueError("%s: ndarray is not C-contiguous" % name)
    return a, a.ctypes.data, a.shape[0]
