"""The drop-in boundary without a GPU: the C-ABI library loads, exports every symbol the header
declares, and every compute entry point fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kmer_mapper_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kmb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def kmb():
    from kmer_mapper_b200 import _build, _lib
    _build.build(verbose=False)        # nvcc cross-compiles sm_100a without a GPU
    return _lib


def test_library_exports_every_declared_symbol(kmb):
    names = declared_symbols()
    assert len(names) >= 30
    handle = kmb.lib()
    for n in names:
        assert hasattr(handle, n), "missing export %s" % n
    # and the ctypes table binds exactly the declared surface
    assert sorted(kmb.SIGNATURES) == names
    assert b"sm_100a" in handle.kmb_version()


def test_library_is_sm100a_only(kmb):
    import subprocess
    from kmer_mapper_b200._build import LIB_PATH
    out = subprocess.run(["cuobjdump", "-lelf", LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_options_round_trip(kmb):
    for name in ("map_reads_blocks_per_sm", "probe_variant", "gathers_in_flight", "use_filter"):
        old = kmb.get_option(name)
        kmb.set_option(name, 3)
        assert kmb.get_option(name) == 3
        kmb.set_option(name, old)
    with pytest.raises(kmb.KmbError):
        kmb.set_option("no_such_option", 1)
    with pytest.raises(kmb.KmbError):
        kmb.set_option("chunk_bytes", 10)


def test_argument_validation_needs_no_device(kmb):
    h = C.c_void_p()
    rc = kmb.lib().kmb_index_create(0, None, None, 5, None, None, None, 0, C.byref(h))
    assert rc == kmb.KMB_ERR_BAD_ARG and b"null" in kmb.lib().kmb_last_error()
    a = np.zeros(4, np.int32)
    rc = kmb.lib().kmb_index_create(0, a.ctypes.data, a.ctypes.data, 0, None, None, None, 0, C.byref(h))
    assert rc == kmb.KMB_ERR_BAD_ARG
    rc = kmb.lib().kmb_mapper_create(None, 1, None, 1000, C.byref(h))
    assert rc == kmb.KMB_ERR_BAD_ARG
    rc = kmb.lib().kmb_codec_actg_from_bytes(0, a.ctypes.data, 5, a.ctypes.data)
    assert rc == kmb.KMB_ERR_BAD_ARG and b"multiple of 4" in kmb.lib().kmb_last_error()
    n, bad = C.c_uint64(), C.c_int64()
    rc = kmb.lib().kmb_hash_reads(0, a.ctypes.data, 4, a.ctypes.data, 1, 32, 0, None, 0, C.byref(n), C.byref(bad))
    assert rc == kmb.KMB_ERR_BAD_ARG and b"k=32" in kmb.lib().kmb_last_error()


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="this test is about hosts without a GPU")
def test_compute_fails_loudly_without_a_gpu(kmb, golden_lookup):
    """No CPU fallback anywhere on the product path."""
    from kmer_mapper_b200.encodings import ACTGTwoBitEncoding
    from kmer_mapper_b200.gpu_counter import GpuCounter
    from kmer_mapper_b200.mapper import in_graph_index, map_kmers_to_graph_index
    from kmer_mapper_b200.util import get_kmer_hashes_from_chunk_sequence
    assert kmb.device_count() == 0
    g = golden_lookup["gpucounter"]
    with pytest.raises(kmb.KmbError):
        map_kmers_to_graph_index(g["index"], g["max_node_id"], g["queries"])
    with pytest.raises(kmb.KmbError):
        in_graph_index(g["index"], g["queries"])
    with pytest.raises(kmb.KmbError):
        get_kmer_hashes_from_chunk_sequence("ACGTACGT", 3)
    with pytest.raises(kmb.KmbError):
        ACTGTwoBitEncoding.from_bytes(np.frombuffer(b"ACGT", np.uint8))
    c = GpuCounter.from_kmers_and_nodes(np.array([1, 2, 3], np.uint64), np.array([10, 11, 12]), 31)
    with pytest.raises(kmb.KmbError):
        c.initialize_cuda(2003)
    # the raw C ABI says the same
    a = np.zeros(4, np.int32)
    k = np.zeros(1, np.uint64)
    f = np.zeros(1, np.uint16)
    h = C.c_void_p()
    rc = kmb.lib().kmb_index_create(0, a.ctypes.data, a.ctypes.data, 4, a.ctypes.data, k.ctypes.data, f.ctypes.data, 1,
                                    C.byref(h))
    assert rc == kmb.KMB_ERR_CUDA and h.value is None


def test_parse_reads_c_abi_needs_no_gpu(kmb):
    """kmb_parse_reads is host code: counts first, exact capacities, carry-over of the incomplete tail."""
    lib = kmb.lib()
    text = np.frombuffer(b"@r1\nACGT\n+\nIIII\n@r2\nGG\n+\n@I\n@r3\nTT", dtype=np.uint8)
    nr, nb, used = C.c_uint64(), C.c_uint64(), C.c_uint64()
    rc = lib.kmb_parse_reads(text.ctypes.data, text.shape[0], 1, 0, 4, None, 0, None, 0, C.byref(nr), C.byref(nb), C.byref(used))
    assert (rc, nr.value, nb.value, used.value) == (0, 2, 6, 28)           # r3 is incomplete: left for the next chunk
    bases = np.zeros(6, np.uint8)
    offsets = np.zeros(3, np.int64)
    rc = lib.kmb_parse_reads(text.ctypes.data, text.shape[0], 1, 0, 4, bases.ctypes.data, 6, offsets.ctypes.data, 3,
                             C.byref(nr), C.byref(nb), C.byref(used))
    assert rc == 0 and bytes(bases) == b"ACGTGG" and list(offsets) == [0, 4, 6]
    rc = lib.kmb_parse_reads(text.ctypes.data, text.shape[0], 1, 0, 4, bases.ctypes.data, 5, offsets.ctypes.data, 3,
                             C.byref(nr), C.byref(nb), C.byref(used))
    assert rc == kmb.KMB_ERR_NOMEM
    bad = np.frombuffer(b"ACGT\nACGT\n+\nIIII\n", dtype=np.uint8)
    rc = lib.kmb_parse_reads(bad.ctypes.data, bad.shape[0], 1, 1, 1, None, 0, None, 0, C.byref(nr), C.byref(nb), C.byref(used))
    assert rc == kmb.KMB_ERR_BAD_ARG
    fa = np.frombuffer(b">a\nAC\nGT\n>b\n\n>c\nTTT", dtype=np.uint8)
    rc = lib.kmb_parse_reads(fa.ctypes.data, fa.shape[0], 0, 1, 1, None, 0, None, 0, C.byref(nr), C.byref(nb), C.byref(used))
    assert (rc, nr.value, nb.value, used.value) == (0, 3, 7, fa.shape[0])
    rc = lib.kmb_parse_reads(fa.ctypes.data, fa.shape[0], 0, 0, 1, None, 0, None, 0, C.byref(nr), C.byref(nb), C.byref(used))
    assert (rc, nr.value, nb.value, used.value) == (0, 2, 4, 13)           # record c may continue in the next chunk


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "kmer_mapper_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "kmer_oracle" not in text, fn
