"""Bit-level primitives of the kernels (kmb_core.cuh), compiled for the host with g++ and checked
against Python integers: exact u64 % modulo (Barrett), SWAR 2-bit encoding, window extraction,
reverse complement, directory word packing.  CPU-only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_harness", "kmb_core_host.cpp")
OUT = os.path.join(HERE, "host_harness", "libkmb_core_host.so")
HDR = os.path.join(os.path.dirname(HERE), "kmer_mapper_b200", "csrc", "kmb_core.cuh")


@pytest.fixture(scope="module")
def core():
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", SRC, "-o", OUT])
    lib = C.CDLL(OUT)
    lib.h_window.restype = C.c_uint64
    lib.h_window.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int]
    lib.h_revcomp.restype = C.c_uint64
    lib.h_revcomp.argtypes = [C.c_uint64, C.c_int]
    lib.h_chain_line.restype = C.c_uint64
    lib.h_chain_line.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
    lib.h_chain_slot.restype = C.c_uint32
    lib.h_chain_extra_lines.restype = C.c_uint32
    lib.h_locate.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.h_sector_header.restype = C.c_uint32
    lib.h_header_count.restype = C.c_uint32
    lib.h_log_bin.restype = C.c_uint32
    lib.h_encode16.restype = C.c_uint32
    return lib


def test_barrett_divmod_is_exact(core):
    rng = np.random.default_rng(1)
    edge = np.array([0, 1, 2, 2 ** 32 - 1, 2 ** 32, 2 ** 62 - 1, 2 ** 62, 2 ** 63, 2 ** 64 - 1, 2 ** 64 - 2], dtype=np.uint64)
    for d in [1, 2, 3, 21, 97, 2003, 65537, 2_000_003, 452_930_477, 1_000_000_007, 2 ** 31 - 1, 2 ** 32 - 1,
              2 ** 32 - 5, 4_000_000_007 % (2 ** 32)]:
        n = np.concatenate([edge, rng.integers(0, 2 ** 64, size=20000, dtype=np.uint64),
                            rng.integers(0, 4 ** 31, size=20000, dtype=np.uint64),
                            # multiples of d and their neighbours: where a one-off quotient would show
                            (rng.integers(0, (2 ** 64 - 1) // d, size=5000, dtype=np.uint64) * np.uint64(d)),
                            (rng.integers(1, (2 ** 64 - 1) // d, size=5000, dtype=np.uint64) * np.uint64(d)) - np.uint64(1)])
        q = np.zeros_like(n)
        r = np.zeros_like(n)
        core.h_divmod(n.ctypes.data_as(C.c_void_p), C.c_int64(n.shape[0]), C.c_uint64(d), q.ctypes.data_as(C.c_void_p),
                      r.ctypes.data_as(C.c_void_p))
        assert np.array_equal(r, n % np.uint64(d)), d
        assert np.array_equal(q, n // np.uint64(d)), d


def _encode_ref(b, n_to_a):
    codes, inv = 0, 0
    for j, ch in enumerate(b):
        c = {65: 0, 97: 0, 67: 1, 99: 1, 71: 2, 103: 2, 84: 3, 116: 3}.get(ch)
        if c is None and n_to_a and ch == 78:
            c = 0
        if c is None:
            inv |= 1 << j
            c = None
        codes |= (c or 0) << (2 * j)
    return codes, inv


def test_swar_encode_all_bytes(core):
    # every byte value in every lane position, both N policies
    for n_to_a in (0, 1):
        for lane in range(16):
            for v in range(256):
                b = bytearray(b"ACGTacgtACGTacgt")
                b[lane] = v
                inv = C.c_uint32(0)
                got = core.h_encode16(bytes(b), n_to_a, C.byref(inv))
                want, winv = _encode_ref(b, n_to_a)
                assert inv.value == winv, (n_to_a, lane, v)
                # codes of invalid lanes are unspecified; compare the valid ones
                mask = 0
                for j in range(16):
                    if not (winv >> j) & 1:
                        mask |= 3 << (2 * j)
                assert (got & mask) == (want & mask), (n_to_a, lane, v)
    rng = np.random.default_rng(2)
    for _ in range(2000):
        b = bytes(rng.choice(np.frombuffer(b"ACGTacgtNn", np.uint8), size=16))
        inv = C.c_uint32(0)
        got = core.h_encode16(b, 1, C.byref(inv))
        want, winv = _encode_ref(b, 1)
        assert inv.value == winv
        if winv == 0:
            assert got == want


def test_window_extraction_is_the_reference_hash(core):
    # tests/test_hashing.py:13-26 convention: first base of the window in the lowest bits
    rng = np.random.default_rng(3)
    codes = rng.integers(0, 4, size=64)
    lo = sum(int(c) << (2 * j) for j, c in enumerate(codes[:32]))
    hi = sum(int(c) << (2 * j) for j, c in enumerate(codes[32:]))
    for k in (1, 2, 15, 16, 17, 21, 30, 31):
        for i in range(32):
            want = sum(int(codes[i + j]) << (2 * j) for j in range(k))
            assert core.h_window(lo, hi, i, k) == want, (k, i)


def test_revcomp(core):
    rng = np.random.default_rng(4)
    for k in (1, 3, 15, 21, 31):
        for _ in range(200):
            codes = rng.integers(0, 4, size=k)
            x = sum(int(c) << (2 * j) for j, c in enumerate(codes))
            rc = [3 - int(c) for c in codes[::-1]]
            want = sum(c << (2 * j) for j, c in enumerate(rc))
            assert core.h_revcomp(x, k) == want
            assert core.h_revcomp(want, k) == x


def test_sector_chain_geometry_headers_bins_and_addressing(core):
    # two slots per 32-byte sector; entry s of a chain lives in the main sector (s < 2) or in overflow sector
    # ovf_base + (s-2)//2; every (sector, slot) pair is used exactly once
    for n_total in (0, 1, 2, 3, 4, 5, 17, 1234):
        extra = core.h_chain_extra_lines(n_total)
        assert extra == max(0, -(-n_total // 2) - 1)
        seen = set()
        for s in range(n_total):
            line = core.h_chain_line(77, 1000, s)
            slot = core.h_chain_slot(s)
            assert 0 <= slot < 2
            assert line == 77 if s < 2 else 1000 <= line < 1000 + extra
            seen.add((line, slot))
        assert len(seen) == n_total
    # sector headers: a count of 0..2, or the chain bit plus the index of the next sector
    for remaining in range(0, 9):
        hdr = core.h_sector_header(remaining, 12345)
        assert core.h_header_count(hdr) == min(remaining, 2)
        assert (hdr == (0x80000000 | 12345)) == (remaining > 2)
    # log bins: node ranges of 2^shift, the last bin takes the rest
    assert [core.h_log_bin(n, 4) for n in (0, 15, 16, 127, 128, 4000)] == [0, 0, 1, 7, 7, 7]
    # addressing: sector / filter word stay in range for any table size, the mask has one or two bits, and
    # k-mer-like keys spread evenly (multiply-shift hashing)
    rng = np.random.default_rng(6)
    out = (C.c_uint32 * 3)()
    for n_main, n_words in ((1, 1), (7, 3), (1000, 77), (250_000_000, 16_000_000), (2 ** 31 - 1, 2 ** 24)):
        for key in [0, 1, 2 ** 62 - 1, 2 ** 64 - 1] + [int(x) for x in rng.integers(0, 2 ** 62, size=300, dtype=np.uint64)]:
            core.h_locate(key, n_main, n_words, 1, out)
            assert out[0] < n_main and out[1] < n_words and bin(out[2]).count("1") in (1, 2)
            core.h_locate(key, n_main, n_words, 0, out)
            assert bin(out[2]).count("1") == 1
    keys = rng.integers(0, 4 ** 31, size=20_000, dtype=np.uint64)
    sec = np.zeros(20_000, np.int64)
    for i, k in enumerate(keys):
        core.h_locate(int(k), 1000, 64, 1, out)
        sec[i] = out[0]
    counts = np.bincount(sec, minlength=1000)
    assert counts.max() < 60 and counts.min() > 2        # mean 20 per sector
