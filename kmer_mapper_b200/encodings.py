"""Drop-in for the reference's legacy 2-bit codec kmer_mapper/encodings.py, computed on the GPU.

Same classes, method names and results (including the quirks SURVEY.md appendix B #14 lists: A,C,T,G
order, ``& 31`` aliasing, pair-wise zeroing of non-ACTG bytes, lower-case output, ``size % 4``
assertion).  Note that this codec's A,C,T,G = 0,1,2,3 order is NOT the order the mapping path
hashes with (A,C,G,T, util.py:72-73); the reference itself never calls it from the CLI.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import check, lib
from .device import current_device


def _run(fn, src: np.ndarray, out: np.ndarray, *extra):
    _lib.require_device()
    check(fn(current_device(), src.ctypes.data, *extra, out.ctypes.data))
    return out


class BaseEncoding:
    """Basic ASCII byte encoding (encodings.py:4-23)."""

    @classmethod
    def from_string(cls, sequence):
        return np.array([ord(c) for c in sequence], dtype=np.uint8)

    @classmethod
    def from_bytes(cls, sequence):
        return sequence

    @classmethod
    def to_bytes(cls, sequence):
        return sequence

    @classmethod
    def to_string(cls, byte_sequence):
        return "".join(chr(b) for b in byte_sequence)


class ACTGTwoBitEncoding:
    letters = ["A", "C", "T", "G"]
    bitcodes = ["00", "01", "10", "11"]
    reverse = np.array([1, 3, 20, 7], dtype=np.uint8)
    # The helper tables and classmethods of encodings.py:30-42 (API surface only: from_bytes above does the same
    # work in one kernel and does not go through them).  Built with an int64 index: the reference's own
    # ``256*reverse[...]`` overflows uint8 under numpy >= 2 (encodings.py:31).
    _lookup_2bytes_to_4bits = np.zeros(256 * 256, dtype=np.uint8)
    _lookup_2bytes_to_4bits[256 * reverse.astype(np.int64)[np.arange(4)[:, None]] + reverse[np.arange(4)]] = \
        np.arange(4)[:, None] * 4 + np.arange(4)
    _shift_4bits = (4 * np.arange(2, dtype=np.uint8))
    _shift_2bits = 2 * np.arange(4, dtype=np.uint8)

    @classmethod
    def convert_2bytes_to_4bits(cls, two_bytes):
        """encodings.py:36-38: an aligned pair of (``& 31``-masked) bases as one uint16 -> its 4-bit code."""
        assert two_bytes.dtype == np.uint16, two_bytes.dtype
        return cls._lookup_2bytes_to_4bits[two_bytes]

    @classmethod
    def join_4bits_to_byte(cls, four_bits):
        """encodings.py:40-42: rows of two 4-bit codes -> one byte each, first code in the low nibble."""
        return np.bitwise_or.reduce(four_bits << cls._shift_4bits, axis=1)

    @classmethod
    def complement(cls, char):
        """encodings.py:44-48: XOR 0b10101010 on every byte (A<->T, C<->G in A,C,T,G order)."""
        a = np.ascontiguousarray(char)
        flat = a.reshape(-1).view(np.uint8)
        out = np.empty_like(flat)
        if flat.size:
            _run(lib().kmb_codec_complement, flat, out, flat.size)
        return out.view(a.dtype).reshape(a.shape)

    @classmethod
    def from_bytes(cls, sequence):
        """encodings.py:51-59."""
        assert sequence.dtype == np.uint8
        assert sequence.size % 4 == 0, sequence.size
        seq = np.ascontiguousarray(sequence).reshape(-1)
        out = np.empty(seq.size // 4, dtype=np.uint8)
        if seq.size:
            _run(lib().kmb_codec_actg_from_bytes, seq, out, seq.size)
        return out

    @classmethod
    def from_string(cls, string):
        byte_repr = np.array([ord(c) for c in string], dtype=np.uint8)
        return cls.from_bytes(byte_repr)

    @classmethod
    def to_string(cls, bits):
        byte_repr = cls.to_bytes(bits)
        return "".join(chr(b) for b in byte_repr)

    @classmethod
    def to_bytes(cls, sequence):
        """encodings.py:70-75: four lower-case letters per packed byte."""
        assert sequence.dtype == np.uint8
        seq = np.ascontiguousarray(sequence).reshape(-1)
        out = np.empty(seq.size * 4, dtype=np.uint8)
        if seq.size:
            _run(lib().kmb_codec_to_bytes, seq, out, seq.size)
        return out


class SimpleEncoding(ACTGTwoBitEncoding):
    # helper table and classmethods of encodings.py:79-93 (API surface; from_bytes is one kernel)
    _lookup_byte_to_2bits = np.zeros(256, dtype=np.uint8)
    _lookup_byte_to_2bits[[97, 65]] = 0
    _lookup_byte_to_2bits[[99, 67]] = 1
    _lookup_byte_to_2bits[[116, 84]] = 2
    _lookup_byte_to_2bits[[103, 71]] = 3
    _shift_2bits = 2 * np.arange(4, dtype=np.uint8)

    @classmethod
    def convert_byte_to_2bits(cls, one_byte):
        """encodings.py:85-88."""
        assert one_byte.dtype == np.uint8, one_byte.dtype
        return cls._lookup_byte_to_2bits[one_byte]

    @classmethod
    def join_2bits_to_byte(cls, two_bits_vector):
        """encodings.py:90-92: rows of four 2-bit codes -> one byte each, first code in the low bits."""
        return np.bitwise_or.reduce(two_bits_vector << cls._shift_2bits, axis=-1)

    @classmethod
    def from_bytes(cls, sequence):
        """encodings.py:96-102: per-byte table a/A,c/C,t/T,g/G -> 0,1,2,3, anything else -> 0."""
        assert sequence.dtype == np.uint8
        assert sequence.size % 4 == 0, sequence.size
        seq = np.ascontiguousarray(sequence).reshape(-1)
        out = np.empty(seq.size // 4, dtype=np.uint8)
        if seq.size:
            _run(lib().kmb_codec_simple_from_bytes, seq, out, seq.size)
        return out


def twobit_swap(number):
    """encodings.py:104-112: reverse the order of all 2-bit groups of every word."""
    a = np.ascontiguousarray(number)
    if a.dtype.itemsize not in (1, 2, 4, 8):
        raise ValueError("twobit_swap: unsupported dtype %s" % a.dtype)
    flat = a.reshape(-1)
    out = np.empty_like(flat)
    if flat.size:
        _lib.require_device()
        check(lib().kmb_codec_twobit_swap(current_device(), flat.ctypes.data, flat.size, a.dtype.itemsize,
                                          out.ctypes.data))
    return out.reshape(a.shape)
