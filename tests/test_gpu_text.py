"""Record parsing on the device (csrc/kmb_textparse.cuh) against the native host parser and an independent Python
parser, record by record, and the text route end to end against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import c_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kmb():
    from kmer_mapper_b200 import _lib
    _lib.require_device()
    return _lib


def _device_parse(kmb, text: bytes, fmt: int):
    a = np.frombuffer(text, dtype=np.uint8)
    bases = np.zeros(len(text) + 16, dtype=np.uint8)
    offsets = np.zeros(len(text) + 2, dtype=np.int64)
    nr, nb = C.c_uint64(), C.c_uint64()
    rc = kmb.lib().kmb_parse_text_device(0, a.ctypes.data if len(text) else None, len(text), fmt, bases.ctypes.data, bases.shape[0],
                                         offsets.ctypes.data, offsets.shape[0], C.byref(nr), C.byref(nb))
    if rc != 0:
        return rc, None
    off = offsets[:nr.value + 1]
    return 0, [bytes(bases[off[i]:off[i + 1]]) for i in range(nr.value)]


def _python_records(text: bytes, fmt: int):
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    reads = []
    if fmt == 1:
        while lines and lines[-1] == b"" and len(lines) % 4 != 0:
            lines.pop()
        if len(lines) % 4 == 3:
            lines.append(b"")
        for i in range(0, len(lines), 4):
            reads.append(lines[i + 1].rstrip(b"\r"))
    else:
        cur = None
        for ln in lines:
            ln = ln.rstrip(b"\r")
            if ln.startswith(b">"):
                if cur is not None:
                    reads.append(cur)
                cur = b""
            elif cur is not None:
                cur += ln
        if cur is not None:
            reads.append(cur)
    return reads


CASES = [
    (1, b"@a\nACGT\n+\n@III\n@b\nGG\n+\n@I\n@c\nT\n+\nI"),                 # quality lines starting with '@', no final newline
    (1, b"@a\r\nACGT\r\n+\r\nIIII\r\n@b\r\nAC\r\n+\r\nII\r\n"),               # CRLF
    (1, b"@a\nACGT\n+\nIIII\n@empty\n\n+\n\n@c\nTT\n+\nII\n"),               # an empty read
    (1, b"@a\nACGT\n+\nIIII\n@last\n\n+\n"),                                 # empty read at the very end, empty quality line
    (1, b"@a\nACGT\n+\nIIII\n\n\n"),                                         # blank lines after the last record
    (0, b">r1 desc\r\nACGT\r\nAC\r\n>r2\n\n>r3\nGGGTTT\n>r4\nA"),               # CRLF, empty read, multi-line, no final newline
    (0, b">only\n"),
    (0, b">x\nAC\n\n\nGT\n>y\n"),
    (0, b""),
    (1, b""),
]


@pytest.mark.parametrize("fmt,text", CASES)
def test_device_parser_hand_made_cases(kmb, fmt, text):
    rc, got = _device_parse(kmb, text, fmt)
    assert rc == 0
    assert got == _python_records(text, fmt)


def test_device_parser_rejects_malformed_text(kmb):
    for fmt, text in ((1, b"ACGT\nACGT\n+\nIIII\n"), (1, b"@a\nACGT\nIIII\nIIII\n"), (1, b"@a\nACGT\n+\nIIII\n@b\nAC\n"),
                      (0, b"ACGT\n>r\nAC\n")):
        rc, _ = _device_parse(kmb, text, fmt)
        assert rc == kmb.KMB_ERR_BAD_ARG, (fmt, text)


@pytest.mark.parametrize("fmt", [0, 1])
def test_device_parser_equals_host_parser_on_large_hostile_text(kmb, fmt):
    """Ragged reads, odd FASTA wrapping, hostile quality characters, many empty lines (more lines than the first
    capacity guess: the parse is repeated with room for the worst case), > 1000 lines per block boundary cases."""
    from kmer_mapper_b200 import synthetic
    rng = np.random.default_rng(3)
    g = synthetic.make_genome(100_000, 5)
    bases, offsets = synthetic.make_reads(g, 30_000, 120, seed=6, ragged=True)
    want = [bytes(bases[offsets[r]:offsets[r + 1]]) for r in range(30_000)]
    qual = np.frombuffer(b"@+>I#", np.uint8)
    parts = []
    for r, seq in enumerate(want):
        if fmt == 1:
            parts.append(b"@read%d +x\n%s\n+\n%s\n" % (r, seq, bytes(rng.choice(qual, size=len(seq)))))
        else:
            w = int(rng.integers(1, 90))
            parts.append(b">read%d\n" % r + b"".join(seq[i:i + w] + b"\n" for i in range(0, len(seq), w)))
            if r % 50 == 0:
                parts.append(b"\n" * 40)       # runs of empty lines inside a record
    text = b"".join(parts)
    rc, got = _device_parse(kmb, text, fmt)
    assert rc == 0 and got == want
    if fmt == 0:       # a text that is mostly newlines exceeds one line per 8 bytes
        text2 = b">r\n" + b"A\n" * 200_000
        rc, got = _device_parse(kmb, text2, 0)
        assert rc == 0 and got == [b"A" * 200_000]


def test_map_text_route_vs_oracle_in_many_chunks(kmb, tmp_path):
    """Mapper.map_text over whole-record windows of FASTQ / FASTA / FASTQ.gz text (reader.text_chunks) == oracle."""
    from kmer_mapper_b200 import synthetic
    from kmer_mapper_b200.device import DeviceIndex, Mapper
    from kmer_mapper_b200.reader import ParallelGzip, open_reads
    k = 31
    g = synthetic.make_genome(300_000, 11)
    idx = synthetic.make_index(g, 40_000, k, 30_000, 200_003, 12, n_hot_nodes=1100)
    bases, offsets = synthetic.make_reads(g, 12_000, 150, seed=13, n_rate=0.01, lower_rate=0.3, ragged=True)
    want, n_kmers = c_oracle.map_reads(idx, idx.max_node_id(), bases, offsets, k, n_threads=4)
    synthetic.write_fasta(str(tmp_path / "r.fa"), bases, offsets, line_width=70)
    synthetic.write_fastq(str(tmp_path / "r.fq"), bases, offsets)
    synthetic.write_fastq(str(tmp_path / "r.fq.gz"), bases, offsets, members=9)
    di = DeviceIndex.from_index(idx)
    old = ParallelGzip.BATCH_BYTES, ParallelGzip.MAX_MEMBER_BYTES
    ParallelGzip.BATCH_BYTES, ParallelGzip.MAX_MEMBER_BYTES = 1_200_000, 1_000_000
    try:
        for name in ("r.fa", "r.fq", "r.fq.gz"):
            for chunk in (50_000, 700_000, 1 << 26):
                m = Mapper(di)
                reads = open_reads(str(tmp_path / name), n_threads=3)
                n = 0
                for tc in reads.text_chunks(min_chunk_size=chunk):
                    m.map_text(tc, reads.format, k)
                    n += 1
                got = m.counts()
                assert np.array_equal(got, want), (name, chunk)
                assert m.stats()[0] == n_kmers
                assert chunk > 100_000 or n >= 3
                m.close()
                reads.close()
    finally:
        ParallelGzip.BATCH_BYTES, ParallelGzip.MAX_MEMBER_BYTES = old
    # an invalid base is reported through the usual path
    from kmer_mapper_b200._lib import InvalidBaseError
    m = Mapper(di)
    m.map_text(b">r1\nACGTACGTACGTACGTACGTACGTACGTACGTACGTRACGT\n", "fasta", k)
    with pytest.raises(InvalidBaseError):
        m.sync()
