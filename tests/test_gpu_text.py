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


# ---- gzip members inflated on the device (csrc/kmb_gzdev.cuh, kmb_mapper_map_gz) ------------------------------------
def _gz_members(blobs, levels):
    import zlib
    out = []
    for i, blob in enumerate(blobs):
        c = zlib.compressobj(levels[i % len(levels)], zlib.DEFLATED, 31)
        out.append(c.compress(blob) + c.flush())
    return out


def _fastq_text(bases, offsets, rng=None, lo=0, hi=None):
    hi = len(offsets) - 1 if hi is None else hi
    parts = []
    for r in range(lo, hi):
        seq = bytes(bases[offsets[r]:offsets[r + 1]])
        q = b"I" * len(seq) if rng is None else bytes(rng.integers(33, 74, size=len(seq), dtype=np.uint8))
        parts.append(b"@read%d some text\n%s\n+\n%s\n" % (r, seq, q))
    return b"".join(parts)


def _cut(text, sizes):
    """text cut into pieces of the given sizes (cycled), anywhere -- not at record boundaries."""
    out, p, i = [], 0, 0
    while p < len(text):
        n = sizes[i % len(sizes)]
        out.append(text[p:p + n])
        p += n
        i += 1
    return out


@pytest.fixture(scope="module")
def gz_case():
    from kmer_mapper_b200 import synthetic
    from kmer_mapper_b200.device import DeviceIndex
    k = 31
    g = synthetic.make_genome(300_000, 21)
    idx = synthetic.make_index(g, 40_000, k, 30_000, 200_003, 22, n_hot_nodes=1100)
    bases, offsets = synthetic.make_reads(g, 20_000, 150, seed=23, n_rate=0.01, lower_rate=0.2, ragged=True)
    want, n_kmers = c_oracle.map_reads(idx, idx.max_node_id(), bases, offsets, k, n_threads=4)
    return dict(k=k, idx=idx, di=DeviceIndex.from_index(idx), bases=bases, offsets=offsets, want=want, n_kmers=n_kmers)


def _gz_stats(kmb):
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    kmb.check(kmb.lib().kmb_gz_device_stats(C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


@pytest.mark.parametrize("name,sizes,levels,quality", [
    ("bgzf_like", [65_280], [6], "random"),                       # dynamic Huffman blocks, few matches
    ("constant_quality", [200_000], [1, 9], "const"),             # long overlapping matches (distance 1)
    ("stored_and_fixed", [3_000, 40, 70_000, 1], [0, 1, 6], "const"),   # stored blocks, tiny members (fixed Huffman), 1-byte members
    ("large_members", [1_500_000], [6], "random"),
])
def test_map_gz_on_device_vs_oracle(kmb, gz_case, name, sizes, levels, quality):
    """Multi-member FASTQ.gz -> counts with every member inflated by a GPU warp == oracle; the device did it all (no
    batch redone by the host decoder), in one batch and in many small ones (records straddle members and batches)."""
    from kmer_mapper_b200.device import Mapper
    rng = np.random.default_rng(5) if quality == "random" else None
    text = _fastq_text(gz_case["bases"], gz_case["offsets"], rng)
    gz = b"".join(_gz_members(_cut(text, sizes), levels))
    for batch in (256 << 20, 1 << 20):
        kmb.set_option("gz_device_batch_bytes", batch)
        kmb.set_option("gz_device_max_mean_member_bytes", 16 << 20)     # also the large members go to the device here
        try:
            before = _gz_stats(kmb)
            m = Mapper(gz_case["di"])
            resume = m.map_gz(gz, "fastq", gz_case["k"])
            got = m.counts()
            after = _gz_stats(kmb)
            assert resume == len(gz)
            assert np.array_equal(got, gz_case["want"]), (name, batch)
            assert m.stats()[0] == gz_case["n_kmers"]
            assert after[1] - before[1] == len(text) and after[2] == before[2], (before, after)
            m.close()
        finally:
            kmb.set_option("gz_device_batch_bytes", 256 << 20)
            kmb.set_option("gz_device_max_mean_member_bytes", 512 << 10)
    if name == "large_members":      # by default a file of few large members is left to the host decoders
        m = Mapper(gz_case["di"])
        assert m.map_gz(gz, "fastq", gz_case["k"]) == 0 and not m.counts().any()
        m.close()


def test_map_gz_fasta_and_shards(kmb, gz_case):
    """FASTA.gz (multi-line records), and the batches dealt to two shards: the two count arrays add up to the oracle's."""
    from kmer_mapper_b200.device import Mapper
    b, o = gz_case["bases"], gz_case["offsets"]
    text = b"".join(b">r%d\n" % r + b"".join(bytes(b[i:min(i + 61, o[r + 1])]) + b"\n" for i in range(o[r], o[r + 1], 61)) for r in range(len(o) - 1))
    gz = b"".join(_gz_members(_cut(text, [50_000, 7_000]), [6, 1]))
    kmb.set_option("gz_device_batch_bytes", 1 << 20)
    try:
        total = np.zeros_like(gz_case["want"])
        n = 0
        for shard in range(2):
            m = Mapper(gz_case["di"])
            assert m.map_gz(gz, "fasta", gz_case["k"], shard_index=shard, shard_count=2) == len(gz)
            c = m.counts()
            assert c.any()
            total += c
            n += m.stats()[0]
            m.close()
        assert np.array_equal(total, gz_case["want"]) and n == gz_case["n_kmers"]
    finally:
        kmb.set_option("gz_device_batch_bytes", 256 << 20)


def test_map_gz_hands_over_to_the_host_decoders(kmb, gz_case, tmp_path):
    """A member too large for one warp (gz_device_max_member_bytes) ends the device's part: resume offset = a member
    start, the host decoders continue there and every record is mapped exactly once (CLI route, 1 and 2 shards).  A plain
    single-member .gz is left to the host entirely."""
    from kmer_mapper_b200.command_line_interface import map_file_text
    from kmer_mapper_b200.device import Mapper
    from kmer_mapper_b200.reader import open_reads
    text = _fastq_text(gz_case["bases"], gz_case["offsets"])
    n = len(text)
    a, b = n // 2, n // 2 + 1_300_000
    assert n > 3_000_000
    pieces = _cut(text[:a], [150_000]) + [text[a:b]] + _cut(text[b:], [100_000])
    members = _gz_members(pieces, [6])
    path = str(tmp_path / "mixed.fq.gz")
    with open(path, "wb") as f:
        f.write(b"".join(members))
    from kmer_mapper_b200.reader import ParallelGzip
    old = ParallelGzip.BATCH_BYTES, ParallelGzip.MAX_MEMBER_BYTES
    ParallelGzip.BATCH_BYTES, ParallelGzip.MAX_MEMBER_BYTES = 1_500_000, 1_400_000     # several host blocks, too
    kmb.set_option("gz_device_max_member_bytes", 1_000_000)
    kmb.set_option("gz_device_batch_bytes", 1 << 20)
    try:
        for world in (1, 2):
            total = np.zeros_like(gz_case["want"])
            for rank in range(world):
                m = Mapper(gz_case["di"])
                reads = open_reads(path, n_threads=3)
                n_chunks, host_from = map_file_text(m, reads, gz_case["k"], rank=rank, world_size=world, chunk_bytes=400_000)
                assert host_from is not None and 0 < host_from < os.path.getsize(path)
                total += m.counts()
                m.close()
                reads.close()
            assert np.array_equal(total, gz_case["want"]), world
        # single member: nothing for the device
        single = str(tmp_path / "single.fq.gz")
        with open(single, "wb") as f:
            f.write(_gz_members([text], [6])[0])
        m = Mapper(gz_case["di"])
        reads = open_reads(single, n_threads=3)
        n_chunks, host_from = map_file_text(m, reads, gz_case["k"])
        assert host_from == 0 and n_chunks >= 1
        assert np.array_equal(m.counts(), gz_case["want"])
        m.close()
        reads.close()
    finally:
        ParallelGzip.BATCH_BYTES, ParallelGzip.MAX_MEMBER_BYTES = old
        kmb.set_option("gz_device_max_member_bytes", 16 << 20)
        kmb.set_option("gz_device_batch_bytes", 256 << 20)


def test_map_gz_corrupt_member_is_an_error_and_wrong_crc_is_caught(kmb, gz_case):
    from kmer_mapper_b200._lib import KmbError
    from kmer_mapper_b200.device import Mapper
    text = _fastq_text(gz_case["bases"], gz_case["offsets"], hi=4000)
    members = _gz_members(_cut(text, [100_000]), [6])
    # a flipped CRC byte in one trailer: the device decodes fine, the CRC comparison fails, the host decoder fails too
    bad = bytearray(b"".join(members))
    pos = len(members[0]) + len(members[1]) - 8
    bad[pos] ^= 0x55
    m = Mapper(gz_case["di"])
    with pytest.raises(KmbError):
        m.map_gz(bytes(bad), "fastq", gz_case["k"])
    m.close()
    # garbage in the middle of a deflate stream
    bad = bytearray(b"".join(members))
    mid = len(members[0]) + len(members[1]) // 2
    bad[mid:mid + 64] = bytes(range(64))
    m = Mapper(gz_case["di"])
    with pytest.raises(KmbError):
        m.map_gz(bytes(bad), "fastq", gz_case["k"])
    m.close()
    # not a gzip file at all
    m = Mapper(gz_case["di"])
    with pytest.raises(KmbError):
        m.map_gz(b"@r\nACGT\n+\nIIII\n" * 10, "fastq", gz_case["k"])
    m.close()
