"""Host-side logic that needs no GPU: index data model, ragged sequences, reader, CLI surface, sharding."""
import argparse
import gzip
import os

import numpy as np
import pytest

from kmer_mapper_b200 import distributed, synthetic
from kmer_mapper_b200.kmer_index import KmerIndex
from kmer_mapper_b200.reader import open_reads
from kmer_mapper_b200.sequences import RaggedSequence, as_ragged
from oracle import oracle


def test_kmer_index_construction_matches_oracle_restatement(tmp_path):
    rng = np.random.default_rng(1)
    keys = rng.integers(0, 4 ** 10, size=5000, dtype=np.uint64)
    nodes = rng.integers(0, 900, size=5000)
    a = KmerIndex.from_flat_kmers(hashes=keys, nodes=nodes, modulo=1009)
    a.convert_to_int32()
    b = oracle.index_from_flat_kmers(keys, nodes, 1009)
    for attr in ("_hashes_to_index", "_n_kmers", "_nodes", "_kmers", "_frequencies"):
        assert np.array_equal(getattr(a, attr), getattr(b, attr)), attr
        assert getattr(a, attr).dtype == getattr(b, attr).dtype
    assert a.max_node_id() == b.max_node_id()
    # the reference's unit-test shape (tests/test_mapping.py:31-44): ACT,CTT,cCG,ATT -> nodes 0..3, modulo 21
    def h(s):
        return sum("ACGT".index(c) << (2 * j) for j, c in enumerate(s.upper()))
    idx = KmerIndex.from_flat_kmers(hashes=np.array([h(s) for s in ("ACT", "CTT", "cCG", "ATT")], np.uint64),
                                    nodes=np.arange(4), modulo=21)
    idx.convert_to_int32()
    assert idx.get(h("ccg"))[0][0] == 2          # tests/test_mapping.py:40
    assert idx.get(h("GGG")) is None
    # .npz round trip, with and without the ".npz" suffix in the name (KmerIndex.from_file tries both)
    p = str(tmp_path / "index")
    a.to_file(p)
    for name in (p, p + ".npz"):
        c = KmerIndex.from_file(name)
        c.convert_to_int32()
        c.remove_ref_offsets()
        for attr in ("_hashes_to_index", "_n_kmers", "_nodes", "_kmers", "_frequencies"):
            assert np.array_equal(getattr(a, attr), getattr(c, attr))
        assert c._modulo == 1009
    np.savez(str(tmp_path / "bad.npz"), kmers=keys)
    with pytest.raises(KeyError):
        KmerIndex.from_file(str(tmp_path / "bad.npz"))


def test_as_ragged_inputs():
    r = as_ragged([b"ACG", "TT", b""])
    assert len(r) == 3 and list(r.offsets) == [0, 3, 5, 5] and bytes(r.bases) == b"ACGTT"
    assert bytes(as_ragged("ACGT").bases) == b"ACGT"
    b = np.frombuffer(b"ACGTAC", np.uint8)
    assert list(as_ragged(b).offsets) == [0, 6]
    r2 = as_ragged((b, np.array([0, 2, 6])))
    assert bytes(r2[1]) == b"GTAC" and list(r2.lengths) == [2, 4]

    class FakeBnp:  # bionumpy-style ragged array: .ravel() + .shape[-1] row lengths
        shape = (2, np.array([2, 4]))

        def ravel(self):
            return b

    r3 = as_ragged(FakeBnp())
    assert list(r3.offsets) == [0, 2, 6]


def _python_parse(path):
    """Independent, line-by-line parser (the cross-check SURVEY.md 8c asks for)."""
    opener = gzip.open if path.endswith(".gz") else open
    reads = []
    with opener(path, "rb") as f:
        lines = f.read().split(b"\n")
    name = path[:-3] if path.endswith(".gz") else path
    if name.endswith((".fq", ".fastq")):
        i = 0
        while i + 1 < len(lines):
            if lines[i].startswith(b"@") and i + 3 < len(lines) + 1:
                reads.append(lines[i + 1].rstrip(b"\r"))
                i += 4
            else:
                break
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


@pytest.mark.parametrize("suffix,writer", [(".fa", "fasta"), (".fasta.gz", "fasta"), (".fq", "fastq"), (".fq.gz", "fastq"),
                                           (".fastq", "fastq")])
def test_reader_matches_independent_parser(tmp_path, suffix, writer):
    g = synthetic.make_genome(50_000, 3)
    bases, offsets = synthetic.make_reads(g, 700, 150, seed=4, n_rate=0.02, lower_rate=0.3, ragged=True)
    path = str(tmp_path / ("reads" + suffix))
    if writer == "fasta":
        synthetic.write_fasta(path, bases, offsets, line_width=60 if suffix == ".fa" else 0)
    else:
        synthetic.write_fastq(path, bases, offsets, members=5)
    want = _python_parse(path)
    assert len(want) == 700
    assert b"".join(want) == bytes(bases)
    for chunk_size in (97, 4096, 10_000_000):
        got = []
        n_chunks = 0
        for chunk in open_reads(path, pinned=False).read_chunks(min_chunk_size=chunk_size):
            s = chunk.sequence
            assert s.offsets[0] == 0 and s.offsets[-1] == s.bases.shape[0]
            got += [bytes(s[i]) for i in range(len(s))]
            n_chunks += 1
        assert got == want, (suffix, chunk_size)
        if chunk_size == 97 and not suffix.endswith(".gz"):   # a gzip member is never split into smaller blocks
            assert n_chunks > 10


@pytest.mark.parametrize("fmt", ["fastq", "fasta"])
def test_reader_parallel_pieces_resynchronise_on_hostile_text(tmp_path, fmt):
    """Chunks above 1 MB are cut into per-thread pieces at record boundaries found by resynchronisation; make
    that hard: quality lines that start with '@' or '+' and contain '>' , FASTA wrapped at odd widths."""
    rng = np.random.default_rng(9)
    g = synthetic.make_genome(100_000, 5)
    bases, offsets = synthetic.make_reads(g, 25_000, 160, seed=6, ragged=True)
    want = [bytes(bases[offsets[r]:offsets[r + 1]]) for r in range(25_000)]
    path = str(tmp_path / ("hostile." + ("fq" if fmt == "fastq" else "fa")))
    qual_alphabet = np.frombuffer(b"@+>I#", np.uint8)
    with open(path, "wb") as f:
        for r, seq in enumerate(want):
            if fmt == "fastq":
                q = bytes(rng.choice(qual_alphabet, size=len(seq)))
                f.write(b"@read%d +x\n%s\n+\n%s\n" % (r, seq, q))
            else:
                w = int(rng.integers(1, 90))
                f.write(b">read%d\n" % r)
                for i in range(0, len(seq), w):
                    f.write(seq[i:i + w] + b"\n")
    assert os.path.getsize(path) > 2_000_000
    for chunk_size in (1_100_000, 1_700_000, 50_000_000):
        for n_threads in (1, 3, 16):
            got = []
            for chunk in open_reads(path, pinned=False, n_threads=n_threads).read_chunks(min_chunk_size=chunk_size):
                s = chunk.sequence
                assert s.offsets[0] == 0 and s.offsets[-1] == s.bases.shape[0]
                lens = np.diff(s.offsets)
                assert (lens >= 0).all()
                got += [bytes(s[i]) for i in range(len(s))]
            assert got == want, (chunk_size, n_threads)


def test_gz_shards_of_ranks_with_different_thread_counts_partition_the_reads(tmp_path, monkeypatch):
    """The ranks of a multi-GPU job take every world_size-th block of a .gz.  The block boundaries must not depend
    on a rank's CPU share: with a capacity that grew with n_threads, ranks with 2 and 12 threads cut the stream at
    different places and reads were dropped or counted twice (round-1 advisor finding)."""
    from kmer_mapper_b200.reader import ParallelGzip
    g = synthetic.make_genome(50_000, 31)
    bases, offsets = synthetic.make_reads(g, 24_000, 100, seed=32)
    want = [bytes(bases[offsets[r]:offsets[r + 1]]) for r in range(24_000)]
    path = str(tmp_path / "many_members.fq.gz")
    synthetic.write_fastq(path, bases, offsets, members=60)      # ~90 KB of text per member
    monkeypatch.setattr(ParallelGzip, "BATCH_BYTES", 300_000)    # a few members per block: ~20 blocks
    monkeypatch.setattr(ParallelGzip, "MAX_MEMBER_BYTES", 200_000)
    for threads in ([4, 4], [2, 12], [1, 16, 3]):
        world = len(threads)
        got, n_chunks = [], 0
        per_rank = []
        for rank, nt in enumerate(threads):
            mine = []
            for chunk in open_reads(path, pinned=False, n_threads=nt).read_chunks(min_chunk_size=50_000, rank=rank, world_size=world):
                s = chunk.sequence
                mine += [bytes(s[i]) for i in range(len(s))]
                n_chunks += 1
            per_rank.append(mine)
        assert n_chunks >= 6
        assert sum(len(m) for m in per_rank) == len(want), threads          # nothing dropped, nothing doubled
        assert sorted(x for m in per_rank for x in m) == sorted(want), threads


def test_parallel_gzip_members_fallback_and_hostile_payload(tmp_path):
    import zlib
    from kmer_mapper_b200.reader import ParallelGzip
    rng = np.random.default_rng(3)
    text = bytes(rng.choice(np.frombuffer(b"ACGT\n@+I", np.uint8), size=3_000_000))
    # many members (bgzip-like), cut at arbitrary places
    cuts = [0] + sorted(rng.integers(1, len(text), size=40).tolist()) + [len(text)]
    multi = str(tmp_path / "multi.gz")
    with open(multi, "wb") as f:
        for a, b in zip(cuts[:-1], cuts[1:]):
            f.write(gzip.compress(text[a:b], compresslevel=1))
    for threads in (1, 4):
        for block in (1 << 16, 1 << 22):
            blocks = list(ParallelGzip(multi, threads).blocks(block))
            assert blocks[-1] == b"" and b"".join(blocks) == text
            assert all(len(x) >= block for x in blocks[:-2])
    # one big member: falls back to sequential streaming
    single = str(tmp_path / "single.gz")
    open(single, "wb").write(gzip.compress(text, compresslevel=1))
    assert b"".join(ParallelGzip(single, 4, max_member_bytes=1 << 20).blocks(1 << 18)) == text
    # a big member in the middle of small ones
    mixed = str(tmp_path / "mixed.gz")
    open(mixed, "wb").write(gzip.compress(text[:1000]) + gzip.compress(text[1000:2_000_000]) + gzip.compress(text[2_000_000:]))
    assert b"".join(ParallelGzip(mixed, 3, max_member_bytes=1 << 20).blocks(1 << 18)) == text
    # the member magic inside the compressed stream (stored blocks keep the payload verbatim)
    payload = b"@r\nACGT\n+\n\x1f\x8b\x08\x00II\n" * 2000
    co = zlib.compressobj(0, zlib.DEFLATED, 31)
    hostile = str(tmp_path / "hostile.gz")
    open(hostile, "wb").write(co.compress(payload) + co.flush() + gzip.compress(b"tail"))
    assert b"".join(ParallelGzip(hostile, 4).blocks(1 << 20)) == payload + b"tail"
    empty = str(tmp_path / "empty.gz")
    open(empty, "wb").write(gzip.compress(b""))
    assert b"".join(ParallelGzip(empty, 2).blocks(10)) == b""
    bad = str(tmp_path / "bad.gz")
    open(bad, "wb").write(gzip.compress(text[:5000])[:-20] + b"garbage-not-gzip")
    with pytest.raises(OSError):
        list(ParallelGzip(bad, 2).blocks(1 << 16))


def test_reader_edge_cases(tmp_path):
    p = str(tmp_path / "e.fa")
    open(p, "wb").write(b">r1 desc\r\nACGT\r\nAC\r\n>r2\n\n>r3\nGGGTTT\n>r4\nA")     # CRLF, empty read, no final newline
    for cs in (1, 5, 1000):
        got = []
        for c in open_reads(p, pinned=False).read_chunks(cs):
            got += [bytes(c.sequence[i]) for i in range(len(c.sequence))]
        assert got == [b"ACGTAC", b"", b"GGGTTT", b"A"]
    q = str(tmp_path / "e.fq")
    open(q, "wb").write(b"@a\nACGT\n+\n@III\n@b\nGG\n+\n@I\n@c\nT\n+\nI")            # quality lines starting with '@'
    for cs in (1, 7, 1000):
        got = []
        for c in open_reads(q, pinned=False).read_chunks(cs):
            got += [bytes(c.sequence[i]) for i in range(len(c.sequence))]
        assert got == [b"ACGT", b"GG", b"T"]
    empty = str(tmp_path / "empty.fa")
    open(empty, "wb").close()
    assert list(open_reads(empty, pinned=False).read_chunks(10)) == []
    with pytest.raises(RuntimeError):
        open_reads(str(tmp_path / "reads.bam"))
    bad = str(tmp_path / "bad.fq")
    open(bad, "wb").write(b"ACGT\nACGT\n+\nIIII\n")
    with pytest.raises(ValueError):
        list(open_reads(bad, pinned=False).read_chunks(100))


def test_cli_without_arguments_prints_help_and_exits_1():
    # command_line_interface.py:185-187
    from kmer_mapper_b200 import command_line_interface as cli
    with pytest.raises(SystemExit) as e:
        cli.run_argument_parser([])
    assert e.value.code == 1


def test_cli_defaults_via_parser(monkeypatch):
    from kmer_mapper_b200 import command_line_interface as cli
    seen = {}
    monkeypatch.setattr(cli, "map_bnp", lambda args: seen.setdefault("a", args))
    # set_defaults(func=map_bnp) captured the original function object; intercept at parse time instead
    real_parse = argparse.ArgumentParser.parse_args

    def fake_parse(self, argv=None, namespace=None):
        ns = real_parse(self, argv, namespace)
        ns.func = lambda args: seen.setdefault("a", args)
        return ns

    monkeypatch.setattr(argparse.ArgumentParser, "parse_args", fake_parse)
    cli.run_argument_parser(["map", "-i", "idx.npz", "-f", "reads.fa", "-o", "out"])
    a = seen["a"]
    assert (a.kmer_size, a.n_threads, a.chunk_size, a.max_hits_per_kmer, a.gpu, a.gpu_hash_map_size,
            a.map_reverse_complements, a.debug, a.index_bundle) == (31, 16, 2500000, 1000, False, 0, False, None, None)
    seen.clear()
    cli.run_argument_parser(["map", "-i", "x", "-f", "r.fq.gz", "-o", "o", "-k", "21", "-g", "False", "-r", "0", "-d", "1",
                             "-c", "10000000", "-t", "4", "-I", "5", "-s", "100"])
    a = seen["a"]
    assert a.gpu is True and a.map_reverse_complements is True     # type=bool: "False" and "0" are truthy (cli:175,180)
    assert (a.kmer_size, a.chunk_size, a.n_threads, a.max_hits_per_kmer, a.gpu_hash_map_size, a.debug) == \
        (21, 10000000, 4, 5, 100, "1")


def test_index_argument_errors(tmp_path):
    from kmer_mapper_b200.util import _get_kmer_index_from_args
    with pytest.raises(SystemExit) as e:
        _get_kmer_index_from_args(argparse.Namespace(kmer_index=None, index_bundle=None))
    assert e.value.code == 1                                        # util.py:47-49
    with pytest.raises(NotImplementedError):
        _get_kmer_index_from_args(argparse.Namespace(kmer_index=None, index_bundle="bundle.npz"))
    idx = oracle.index_from_flat_kmers(np.array([1, 2, 3], np.uint64), np.array([1, 2, 3]), 11)
    assert _get_kmer_index_from_args(argparse.Namespace(kmer_index=idx)) is idx   # loaded object accepted (util.py:40-44)


def test_sharding_covers_every_read_once():
    offsets = np.concatenate([[0], np.cumsum(np.random.default_rng(0).integers(0, 200, size=1001))])
    for world in (1, 2, 3, 4, 8):
        seen = []
        for r in range(world):
            lo, hi, b0, b1 = distributed.shard_reads(offsets, r, world)
            assert b0 == offsets[lo] and b1 == offsets[hi]
            seen += list(range(lo, hi))
        assert seen == list(range(1001))
        assert sum(distributed.chunk_belongs_to_rank(i, r, world) for i in range(50) for r in range(world)) == 50


# ---------------------------------------------------------------------------------------------
# packed transport: the host-side 2-bit encoder (kmb_pack_bases, kmb_hostpack.cpp) against the oracle's
# N policy + DNAEncoding restatement (oracle.replace_n_with_a / encode_bases)
# ---------------------------------------------------------------------------------------------
def _pack(bases, flags=0, threads=0):
    import ctypes as C
    from kmer_mapper_b200 import _lib
    n = len(bases)
    cap = (n + 15) // 16 + 4
    words = np.full(cap, 0xDEADBEEF, dtype=np.uint32)
    bad = C.c_int64(-5)
    rc = _lib.lib().kmb_pack_bases(bases.ctypes.data, n, flags, threads, words.ctypes.data, cap, C.byref(bad))
    return rc, words, bad.value


def _oracle_words(bases, n_to_a=True):
    codes = oracle.encode_bases(oracle.replace_n_with_a(bases) if n_to_a else bases).astype(np.uint64)
    cap = (len(bases) + 15) // 16 + 4
    padded = np.zeros(cap * 16, dtype=np.uint64)
    padded[:len(codes)] = codes
    return (padded.reshape(-1, 16) << (2 * np.arange(16, dtype=np.uint64))).sum(axis=1).astype(np.uint32)


@pytest.fixture
def streaming(request):
    from kmer_mapper_b200 import _lib
    _lib.set_option("host_pack_streaming", request.param)
    yield request.param
    _lib.set_option("host_pack_streaming", 1)


@pytest.mark.parametrize("streaming", [0, 1], indirect=True)   # 1: non-temporal stores where the destination is 16-byte aligned
@pytest.mark.parametrize("threads", [1, 0, 3])
def test_host_pack_matches_oracle_encoding(threads, streaming):
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)
    for n in (0, 1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 1000, 4097, 65_536 * 16 + 7, (1 << 22) + 5):
        bases = rng.choice(alphabet, size=n)
        rc, words, bad = _pack(bases, 0, threads)
        assert (rc, bad) == (0, -1)
        assert np.array_equal(words, _oracle_words(bases)), n
        # an unaligned start inside a bigger buffer (chunks begin at arbitrary read offsets)
        if n > 40:
            rc, words, bad = _pack(bases[3:], 0, threads)
            assert rc == 0 and np.array_equal(words, _oracle_words(bases[3:])), n


def test_host_pack_reports_first_invalid_byte_like_the_oracle():
    rng = np.random.default_rng(6)
    for n in (20, 1000, 300_000, (1 << 21) + 11):
        bases = rng.choice(np.frombuffer(b"ACGTacgt", dtype=np.uint8), size=n)
        for junk in (b"n", b"X", b"@", b"\x00", b"\xff", b"N"):
            flags = 2 if junk == b"N" else 0           # KMB_FLAG_NO_N_TO_A: 'N' is then invalid as well
            b = bases.copy()
            pos = np.sort(rng.integers(0, n, size=3))
            b[pos] = junk[0]
            rc, _, bad = _pack(b, flags)
            with pytest.raises(oracle.InvalidBaseError) as e:
                oracle.encode_bases(b if flags else oracle.replace_n_with_a(b))
            assert rc == -2 and bad == e.value.offset == pos[0]
    import ctypes as C
    from kmer_mapper_b200 import _lib
    words = np.zeros(4, np.uint32)
    assert _lib.lib().kmb_pack_bases(bases.ctypes.data, 100, 0, 1, words.ctypes.data, 4, None) == -1   # capacity


def test_host_pack_worker_pool_survives_fork():
    """bench.py and the reference-style CLI fork workers; a child must not wait for the parent's pool threads."""
    import multiprocessing as mp
    bases = np.frombuffer(b"ACGT" * (1 << 20), dtype=np.uint8).copy()
    want = _oracle_words(bases)
    assert np.array_equal(_pack(bases, 0, 4)[1], want)          # the pool now exists in this process
    ctx = mp.get_context("fork")
    q = ctx.Queue()

    def child():
        q.put(bool(np.array_equal(_pack(bases, 0, 4)[1], want)))

    p = ctx.Process(target=child)
    p.start()
    p.join(60)
    assert p.exitcode == 0 and q.get(timeout=5) is True


@pytest.mark.parametrize("fmt", ["fastq", "fasta", "fastq.gz"])
def test_reader_rank_shards_cover_every_read_once_in_order(tmp_path, fmt):
    """Multi-GPU CLI: rank r of world_size parses the records that start in its 1/world_size of a plain file's
    bytes (found by resynchronisation on hostile text), or every world_size-th chunk of a .gz; concatenated in rank
    order (plain) / chunk order (.gz) the shards are the file."""
    rng = np.random.default_rng(11)
    g = synthetic.make_genome(60_000, 8)
    n = 6_000
    bases, offsets = synthetic.make_reads(g, n, 120, seed=12, ragged=True)
    want = [bytes(bases[offsets[r]:offsets[r + 1]]) for r in range(n)]
    path = str(tmp_path / ("shards." + {"fastq": "fq", "fasta": "fa", "fastq.gz": "fq.gz"}[fmt]))
    if fmt == "fastq.gz":
        synthetic.write_fastq(path, bases, offsets, members=7)
    else:
        qual_alphabet = np.frombuffer(b"@+>I#", np.uint8)
        with open(path, "wb") as f:
            for r, seq in enumerate(want):
                if fmt == "fastq":
                    f.write(b"@read%d\n%s\n+\n%s\n" % (r, seq, bytes(rng.choice(qual_alphabet, size=len(seq)))))
                else:
                    w = int(rng.integers(1, 70))
                    f.write(b">read%d\n" % r)
                    for i in range(0, len(seq), w):
                        f.write(seq[i:i + w] + b"\n")
    for world in (1, 2, 3, 7, 64):
        per_rank = []
        for rank in range(world):
            mine = []
            for chunk in open_reads(path, pinned=False).read_chunks(min_chunk_size=20_000, rank=rank, world_size=world):
                s = chunk.sequence
                mine.append([bytes(s[i]) for i in range(len(s))])
            per_rank.append(mine)
        if fmt == "fastq.gz":      # round-robin over chunks: interleave them back
            n_chunks = sum(len(m) for m in per_rank)
            got = []
            for i in range(n_chunks):
                got += per_rank[i % world][i // world]
        else:
            got = [read for mine in per_rank for chunk in mine for read in chunk]
            if world in (2, 3):    # every rank has a real share
                assert all(sum(len(c) for c in mine) > n // (2 * world) for mine in per_rank)
        assert got == want, (fmt, world)


# ---------------------------------------------------------------------------------------------
# the native single-stream gzip decoder (kmb_gzstream_*, csrc/kmb_inflate.cpp) against zlib
# ---------------------------------------------------------------------------------------------
def _gzstream(blob, cap=65536, threads=2):
    import ctypes as C
    from kmer_mapper_b200 import _lib
    lib = _lib.lib()
    head = 1 << 16
    gz = np.frombuffer(blob, dtype=np.uint8) if blob else np.zeros(0, np.uint8)
    h = C.c_void_p()
    assert lib.kmb_gzstream_open(gz.ctypes.data if blob else None, len(blob), threads, C.byref(h)) == 0
    out, tail = [], b""
    try:
        while True:
            buf = np.empty(head + cap, dtype=np.uint8)
            if tail:
                buf[head - len(tail):head] = np.frombuffer(tail, dtype=np.uint8)
            produced, finished = C.c_uint64(), C.c_int()
            rc = lib.kmb_gzstream_read(h, buf.ctypes.data + head, cap, len(tail), C.byref(produced), C.byref(finished))
            if rc != 0:
                raise OSError(lib.kmb_gzstream_error(h).decode())
            piece = buf[head:head + produced.value].tobytes()
            out.append(piece)
            tail = (tail + piece)[-32768:]
            if finished.value:
                return b"".join(out)
    finally:
        lib.kmb_gzstream_close(h)


def test_gzstream_decoder_matches_zlib_on_every_block_type():
    import io
    import zlib
    rng = np.random.default_rng(21)
    g = synthetic.make_genome(30_000, 2)
    bases, offsets = synthetic.make_reads(g, 4_000, 150, seed=3, ragged=True)
    fastq = b"".join(b"@r%d\n%s\n+\n%s\n" % (r, bytes(bases[offsets[r]:offsets[r + 1]]),
                                                bytes(rng.choice(np.frombuffer(b"FFFF:,#", np.uint8), size=offsets[r + 1] - offsets[r])))
                     for r in range(4_000))
    payloads = [b"", b"A", b"hello hello hello world\n" * 3, b"\0" * 300_000, rng.integers(0, 256, 200_000, dtype=np.uint8).tobytes(),
                fastq, (b"ACGT" * 64 + b"N") * 3000, fastq[:50_000] + rng.integers(0, 256, 70_000, dtype=np.uint8).tobytes() + b"x" * 90_000]
    for data in payloads:
        for level in (0, 1, 6, 9):                       # stored, fast and best dynamic blocks
            assert _gzstream(gzip.compress(data, compresslevel=level)) == data
        co = zlib.compressobj(6, zlib.DEFLATED, 31, 9, zlib.Z_FIXED)    # fixed Huffman blocks
        assert _gzstream(co.compress(data) + co.flush(), cap=1 << 20) == data
    # header with a file name, several members (one empty), zero padding after the last
    b = io.BytesIO()
    with gzip.GzipFile(filename="reads.fq", mode="wb", fileobj=b, mtime=0) as f:
        f.write(b"first member\n" * 1000)
    blob = b.getvalue() + gzip.compress(b"second\n" * 5000, 1) + gzip.compress(b"") + gzip.compress(fastq, 9) + b"\0" * 37
    assert _gzstream(blob) == b"first member\n" * 1000 + b"second\n" * 5000 + fastq


def test_gzstream_decoder_rejects_corrupt_and_truncated_streams():
    rng = np.random.default_rng(22)
    data = bytes(rng.choice(np.frombuffer(b"ACGT\n@+F", np.uint8), size=400_000))
    base = gzip.compress(data, 6)
    for t in range(120):
        b2 = bytearray(base)
        if t % 3 == 0:
            b2 = b2[:int(rng.integers(1, len(b2) - 1))]
        elif t % 3 == 1:
            for _ in range(3):
                b2[int(rng.integers(10, len(b2)))] ^= 1 << int(rng.integers(0, 8))
        else:
            p = int(rng.integers(10, len(b2) - 100))
            b2[p:p + 50] = rng.integers(0, 256, 50, dtype=np.uint8).tobytes()
        with pytest.raises(OSError):       # a wrong bit either breaks the code stream or fails the CRC-32 of the trailer
            _gzstream(bytes(b2))
    with pytest.raises(OSError):
        _gzstream(b"this is not gzip at all, not even close....")


def test_encodings_helper_classmethods_match_the_reference(golden_encodings):
    """encodings.py:36-42, 85-93: the table/bit-join helpers, against vectors generated by the reference module."""
    from kmer_mapper_b200.encodings import ACTGTwoBitEncoding, SimpleEncoding
    e = golden_encodings
    masked = e["seq_any"] & 31
    four = ACTGTwoBitEncoding.convert_2bytes_to_4bits(masked.view(np.uint16))
    assert np.array_equal(four, e["helper_2bytes_to_4bits"])
    assert np.array_equal(ACTGTwoBitEncoding.join_4bits_to_byte(four.reshape(-1, 2)), e["helper_join_4bits"])
    assert np.array_equal(e["helper_join_4bits"], e["actg_from_bytes_any"])          # the helpers compose to from_bytes
    two = SimpleEncoding.convert_byte_to_2bits(e["seq_any"])
    assert np.array_equal(two, e["helper_byte_to_2bits"])
    assert np.array_equal(SimpleEncoding.join_2bits_to_byte(two.reshape(-1, 4)), e["helper_join_2bits"])
    assert np.array_equal(e["helper_join_2bits"], e["simple_from_bytes_any"])
    with pytest.raises(AssertionError):
        ACTGTwoBitEncoding.convert_2bytes_to_4bits(masked)                           # dtype check (:37)


def test_chunks_stay_valid_for_a_few_chunks_and_copy_makes_them_independent(tmp_path):
    """A chunk is a view into one of N_CHUNK_BUFFERS rotating buffers (reader.ReadChunk): valid while the next
    N_CHUNK_BUFFERS - 1 chunks are made; chunk.copy() is what a consumer that keeps chunks must take."""
    from kmer_mapper_b200.reader import N_CHUNK_BUFFERS
    g = synthetic.make_genome(20_000, 41)
    bases, offsets = synthetic.make_reads(g, 4_000, 80, seed=42)
    want = [bytes(bases[offsets[r]:offsets[r + 1]]) for r in range(4_000)]
    path = str(tmp_path / "many_chunks.fa")
    synthetic.write_fasta(path, bases, offsets)
    kept = [c.copy() for c in open_reads(path, pinned=False, n_threads=2).read_chunks(min_chunk_size=20_000)]
    assert len(kept) > 2 * N_CHUNK_BUFFERS
    got = [bytes(c.sequence[i]) for c in kept for i in range(len(c))]
    assert got == want
    # without copy(): a window of N_CHUNK_BUFFERS - 1 chunks behind the newest is still intact
    window, got = [], []
    for c in open_reads(path, pinned=False, n_threads=2).read_chunks(min_chunk_size=20_000):
        window.append(c)
        if len(window) == N_CHUNK_BUFFERS:
            old = window.pop(0)
            got += [bytes(old.sequence[i]) for i in range(len(old))]
    for old in window:
        got += [bytes(old.sequence[i]) for i in range(len(old))]
    assert got == want
