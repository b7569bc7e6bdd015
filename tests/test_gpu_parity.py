"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the committed golden fixtures.
Bit-exact: everything here is integer / byte work.  Run on the B200 box: pytest -m gpu."""
import itertools

import numpy as np
import pytest

from oracle import c_oracle, oracle

pytestmark = pytest.mark.gpu

LOOKUP_CASES = ["gpucounter", "test_mapping", "long_buckets", "long_buckets_cut5", "sparse_k31",
                "full64", "empty_query", "malformed_directory"]


@pytest.fixture(scope="module")
def kmb():
    from kmer_mapper_b200 import _lib
    _lib.require_device()  # fail loudly: these tests are meaningless without the CUDA library + a GPU
    # the count arrays of these tests are small enough for the direct-reduction route (no hit log); the hit log and its
    # apply pass are what the full-size runs use, so that is what this module exercises unless a variant says otherwise
    _lib.set_option("direct_counts_max_nodes", 0)
    yield _lib
    for name, v in (("probe_variant", 1), ("use_filter", -1), ("gathers_in_flight", 4), ("filter_l2_budget_bytes", 60 << 20),
                    ("apply_window_log2", 0), ("direct_counts_max_nodes", 4 << 20),
                    ("log_max_entries", 2048 << 20), ("chunk_bytes", 64 << 20), ("host_pack", -1), ("host_threads", 0), ("read_table", -1)):
        _lib.set_option(name, v)


# apply_window_log2: a tiny apply window forces several windows per node range (the large-count-array path) on small inputs
VARIANTS = [dict(probe_variant=1, use_filter=1, gathers_in_flight=4),
            dict(probe_variant=1, use_filter=0, gathers_in_flight=4),
            dict(probe_variant=1, use_filter=1, gathers_in_flight=2),
            dict(probe_variant=1, use_filter=0, gathers_in_flight=2),
            dict(probe_variant=1, use_filter=1, gathers_in_flight=4, filter_l2_budget_bytes=512),
            dict(probe_variant=1, use_filter=1, gathers_in_flight=4, log_max_entries=4096),
            dict(probe_variant=1, use_filter=1, gathers_in_flight=4, apply_window_log2=10),
            dict(probe_variant=1, use_filter=1, gathers_in_flight=4, direct_counts_max_nodes=4 << 20),   # no hit log: direct reductions
            dict(probe_variant=0, use_filter=1),
            dict(probe_variant=0, use_filter=0)]
# the fused reads kernel over the minimizer-bucketed read-path table (k = 31 only; opt-in)
READ_TABLE_VARIANTS = [dict(read_table=1, use_filter=1), dict(read_table=1, use_filter=0),
                       dict(read_table=1, use_filter=1, filter_l2_budget_bytes=512),
                       dict(read_table=1, use_filter=1, log_max_entries=4096),
                       dict(read_table=1, use_filter=1, direct_counts_max_nodes=4 << 20)]


def _fresh(index):
    """Drop the cached device copy so that index-creation options (use_filter) take effect."""
    if hasattr(index, "_kmb_device_index"):
        del index._kmb_device_index
    return index


def _set(kmb, variant):
    kmb.set_option("filter_l2_budget_bytes", 60 << 20)
    kmb.set_option("log_max_entries", 2048 << 20)
    kmb.set_option("read_table", 0)
    kmb.set_option("gathers_in_flight", 4)
    kmb.set_option("apply_window_log2", 23)
    kmb.set_option("direct_counts_max_nodes", 0)
    for k, v in variant.items():
        kmb.set_option(k, v)


# ---------------------------------------------------------------------------------------------
# L1 / L2: map_kmers_to_graph_index, in_graph_index  (mapper.pyx:19-72, :81-190)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", VARIANTS, ids=lambda v: "-".join("%s%d" % (k[0], x) for k, x in v.items()))
@pytest.mark.parametrize("name", LOOKUP_CASES)
def test_lookup_golden_fixtures(kmb, golden_lookup, name, variant):
    from kmer_mapper_b200.mapper import in_graph_index, in_graph_index_no_memory_maps, map_kmers_to_graph_index
    g = golden_lookup[name]
    _set(kmb, variant)
    idx = _fresh(g["index"])
    got = map_kmers_to_graph_index(idx, g["max_node_id"], g["queries"], g["cutoff"])
    assert got.dtype == np.uint32 and got.shape == (g["max_node_id"] + 1,)
    assert np.array_equal(got, g["ref_counts"])
    assert np.array_equal(in_graph_index(idx, g["queries"]), g["ref_member"])
    assert np.array_equal(in_graph_index_no_memory_maps(idx, g["queries"], 5), g["ref_member"])


def test_reference_known_answer_through_gpu_counter(kmb):
    # the reference's own vector, tests/test_gpucounter.py:41-48
    from kmer_mapper_b200.gpu_counter import GpuCounter
    kmers = np.array([1, 2, 3], dtype=np.uint64)
    nodes = np.array([10, 11, 12])
    counter = GpuCounter.from_kmers_and_nodes(kmers, nodes, 31)
    counter.initialize_cuda(2003)
    counter.count(np.array([1, 1, 1, 2, 3, 1, 3], dtype=np.uint64))
    node_counts = counter.get_node_counts(15)
    assert np.all(node_counts[[10, 11, 12]] == [4, 1, 2])
    assert node_counts.shape == (15,)
    # cumulative across calls (gpu_counter.py:23-24), auto capacity (cli:178 default 0)
    counter2 = GpuCounter.from_kmers_and_nodes(kmers, nodes, 31)
    counter2.initialize_cuda(0)
    for _ in range(3):
        counter2.count(np.array([1, 1, 1, 2, 3, 1, 3], dtype=np.uint64))
    assert np.all(counter2.get_node_counts(15)[[10, 11, 12]] == [12, 3, 6])


def test_gpu_counter_matches_restated_semantics(kmb):
    from kmer_mapper_b200.gpu_counter import GpuCounter
    rng = np.random.default_rng(11)
    k = 9
    keys = rng.integers(0, 4 ** k, size=3000, dtype=np.uint64)      # duplicates: a k-mer on several nodes
    nodes = rng.integers(0, 700, size=3000)
    queries = np.concatenate([rng.choice(keys, 5000), rng.integers(0, 4 ** k, size=5000, dtype=np.uint64)])
    for rc in (False, True):
        c = GpuCounter.from_kmers_and_nodes(keys, nodes, k)
        c.initialize_cuda(0)
        c.count(queries[:4000], count_revcomps=rc)
        c.count(queries[4000:], count_revcomps=rc)
        want = oracle.gpu_counter_node_counts(keys, nodes, queries, min_nodes=800, count_revcomps=rc, k=k)
        got = c.get_node_counts(800)
        assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize("variant", VARIANTS[:8], ids=lambda v: "-".join("%s%d" % (k[0], x) for k, x in v.items()))
def test_lookup_random_indexes_vs_oracle(kmb, variant):
    from kmer_mapper_b200.mapper import in_graph_index, map_kmers_to_graph_index
    _set(kmb, variant)
    rng = np.random.default_rng(7)
    for trial in range(12):
        n = int(rng.integers(1, 30000))
        modulo = int(rng.choice([1, 2, 21, 97, 1009, 65537, 1000003, 4000037]))
        bits = int(rng.choice([6, 20, 62, 64]))
        hi = 2 ** bits
        keys = rng.integers(0, hi, size=n, dtype=np.uint64) if bits < 64 else rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
        nodes = rng.integers(0, 1 + int(rng.integers(1, 5000)), size=n)
        idx = oracle.index_from_flat_kmers(keys, nodes, modulo)
        q = np.concatenate([rng.choice(keys, 20000),
                            rng.integers(0, hi, size=20000, dtype=np.uint64) if bits < 64 else
                            rng.integers(0, 2 ** 64, size=20000, dtype=np.uint64)])
        mx = idx.max_node_id() + int(rng.integers(0, 3))
        cutoff = int(rng.choice([0, 1, 2, 1000, 70000, -1]))
        want = c_oracle.map_kmers_to_graph_index(idx, mx, q, cutoff)
        got = map_kmers_to_graph_index(idx, mx, q, cutoff)
        assert np.array_equal(got, want), (trial, n, modulo, bits, cutoff)
        assert np.array_equal(in_graph_index(idx, q), c_oracle.in_graph_index(idx, q))


def test_lookup_argument_errors_match_reference_behaviour(kmb, golden_lookup):
    from kmer_mapper_b200._lib import KmbError
    from kmer_mapper_b200.mapper import map_kmers_to_graph_index
    g = golden_lookup["gpucounter"]
    with pytest.raises(ValueError):     # uint64_t[::1] memoryview cast, mapper.pyx:19
        map_kmers_to_graph_index(g["index"], 14, g["queries"].astype(np.int64))
    with pytest.raises(ValueError):
        map_kmers_to_graph_index(g["index"], 14, np.zeros(10, np.uint64)[::2])
    with pytest.raises(KmbError):       # the reference would write out of bounds (boundscheck off)
        map_kmers_to_graph_index(g["index"], 5, g["queries"])
    bad = oracle.OracleIndex(np.array([0, 5], np.int32), np.array([1, 3], np.int32), np.array([0, 1], np.int32),
                             np.array([1, 2], np.uint64), np.array([1, 1], np.uint16), 2)
    with pytest.raises(KmbError):       # bucket outside the entry arrays
        map_kmers_to_graph_index(bad, 5, g["queries"])
    assert map_kmers_to_graph_index(g["index"], 14, np.zeros(0, np.uint64)).sum() == 0


def test_map_kmers_device_buffers_and_accumulation(kmb, golden_lookup):
    import torch
    from kmer_mapper_b200.device import DeviceIndex, Mapper
    g = golden_lookup["sparse_k31"]
    di = DeviceIndex.from_index(_fresh(g["index"]))
    counts = torch.zeros(g["max_node_id"] + 1, dtype=torch.int32, device="cuda")
    m = Mapper(di, g["max_node_id"] + 1, g["cutoff"], counts_tensor=counts)
    q = torch.from_numpy(g["queries"].view(np.int64)).cuda()
    m.set_stream(torch.cuda.current_stream())
    m.map_kmers(q)
    m.map_kmers(g["queries"])           # host buffer on top: counts accumulate
    m.sync()
    torch.cuda.synchronize()
    got = counts.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, g["ref_counts"] * 2)
    assert m.stats() == (2 * g["queries"].shape[0], 2 * int(g["ref_counts"].sum()))
    m.reset()
    assert m.counts().sum() == 0 and m.stats() == (0, 0)
    m.close()


def test_two_mappers_on_one_index_and_overflow_chains(kmb):
    """The slot counters live inside the index lines: a second concurrent mapper must get its own copy,
    and a bucket with far more than ten entries must be walked through its whole chain."""
    from kmer_mapper_b200.device import DeviceIndex, Mapper
    rng = np.random.default_rng(21)
    keys = rng.integers(0, 4 ** 12, size=4000, dtype=np.uint64)
    keys = np.concatenate([keys, np.full(777, keys[3], np.uint64), np.full(35, keys[9], np.uint64)])
    nodes = rng.integers(0, 3000, size=keys.shape[0])
    idx = oracle.index_from_flat_kmers(keys, nodes, 211)          # ~23 entries per bucket: chains everywhere
    di = DeviceIndex.from_index(_fresh(idx))
    assert di.n_overflow_lines > 0 and di.n_live_entries == keys.shape[0]
    q1 = np.concatenate([rng.choice(keys, 30000), rng.integers(0, 4 ** 12, size=30000, dtype=np.uint64)])
    q2 = rng.choice(keys, 10000)
    a, b = Mapper(di, 3000, 1000), Mapper(di, 3000, 5)
    a.map_kmers(q1)
    b.map_kmers(q2)
    a.map_kmers(q2)
    assert np.array_equal(b.counts(), c_oracle.map_kmers_to_graph_index(idx, 2999, q2, 5))
    assert np.array_equal(a.counts(), c_oracle.map_kmers_to_graph_index(idx, 2999, np.concatenate([q1, q2]), 1000))
    a.close()
    b.close()
    c = Mapper(di, 3000, 1000)                                     # the master copy is clean again
    c.map_kmers(q2)
    assert np.array_equal(c.counts(), c_oracle.map_kmers_to_graph_index(idx, 2999, q2, 1000))
    c.close()


# ---------------------------------------------------------------------------------------------
# H1: hashing (util.py:71-75)
# ---------------------------------------------------------------------------------------------
def test_hash_formula_golden(kmb, golden_hashing):
    from kmer_mapper_b200.util import get_kmer_hashes_from_chunk_sequence
    g = golden_hashing
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[g["numeric"]]
    h = get_kmer_hashes_from_chunk_sequence((bases, np.array([0, bases.shape[0]], np.int64)), int(g["k"]))
    assert h.dtype == np.uint64 and np.array_equal(h, g["hashes"])


def test_hashing_ragged_case_n_policy_and_errors(kmb):
    from kmer_mapper_b200._lib import InvalidBaseError
    from kmer_mapper_b200.sequences import RaggedSequence
    from kmer_mapper_b200.util import get_kmer_hashes_from_chunk_sequence as gk
    reads = [b"ACGTNACGTTTGACCA", b"ac", b"", b"gattacaNNgattaca", b"TTT", b"A" * 70 + b"C" * 70, b"", b"g"]
    seq = RaggedSequence.from_strings(reads)
    for k in (1, 2, 3, 5, 16, 17, 31):
        want = oracle.kmer_hashes_loops(seq.bases, seq.offsets, k, n_to_a=True)
        got = gk(seq, k, n_to_a=True)
        assert got.dtype == np.uint64 and np.array_equal(got, want), k
    # mixed case hashes like upper case (tests/test_mapping.py:33,40)
    assert np.array_equal(gk([b"cCG"], 3), gk([b"CCG"], 3))
    assert gk("ACG", 3)[0] == 0 + 1 * 4 + 2 * 16
    # the bare util function has no N policy: N raises, like every other byte
    with pytest.raises(InvalidBaseError) as e:
        gk(seq, 3)
    assert e.value.offset == 4
    with pytest.raises(InvalidBaseError) as e:
        gk([b"ACGT", b"ACGnT"], 3, n_to_a=True)   # lower-case n is not covered by the N policy (cli:41)
    assert e.value.offset == 7
    with pytest.raises(InvalidBaseError) as e:
        gk([b"AC", b"R"], 5, n_to_a=True)         # invalid even where no window starts
    assert e.value.offset == 2
    assert gk([b"AC", b""], 3).shape == (0,)


@pytest.mark.parametrize("k", [1, 15, 21, 31])
def test_hashing_random_reads_vs_oracle_host_and_device(kmb, k):
    import torch
    from kmer_mapper_b200 import synthetic as S
    from kmer_mapper_b200.util import get_kmer_hashes_from_chunk_sequence as gk
    g = S.make_genome(200_000, 5)
    for ragged, n_reads, L in ((False, 3000, 150), (True, 4000, 90), (False, 7, 10_000)):
        bases, offsets = S.make_reads(g, n_reads, L, seed=k + L, n_rate=0.02, lower_rate=0.5, ragged=ragged)
        want = c_oracle.kmer_hashes(bases, offsets, k, n_to_a=True)
        assert np.array_equal(gk((bases, offsets), k, n_to_a=True), want)
        tb, to = torch.from_numpy(bases).cuda(), torch.from_numpy(offsets).cuda()
        got = gk((tb, to), k, n_to_a=True)
        assert got.is_cuda and np.array_equal(got.cpu().numpy().view(np.uint64), want)


# ---------------------------------------------------------------------------------------------
# M1: fused reads -> counts (command_line_interface.py:32-56)
# ---------------------------------------------------------------------------------------------
def _small_world(k, seed, n_entries=60_000, modulo=262_147, zipf=False, hot=1200):
    from kmer_mapper_b200 import synthetic as S
    g = S.make_genome(400_000, seed)
    idx = S.make_index(g, n_entries, k, 50_000, modulo, seed + 1, n_hot_nodes=hot, zipf_nodes=zipf)
    return g, idx


@pytest.mark.parametrize("k,variant", [(k, v) for k in (31, 21, 15, 5) for v in VARIANTS[:8]] +
                         [(31, v) for v in READ_TABLE_VARIANTS],
                         ids=lambda x: str(x) if isinstance(x, int) else "-".join("%s%d" % (k[0], v) for k, v in x.items()))
def test_map_reads_vs_oracle(kmb, k, variant):
    import torch
    from kmer_mapper_b200 import synthetic as S
    from kmer_mapper_b200.device import DeviceIndex, Mapper
    _set(kmb, variant)
    g, idx = _small_world(k, 100 + k, zipf=(k == 15))
    di = DeviceIndex.from_index(_fresh(idx))
    assert (di.filter_bytes > 0) == bool(variant["use_filter"])
    mx = idx.max_node_id()
    for ragged, n_reads, L, n_rate in ((False, 20_000, 150, 0.0), (True, 30_000, 120, 0.03), (False, 40, 10_000, 0.05),
                                       (True, 3_000, 40, 0.0)):
        bases, offsets = S.make_reads(g, n_reads, L, seed=7 * k + L, n_rate=n_rate, lower_rate=0.5 if n_rate else 0.0,
                                      ragged=ragged)
        want, n_want = c_oracle.map_reads(idx, mx, bases, offsets, k, n_threads=4)
        m = Mapper(di, mx + 1)
        for host_pack in (1, 0):                            # host buffers: 2-bit packed transport, then plain ASCII
            kmb.set_option("host_pack", host_pack)
            m.map_reads(bases, offsets, k)
            got = m.counts()
            assert np.array_equal(got, want)
            assert m.stats() == (n_want, int(want.astype(np.uint64).sum()))
            # the variant really ran the kernel it is named after
            assert kmb.get_option("last_reads_kernel") == (1 if variant.get("read_table") == 1 and k == 31 else 0)
            m.reset()
        kmb.set_option("host_pack", -1)
        tb, to = torch.from_numpy(bases).cuda(), torch.from_numpy(offsets).cuda()
        m.map_reads(tb, to, k)                              # device buffers (in place)
        assert np.array_equal(m.counts(), want)
        m.close()


def test_read_table_is_chosen_by_itself_for_indexes_that_are_not_small(kmb):
    """read_table = -1: the minimizer-bucketed table serves k = 31 reads without reverse complements when the index has
    at least read_table_min_entries live entries (threshold lowered for the test), whatever its key filter looks like."""
    from kmer_mapper_b200 import synthetic as S
    from kmer_mapper_b200.device import DeviceIndex, Mapper
    k = 31
    g, idx = _small_world(k, 700)
    mx = idx.max_node_id()
    bases, offsets = S.make_reads(g, 4_000, 150, seed=5, ragged=True)
    want, _ = c_oracle.map_reads(idx, mx, bases, offsets, k, n_threads=4)
    try:
        for budget, min_entries, rc, expect in ((512, 1, False, 1), (60 << 20, 1, False, 1), (512, 1 << 40, False, 0),
                                                (60 << 20, 1 << 40, False, 0), (512, 1, True, 0)):
            _set(kmb, dict(filter_l2_budget_bytes=budget, read_table=-1))
            kmb.set_option("read_table_min_entries", min_entries)
            m = Mapper(DeviceIndex.from_index(_fresh(idx)), mx + 1)
            m.map_reads(bases, offsets, k, revcomp=rc)
            got = m.counts()
            assert kmb.get_option("last_reads_kernel") == expect, (budget, min_entries, rc)
            if not rc:
                assert np.array_equal(got, want)
            m.close()
    finally:
        kmb.set_option("read_table_min_entries", 8 << 20)
        kmb.set_option("filter_l2_budget_bytes", 60 << 20)
        kmb.set_option("read_table", -1)


@pytest.mark.parametrize("host_pack,host_threads,read_table", [(1, 0, 0), (1, 3, 0), (0, 0, 0), (1, 0, 1), (0, 0, 1), (2, 0, 0), (-1, 4, 0)])
def test_map_reads_chunked_host_path_and_linearity(kmb, host_pack, host_threads, read_table):
    """host_pack 2 / -1 with a PINNED source: the hybrid transport -- some chunks packed by the cores, the others as
    ASCII straight from the caller's buffer, whichever pipe is free -- must give the same counts as either alone."""
    from kmer_mapper_b200 import synthetic as S
    from kmer_mapper_b200.device import DeviceIndex, Mapper
    kmb.set_option("host_pack", host_pack)
    kmb.set_option("host_threads", host_threads)
    kmb.set_option("read_table", read_table)
    k = 31
    g, idx = _small_world(k, 300)
    di = DeviceIndex.from_index(_fresh(idx))
    mx = idx.max_node_id()
    bases, offsets = S.make_reads(g, 50_000, 150, seed=9, ragged=True, n_rate=0.01)
    want, _ = c_oracle.map_reads(idx, mx, bases, offsets, k, n_threads=4)
    hybrid = host_pack in (2, -1)
    if hybrid:
        import torch
        pb, po = torch.from_numpy(bases).pin_memory(), torch.from_numpy(offsets).pin_memory()
        bases, offsets = pb.numpy(), po.numpy()
    kmb.set_option("chunk_bytes", 1 << 16)                  # ~120 chunks: exercises the slot ring
    m = Mapper(di, mx + 1)
    before = kmb.get_option("host_chunks_packed"), kmb.get_option("host_chunks_ascii")
    m.map_reads(bases, offsets, k)
    assert np.array_equal(m.counts(), want)
    if hybrid:     # both pipes were used
        assert kmb.get_option("host_chunks_packed") > before[0] and kmb.get_option("host_chunks_ascii") > before[1]
    # linearity: mapping two halves into the same mapper == mapping the whole (additive map-reduce, cli:124-130)
    m.reset()
    half = 25_000
    m.map_reads(bases[:offsets[half]], offsets[:half + 1], k)
    m.map_reads(bases[offsets[half]:], offsets[half:] - offsets[half], k)
    assert np.array_equal(m.counts(), want)
    m.close()
    kmb.set_option("chunk_bytes", 64 << 20)
    kmb.set_option("host_pack", -1)
    kmb.set_option("host_threads", 0)
    kmb.set_option("read_table", -1)


def test_map_reads_reverse_complement_flag(kmb):
    from kmer_mapper_b200 import synthetic as S
    from kmer_mapper_b200.device import DeviceIndex, Mapper
    k = 21
    g, idx = _small_world(k, 400)
    mx = idx.max_node_id()
    bases, offsets = S.make_reads(g, 5_000, 150, seed=3)
    # reverse-complement the reads on the host: comp = 3 - code, order reversed
    comp = np.zeros(256, np.uint8)
    comp[list(b"ACGTacgt")] = list(b"TGCAtgca")
    rc_bases = np.concatenate([comp[bases[offsets[r]:offsets[r + 1]]][::-1] for r in range(len(offsets) - 1)])
    fwd, _ = c_oracle.map_reads(idx, mx, bases, offsets, k)
    rev, _ = c_oracle.map_reads(idx, mx, rc_bases, offsets, k)
    assert rev.sum() < fwd.sum()        # the index holds forward k-mers only
    m = Mapper(DeviceIndex.from_index(_fresh(idx)), mx + 1)
    m.map_reads(rc_bases, offsets, k, revcomp=True)         # -r: both strands of every window
    assert np.array_equal(m.counts(), fwd + rev)
    # same through ready-made hashes
    m.reset()
    m.map_kmers(c_oracle.kmer_hashes(rc_bases, offsets, k), revcomp=True, k=k)
    assert np.array_equal(m.counts(), fwd + rev)
    m.close()


@pytest.mark.parametrize("host_pack", [1, 0])
def test_map_reads_invalid_base_reports_offset(kmb, host_pack):
    from kmer_mapper_b200 import synthetic as S
    from kmer_mapper_b200._lib import InvalidBaseError
    from kmer_mapper_b200.device import DeviceIndex, Mapper
    kmb.set_option("host_pack", host_pack)
    g, idx = _small_world(31, 500, n_entries=2000, hot=0)
    bases, offsets = S.make_reads(g, 3000, 100, seed=1)
    bases[123_457] = ord("n")
    bases[200_001] = ord("R")
    kmb.set_option("chunk_bytes", 1 << 16)
    m = Mapper(DeviceIndex.from_index(_fresh(idx)), idx.max_node_id() + 1)
    m.map_reads(bases, offsets, 31)
    with pytest.raises(InvalidBaseError) as e:
        m.counts()
    assert e.value.offset == 123_457
    m.reset()
    m.map_reads(bases, offsets, 31, n_to_a=False)           # without the N policy 'N' is invalid too
    bases2 = bases.copy()
    bases2[[123_457, 200_001]] = ord("A")
    bases2[77] = ord("N")
    m.reset()
    m.map_reads(bases2, offsets, 31, n_to_a=False)
    with pytest.raises(InvalidBaseError) as e:
        m.sync()
    assert e.value.offset == 77
    m.close()
    kmb.set_option("chunk_bytes", 64 << 20)
    kmb.set_option("host_pack", -1)


def test_map_cpu_chunk_worker_shape(kmb):
    from kmer_mapper_b200 import synthetic as S
    from kmer_mapper_b200.command_line_interface import map_cpu
    from kmer_mapper_b200.sequences import RaggedSequence
    g, idx = _small_world(31, 600)
    bases, offsets = S.make_reads(g, 2000, 150, seed=2, n_rate=0.01)
    got = map_cpu({"kmer_size": 31}, idx, RaggedSequence(bases, offsets))
    want = oracle.map_reads(idx, idx.max_node_id(), bases, offsets, 31)
    assert got.dtype == np.uint32 and np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------
# E1: legacy codec (encodings.py)
# ---------------------------------------------------------------------------------------------
def test_legacy_codec_golden(kmb, golden_encodings):
    from kmer_mapper_b200.encodings import ACTGTwoBitEncoding, BaseEncoding, SimpleEncoding, twobit_swap
    e = golden_encodings
    for nm in ("any", "acgt"):
        assert np.array_equal(ACTGTwoBitEncoding.from_bytes(e["seq_" + nm]), e["actg_from_bytes_" + nm])
        assert np.array_equal(SimpleEncoding.from_bytes(e["seq_" + nm]), e["simple_from_bytes_" + nm])
    assert np.array_equal(ACTGTwoBitEncoding.to_bytes(e["actg_from_bytes_acgt"]), e["actg_to_bytes"])
    assert np.array_equal(SimpleEncoding.to_bytes(e["simple_from_bytes_acgt"]), e["simple_to_bytes"])
    c64 = ACTGTwoBitEncoding.complement(e["words64"])
    assert c64.dtype == e["words64"].dtype and np.array_equal(c64, e["complement64"])
    assert np.array_equal(ACTGTwoBitEncoding.complement(e["actg_from_bytes_acgt"]), e["complement8"])
    assert np.array_equal(twobit_swap(e["words64"]), e["twobit_swap64"])
    assert np.array_equal(twobit_swap(e["words32"]), e["twobit_swap32"])
    assert np.array_equal(ACTGTwoBitEncoding.from_string("ACTGACTG"), e["from_string_ACTGACTG"])
    assert ACTGTwoBitEncoding.to_string(ACTGTwoBitEncoding.from_string("ACTGACTG")) == "actgactg"   # lower case (:75)
    assert BaseEncoding.to_string(BaseEncoding.from_string("ACGT")) == "ACGT"
    with pytest.raises(AssertionError):
        ACTGTwoBitEncoding.from_bytes(np.zeros(5, np.uint8))                                          # size % 4 (:53)
    # reverse-complement hash identity of this codec: swap(complement(h)) >> 2(32-k)
    rng = np.random.default_rng(1)
    w = rng.integers(0, 2 ** 62, size=1000, dtype=np.uint64)
    assert np.array_equal(twobit_swap(ACTGTwoBitEncoding.complement(w)), oracle.twobit_swap(oracle.actg_complement(w)))
    big = rng.integers(0, 256, size=4 * 100_003, dtype=np.uint8)
    assert np.array_equal(ACTGTwoBitEncoding.from_bytes(big), oracle.actg_from_bytes(big))
    assert np.array_equal(SimpleEncoding.from_bytes(big), oracle.simple_from_bytes(big))
