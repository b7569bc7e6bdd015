"""Pin the CPU oracle (oracle/) before anything is compared against it (CPU-only tests).

Sources of truth, strongest first: the compiled unmodified reference mapper.pyx (oracle/_ref, present
when oracle/build_ref.py ran in the build container), the committed golden fixtures generated from
it (tests/golden), the reference's own known-answer vector tests/test_gpucounter.py:41-48.
"""
import numpy as np
import pytest

from oracle import c_oracle, oracle, ref_loader

LOOKUP_CASES = ["gpucounter", "test_mapping", "long_buckets", "long_buckets_cut5", "sparse_k31",
                "full64", "empty_query", "malformed_directory"]


def test_reference_golden_vector_gpucounter():
    # tests/test_gpucounter.py:41-48: kmers [1,2,3] on nodes [10,11,12], query [1,1,1,2,3,1,3] -> [4,1,2]
    idx = oracle.index_from_flat_kmers(np.array([1, 2, 3], np.uint64), np.array([10, 11, 12]), 2003)
    q = np.array([1, 1, 1, 2, 3, 1, 3], np.uint64)
    for fn in (oracle.map_kmers_to_graph_index, oracle.map_kmers_to_graph_index_loops,
               c_oracle.map_kmers_to_graph_index, oracle.count_kmers_bruteforce):
        counts = fn(idx, 14, q)
        assert counts.dtype == np.uint32 and counts.shape == (15,)
        assert list(counts[[10, 11, 12]]) == [4, 1, 2]
        assert counts.sum() == 7
    nc = oracle.gpu_counter_node_counts(np.array([1, 2, 3]), np.array([10, 11, 12]), q, min_nodes=15)
    assert list(nc[[10, 11, 12]]) == [4, 1, 2] and nc.shape == (15,)


@pytest.mark.parametrize("name", LOOKUP_CASES)
def test_oracle_matches_golden_fixture(golden_lookup, name):
    g = golden_lookup[name]
    for fn in (oracle.map_kmers_to_graph_index, c_oracle.map_kmers_to_graph_index):
        got = fn(g["index"], g["max_node_id"], g["queries"], g["cutoff"])
        assert got.dtype == np.uint32
        assert np.array_equal(got, g["ref_counts"]), name
    assert np.array_equal(oracle.in_graph_index(g["index"], g["queries"]), g["ref_member"])
    assert np.array_equal(c_oracle.in_graph_index(g["index"], g["queries"]), g["ref_member"])
    if g["queries"].shape[0] <= 8000:
        got = oracle.map_kmers_to_graph_index_loops(g["index"], g["max_node_id"], g["queries"], g["cutoff"])
        assert np.array_equal(got, g["ref_counts"])


@pytest.mark.parametrize("name", ["gpucounter", "test_mapping", "long_buckets", "sparse_k31", "full64"])
def test_bruteforce_counter_agrees_on_wellformed_indexes(golden_lookup, name):
    g = golden_lookup[name]
    got = oracle.count_kmers_bruteforce(g["index"], g["max_node_id"], g["queries"], g["cutoff"])
    assert np.array_equal(got, g["ref_counts"])


def test_oracle_matches_compiled_reference_on_random_indexes():
    ref = ref_loader.load_reference_mapper()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference in the build container)")
    rng = np.random.default_rng(7)
    for trial in range(20):
        n = int(rng.integers(1, 3000))
        modulo = int(rng.choice([1, 2, 21, 97, 1009, 65537, 1000003]))
        bits = int(rng.choice([6, 20, 62, 64]))
        keys = rng.integers(0, 2 ** bits, size=n, dtype=np.uint64, endpoint=False) if bits < 64 else \
            rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
        nodes = rng.integers(0, 1 + int(rng.integers(1, 5000)), size=n)
        idx = oracle.index_from_flat_kmers(keys, nodes, modulo)
        q = np.concatenate([rng.choice(keys, 2000), rng.integers(0, 2 ** bits if bits < 64 else 2 ** 64,
                                                                 size=2000, dtype=np.uint64)])
        mx = idx.max_node_id() + int(rng.integers(0, 3))
        cutoff = int(rng.choice([0, 1, 2, 1000, 70000]))
        want = ref.map_kmers_to_graph_index(idx, mx, q, cutoff)
        assert np.array_equal(oracle.map_kmers_to_graph_index(idx, mx, q, cutoff), want)
        assert np.array_equal(c_oracle.map_kmers_to_graph_index(idx, mx, q, cutoff), want)
        assert np.array_equal(oracle.in_graph_index(idx, q), ref.in_graph_index(idx, q))
        assert np.array_equal(c_oracle.in_graph_index(idx, q), ref.in_graph_index_no_memory_maps(idx, q))


def test_hash_formula_golden(golden_hashing):
    # tests/test_hashing.py:13-26: get_kmers(arange(35) % 4, 31), LSB-first
    g = golden_hashing
    k = int(g["k"])
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[g["numeric"]]
    offsets = np.array([0, bases.shape[0]])
    for fn in (oracle.kmer_hashes, oracle.kmer_hashes_loops, c_oracle.kmer_hashes):
        h = fn(bases, offsets, k)
        assert h.dtype == np.uint64
        assert np.array_equal(h, g["hashes"])
    # complement-and-mask identity (:18) and the convolve-based reverse complement (:22-26); in the
    # A,C,G,T=0..3 code the true complement is 3-c (bitwise NOT), the reference's (c+2)%4 line is the
    # A,C,T,G-era formula and is pinned as written
    comp = (~g["hashes"]) & np.uint64(4 ** k - 1)
    assert np.array_equal(comp, g["complement_masked"])
    rc = oracle.reverse_complement_hashes(g["hashes"], k)
    comp_bases = np.frombuffer(b"TGCA", dtype=np.uint8)[g["numeric"]][::-1]
    assert np.array_equal(oracle.kmer_hashes(comp_bases, offsets, k)[::-1], rc)


def test_hashing_ragged_case_and_n_policy():
    reads = [b"ACGTNACGTTTGACCA", b"ac", b"", b"gattacaNNgattaca", b"TTT"]
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8)
    offsets = np.cumsum([0] + [len(r) for r in reads])
    for k in (1, 3, 5, 16, 17, 31):
        want = oracle.kmer_hashes_loops(bases, offsets, k)
        assert want.shape[0] == sum(max(0, len(r) - k + 1) for r in reads)
        assert np.array_equal(oracle.kmer_hashes(bases, offsets, k), want)
        assert np.array_equal(c_oracle.kmer_hashes(bases, offsets, k), want)
    # case-insensitive (tests/test_mapping.py:33,40: "cCG" hashed like "ccg")
    up = np.frombuffer(b"".join(reads).upper(), dtype=np.uint8)
    assert np.array_equal(oracle.kmer_hashes(up, offsets, 3), oracle.kmer_hashes(bases, offsets, 3))
    # N -> A only for upper-case N, and only when the policy is on (cli:40-41)
    with pytest.raises(ValueError):
        oracle.kmer_hashes(bases, offsets, 3, n_to_a=False)
    with pytest.raises(ValueError):
        c_oracle.kmer_hashes(np.frombuffer(b"ACGnT", np.uint8), np.array([0, 5]), 3)
    with pytest.raises(oracle.InvalidBaseError) as e:
        oracle.kmer_hashes(np.frombuffer(b"ACGTRACGT", np.uint8), np.array([0, 9]), 3)
    assert e.value.offset == 4


def test_map_reads_port_equals_hash_then_lookup():
    rng = np.random.default_rng(3)
    genome = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=20000)
    k = 15
    off = np.array([0, genome.shape[0]])
    gk = oracle.kmer_hashes(genome, off, k)
    pick = rng.choice(gk.shape[0], 3000, replace=False)
    idx = oracle.index_from_flat_kmers(gk[pick], rng.integers(0, 900, size=3000), 4099)
    starts = rng.integers(0, genome.shape[0] - 100, size=200)
    lens = rng.integers(0, 100, size=200)
    bases = np.concatenate([genome[s:s + l] for s, l in zip(starts, lens)]).copy()
    bases[rng.random(bases.shape[0]) < 0.03] = ord("N")
    lower = (rng.random(bases.shape[0]) < 0.5) & (bases != ord("N"))   # lower-case n is an error
    bases[lower] = np.frombuffer(bytes(bases[lower]).lower(), np.uint8)
    offsets = np.concatenate([[0], np.cumsum(lens)])
    want = oracle.map_reads(idx, 899, bases, offsets, k)
    for t in (1, 4):
        got, n = c_oracle.map_reads(idx, 899, bases, offsets, k, n_threads=t)
        assert np.array_equal(got, want)
        assert n == sum(max(0, l - k + 1) for l in lens)
    assert want.sum() > 0


def test_legacy_codec_restatement_matches_reference_vectors(golden_encodings):
    e = golden_encodings
    for nm in ("any", "acgt"):
        assert np.array_equal(oracle.actg_from_bytes(e["seq_" + nm]), e["actg_from_bytes_" + nm])
        assert np.array_equal(oracle.simple_from_bytes(e["seq_" + nm]), e["simple_from_bytes_" + nm])
    assert np.array_equal(oracle.actg_to_bytes(e["actg_from_bytes_acgt"]), e["actg_to_bytes"])
    assert np.array_equal(oracle.actg_to_bytes(e["simple_from_bytes_acgt"]), e["simple_to_bytes"])
    assert np.array_equal(oracle.actg_complement(e["words64"]), e["complement64"])
    assert np.array_equal(oracle.actg_complement(e["actg_from_bytes_acgt"]), e["complement8"])
    assert np.array_equal(oracle.twobit_swap(e["words64"]), e["twobit_swap64"])
    assert np.array_equal(oracle.twobit_swap(e["words32"]), e["twobit_swap32"])
    assert np.array_equal(oracle.actg_from_bytes(np.frombuffer(b"ACTGACTG", np.uint8)), e["from_string_ACTGACTG"])


def test_counter_index_route_restatement_equals_bruteforce():
    """CounterKmerIndex route (command_line_interface.py:46-49,118-119,133-138): values per unique key summed over
    chunks, then scattered onto nodes -- against a dictionary count."""
    rng = np.random.default_rng(5)
    kmers = rng.integers(0, 50, size=40).astype(np.uint64)        # duplicates: several entries share a key
    nodes = rng.integers(0, 12, size=40)
    chunks = [rng.integers(0, 60, size=n).astype(np.uint64) for n in (0, 17, 300)]
    values, node_counts = oracle.counter_index_route(kmers, nodes, chunks, min_nodes=15)
    allq = np.concatenate(chunks)
    unique = np.unique(kmers)
    assert np.array_equal(values, [int((allq == u).sum()) for u in unique])
    want = np.zeros(15)
    for km, nd in zip(kmers, nodes):
        want[nd] += int((allq == km).sum())
    assert np.array_equal(node_counts, want) and node_counts.dtype == np.float64
