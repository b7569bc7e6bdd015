"""End to end through the reference-shaped CLI on the GPU: files in, <out>.npy out, compared with the oracle
run on an independently parsed copy of the same files."""
import argparse
import os

import numpy as np
import pytest

from oracle import c_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world(tmp_path_factory):
    from kmer_mapper_b200 import _lib, synthetic
    _lib.require_device()
    d = tmp_path_factory.mktemp("cli")
    k = 31
    g = synthetic.make_genome(300_000, 11)
    idx = synthetic.make_index(g, 40_000, k, 30_000, 200_003, 12, n_hot_nodes=1100)
    idx.to_file(str(d / "index.npz"))
    bases, offsets = synthetic.make_reads(g, 6_000, 150, seed=13, n_rate=0.01, lower_rate=0.3, ragged=True)
    synthetic.write_fasta(str(d / "reads.fa"), bases, offsets, line_width=70)
    synthetic.write_fastq(str(d / "reads.fq.gz"), bases, offsets, members=4)
    want, n_kmers = c_oracle.map_reads(idx, idx.max_node_id(), bases, offsets, k, n_threads=4)
    assert want.sum() > 1000
    return dict(dir=d, idx=idx, want=want, k=k, bases=bases, offsets=offsets)


@pytest.mark.parametrize("parse", ["device", "host"])
@pytest.mark.parametrize("reads,chunk", [("reads.fa", 2_500_000), ("reads.fa", 20_000), ("reads.fq.gz", 50_000)])
def test_cli_map_writes_reference_shaped_output(world, reads, chunk, parse, monkeypatch):
    """Both routes of the CLI: records parsed by GPU kernels from the raw text (default), or by the native host parser."""
    from kmer_mapper_b200.command_line_interface import PARSE_ENV, run_argument_parser
    monkeypatch.setenv(PARSE_ENV, parse)
    d = world["dir"]
    out = str(d / ("out_%s_%d_%s" % (reads.replace(".", "_"), chunk, parse)))
    run_argument_parser(["map", "-i", str(d / "index.npz"), "-f", str(d / reads), "-o", out, "-k", str(world["k"]),
                         "-c", str(chunk), "-t", "3"])
    got = np.load(out + ".npy")                    # np.save appends .npy (command_line_interface.py:149)
    assert got.dtype == np.uint32 and got.shape == (world["idx"].max_node_id() + 1,)
    assert np.array_equal(got, world["want"])


def test_map_bnp_programmatic_call_returns_counts(world):
    # KAGE-style call: a Namespace carrying a loaded index object and output_file=None (cli:146-147, util.py:40-44)
    from kmer_mapper_b200.command_line_interface import map_bnp
    from kmer_mapper_b200.kmer_index import KmerIndex
    d = world["dir"]
    idx = KmerIndex.from_file(str(d / "index.npz"))
    args = argparse.Namespace(kmer_index=idx, index_bundle=None, reads=str(d / "reads.fq.gz"), kmer_size=world["k"],
                              n_threads=1, chunk_size=1_000_000, output_file=None, debug=None, max_hits_per_kmer=1000,
                              gpu=True, gpu_hash_map_size=0, map_reverse_complements=False, func=map_bnp)
    got = map_bnp(args)
    assert np.array_equal(got, world["want"])
    # the CPU route refuses reverse complements (cli:107); the GPU route accepts the flag
    args = argparse.Namespace(**{**vars(args), "gpu": False, "map_reverse_complements": True, "func": map_bnp})
    with pytest.raises(AssertionError):
        map_bnp(args)


def test_max_hits_flag_is_ignored_like_the_reference_unless_opted_in(world, monkeypatch):
    """-I is parsed and unused in the reference (cli:51,173): the cut-off stays 1000.  The opt-in environment
    variable makes it real; the fixture index holds one k-mer on 1100 nodes, so a cut-off of 2000 counts more."""
    from kmer_mapper_b200.command_line_interface import HONOUR_MAX_HITS_ENV, run_argument_parser
    d = world["dir"]
    base = ["map", "-i", str(d / "index.npz"), "-f", str(d / "reads.fa"), "-k", str(world["k"]), "-I", "2000"]
    monkeypatch.delenv(HONOUR_MAX_HITS_ENV, raising=False)
    run_argument_parser(base + ["-o", str(d / "maxhits_default")])
    assert np.array_equal(np.load(str(d / "maxhits_default.npy")), world["want"])
    monkeypatch.setenv(HONOUR_MAX_HITS_ENV, "1")
    run_argument_parser(base + ["-o", str(d / "maxhits_opt_in")])
    got = np.load(str(d / "maxhits_opt_in.npy"))
    want, _ = c_oracle.map_reads(world["idx"], world["idx"].max_node_id(), world["bases"], world["offsets"], world["k"],
                                 max_index_lookup_frequency=2000, n_threads=4)
    assert np.array_equal(got, want)


def test_map_gpu_signature_and_invalid_reads(world, tmp_path):
    from kmer_mapper_b200._lib import InvalidBaseError
    from kmer_mapper_b200.command_line_interface import map_gpu
    from kmer_mapper_b200.reader import open_reads
    d = world["dir"]
    chunks = open_reads(str(d / "reads.fa")).read_chunks(min_chunk_size=100_000)
    got = map_gpu(world["idx"], chunks, world["k"], 0, False)
    assert np.array_equal(got, world["want"])
    bad = tmp_path / "bad.fa"
    bad.write_bytes(b">r1\nACGTACGTACGTACGTACGTACGTACGTACGTACGTRACGT\n")
    with pytest.raises(InvalidBaseError):
        map_gpu(world["idx"], open_reads(str(bad)).read_chunks(1000), world["k"], 0, False)


def test_counter_kmer_index_route_vs_oracle(world, tmp_path):
    """The CounterKmerIndex route (command_line_interface.py:46-49, 118-119, 133-138): per-chunk value arrays from
    map_cpu, their sum, and the final get_node_counts -- through map_cpu directly, through map_bnp with a loaded
    object, and through the CLI with a counter-index file (util.py:63-66 fallback)."""
    from oracle import oracle
    from kmer_mapper_b200.command_line_interface import map_bnp, map_cpu, run_argument_parser
    from kmer_mapper_b200.counter_index import CounterKmerIndex
    idx, k = world["idx"], world["k"]
    bases, offsets = world["bases"], world["offsets"]
    n = len(offsets) - 1
    cuts = [0, n // 3, n // 3, n]                      # three chunks, one of them empty
    chunks = [(np.ascontiguousarray(bases[offsets[a]:offsets[b]]), np.ascontiguousarray(offsets[a:b + 1] - offsets[a]))
              for a, b in zip(cuts[:-1], cuts[1:])]
    hashes = [c_oracle.kmer_hashes(b, o, k) for b, o in chunks]
    want_values, want_nodes = oracle.counter_index_route(idx._kmers, idx._nodes, hashes)
    cki = CounterKmerIndex.from_kmer_index(idx)
    total = np.zeros(cki.counter.n_keys, dtype=np.uint32)
    for chunk in chunks:
        got = map_cpu({"kmer_size": k}, cki, chunk)
        assert got.dtype == np.uint32 and got.shape == total.shape
        total += got
    assert np.array_equal(total, want_values)
    cki.counter._values = total                        # command_line_interface.py:136
    assert np.array_equal(cki.get_node_counts(), want_nodes)
    # programmatic run with the object, and the CLI with a file that is not a KmerIndex archive
    d = world["dir"]
    args = argparse.Namespace(kmer_index=CounterKmerIndex.from_kmer_index(idx), index_bundle=None, reads=str(d / "reads.fq.gz"),
                              kmer_size=k, n_threads=1, chunk_size=200_000, output_file=None, debug=None,
                              max_hits_per_kmer=1000, gpu=False, gpu_hash_map_size=0, map_reverse_complements=False, func=map_bnp)
    assert np.array_equal(map_bnp(args), want_nodes)
    cki.to_file(str(tmp_path / "counter_index.npz"))
    out = str(tmp_path / "counter_out")
    run_argument_parser(["map", "-i", str(tmp_path / "counter_index.npz"), "-f", str(d / "reads.fa"), "-o", out, "-k", str(k),
                         "-c", "100000"])
    assert np.array_equal(np.load(out + ".npy"), want_nodes)
