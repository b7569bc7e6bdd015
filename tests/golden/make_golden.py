"""Generate the committed golden fixtures by RUNNING THE REFERENCE ITSELF in the build container.

  * lookup/count/membership vectors: outputs of the reference's own Cython kernel
    (/root/reference/kmer_mapper/mapper.pyx compiled unmodified by oracle/build_ref.py).
  * legacy codec vectors: outputs of /root/reference/kmer_mapper/encodings.py, exec'd from where it
    lies with the one-token numpy-2 upcast SURVEY.md appendix A describes (``256*reverse[...]``
    overflows uint8 under numpy >= 2; semantics unchanged).
  * hashing: the formula case of the reference's tests/test_hashing.py:13-26
    (numeric sequence arange(35) % 4, k = 31), evaluated with Python integers -- bionumpy itself is
    not installable here, so these vectors pin the formula, not bionumpy's output.

/root/reference does not exist on the GPU box; only the .npz files this script writes travel.
Run from the repo root:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import build_ref, ref_loader  # noqa: E402
from oracle.oracle import OracleIndex, index_from_flat_kmers  # noqa: E402


def load_reference_encodings():
    src = open("/root/reference/kmer_mapper/encodings.py").read()
    needle = "256*reverse[np.arange(4)[:, None]]"
    assert needle in src
    src = src.replace(needle, "256*reverse.astype(np.int64)[np.arange(4)[:, None]]")
    mod = types.ModuleType("reference_encodings")
    exec(compile(src, "/root/reference/kmer_mapper/encodings.py", "exec"), mod.__dict__)
    return mod


def lookup_cases(rng):
    cases = {}

    # 0: the reference's own golden vector, tests/test_gpucounter.py:41-48, table size 2003
    idx = index_from_flat_kmers(np.array([1, 2, 3], np.uint64), np.array([10, 11, 12]), 2003)
    cases["gpucounter"] = (idx, 14, np.array([1, 1, 1, 2, 3, 1, 3], np.uint64), 1000)

    # 1: the shape of tests/test_mapping.py:31-44: 4 k-mers ACT,CTT,cCG,ATT -> nodes 0..3, modulo 21,
    #    max_node_id 100, queries = the node k-mers themselves (hash = sum code*4^j, ACGT=0..3)
    def h(s):
        return sum("ACGT".index(c) << (2 * j) for j, c in enumerate(s.upper()))
    km = np.array([h(s) for s in ("ACT", "CTT", "cCG", "ATT")], np.uint64)
    idx = index_from_flat_kmers(km, np.arange(4), 21)
    cases["test_mapping"] = (idx, 100, km, 1000)

    # 2: tiny modulo => long buckets with many different keys, multi-node k-mers, hot frequency
    keys = rng.integers(0, 4 ** 15, size=300, dtype=np.uint64)
    keys = np.concatenate([keys, np.repeat(keys[:20], 3), np.full(1200, keys[5], np.uint64)])
    nodes = rng.integers(0, 500, size=keys.shape[0])
    idx = index_from_flat_kmers(keys, nodes, 97)
    q = np.concatenate([rng.choice(keys, 4000), rng.integers(0, 4 ** 15, size=4000, dtype=np.uint64)])
    cases["long_buckets"] = (idx, 499, q, 1000)
    cases["long_buckets_cut5"] = (idx, 499, q, 5)

    # 3: k=31-sized keys, sparse directory (load 0.2), 10 % hits, frequency field set by hand
    keys = rng.integers(0, 4 ** 31, size=20000, dtype=np.uint64)
    nodes = rng.integers(0, 16000, size=keys.shape[0])
    freq = rng.choice(np.array([1, 2, 999, 1000, 1001, 65535], np.uint16), size=keys.shape[0])
    idx = index_from_flat_kmers(keys, nodes, 100003, frequencies=freq)
    q = np.concatenate([rng.choice(keys, 5000), rng.integers(0, 4 ** 31, size=45000, dtype=np.uint64)])
    rng.shuffle(q)
    cases["sparse_k31"] = (idx, int(nodes.max()), q, 1000)

    # 4: keys using all 64 bits, max_node_id larger than any node, empty query
    keys = rng.integers(0, 2 ** 64, size=5000, dtype=np.uint64)
    nodes = rng.integers(0, 70000, size=keys.shape[0])
    idx = index_from_flat_kmers(keys, nodes, 65537)
    q = np.concatenate([rng.choice(keys, 3000), rng.integers(0, 2 ** 64, size=3000, dtype=np.uint64)])
    cases["full64"] = (idx, 80000, q, 1000)
    cases["empty_query"] = (idx, 80000, np.zeros(0, np.uint64), 1000)

    # 5: malformed-but-in-bounds directory: overlapping buckets and entries filed under the wrong bucket
    keys = rng.integers(0, 1000, size=64, dtype=np.uint64)
    nodes = rng.integers(0, 50, size=64)
    modulo = 13
    h2i = rng.integers(0, 40, size=modulo).astype(np.int32)
    nk = rng.integers(0, 20, size=modulo).astype(np.int32)
    idx = OracleIndex(h2i, nk, nodes, keys, rng.integers(1, 3, size=64), modulo)
    q = rng.integers(0, 1000, size=20000, dtype=np.uint64)
    cases["malformed_directory"] = (idx, 49, q, 1000)
    return cases


def main():
    so = build_ref.build_reference_mapper()
    assert so, "needs /root/reference"
    ref = ref_loader.load_reference_mapper()
    rng = np.random.default_rng(20261018)

    out = {}
    for name, (idx, max_node, q, cutoff) in lookup_cases(rng).items():
        counts = ref.map_kmers_to_graph_index(idx, max_node, q, cutoff)
        member = ref.in_graph_index(idx, q)
        member2 = ref.in_graph_index_no_memory_maps(idx, q)
        assert counts.dtype == np.uint32 and counts.shape[0] == max_node + 1
        assert np.array_equal(member, member2)
        for key, val in (("hashes_to_index", idx._hashes_to_index), ("n_kmers", idx._n_kmers),
                         ("nodes", idx._nodes), ("kmers", idx._kmers), ("frequencies", idx._frequencies),
                         ("modulo", np.array(idx._modulo, np.uint64)), ("max_node_id", np.array(max_node)),
                         ("cutoff", np.array(cutoff)), ("queries", q),
                         ("ref_counts", counts), ("ref_member", member)):
            out[name + "/" + key] = val
    g = out["gpucounter/ref_counts"]
    assert list(g[[10, 11, 12]]) == [4, 1, 2] and g.shape[0] == 15, g   # tests/test_gpucounter.py:47
    np.savez_compressed(os.path.join(HERE, "golden_lookup.npz"), **out)
    print("golden_lookup.npz:", len(out), "arrays")

    enc = load_reference_encodings()
    e = {}
    letters = np.frombuffer(b"ACGTacgtNn!#4'XRYKM-*\n\r \x00\xff", dtype=np.uint8)
    seq = rng.choice(letters, size=4096).astype(np.uint8)
    acgt = rng.choice(np.frombuffer(b"ACGTacgt", dtype=np.uint8), size=4096).astype(np.uint8)
    words = rng.integers(0, 2 ** 64, size=512, dtype=np.uint64)
    words32 = rng.integers(0, 2 ** 32, size=512, dtype=np.uint32)
    e["seq_any"] = seq
    e["seq_acgt"] = acgt
    e["words64"] = words
    e["words32"] = words32
    for nm, s in (("any", seq), ("acgt", acgt)):
        e["actg_from_bytes_" + nm] = enc.ACTGTwoBitEncoding.from_bytes(s)
        e["simple_from_bytes_" + nm] = enc.SimpleEncoding.from_bytes(s)
    # the helper classmethods (encodings.py:36-42, 85-93)
    masked = (seq & 31)
    e["helper_2bytes_to_4bits"] = enc.ACTGTwoBitEncoding.convert_2bytes_to_4bits(masked.view(np.uint16))
    e["helper_join_4bits"] = enc.ACTGTwoBitEncoding.join_4bits_to_byte(e["helper_2bytes_to_4bits"].reshape(-1, 2))
    e["helper_byte_to_2bits"] = enc.SimpleEncoding.convert_byte_to_2bits(seq)
    e["helper_join_2bits"] = enc.SimpleEncoding.join_2bits_to_byte(e["helper_byte_to_2bits"].reshape(-1, 4))
    e["actg_to_bytes"] = enc.ACTGTwoBitEncoding.to_bytes(e["actg_from_bytes_acgt"])
    e["simple_to_bytes"] = enc.SimpleEncoding.to_bytes(e["simple_from_bytes_acgt"])
    e["complement64"] = enc.ACTGTwoBitEncoding.complement(words)
    e["complement8"] = enc.ACTGTwoBitEncoding.complement(e["actg_from_bytes_acgt"])
    e["twobit_swap64"] = enc.twobit_swap(words)
    e["twobit_swap32"] = enc.twobit_swap(words32)
    e["from_string_ACTGACTG"] = enc.ACTGTwoBitEncoding.from_string("ACTGACTG")
    e["base_from_string"] = enc.BaseEncoding.from_string("ACGTN")
    assert enc.ACTGTwoBitEncoding.to_string(e["from_string_ACTGACTG"]) == "actgactg"
    assert int(enc.twobit_swap(np.array([0x0123456789ABCDEF], np.uint64))[0]) == 0xFB73EA62D951C840
    np.savez_compressed(os.path.join(HERE, "golden_encodings.npz"), **e)
    print("golden_encodings.npz:", len(e), "arrays")

    # tests/test_hashing.py:13-26, Python-integer evaluation of the formula
    k = 31
    numeric = [i % 4 for i in range(35)]
    hashes = [sum(numeric[p + j] << (2 * j) for j in range(k)) for p in range(35 - k + 1)]
    comp = [(~x) & (4 ** k - 1) for x in hashes]                       # :18 complement & mask (ACGT: 3-c)
    rc_codes = [(c + 2) % 4 for c in numeric][::-1]                    # :22 as written in the reference
    conv = [sum(rc_codes[p + j] << (2 * (k - 1 - j)) for j in range(k)) for p in range(35 - k + 1)]
    # np.convolve(a, w, 'valid')[n] = sum_m a[n+m] w[k-1-m]; the reference prints it reversed (:26)
    np.savez_compressed(os.path.join(HERE, "golden_hashing.npz"),
                        numeric=np.array(numeric, np.uint8), k=np.array(k),
                        hashes=np.array(hashes, np.uint64), complement_masked=np.array(comp, np.uint64),
                        convolve_rev=np.array(conv[::-1], np.uint64))
    print("golden_hashing.npz: ok")


if __name__ == "__main__":
    main()
