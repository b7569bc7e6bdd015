"""Seeded synthetic genomes, indexes and reads of the benchmark shapes.

Two back ends with the same structure: numpy on the host (tests, CLI fixtures, the CPU baseline's
bounded sample) and torch on the GPU (full-size bench inputs: a 100 M-entry index takes seconds on
the device and minutes on the host).  This is input generation -- plumbing around the measured
path, never inside a timed region.
"""
from __future__ import annotations

import gzip

import numpy as np

from .kmer_index import KmerIndex

ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)


# ---------------------------------------------------------------------------------------------
# numpy (host)
# ---------------------------------------------------------------------------------------------
def make_genome(length: int, seed: int) -> np.ndarray:
    """uint8 codes 0..3 (A,C,G,T), uniform."""
    return np.random.default_rng(seed).integers(0, 4, size=int(length), dtype=np.uint8)


def kmers_at(genome: np.ndarray, positions: np.ndarray, k: int) -> np.ndarray:
    """hash = sum_j code[p+j] * 4**j (util.py:71-75 / tests/test_hashing.py:13-26 convention)."""
    positions = np.asarray(positions, dtype=np.int64)
    h = np.zeros(positions.shape[0], dtype=np.uint64)
    for j in range(k):
        h |= genome[positions + j].astype(np.uint64) << np.uint64(2 * j)
    return h


def make_index(genome, n_entries, k, n_nodes, modulo, seed, n_hot_nodes=0, zipf_nodes=False) -> KmerIndex:
    """k-mers at n_entries distinct random genome positions, nodes ~ U[0, n_nodes) (or Zipf s=1 for the
    atomic-contention configs), frequencies = multiplicity of the key among the entries; optionally
    one extra k-mer placed on ``n_hot_nodes`` nodes to exercise the frequency > 1000 cut-off."""
    rng = np.random.default_rng(seed)
    n_pos = genome.shape[0] - k + 1
    positions = rng.choice(n_pos, size=int(n_entries), replace=False) if n_entries <= n_pos else \
        rng.integers(0, n_pos, size=int(n_entries))
    keys = kmers_at(genome, positions, k)
    if zipf_nodes:
        u = rng.random(keys.shape[0])
        nodes = np.minimum((np.exp(u * np.log(n_nodes)) - 1.0).astype(np.int64), n_nodes - 1)
    else:
        nodes = rng.integers(0, n_nodes, size=keys.shape[0])
    if n_hot_nodes:
        hot_key = kmers_at(genome, np.array([int(rng.integers(0, n_pos))]), k)
        keys = np.concatenate([keys, np.repeat(hot_key, n_hot_nodes)])
        nodes = np.concatenate([nodes, rng.integers(0, n_nodes, size=n_hot_nodes)])
    idx = KmerIndex.from_flat_kmers(hashes=keys, nodes=nodes, modulo=modulo)
    idx.convert_to_int32()
    return idx


def make_reads(genome, n_reads, read_len, seed, error_rate=0.005, n_rate=0.0, lower_rate=0.0,
               ragged=False):
    """(bases uint8[B] ASCII, offsets int64[R+1]): forward-strand substrings at uniform positions with
    substitution errors; optional upper-case N and lower-casing (config 5); ``ragged`` draws read lengths
    from U[0, read_len] instead of a constant."""
    rng = np.random.default_rng(seed)
    n_reads = int(n_reads)
    lens = rng.integers(0, read_len + 1, size=n_reads) if ragged else np.full(n_reads, read_len, dtype=np.int64)
    starts = rng.integers(0, genome.shape[0] - read_len, size=n_reads)
    offsets = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    read_of = np.repeat(np.arange(n_reads), lens)
    pos_in = np.arange(total) - offsets[read_of]
    codes = genome[starts[read_of] + pos_in]
    err = rng.random(total) < error_rate
    codes = np.where(err, (codes + rng.integers(1, 4, size=total, dtype=np.uint8)) & 3, codes).astype(np.uint8)
    bases = ASCII[codes]
    if lower_rate:
        low = rng.random(total) < lower_rate
        bases = np.where(low, bases | 0x20, bases).astype(np.uint8)
    if n_rate:
        bases = np.where(rng.random(total) < n_rate, np.uint8(ord("N")), bases).astype(np.uint8)
    return np.ascontiguousarray(bases), offsets


def write_fasta(path, bases, offsets, line_width=0):
    """2-line FASTA (or wrapped at ``line_width``); ``.gz`` suffix -> gzip."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        for r in range(len(offsets) - 1):
            seq = bytes(bases[offsets[r]:offsets[r + 1]])
            f.write(b">read%d\n" % r)
            if line_width:
                for i in range(0, max(len(seq), 1), line_width):
                    f.write(seq[i:i + line_width] + b"\n")
            else:
                f.write(seq + b"\n")


def write_fastq(path, bases, offsets, members=1):
    """4-line FASTQ, constant quality; ``.gz`` -> gzip with ``members`` concatenated members (still a
    valid .gz for any reader, and inflatable in parallel)."""
    n = len(offsets) - 1
    gz = str(path).endswith(".gz")
    bounds = np.linspace(0, n, (members if gz else 1) + 1).astype(int)
    with open(path, "wb") as raw:
        for a, b in zip(bounds[:-1], bounds[1:]):
            parts = []
            for r in range(a, b):
                seq = bytes(bases[offsets[r]:offsets[r + 1]])
                parts.append(b"@read%d\n%s\n+\n%s\n" % (r, seq, b"I" * len(seq)))
            blob = b"".join(parts)
            raw.write(gzip.compress(blob, compresslevel=1) if gz else blob)


# ---------------------------------------------------------------------------------------------
# torch (device) -- full-size bench inputs
# ---------------------------------------------------------------------------------------------
def t_make_genome(length, seed, device="cuda"):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.randint(0, 4, (int(length),), dtype=torch.uint8, device=device, generator=g)


def t_kmers_at(genome, positions, k):
    import torch
    h = torch.zeros(positions.shape[0], dtype=torch.int64, device=genome.device)
    for j in range(k):
        h |= genome[positions + j].to(torch.int64) << (2 * j)
    return h  # int64 bit pattern == uint64 (k <= 31 -> < 2**62)


def t_make_index(genome, n_entries, k, n_nodes, modulo, seed, zipf_nodes=False):
    """Device-side construction of the six index arrays (same rule as KmerIndex.from_flat_kmers).
    Returns a dict of torch tensors: hashes_to_index/n_kmers/nodes int32, kmers int64 (uint64 bits),
    frequencies uint16, plus modulo."""
    import torch
    dev = genome.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n_pos = genome.shape[0] - k + 1
    # ~n_entries distinct positions: Bernoulli thinning is enough for a synthetic index
    positions = torch.randint(0, n_pos, (int(n_entries * 1.06),), device=dev, generator=g)
    positions = torch.unique(positions)
    perm = torch.randperm(positions.shape[0], device=dev, generator=g)[:int(n_entries)]
    positions = positions[perm]
    del perm
    keys = t_kmers_at(genome, positions, k)
    del positions
    if zipf_nodes:
        u = torch.rand(keys.shape[0], device=dev, generator=g, dtype=torch.float64)
        nodes = torch.clamp((torch.exp(u * float(np.log(n_nodes))) - 1.0).to(torch.int64), max=n_nodes - 1)
    else:
        nodes = torch.randint(0, int(n_nodes), (keys.shape[0],), device=dev, generator=g)
    h = keys % int(modulo)
    h, order = torch.sort(h, stable=True)
    keys = keys[order]
    nodes = nodes[order].to(torch.int32)
    del order
    n_kmers = torch.bincount(h, minlength=int(modulo))
    hashes_to_index = (torch.cumsum(n_kmers, 0) - n_kmers).to(torch.int32)
    n_kmers = n_kmers.to(torch.int32)
    del h
    _, inv, cnt = torch.unique(keys, return_inverse=True, return_counts=True)
    freq = torch.clamp(cnt[inv], max=65535).to(torch.int32).to(torch.uint16)
    del inv, cnt
    return dict(hashes_to_index=hashes_to_index, n_kmers=n_kmers, nodes=nodes, kmers=keys, frequencies=freq,
                modulo=int(modulo))


def t_make_reads(genome, n_reads, read_len, seed, error_rate=0.005, n_rate=0.0, lower_rate=0.0, slice_reads=1 << 20):
    """(bases uint8[B] on the device, offsets int64[R+1] on the device), constant read length."""
    import torch
    dev = genome.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n_reads = int(n_reads)
    bases = torch.empty(n_reads * read_len, dtype=torch.uint8, device=dev)
    ascii_lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    ar = torch.arange(read_len, device=dev)
    for s in range(0, n_reads, slice_reads):
        m = min(slice_reads, n_reads - s)
        starts = torch.randint(0, genome.shape[0] - read_len, (m,), device=dev, generator=g)
        codes = genome[(starts[:, None] + ar[None, :]).reshape(-1)]
        r = torch.rand(codes.shape[0], device=dev, generator=g)
        sub = torch.randint(1, 4, (codes.shape[0],), dtype=torch.uint8, device=dev, generator=g)
        codes = torch.where(r < error_rate, (codes + sub) & 3, codes)
        b = ascii_lut[codes.to(torch.int64)]
        if lower_rate:
            low = torch.rand(b.shape[0], device=dev, generator=g) < lower_rate
            b = torch.where(low, b | 0x20, b)
        if n_rate:
            isn = torch.rand(b.shape[0], device=dev, generator=g) < n_rate
            b = torch.where(isn, torch.full_like(b, ord("N")), b)
        bases[s * read_len:(s + m) * read_len] = b
    offsets = torch.arange(n_reads + 1, device=dev, dtype=torch.int64) * read_len
    return bases, offsets


class TensorIndex:
    """Duck-typed index (mapper.pyx:22-29 attribute names) over device tensors, for DeviceIndex.from_index."""

    def __init__(self, d):
        self._hashes_to_index = d["hashes_to_index"]
        self._n_kmers = d["n_kmers"]
        self._nodes = d["nodes"]
        self._kmers = d["kmers"]
        self._frequencies = d["frequencies"]
        self._modulo = d["modulo"]

    def max_node_id(self):
        return int(self._nodes.max().item())

    def to_host(self) -> KmerIndex:
        idx = KmerIndex(self._hashes_to_index.cpu().numpy(), self._n_kmers.cpu().numpy(), self._nodes.cpu().numpy(),
                        None, self._kmers.cpu().numpy().view(np.uint64), self._modulo,
                        self._frequencies.cpu().numpy().view(np.uint16))
        return idx
