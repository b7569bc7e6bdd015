"""Throughput of the host-side 2-bit encoder (kmb_pack_bases) per thread count, and of the packed host path of
map_reads per chunk size.  Usage: python tools/host_pack_throughput.py [--gpu]   (JSON lines on stdout)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmer_mapper_b200 import _lib  # noqa: E402


def pack_rate(n, threads, reps=5):
    L = _lib.lib()
    rng = np.random.default_rng(1)
    bases = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    cap = (n + 15) // 16 + 4
    words = np.zeros(cap, dtype=np.uint32)
    bad = C.c_int64()
    L.kmb_pack_bases(bases.ctypes.data, n, 0, threads, words.ctypes.data, cap, C.byref(bad))
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        L.kmb_pack_bases(bases.ctypes.data, n, 0, threads, words.ctypes.data, cap, C.byref(bad))
        best = min(best, time.perf_counter() - t)
    return n / best / 1e9


def main():
    cpus = len(os.sched_getaffinity(0))
    for n in (1 << 26, 1 << 30):
        for threads in sorted({1, 2, 4, 8, cpus // 2, cpus}):
            if threads < 1:
                continue
            print(json.dumps({"op": "kmb_pack_bases", "bases": n, "threads": threads,
                              "GBps": round(pack_rate(n, threads), 2)}), flush=True)
    if "--gpu" not in sys.argv:
        return
    import torch
    from kmer_mapper_b200 import synthetic as S
    from kmer_mapper_b200.device import DeviceIndex, Mapper
    k, L, n_reads = 31, 150, 20_000_000
    dev = torch.device("cuda:0")
    genome = S.t_make_genome(1 << 28, seed=3, device=dev)
    index = S.TensorIndex(S.t_make_index(genome, 20_000_000, k, 20_000_000, modulo=90_000_001, seed=4))
    di = DeviceIndex.from_index(index)
    bases_d, offsets_d = S.t_make_reads(genome, n_reads, L, seed=5)
    bases = torch.empty(bases_d.shape, dtype=torch.uint8, pin_memory=True).copy_(bases_d)
    offsets = torch.empty(offsets_d.shape, dtype=torch.int64, pin_memory=True).copy_(offsets_d)
    pageable = bases.numpy().copy()
    torch.cuda.synchronize()
    n_kmers = n_reads * (L - k + 1)
    m = Mapper(di, index.max_node_id() + 1)
    for host_pack, chunk_mb, threads, src in ((0, 64, 0, "pinned"), (1, 16, 0, "pinned"), (1, 64, 0, "pinned"),
                                              (1, 256, 0, "pinned"), (1, 64, 8, "pinned"), (1, 64, 0, "pageable"),
                                              (0, 64, 0, "pageable")):
        _lib.set_option("host_pack", host_pack)
        _lib.set_option("chunk_bytes", chunk_mb << 20)
        _lib.set_option("host_threads", threads)
        b = bases.numpy() if src == "pinned" else pageable
        best = 1e9
        for _ in range(3):
            m.reset()
            torch.cuda.synchronize()
            t = time.perf_counter()
            m.map_reads(b, offsets.numpy(), k)
            m.sync()
            best = min(best, time.perf_counter() - t)
        print(json.dumps({"op": "map_reads(host buffers)", "host_pack": host_pack, "chunk_MB": chunk_mb,
                          "host_threads": threads or cpus, "source": src, "kmers": n_kmers, "ms": round(best * 1e3, 2),
                          "GKps": round(n_kmers / best / 1e9, 2)}), flush=True)
    m.close()


if __name__ == "__main__":
    main()
