#!/usr/bin/env python
"""Device-resident throughput of the other drop-in operators on the config-2 index: the hashing API
(util.py:71-75), map_kmers_to_graph_index on ready-made hashes (mapper.pyx:19) and in_graph_index (:81)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from kmer_mapper_b200 import _lib  # noqa: E402
from kmer_mapper_b200.device import DeviceIndex, Mapper  # noqa: E402
from kmer_mapper_b200.util import get_kmer_hashes_from_chunk_sequence  # noqa: E402

w = bench.workload("config2", 1.0)
w["reads"] = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
tindex, bases, offsets = bench.generate(w, 0, torch.device("cuda", 0))
di = DeviceIndex.from_index(tindex, device=0)
n_counts = tindex.max_node_id() + 1
k = w["k"]


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


dt, hashes = timed(lambda: get_kmer_hashes_from_chunk_sequence((bases, offsets), k, n_to_a=True))
n = hashes.shape[0]
print(json.dumps(dict(op="get_kmer_hashes_from_chunk_sequence (device in, device out)", kmers=n, ms=dt * 1e3, GKps=n / dt / 1e9)), flush=True)
_lib.set_option("time_kernels", 1)
m = Mapper(di, n_counts)


def run_map():
    m.reset()
    m.map_kmers(hashes)
    m.flush()
    m.sync()


dt, _ = timed(run_map)
print(json.dumps(dict(op="map_kmers (device uint64 k-mers -> node counts, incl. log apply)", kmers=n, ms=dt * 1e3, GKps=n / dt / 1e9)), flush=True)
ref = Mapper(di, n_counts)
ref.map_reads(bases, offsets, k)
same = bool((ref.counts() == m.counts()).all())
dt, mask = timed(lambda: di.in_graph_index(hashes))
print(json.dumps(dict(op="in_graph_index (device in/out)", kmers=n, ms=dt * 1e3, GKps=n / dt / 1e9, members=float(mask.float().mean().item()),
                      map_kmers_equals_map_reads=same)), flush=True)

# Fixed cost of one drop-in call (the reference's calling pattern: one map_kmers_to_graph_index per 2.5 MB chunk,
# command_line_interface.py:51): a tiny batch of host k-mers through the public function, which returns a fresh
# uint32[max_node_id + 1] every time (mapper.pyx:37), so the 320 MB read-back is part of the contract; the call on a
# device batch with the counts left on the device shows what the call itself costs (cached Mapper, reset, launch).
import numpy as np  # noqa: E402

from kmer_mapper_b200.mapper import map_kmers_to_graph_index  # noqa: E402

small = hashes[:1000].cpu().numpy()
map_kmers_to_graph_index(tindex, n_counts - 1, small)
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    out = map_kmers_to_graph_index(tindex, n_counts - 1, small)
dt_host = (time.perf_counter() - t0) / reps
small_dev = hashes[:1000]
m2 = Mapper(di, n_counts)
m2.reset(); m2.map_kmers(small_dev); m2.flush(); m2.sync()
t0 = time.perf_counter()
reps = 200
for _ in range(reps):
    m2.map_kmers(small_dev)
    m2.flush()
m2.sync()
dt_dev = (time.perf_counter() - t0) / reps
t0 = time.perf_counter()
for _ in range(reps):
    m2.reset()
    m2.map_kmers(small_dev)
    m2.flush()
m2.sync()
dt_dev_reset = (time.perf_counter() - t0) / reps
print(json.dumps(dict(op="fixed cost per call, 1000 k-mers, %d nodes" % n_counts,
                      map_kmers_to_graph_index_host_in_fresh_array_out_us=dt_host * 1e6,
                      of_which_counts_bytes_read_back=int(out.nbytes),
                      mapper_map_kmers_plus_flush_us=dt_dev * 1e6,
                      mapper_reset_map_kmers_flush_us=dt_dev_reset * 1e6)), flush=True)
