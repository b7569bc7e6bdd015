#!/usr/bin/env python
"""L2 policy variants of the fused kernel on the config-2 index (run plain for timings, or under
ncu -k regex:kmb_map_reads --metrics ... to see DRAM sectors and RED hit rates per variant)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from kmer_mapper_b200 import _lib  # noqa: E402
from kmer_mapper_b200.device import DeviceIndex, Mapper  # noqa: E402

VARIANTS = [  # (policy_filter, policy_line, l2_persist)
    (2, 0, 1), (2, 0, 0), (2, 1, 1), (0, 0, 0)]


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    w = bench.workload("config2", 1.0)
    w["reads"] = n_reads
    tindex, bases, offsets = bench.generate(w, 0, torch.device("cuda", 0))
    di = DeviceIndex.from_index(tindex, device=0)
    n_counts = tindex.max_node_id() + 1
    _lib.set_option("time_kernels", 1)
    ref = None
    for v in VARIANTS:
        for name, val in zip(("policy_filter", "policy_line", "l2_persist"), v):
            _lib.set_option(name, val)
        m = Mapper(di, n_counts)
        m.map_reads(bases, offsets, w["k"])     # warm-up: brings the filter into L2 under this policy set
        m.flush()
        m.kernel_time()
        m.reset()
        m.map_reads(bases, offsets, w["k"])
        m.flush()
        ms, n = m.kernel_time()
        nk, nc = m.stats()
        c = m.counts()
        if ref is None:
            ref = c
        print(json.dumps(dict(variant=v, kernel_ms=ms / n, GKps=nk / (ms / n) / 1e6, same=bool((c == ref).all()))), flush=True)
        m.close()


if __name__ == "__main__":
    main()
