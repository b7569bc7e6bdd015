#!/usr/bin/env python
"""Where does the fused kernel's time go?  Switch parts of it off (results become wrong, timing only)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from kmer_mapper_b200 import _lib  # noqa: E402
from kmer_mapper_b200.device import DeviceIndex, Mapper  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
w = bench.workload(sys.argv[2] if len(sys.argv) > 2 else "config2", 1.0)
w["reads"] = n_reads
tindex, bases, offsets = bench.generate(w, 0, torch.device("cuda", 0))
di = DeviceIndex.from_index(tindex, device=0)
n_counts = tindex.max_node_id() + 1
_lib.set_option("time_kernels", 1)
names = {0: "full", 1: "no RED", 8: "no key loads (and no RED)", 2: "no line loads", 4: "no filter loads (no candidates)",
         6: "compute only"}
for u in (2, 4):
    _lib.set_option("gathers_in_flight", u)
    for ab in (0, 1, 2, 4):
        _lib.set_option("ablate", ab)
        m = Mapper(di, n_counts)
        m.map_reads(bases, offsets, w["k"])
        m.flush()
        m.kernel_time()
        m.reset()
        m.map_reads(bases, offsets, w["k"])
        m.flush()
        ms, n = m.kernel_time()
        nk, _ = m.stats()
        print(json.dumps(dict(U=u, ablate=names[ab], kernel_ms=round(ms / n, 2), GKps=round(nk / (ms / n) / 1e6, 1))), flush=True)
        m.close()
_lib.set_option("ablate", 0)
