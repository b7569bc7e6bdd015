#!/usr/bin/env python
"""What bounds the random gather?  (a) number of SMs issuing, (b) L2 fetch granularity."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmer_mapper_b200 import _lib  # noqa: E402


def run(table_bytes, n_loads, load_bytes, unroll, threads, bps):
    ms = C.c_float(0)
    _lib.check(_lib.lib().kmb_bench_gather(0, table_bytes, n_loads, load_bytes, unroll, threads, bps, C.byref(ms)))
    return ms.value


_lib.require_device()
T = 452_930_477 * 8
N = 1 << 29
for gran in (0, 32, 64, 128):
    _lib.set_option("l2_fetch_granularity", gran)
    for blocks in (0, 37, 74, 148, 296, 592):
        _lib.set_option("bench_grid_blocks", blocks)
        for lb in (8, 32):
            ms = run(T, N, lb, 8, 256, 8)
            print(json.dumps(dict(gran=gran, grid_blocks=blocks, load_bytes=lb, ms=ms, Ggathers_per_s=N / ms / 1e6)), flush=True)
