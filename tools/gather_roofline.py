#!/usr/bin/env python
"""Random-gather micro-roofline: uniform random 8/16/32-byte loads from a table of the
directory's footprint, device-timed.  Prints one JSON line per shape; the best 8-byte figure is the
denominator for `fraction_of_gather_roofline`."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmer_mapper_b200 import _lib  # noqa: E402


def run(table_bytes, n_loads, load_bytes, unroll, threads, blocks_per_sm):
    ms = C.c_float(0)
    _lib.check(_lib.lib().kmb_bench_gather(0, table_bytes, n_loads, load_bytes, unroll, threads, blocks_per_sm, C.byref(ms)))
    return ms.value


def main():
    _lib.require_device()
    sizes = [int(x) for x in sys.argv[1:]] or [452_930_477 * 8, 57 << 20, 1_000_000_007 * 8]
    n_loads = 1 << 30
    for table_bytes in sizes:
        best = None
        for load_bytes in (8, 16, 32):
            for unroll, threads, bps in ((4, 256, 8), (8, 256, 8), (16, 256, 4), (8, 512, 4), (16, 512, 2), (8, 1024, 2)):
                ms = run(table_bytes, n_loads, load_bytes, unroll, threads, bps)
                rec = dict(table_bytes=table_bytes, load_bytes=load_bytes, unroll=unroll, threads=threads, blocks_per_sm=bps,
                           ms=ms, gathers_per_s=n_loads / ms * 1e3, sector_GBps=n_loads * 32 / ms * 1e3 / 1e9)
                print(json.dumps(rec), flush=True)
                if load_bytes == 8 and (best is None or rec["gathers_per_s"] > best["gathers_per_s"]):
                    best = rec
        print(json.dumps(dict(best_8B=best)), flush=True)


if __name__ == "__main__":
    main()
