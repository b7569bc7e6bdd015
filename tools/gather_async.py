#!/usr/bin/env python
"""Random 32-byte-sector gathers: plain loads (LDG.256 equivalents) against cp.async (LDGSTS) into shared memory.
Answers: what rate of random DRAM fetches does each path sustain on one B200?  (kmb_bench_gather, modes 0 / 6 / 7)"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmer_mapper_b200 import _lib  # noqa: E402


def run(table_bytes, n_loads, load_bytes, unroll, threads, bps, mode):
    _lib.set_option("bench_load_mode", mode)
    ms = C.c_float(0)
    _lib.check(_lib.lib().kmb_bench_gather(0, table_bytes, n_loads, load_bytes, unroll, threads, bps, C.byref(ms)))
    _lib.set_option("bench_load_mode", 0)
    return ms.value


def main():
    _lib.require_device()
    table_bytes = 8 << 30
    n_loads = 1 << 29
    for table_bytes in (8 << 30, 3 << 30):
      for mode, lb, name in ((0, 8, "ld.global 8 B"), (0, 16, "ld.global 16 B"), (0, 32, "ld.global 2 x 16 B (registers)"), (8, 32, "ld.global 1 x 32 B (LDG.256, L2::64B)"),
                           (6, 32, "cp.async 2 x 16 B"), (7, 32, "cp.async 2 x 16 B, L2::64B")):
        for unroll, threads, bps in ((4, 256, 6), (2, 256, 3), (1, 256, 2)):
            ms = run(table_bytes, n_loads, lb, unroll, threads, bps, mode)
            name = name + " table %d GB" % (table_bytes >> 30)
            print(json.dumps(dict(path=name, unroll=unroll, threads=threads, blocks_per_sm=bps, in_flight_per_sm=unroll * threads * bps,
                                  ms=round(ms, 2), gathers_per_s=round(n_loads / ms * 1e3 / 1e9, 2))), flush=True)


if __name__ == "__main__":
    main()
