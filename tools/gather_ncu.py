#!/usr/bin/env python
"""Run under `ncu --metrics dram__sectors_read.sum,...`: how many DRAM sectors does one random 8-byte load cost?"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmer_mapper_b200 import _lib  # noqa: E402

_lib.require_device()
T = 452_930_477 * 8
N = 1 << 28
for gran in (0, 32, 128):
    _lib.set_option("l2_fetch_granularity", gran)
    for lb in (8, 32):
        ms = C.c_float(0)
        _lib.check(_lib.lib().kmb_bench_gather(0, T, N, lb, 8, 256, 8, C.byref(ms)))
        print(gran, lb, ms.value, N / ms.value / 1e6, flush=True)
