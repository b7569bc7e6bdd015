#!/usr/bin/env python
"""DRAM sectors per random 8-byte load for different load flavours (run under ncu with dram__sectors_read.sum)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmer_mapper_b200 import _lib  # noqa: E402

_lib.require_device()
T = 452_930_477 * 8
N = 1 << 28
for mode in (0, 1, 2, 3, 4, 5):
    _lib.set_option("bench_load_mode", mode)
    ms = C.c_float(0)
    _lib.check(_lib.lib().kmb_bench_gather(0, T, N, 8, 8, 256, 8, C.byref(ms)))
    print(mode, ms.value, N / ms.value / 1e6, flush=True)
