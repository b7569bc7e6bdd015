#!/usr/bin/env python
"""Random 8-byte gathers from tables of growing size: where does the L2 stop holding the table?"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmer_mapper_b200 import _lib  # noqa: E402

_lib.require_device()
for mb in (16, 32, 48, 56, 60, 64, 68, 72, 80, 88, 96, 104, 112, 120, 128, 144, 160, 192, 256):
    ms = C.c_float(0)
    n = 1 << 29
    _lib.check(_lib.lib().kmb_bench_gather(0, mb << 20, n, 8, 8, 256, 8, C.byref(ms)))
    print(json.dumps(dict(table_MB=mb, ms=round(ms.value, 3), Ggathers_per_s=round(n / ms.value / 1e6, 1))), flush=True)
