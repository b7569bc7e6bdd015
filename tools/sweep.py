#!/usr/bin/env python
"""Tuning sweep on one GPU: build one workload, time the resident fused path under every option set.
    python tools/sweep.py [--workload config2] [--scale 1.0] [--steps 3]"""
import argparse
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from kmer_mapper_b200 import _lib  # noqa: E402
from kmer_mapper_b200.device import DeviceIndex, Mapper  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--reads", type=int, default=0)
    ap.add_argument("--grid", default="default")
    a = ap.parse_args()
    w = bench.workload(a.workload, a.scale)
    if a.reads:
        w["reads"] = a.reads
    tindex, bases, offsets = bench.generate(w, 0, torch.device("cuda", 0))
    n_counts = tindex.max_node_id() + 1
    ref_counts = None
    grids = {
        "default": dict(use_filter=[1, 0], gathers_in_flight=[4, 8, 16], map_reads_blocks_per_sm=[0]),
        "occupancy": dict(use_filter=[1], gathers_in_flight=[4, 8, 16], map_reads_blocks_per_sm=[1, 2, 3, 4, 5, 6]),
        "v21": dict(use_filter=[1], gathers_in_flight=[2, 4], map_reads_blocks_per_sm=[0, 2]),
        "filter": dict(gathers_in_flight=[2], filter_l2_budget_bytes=[40 << 20, 48 << 20, 52 << 20, 56 << 20, 60 << 20], use_filter=[1]),
        "size": dict(gathers_in_flight=[2], filter_l2_budget_bytes=[54 << 20, 57 << 20], sectors_per_100_entries=[200, 226, 250, 300], use_filter=[1]),
        "persist": dict(use_filter=[1], gathers_in_flight=[8], map_reads_blocks_per_sm=[0], l2_persist=[0, 1]),
        "r2": dict(apply_window_log2=[0]),
        "filt": dict(use_filter=[1, 0]),
        "slabs": dict(apply_slabs_per_sm=[8, 16, 32]),
        "carve": dict(map_carveout=[-1, 100, 72, 58]),
        "cta2": dict(map_reads_blocks_per_sm=[0, 2]),
        "u": dict(gathers_in_flight=[4, 8]),
        "window3": dict(apply_window_log2=[26, 25, 24, 23]),
        "window": dict(apply_window_log2=[26, 25, 24, 23, 22, 21]),
    }[a.grid]
    names = list(grids)
    last_filter = None
    di = None
    for combo in itertools.product(*[grids[n] for n in names]):
        opts = dict(zip(names, combo))
        for n, v in opts.items():
            _lib.set_option(n, v)
        if di is None or (opts.get("use_filter"), opts.get("filter_l2_budget_bytes"), opts.get("sectors_per_100_entries")) != last_filter:
            if hasattr(tindex, "_kmb_device_index"):
                del tindex._kmb_device_index
            di = None
            torch.cuda.empty_cache()
            di = DeviceIndex.from_index(tindex, device=0)
            last_filter = (opts.get("use_filter"), opts.get("filter_l2_budget_bytes"), opts.get("sectors_per_100_entries"))
        _lib.set_option("time_kernels", 1)
        m = Mapper(di, n_counts)
        for _ in range(2):
            m.reset()
            m.map_reads(bases, offsets, w["k"])
        m.kernel_time()
        m.apply_time()
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
        for _ in range(a.steps):
            m.reset()
            m.map_reads(bases, offsets, w["k"])
            m.flush()
        m.sync()
        step_ms = (time.perf_counter() - t0) * 1e3 / a.steps
        ms, n = m.kernel_time()
        ams, an = m.apply_time()
        nk, nc = m.stats()
        ncand = m.candidates()
        c = m.counts()
        if ref_counts is None:
            ref_counts = c
        same = bool((c == ref_counts).all())
        m.close()
        print(json.dumps(dict(opts=opts, cand_per_kmer=round(ncand / max(nk, 1), 4), hits_per_kmer=round(nc / max(nk, 1), 4), apply_ms=ams / max(an, 1), kernel_ms=ms / n, kernel_GKps=nk / (ms / n) / 1e6, step_ms=step_ms, step_GKps=nk / step_ms / 1e6,
                              filter_bytes=di.filter_bytes, overflow_lines=di.n_overflow_lines,
                              counts_equal_first=same)), flush=True)


if __name__ == "__main__":
    main()
