#!/usr/bin/env python
"""Where does file -> counts spend its time?  FASTQ text through the device-side parser (kmb_mapper_map_text), with the
library's own host-side timers: waiting for a slot, staging pageable text into pinned memory, waiting for H2D + parse.
    python tools/text_route_profile.py [n_reads]"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from kmer_mapper_b200 import _lib  # noqa: E402
from kmer_mapper_b200.device import DeviceIndex, Mapper  # noqa: E402
from kmer_mapper_b200.reader import open_reads  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
w = bench.workload("config2", 1.0)
w["reads"] = n_reads
tindex, bases, offsets = bench.generate(w, 0, torch.device("cuda", 0))
di = DeviceIndex.from_index(tindex, device=0)
n_counts = tindex.max_node_id() + 1
d = tempfile.mkdtemp(prefix="kmb_files_", dir="/dev/shm")
fq, fqgz, _bgzf = bench.write_fastq_files(d, bases, n_reads, w["read_len"])
names = ("text_us_slot", "text_us_stage", "text_us_wait", "text_us_alloc")
for path in (fq, fqgz):
    out = torch.empty(n_counts, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    for chunk in (16 << 20, 64 << 20, 256 << 20):
        m = Mapper(di, n_counts)
        for rep in range(2):
            before = {n: _lib.get_option(n) for n in names}
            t0 = time.perf_counter()
            reads = open_reads(path)
            t_iter = 0.0
            t_map = 0.0
            it = reads.text_chunks(min_chunk_size=chunk)
            n_chunks = 0
            while True:
                a = time.perf_counter()
                tc = next(it, None)
                b = time.perf_counter()
                t_iter += b - a
                if tc is None:
                    break
                m.map_text(tc, reads.format, w["k"])
                t_map += time.perf_counter() - b
                n_chunks += 1
            c = time.perf_counter()
            m.counts(out=out)
            t_counts = time.perf_counter() - c
            total = time.perf_counter() - t0
            reads.close()
            m.reset()
            rec = dict(file=os.path.basename(path), chunk_mb=chunk >> 20, rep=rep, chunks=n_chunks, total_s=round(total, 3),
                       reads_per_s=round(n_reads / total / 1e6, 2), reader_s=round(t_iter, 3), map_text_s=round(t_map, 3),
                       counts_s=round(t_counts, 3))
            rec.update({n[8:] + "_s": round((_lib.get_option(n) - before[n]) / 1e6, 3) for n in names})
            print(json.dumps(rec), flush=True)
        m.close()
