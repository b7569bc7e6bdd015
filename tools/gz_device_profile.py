#!/usr/bin/env python
"""FASTQ.gz -> counts: gzip members inflated by GPU warps (kmb_mapper_map_gz) against the host decoders, on the bench's
file (members of ~4 MB of text) and on a BGZF-like one (members of 65280 bytes of text, what bgzip writes).
    python tools/gz_device_profile.py [n_reads]"""
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from kmer_mapper_b200 import _lib  # noqa: E402
from kmer_mapper_b200 import command_line_interface as cli  # noqa: E402
from kmer_mapper_b200.device import DeviceIndex, Mapper  # noqa: E402
from kmer_mapper_b200.reader import open_reads  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
w = bench.workload("config2", 1.0)
w["reads"] = n_reads
tindex, bases, offsets = bench.generate(w, 0, torch.device("cuda", 0))
di = DeviceIndex.from_index(tindex, device=0)
n_counts = tindex.max_node_id() + 1
d = tempfile.mkdtemp(prefix="kmb_files_", dir="/dev/shm")
fq, fqgz, bgzf = bench.write_fastq_files(d, bases, n_reads, w["read_len"])
size = os.path.getsize(fq)
_lib.set_option("gz_device_max_mean_member_bytes", 64 << 20)    # the device route also for the large members

mapper = Mapper(di, n_counts)
mapper.map_reads(bases[:n_reads * w["read_len"]], offsets[:n_reads + 1], w["k"])
want = mapper.counts()
out = torch.empty(n_counts, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
for path in (fqgz, bgzf):
    for route in ("device", "host"):
        os.environ[cli.GZ_ENV] = route
        for rep in range(3):
            mapper.reset()
            t0 = time.perf_counter()
            reads = open_reads(path)
            n_chunks, host_from = cli.map_file_text(mapper, reads, w["k"])
            t1 = time.perf_counter()
            mapper.counts(out=out)
            dt = time.perf_counter() - t0
            reads.close()
            a, b, c = (_lib.C.c_uint64() for _ in range(3))
            _lib.lib().kmb_gz_device_stats(_lib.C.byref(a), _lib.C.byref(b), _lib.C.byref(c))
            print(json.dumps(dict(file=os.path.basename(path), gz_bytes=os.path.getsize(path), text_bytes=size, route=route, rep=rep,
                                  seconds=round(dt, 4), submit_s=round(t1 - t0, 4), reads_per_s_M=round(n_reads / dt / 1e6, 2),
                                  text_GBps=round(size / dt / 1e9, 2), counts_equal=bool(np.array_equal(out, want)),
                                  host_from=host_from, device_members=a.value, host_redone_batches=c.value)), flush=True)
mapper.close()
