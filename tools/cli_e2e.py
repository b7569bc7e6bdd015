#!/usr/bin/env python
"""End-to-end reads/s of the CLI: FASTA / FASTQ / FASTQ.gz file -> <out>.npy, wall clock from file open to counts
on the host.  Index: config-2 shape at 1/10 scale (10 M entries) so that the .npz round trip stays short."""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from kmer_mapper_b200 import synthetic  # noqa: E402
from kmer_mapper_b200.command_line_interface import run_argument_parser  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
w = bench.workload("config2", 0.1)
w["reads"] = n_reads
d = tempfile.mkdtemp(prefix="kmb_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
tindex, bases, offsets = bench.generate(w, 0, torch.device("cuda", 0))
tindex.to_host().to_file(os.path.join(d, "index.npz"))
hb, ho = bases.cpu().numpy(), offsets.cpu().numpy()
del tindex, bases, offsets
torch.cuda.empty_cache()
t = time.perf_counter()
synthetic.write_fastq(os.path.join(d, "reads.fq"), hb, ho)
synthetic.write_fastq(os.path.join(d, "reads.fq.gz"), hb, ho, members=64)      # multi-member, like bgzip output
os.system("gzip -1 -c %s/reads.fq > %s/reads_single.fq.gz" % (d, d))              # one member: sequential inflate
synthetic.write_fasta(os.path.join(d, "reads.fa"), hb, ho)
print("wrote files in %.1f s" % (time.perf_counter() - t), file=sys.stderr)
ref = None
for name in ("reads.fa", "reads.fq", "reads.fq.gz", "reads_single.fq.gz"):
    for chunk in (2_500_000, 10_000_000, 64_000_000):
        out = os.path.join(d, "out")
        t0 = time.perf_counter()
        run_argument_parser(["map", "-i", os.path.join(d, "index.npz"), "-f", os.path.join(d, name), "-o", out, "-k", "31",
                             "-c", str(chunk)])
        dt = time.perf_counter() - t0
        c = np.load(out + ".npy")
        if ref is None:
            ref = c
        print(json.dumps(dict(file=name, file_MB=round(os.path.getsize(os.path.join(d, name)) / 1e6), chunk_bytes=chunk,
                              seconds=round(dt, 3), reads_per_s=round(n_reads / dt), Mbases_per_s=round(hb.shape[0] / dt / 1e6),
                              counts_equal_first=bool(np.array_equal(c, ref)), host_cores=os.cpu_count())), flush=True)
import shutil
shutil.rmtree(d, ignore_errors=True)
