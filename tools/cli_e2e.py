#!/usr/bin/env python
"""End-to-end reads/s of the CLI: FASTA / FASTQ / FASTQ.gz file -> <out>.npy.

Two clocks per run: `seconds` = the whole `kmer_mapper map ...` call (index .npz load + device index build +
streaming + counts to disk), `map_seconds` = the streaming part alone, taken from the CLI's own log line
("Time spent only on hashing and counting hashes", command_line_interface.py:139 in the reference).
Index: config-2 shape at 1/10 scale (10 M entries) so that the .npz round trip stays short.
Usage: python tools/cli_e2e.py [n_reads_plain] [n_reads_gz]
"""
import json
import logging
import os
import re
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from kmer_mapper_b200 import synthetic  # noqa: E402
from kmer_mapper_b200.command_line_interface import run_argument_parser  # noqa: E402


def write_fixed_length(path, bases, n, L, fastq):
    """Constant-length reads, one vectorised write (the per-read Python loop of synthetic.write_* takes minutes
    for tens of millions of reads)."""
    ids = ((np.arange(n, dtype=np.int64)[:, None] // 10 ** np.arange(8, -1, -1)) % 10 + 48).astype(np.uint8)
    nl = np.full((n, 1), 10, np.uint8)
    parts = [np.full((n, 1), ord("@" if fastq else ">"), np.uint8), ids, nl, bases.reshape(n, L), nl]
    if fastq:
        parts += [np.full((n, 1), ord("+"), np.uint8), nl, np.full((n, L), ord("I"), np.uint8), nl]
    np.concatenate(parts, axis=1).tofile(path)


class _MapSeconds(logging.Handler):
    def __init__(self):
        super().__init__(logging.INFO)
        self.seconds = None

    def emit(self, record):
        m = re.search(r"Time spent only on hashing and counting hashes: ([0-9.]+)", record.getMessage())
        if m:
            self.seconds = float(m.group(1))


def main():
    n_plain = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
    n_gz = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
    w = bench.workload("config2", 0.1)
    w["reads"] = n_plain
    L = w["read_len"]
    d = tempfile.mkdtemp(prefix="kmb_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    tindex, bases, offsets = bench.generate(w, 0, torch.device("cuda", 0))
    tindex.to_host().to_file(os.path.join(d, "index.npz"))
    hb, ho = bases.cpu().numpy(), offsets.cpu().numpy()
    del tindex, bases, offsets
    torch.cuda.empty_cache()
    t = time.perf_counter()
    write_fixed_length(os.path.join(d, "reads.fa"), hb, n_plain, L, fastq=False)
    write_fixed_length(os.path.join(d, "reads.fq"), hb, n_plain, L, fastq=True)
    # members of ~2000 reads (0.6 MB of text): between bgzip's 64 KB blocks and the multi-MB members of `cat *.gz`
    synthetic.write_fastq(os.path.join(d, "reads.fq.gz"), hb[:n_gz * L], ho[:n_gz + 1], members=max(64, n_gz // 2000))
    write_fixed_length(os.path.join(d, "gz_plain.fq"), hb[:n_gz * L], n_gz, L, fastq=True)
    os.system("gzip -1 -c %s/gz_plain.fq > %s/reads_single.fq.gz" % (d, d))                       # one member
    print("wrote files in %.1f s" % (time.perf_counter() - t), file=sys.stderr)
    handler = _MapSeconds()
    logging.getLogger().addHandler(handler)
    logging.getLogger().setLevel(logging.INFO)
    ref = {}
    for name, n_reads in (("reads.fa", n_plain), ("reads.fq", n_plain), ("reads.fq.gz", n_gz), ("reads_single.fq.gz", n_gz)):
        for chunk in (10_000_000, 64_000_000, 256_000_000):
            out = os.path.join(d, "out")
            handler.seconds = None
            t0 = time.perf_counter()
            run_argument_parser(["map", "-i", os.path.join(d, "index.npz"), "-f", os.path.join(d, name), "-o", out,
                                 "-k", "31", "-c", str(chunk)])
            dt = time.perf_counter() - t0
            c = np.load(out + ".npy")
            ref.setdefault(n_reads, c)
            ms = handler.seconds or dt
            print(json.dumps(dict(file=name, file_MB=round(os.path.getsize(os.path.join(d, name)) / 1e6), reads=n_reads,
                                  chunk_bytes=chunk, seconds=round(dt, 3), map_seconds=round(ms, 3),
                                  reads_per_s=round(n_reads / dt), map_reads_per_s=round(n_reads / ms),
                                  map_Mbases_per_s=round(n_reads * L / ms / 1e6),
                                  counts_equal_first=bool(np.array_equal(c, ref[n_reads])),
                                  host_cores=len(os.sched_getaffinity(0)))), flush=True)
    shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
