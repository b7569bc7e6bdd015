#!/usr/bin/env python
"""Multi-GPU CLI check: `python -m kmer_mapper_b200 map` as one process and as N torchrun ranks (byte-range shards of
plain files, every N-th chunk of a .gz, one all-reduce of the counts) must write identical <out>.npy files.
Usage: python tools/cli_ranks_check.py [n_ranks] [n_reads]      (needs n_ranks GPUs; JSON lines on stdout)"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from kmer_mapper_b200 import synthetic  # noqa: E402
from tools.cli_e2e import write_fixed_length  # noqa: E402


def main():
    n_ranks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
    w = bench.workload("config2", 0.05)
    w["reads"] = n_reads
    L = w["read_len"]
    d = tempfile.mkdtemp(prefix="kmb_ranks_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    tindex, bases, offsets = bench.generate(w, 0, torch.device("cuda", 0))
    tindex.to_host().to_file(os.path.join(d, "index.npz"))
    hb, ho = bases.cpu().numpy(), offsets.cpu().numpy()
    del tindex, bases, offsets
    torch.cuda.empty_cache()
    write_fixed_length(os.path.join(d, "reads.fa"), hb, n_reads, L, fastq=False)
    write_fixed_length(os.path.join(d, "reads.fq"), hb, n_reads, L, fastq=True)
    n_gz = n_reads // 8
    synthetic.write_fastq(os.path.join(d, "reads.fq.gz"), hb[:n_gz * L], ho[:n_gz + 1], members=16)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root)
    ok = True
    for name in ("reads.fa", "reads.fq", "reads.fq.gz"):
        args = ["map", "-i", os.path.join(d, "index.npz"), "-f", os.path.join(d, name), "-k", "31", "-c", "10000000"]
        t0 = time.perf_counter()
        subprocess.check_call([sys.executable, "-m", "kmer_mapper_b200"] + args + ["-o", os.path.join(d, "one")], env=env,
                              stdout=subprocess.DEVNULL)
        t1 = time.perf_counter()
        subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n_ranks),
                               "--master-addr", "127.0.0.1", "--master-port", "29631", "-m", "kmer_mapper_b200"] + args +
                              ["-o", os.path.join(d, "many")], env=env, stdout=subprocess.DEVNULL)
        t2 = time.perf_counter()
        a, b = np.load(os.path.join(d, "one.npy")), np.load(os.path.join(d, "many.npy"))
        same = bool(a.dtype == b.dtype and np.array_equal(a, b))
        ok &= same
        print(json.dumps(dict(file=name, ranks=n_ranks, counts_identical=same, total_counts=int(a.astype(np.uint64).sum()),
                              seconds_one_process=round(t1 - t0, 2), seconds_ranks=round(t2 - t1, 2))), flush=True)
    shutil.rmtree(d, ignore_errors=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
