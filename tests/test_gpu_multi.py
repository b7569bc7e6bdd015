"""The multi-GPU path on real GPUs: reads sharded over two ranks (one process per GPU, torchrun), private count arrays,
one all-reduce through the C ABI (kmb_mapper_allreduce = ncclAllReduce(ncclUint32, ncclSum)) -- the reference's additive
map-reduce (command_line_interface.py:124-130).  Every result is compared with the ORACLE on all the reads, not with a
one-rank GPU run.  Skipped below two devices (gpurun --gpus 2 runs them)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from oracle import c_oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_rank_worker.py")


def _n_devices():
    from kmer_mapper_b200 import _lib
    return _lib.device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(n, *worker_args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), WORKER] + [str(a) for a in worker_args]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + "\n" + out.stderr[-3000:]


@pytest.fixture(scope="module")
def job(tmp_path_factory):
    if _n_devices() < 2:
        pytest.skip("needs two GPUs")
    from kmer_mapper_b200 import synthetic
    d = tmp_path_factory.mktemp("multi")
    k = 31
    g = synthetic.make_genome(400_000, 21)
    idx = synthetic.make_index(g, 60_000, k, 45_000, 300_007, 22, n_hot_nodes=1200)
    idx.to_file(str(d / "index.npz"))
    bases, offsets = synthetic.make_reads(g, 9_000, 150, seed=23, n_rate=0.01, lower_rate=0.3, ragged=True)
    np.savez(str(d / "reads.npz"), bases=bases, offsets=offsets, k=k)
    synthetic.write_fasta(str(d / "reads.fa"), bases, offsets, line_width=70)
    synthetic.write_fastq(str(d / "reads.fq"), bases, offsets)
    synthetic.write_fastq(str(d / "reads.fq.gz"), bases, offsets, members=6)
    want, n_kmers = c_oracle.map_reads(idx, idx.max_node_id(), bases, offsets, k, n_threads=4)
    assert want.sum() > 1000
    return dict(dir=d, want=want, n_kmers=n_kmers)


def test_two_ranks_allreduce_through_the_c_abi_vs_oracle(job):
    out = str(job["dir"] / "api_counts.npy")
    _torchrun(2, "api", job["dir"], out)
    got = np.load(out)
    assert got.dtype == np.uint32 and np.array_equal(got, job["want"])
    stats = sum(np.load(out + ".rank%d.stats.npy" % r) for r in range(2))
    assert stats[0] == job["n_kmers"]                      # every window was mapped by exactly one rank
    assert stats[1] == int(job["want"].astype(np.int64).sum())   # sum of the reduced counts == entries counted by all ranks


@pytest.mark.parametrize("reads", ["reads.fa", "reads.fq", "reads.fq.gz"])
def test_cli_under_torchrun_two_ranks_vs_oracle(job, reads):
    out = str(job["dir"] / ("cli_" + reads.replace(".", "_")))
    _torchrun(2, "cli", job["dir"], reads, out)
    got = np.load(out + ".npy")
    assert got.dtype == np.uint32 and np.array_equal(got, job["want"])
