"""The N>1 path on CPU: world_size-2 gloo job.  Each rank takes its shard of the reads (contiguous
read ranges, and the CLI's round-robin chunks), produces a private uint32 count array -- here with the
CPU oracle standing in for the GPU kernels, which is all a GPU-less box can run -- and the arrays are
summed by distributed.all_reduce_counts.  The result must equal the unsharded count array bit for bit,
including uint32 wrap-around."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from kmer_mapper_b200 import distributed, synthetic
    from oracle import c_oracle
    r, w, _ = distributed.init_process_group(backend="gloo")
    assert (r, w) == (rank, world_size)
    g = synthetic.make_genome(100_000, 1)
    idx = synthetic.make_index(g, 20_000, 21, 5000, 65_537, 2)
    bases, offsets = synthetic.make_reads(g, 4001, 120, seed=3, ragged=True)
    mx = idx.max_node_id()
    # (a) contiguous read ranges
    lo, hi, b0, b1 = distributed.shard_reads(offsets, rank, world_size)
    mine, _ = c_oracle.map_reads(idx, mx, bases[b0:b1], offsets[lo:hi + 1] - b0, 21)
    mine[7] += np.uint32(0xFFFFFFF0)                 # force wrap-around in the sum
    total = distributed.all_reduce_counts(mine.copy())
    # (b) round-robin chunks of 500 reads, as the CLI assigns them
    part = np.zeros(mx + 1, dtype=np.uint32)
    for ci, s in enumerate(range(0, 4001, 500)):
        if distributed.chunk_belongs_to_rank(ci, rank, world_size):
            e = min(s + 500, 4001)
            c, _ = c_oracle.map_reads(idx, mx, bases[offsets[s]:offsets[e]], offsets[s:e + 1] - offsets[s], 21)
            part += c
    import torch
    t = torch.from_numpy(part.view(np.int32))
    total2 = distributed.all_reduce_counts(t).numpy().view(np.uint32)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.stack([total, total2]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharded_counts_equal_unsharded(tmp_path):
    import torch.multiprocessing as mp
    from kmer_mapper_b200 import synthetic
    from oracle import c_oracle
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = synthetic.make_genome(100_000, 1)
    idx = synthetic.make_index(g, 20_000, 21, 5000, 65_537, 2)
    bases, offsets = synthetic.make_reads(g, 4001, 120, seed=3, ragged=True)
    want, _ = c_oracle.map_reads(idx, idx.max_node_id(), bases, offsets, 21)
    wrapped = want.copy()
    wrapped[7] += np.uint32((2 * 0xFFFFFFF0) & 0xFFFFFFFF)
    for r in range(world):
        got = np.load(str(tmp_path / ("r%d.npy" % r)))
        assert got.dtype == np.uint32
        assert np.array_equal(got[0], wrapped)
        assert np.array_equal(got[1], want)
