"""One rank of a multi-GPU test job (started by torchrun from tests/test_gpu_multi.py; not collected by pytest).

    api <dir> <out>            this rank maps its contiguous shard of <dir>/reads.npz through the Mapper and the counts are
                               summed by kmb_mapper_allreduce (NCCL behind the C ABI); rank 0 saves them
    cli <dir> <reads> <out>    the reference-shaped CLI under torchrun (the reader shards the file, one all-reduce)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    mode, d = sys.argv[1], sys.argv[2]
    from kmer_mapper_b200 import distributed
    if mode == "api":
        import torch
        from kmer_mapper_b200.device import DeviceIndex, Mapper
        from kmer_mapper_b200.kmer_index import KmerIndex
        rank, world, local_rank = distributed.init_process_group("nccl")
        torch.cuda.set_device(local_rank)
        idx = KmerIndex.from_file(os.path.join(d, "index.npz"))
        idx.convert_to_int32()
        z = np.load(os.path.join(d, "reads.npz"))
        bases, offsets, k = z["bases"], z["offsets"], int(z["k"])
        r_lo, r_hi, b_lo, b_hi = distributed.shard_reads(offsets, rank, world)
        di = DeviceIndex.from_index(idx, device=local_rank)
        m = Mapper(di)
        comm = distributed.Comm(device=local_rank)
        # two calls, so that the reduction also covers counts accumulated over several chunks
        mid = (r_lo + r_hi) // 2
        for a, b in ((r_lo, mid), (mid, r_hi)):
            if b > a:
                m.map_reads(np.ascontiguousarray(bases[offsets[a]:offsets[b]]), np.ascontiguousarray(offsets[a:b + 1] - offsets[a]), k)
        comm.all_reduce(m)
        got = m.counts()
        n_kmers, n_counted = m.stats()
        if rank == 0:
            np.save(sys.argv[3], got)
        np.save(sys.argv[3] + ".rank%d.stats" % rank, np.array([n_kmers, n_counted], dtype=np.int64))
        comm.close()
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    elif mode == "cli":
        from kmer_mapper_b200.command_line_interface import run_argument_parser
        run_argument_parser(["map", "-i", os.path.join(d, "index.npz"), "-f", os.path.join(d, sys.argv[3]), "-o", sys.argv[4],
                             "-k", "31", "-c", "60000"])
    else:
        raise SystemExit("unknown mode " + mode)


if __name__ == "__main__":
    main()
