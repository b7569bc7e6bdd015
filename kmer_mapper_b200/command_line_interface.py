"""Drop-in for the reference's kmer_mapper/command_line_interface.py: same sub-command, flags,
defaults, log lines, return value and ``<output>.npy`` file; the per-chunk work (N policy, 2-bit
encoding, rolling k-mers, index probe, per-node counts) runs as one fused CUDA kernel per chunk.

Differences from the reference, all deliberate:
  * both the default route and ``--gpu True`` compute on the GPU and both return the CPU route's
    integers: ``uint32[max_node_id+1]``, frequency cut-off 1000 applied, upper-case N -> A applied
    (the reference's GPU route returns unfiltered float64 counts, gpu_counter.py:37);
  * ``-t/--n-threads`` is accepted and ignored (no worker processes; one process per GPU);
  * launched under torchrun, every rank maps the chunks ``i % world_size == rank`` and the count
    arrays are summed by one all-reduce; rank 0 writes the output.
"""
from __future__ import annotations

import argparse
import logging
import os
import sys
import time

import numpy as np

logging.basicConfig(stream=sys.stdout, level=logging.INFO, format='%(asctime)s %(levelname)s: %(message)s')

from . import distributed  # noqa: E402
from .device import DEFAULT_MAX_FREQUENCY, DeviceIndex, Mapper  # noqa: E402
from .reader import open_reads  # noqa: E402
from .sequences import as_ragged  # noqa: E402
from .util import _get_kmer_index_from_args, log_memory_usage_now  # noqa: E402,F401


def main():
    run_argument_parser(sys.argv[1:])


def map_cpu(args, kmer_index, chunk_sequence):
    """command_line_interface.py:32-56: one chunk -> ``uint32[max_node_id+1]``: N -> A (:41), hashes (:42),
    lookup + count (:51).  ``args`` is the dict the reference passes (``kmer_size``); ``chunk_sequence``
    is the ragged sequence itself (the reference passes a shared-memory name for it)."""
    kmer_size = args["kmer_size"]
    t = time.perf_counter()
    seq = as_ragged(chunk_sequence)
    logging.debug("N sequences in chunk: %d" % len(seq))
    di = DeviceIndex.from_index(kmer_index)
    m = Mapper(di, kmer_index.max_node_id() + 1, DEFAULT_MAX_FREQUENCY)
    try:
        m.map_reads(seq.bases, seq.offsets, kmer_size, revcomp=False, n_to_a=True)
        mapped = m.counts()
    finally:
        m.close()
    logging.debug("Chunk of %d reads took %.2f sec" % (len(seq), time.perf_counter() - t))
    return mapped


def map_gpu(index, chunks, k, hash_map_size=0, map_reverse_complements=False, rank=0, world_size=1,
            return_mapper=False):
    """command_line_interface.py:59-79: map an iterable of chunks (objects with ``.sequence``) against the
    index on the current GPU; returns the node counts.  ``hash_map_size`` is accepted for signature
    compatibility (the device index keeps the reference's own modulo-bucketed layout, so there is no
    separate hash-map capacity to choose)."""
    logging.info("Making counter")
    di = DeviceIndex.from_index(index)
    mapper = Mapper(di, di.max_node_id() + 1, DEFAULT_MAX_FREQUENCY)
    logging.info("CUDA counter initialized")
    t_start = time.perf_counter()
    n_reads = 0
    for i, chunk in enumerate(chunks):
        if not distributed.chunk_belongs_to_rank(i, rank, world_size):
            continue
        t0 = time.perf_counter()
        seq = as_ragged(chunk.sequence if hasattr(chunk, "sequence") else chunk)
        mapper.map_reads(seq.bases, seq.offsets, k, revcomp=bool(map_reverse_complements), n_to_a=True)
        n_reads += len(seq)
        logging.debug("GPU: Whole chunk finished in %.5f sec", (time.perf_counter() - t0))
    mapper.sync()
    logging.info("Time spent only on hashing and counting hashes: %.5f" % (time.perf_counter() - t_start))
    if return_mapper:
        return mapper
    counts = mapper.counts()
    mapper.close()
    return counts


def map_bnp(args):
    """command_line_interface.py:82-151."""
    if getattr(args, "debug", None):
        logging.info("Will print debug log")
        logging.getLogger().setLevel(logging.DEBUG)

    k = args.kmer_size
    start_time = time.perf_counter()
    kmer_index = _get_kmer_index_from_args(args)

    n_bytes = os.stat(args.reads).st_size
    if args.reads.endswith(".gz"):
        n_bytes *= 6.5  # rough estimate for gzipped to give a progress
    approx_number_of_chunks = int(n_bytes / args.chunk_size)
    logging.info("N bytes of reads: %d" % n_bytes)
    logging.info("Approx number of chunks of %d bytes: %d" % (args.chunk_size, approx_number_of_chunks))

    if not getattr(args, "gpu", False):
        # the reference's CPU route refuses reverse complements (command_line_interface.py:107)
        assert not getattr(args, "map_reverse_complements", False), \
            "Mapping reverse complements only supported with GPU-mode for now"

    rank, world_size, local_rank = distributed.init_process_group()
    if world_size > 1:
        import torch
        if torch.cuda.is_available():
            torch.cuda.set_device(local_rank)

    file = open_reads(args.reads)
    chunks = file.read_chunks(min_chunk_size=args.chunk_size)
    t_before_map = time.perf_counter()
    if world_size == 1:
        node_counts = map_gpu(kmer_index, chunks, k, getattr(args, "gpu_hash_map_size", 0),
                              getattr(args, "map_reverse_complements", False))
    else:
        node_counts = _map_sharded(kmer_index, chunks, k, getattr(args, "map_reverse_complements", False),
                                   rank, world_size)
    file.close()
    logging.info("Time spent only on hashing and counting hashes: %.4f" % (time.perf_counter() - t_before_map))

    args_dict = vars(args)
    args_dict.pop("func", None)

    if args.output_file is None:
        return node_counts

    if rank == 0:
        np.save(args.output_file, node_counts)
        logging.info("Saved node counts to %s.npy" % args.output_file)
    logging.info("Spent %.3f sec in total mapping kmers using %d threads" % (time.perf_counter() - start_time,
                                                                             args.n_threads))


def _map_sharded(kmer_index, chunks, k, map_reverse_complements, rank, world_size):
    """One rank of a torchrun job: private counts in a torch tensor, one all-reduce at the end."""
    import torch
    di = DeviceIndex.from_index(kmer_index)
    n_counts = di.max_node_id() + 1
    counts = torch.zeros(n_counts, dtype=torch.int32, device="cuda")
    mapper = Mapper(di, n_counts, DEFAULT_MAX_FREQUENCY, counts_tensor=counts)
    for i, chunk in enumerate(chunks):
        if not distributed.chunk_belongs_to_rank(i, rank, world_size):
            continue
        seq = chunk.sequence
        mapper.map_reads(seq.bases, seq.offsets, k, revcomp=bool(map_reverse_complements), n_to_a=True)
    mapper.sync()
    distributed.all_reduce_counts(counts)
    torch.cuda.synchronize()
    out = counts.cpu().numpy().view(np.uint32)
    mapper.close()
    return out


def run_argument_parser(args):
    parser = argparse.ArgumentParser(
        description='Kmer Mapper',
        prog='kmer_mapper',
        formatter_class=lambda prog: argparse.HelpFormatter(prog, max_help_position=50, width=100))

    subparsers = parser.add_subparsers()
    subparser = subparsers.add_parser("map", help="Map reads to a kmer index")
    subparser.add_argument("-i", "--kmer-index", required=False)
    subparser.add_argument("-b", "--index-bundle", required=False)
    subparser.add_argument("-f", "--reads", required=True, help="Reads in .fa, .fq, .fa.gz, or fq.gz format")
    subparser.add_argument("-k", "--kmer-size", required=False, default=31, type=int)
    subparser.add_argument("-t", "--n-threads", required=False, default=16, type=int)
    subparser.add_argument("-c", "--chunk-size", required=False, type=int, default=2500000,
                           help="N bytes to process in each chunk")
    subparser.add_argument("-o", "--output-file", required=True)
    subparser.add_argument("-d", "--debug", required=False, help="Set to True to print debug log")
    subparser.add_argument("-I", "--max-hits-per-kmer", required=False, default=1000, type=int,
                           help="Ignore kmers that have more than this amount of hits in index")
    subparser.add_argument("-g", "--gpu", default=False, type=bool,
                           help="Set to True to use GPU-counting. Experimental."
                           " Requires suitable hardware and dependencies.")
    subparser.add_argument("-s", "--gpu-hash-map-size", default=0, type=int,
                           help="Can be overriden to set GPU hash map size. "
                           "Set to a lower number to decrease GPU memory requirements. Higher number makes things faster")
    subparser.add_argument("-r", "--map-reverse-complements", default=False, type=bool,
                           help="Also count kmers in reverse complement of reads. "
                                "Default False. Not necessary if index contains reverse complements.")
    subparser.set_defaults(func=map_bnp)

    if len(args) == 0:
        parser.print_help()
        sys.exit(1)

    args = parser.parse_args(args)
    return args.func(args)
