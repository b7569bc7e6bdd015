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
import mmap
import os
import sys
import time

import numpy as np

logging.basicConfig(stream=sys.stdout, level=logging.INFO, format='%(asctime)s %(levelname)s: %(message)s')

from . import distributed  # noqa: E402
from .counter_index import CounterKmerIndex  # noqa: E402
from .device import DEFAULT_MAX_FREQUENCY, DeviceIndex, Mapper, borrowed_mapper  # noqa: E402
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
    if isinstance(kmer_index, CounterKmerIndex):
        # command_line_interface.py:46-49: count per unique k-mer; the chunk's result is the counter's value array.
        # The counter is zeroed first, so that the result is THIS chunk's counts and the caller's additive reduce
        # (:124-130) gives the totals.
        counter = kmer_index.counter
        counter._values = np.zeros(counter.n_keys, dtype=np.uint32)
        counter.count_reads(seq.bases, seq.offsets, kmer_size)
        mapped = counter._values
        logging.debug("Mapped with counter. Got values of length %d" % len(mapped))
        return mapped
    di = DeviceIndex.from_index(kmer_index)
    with borrowed_mapper(di, kmer_index.max_node_id() + 1, DEFAULT_MAX_FREQUENCY) as m:
        m.map_reads(seq.bases, seq.offsets, kmer_size, revcomp=False, n_to_a=True)
        mapped = m.counts()
    logging.debug("Chunk of %d reads took %.2f sec" % (len(seq), time.perf_counter() - t))
    return mapped


HONOUR_MAX_HITS_ENV = "KMER_MAPPER_B200_HONOUR_MAX_HITS"


def _frequency_cutoff(args):
    """The reference parses ``-I/--max-hits-per-kmer`` and never uses it: every route calls the lookup with its
    default cut-off of 1000 (command_line_interface.py:51,173; mapper.pyx:19).  Same here -- unless the
    environment variable KMER_MAPPER_B200_HONOUR_MAX_HITS=1 opts in to what the flag's help text promises."""
    if os.environ.get(HONOUR_MAX_HITS_ENV, "") == "1" and _flag(args, "max_hits_per_kmer") is not None:
        return int(args.max_hits_per_kmer)
    return DEFAULT_MAX_FREQUENCY


def map_gpu(index, chunks, k, hash_map_size=0, map_reverse_complements=False, rank=0, world_size=1,
            return_mapper=False, max_index_lookup_frequency=DEFAULT_MAX_FREQUENCY):
    """command_line_interface.py:59-79: map an iterable of chunks (objects with ``.sequence``) against the
    index on the current GPU; returns the node counts.  ``hash_map_size`` is accepted for signature
    compatibility (the device index keeps the reference's own modulo-bucketed layout, so there is no
    separate hash-map capacity to choose)."""
    logging.info("Making counter")
    di = DeviceIndex.from_index(index)
    mapper = Mapper(di, di.max_node_id() + 1, max_index_lookup_frequency)
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


def _flag(args, name, default=None):
    return getattr(args, name, default)


def map_bnp(args):
    """The run itself (command_line_interface.py:82-151): load the index, stream the reads file in chunks
    through the GPU, write ``<output>.npy`` -- or return the counts when ``args.output_file`` is None, which
    is how KAGE calls it with a hand-built Namespace."""
    if _flag(args, "debug"):
        logging.info("Will print debug log")
        logging.getLogger().setLevel(logging.DEBUG)
    t_start = time.perf_counter()
    kmer_size = args.kmer_size
    want_revcomp = bool(_flag(args, "map_reverse_complements", False))
    if not _flag(args, "gpu", False):
        # the reference's CPU route refuses reverse complements (command_line_interface.py:107)
        assert not want_revcomp, "Mapping reverse complements only supported with GPU-mode for now"
    index = _get_kmer_index_from_args(args)

    file_bytes = os.stat(args.reads).st_size * (6.5 if args.reads.endswith(".gz") else 1)  # same rough gz factor
    logging.info("N bytes of reads: %d" % file_bytes)
    logging.info("Approx number of chunks of %d bytes: %d" % (args.chunk_size, int(file_bytes / args.chunk_size)))

    rank, world_size, local_rank = distributed.init_process_group()
    if world_size > 1:
        import torch
        if torch.cuda.is_available():
            torch.cuda.set_device(local_rank)

    reads = open_reads(args.reads, n_threads=distributed.host_threads_per_rank() if world_size > 1 else None)
    t_map = time.perf_counter()
    try:
        if not isinstance(index, CounterKmerIndex) and os.environ.get(PARSE_ENV, "device") != "host":
            # default route: the raw text goes to the GPU and the records are parsed there (kmb_mapper_map_text)
            node_counts = _map_text(index, reads, kmer_size, want_revcomp, rank, world_size, _frequency_cutoff(args),
                                    args.chunk_size if args.chunk_size != REFERENCE_CHUNK_SIZE else DEVICE_TEXT_CHUNK)
            chunk_iter = None
        else:
            chunk_iter = reads.read_chunks(min_chunk_size=args.chunk_size, rank=rank, world_size=world_size)
        if chunk_iter is None:
            pass
        elif isinstance(index, CounterKmerIndex):
            node_counts = _map_counter_index(index, chunk_iter, kmer_size, want_revcomp, rank, world_size)
        elif world_size == 1:
            node_counts = map_gpu(index, chunk_iter, kmer_size, _flag(args, "gpu_hash_map_size", 0), want_revcomp,
                                  max_index_lookup_frequency=_frequency_cutoff(args))
        else:
            node_counts = _map_sharded(index, chunk_iter, kmer_size, want_revcomp, rank, world_size,
                                       _frequency_cutoff(args))
    finally:
        reads.close()
    logging.info("Time spent only on hashing and counting hashes: %.4f" % (time.perf_counter() - t_map))

    vars(args).pop("func", None)  # the reference strips it before handing the dict to its workers (cli:121-122)
    if args.output_file is None:
        return node_counts
    if rank == 0:
        np.save(args.output_file, node_counts)  # numpy appends ".npy" (command_line_interface.py:149)
        logging.info("Saved node counts to %s.npy" % args.output_file)
    logging.info("Spent %.3f sec in total mapping kmers using %d threads" % (time.perf_counter() - t_start,
                                                                             _flag(args, "n_threads", 1)))


GZ_ENV = "KMER_MAPPER_B200_GZ"         # "device" (default): multi-member .gz inflated by GPU warps; "host": by the host decoders
PARSE_ENV = "KMER_MAPPER_B200_PARSE"   # "device" (default): records are parsed by GPU kernels; "host": by the native host parser
REFERENCE_CHUNK_SIZE = 2500000          # the reference's -c default (command_line_interface.py:169), a CPU-worker setting
DEVICE_TEXT_CHUNK = 64 << 20            # what the device route reads per chunk when -c is left at that default


def map_file_text(mapper, reads, k, map_reverse_complements=False, rank=0, world_size=1, chunk_bytes=DEVICE_TEXT_CHUNK):
    """This rank's share of an open reads file (reader.ReadFile) through ``mapper``, records parsed on the device.  Plain
    files: whole-record windows of text, pread into pinned staging (``kmb_mapper_map_text_fd``).  Multi-member .gz
    (bgzip, concatenated members): the COMPRESSED bytes go to the GPU and are inflated there, one member per warp
    (``kmb_mapper_map_gz``); what the device cannot take (a plain single-member .gz) is inflated by the host decoders
    from where it stopped.  Returns (text chunks mapped, offset the host decoders took over at or None)."""
    n_chunks, gz_start, done = 0, 0, False
    if reads.path.lower().endswith(".gz") and os.environ.get(GZ_ENV, "device") != "host" and os.path.getsize(reads.path):
        with open(reads.path, "rb") as f, mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            whole = np.frombuffer(mm, dtype=np.uint8)
            try:
                gz_start = mapper.map_gz(whole, reads.format, k, revcomp=bool(map_reverse_complements), n_to_a=True,
                                         shard_index=rank, shard_count=world_size)
                done = gz_start >= whole.shape[0]
            finally:
                del whole
    if not done:
        for chunk in reads.text_chunks(min_chunk_size=chunk_bytes, rank=rank, world_size=world_size, gz_start=gz_start):
            mapper.map_text(chunk, reads.format, k, revcomp=bool(map_reverse_complements), n_to_a=True)
            n_chunks += 1
    return n_chunks, (None if done else gz_start)


def _map_text(kmer_index, reads, k, map_reverse_complements, rank, world_size, max_index_lookup_frequency, chunk_bytes):
    """File -> pinned staging -> GPU: (inflate,) newline scan, record parsing, 2-bit encoding, k-mers, index probe and
    counts all on the device; the host only cuts the file into whole-record windows or gzip members.  Under torchrun
    every rank takes its share of the file and the count arrays are summed by one all-reduce."""
    di = DeviceIndex.from_index(kmer_index)
    mapper = Mapper(di, di.max_node_id() + 1, max_index_lookup_frequency)
    comm = distributed.Comm(device=di.device) if world_size > 1 else None
    n_chunks, host_from = map_file_text(mapper, reads, k, map_reverse_complements, rank, world_size, chunk_bytes)
    mapper.sync()
    logging.debug("Device-side parsing: %d chunk(s) of text%s" % (n_chunks, "" if not reads.path.lower().endswith(".gz") else
                  "; gzip members inflated on the device" + ("" if host_from is None else " up to byte %d, by the host from there" % host_from)))
    if comm is not None:
        comm.all_reduce(mapper)
    out = mapper.counts()
    mapper.close()
    if comm is not None:
        comm.close()
    return out


def _map_counter_index(kmer_index, chunks, k, map_reverse_complements, rank, world_size):
    """The CounterKmerIndex route (command_line_interface.py:118-119, 133-138): k-mer counts per unique key summed over
    the chunks (and, under torchrun, over the GPUs), then ``counter._values = totals`` and ``get_node_counts()`` once."""
    counter = kmer_index.counter
    counter._values = np.zeros(counter.n_keys, dtype=np.uint32)      # initial_data = zeros_like(counter._values)
    for chunk in chunks:
        seq = as_ragged(chunk.sequence if hasattr(chunk, "sequence") else chunk)
        counter.count_reads(seq.bases, seq.offsets, k, count_revcomps=bool(map_reverse_complements))
    counter._mapper.sync()
    if world_size > 1:
        comm = distributed.Comm(device=counter._index.device)
        comm.all_reduce(counter._mapper)
        counter._mapper.sync()
        comm.close()
    t = time.perf_counter()
    node_counts = kmer_index.get_node_counts()
    logging.info("Time spent getting node counts in the end: %.3f" % (time.perf_counter() - t))
    return node_counts


def _map_sharded(kmer_index, chunks, k, map_reverse_complements, rank, world_size,
                 max_index_lookup_frequency=DEFAULT_MAX_FREQUENCY):
    """One rank of a torchrun job: a private count array per GPU, summed once at the end by one all-reduce
    (``kmb_mapper_allreduce``: NCCL through the C ABI) -- the additive reduce of command_line_interface.py:124-130."""
    di = DeviceIndex.from_index(kmer_index)
    n_counts = di.max_node_id() + 1
    mapper = Mapper(di, n_counts, max_index_lookup_frequency)
    comm = distributed.Comm(device=di.device)
    for chunk in chunks:    # the reader already hands this rank its share only (reader.py: read_chunks)
        seq = chunk.sequence
        mapper.map_reads(seq.bases, seq.offsets, k, revcomp=bool(map_reverse_complements), n_to_a=True)
    mapper.sync()           # an invalid base on any rank raises here, before the collective
    comm.all_reduce(mapper)
    out = mapper.counts()
    mapper.close()
    comm.close()
    return out


# (short flag, long flag, argparse keywords): names, defaults and types exactly as the reference declares them
# (command_line_interface.py:164-182) -- including type=bool for -g and -r, for which ANY non-empty string,
# "False" included, is true, and the untyped -d.  -t and -I are accepted and unused here.
_MAP_FLAGS = (
    ("-i", "--kmer-index", dict(required=False, help="KmerIndex .npz file")),
    ("-b", "--index-bundle", dict(required=False, help="index bundle (not supported by this implementation)")),
    ("-f", "--reads", dict(required=True, help="reads: .fa / .fq, optionally .gz")),
    ("-k", "--kmer-size", dict(required=False, default=31, type=int)),
    ("-t", "--n-threads", dict(required=False, default=16, type=int, help="accepted for compatibility; the GPU path has no worker processes")),
    ("-c", "--chunk-size", dict(required=False, default=2500000, type=int, help="bytes of the reads file per chunk")),
    ("-o", "--output-file", dict(required=True, help="node counts are written to <output-file>.npy")),
    ("-d", "--debug", dict(required=False, help="any value switches the DEBUG log on")),
    ("-I", "--max-hits-per-kmer", dict(required=False, default=1000, type=int,
                                       help="parsed and ignored, like the reference (the cut-off is always 1000), unless "
                                            "KMER_MAPPER_B200_HONOUR_MAX_HITS=1 is set in the environment")),
    ("-g", "--gpu", dict(default=False, type=bool, help="both routes run on the GPU here; -g additionally allows -r")),
    ("-s", "--gpu-hash-map-size", dict(default=0, type=int, help="accepted for compatibility")),
    ("-r", "--map-reverse-complements", dict(default=False, type=bool,
                                             help="also count the reverse complement of every read k-mer")),
)


def run_argument_parser(args):
    parser = argparse.ArgumentParser(prog="kmer_mapper", description="Kmer Mapper",
                                     formatter_class=lambda prog: argparse.HelpFormatter(prog, max_help_position=50, width=100))
    commands = parser.add_subparsers()
    map_cmd = commands.add_parser("map", help="Map reads to a kmer index")
    for short, long_name, kwargs in _MAP_FLAGS:
        map_cmd.add_argument(short, long_name, **kwargs)
    map_cmd.set_defaults(func=map_bnp)
    if len(args) == 0:  # command_line_interface.py:185-187
        parser.print_help()
        sys.exit(1)
    parsed = parser.parse_args(args)
    return parsed.func(parsed)
