"""Read sharding across GPUs and the final count reduction.

The reference's only parallelism is data parallelism over read chunks with an additive reduction
of the per-worker ``uint32`` count arrays (command_line_interface.py:124-130).  Here: one process
per GPU (torchrun), every rank holds a full replica of the device index and a private count array,
rank r maps the chunks ``i % world_size == r`` with no communication, and the arrays are summed once
at the end by a single all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests).  uint32 addition
wraps mod 2**32, so any reduction order gives identical bits; the all-reduce is issued on the int32
view of the same memory (two's-complement addition is the same operation).
"""
from __future__ import annotations

import os

import numpy as np


def world():
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when not launched by it."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def host_threads_per_rank() -> int:
    """CPU threads each rank of this node may use for host-side work (the 2-bit encoder of host input, the read
    parser): the CPUs of the affinity mask shared fairly between the ranks torchrun started on this node."""
    local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
    return max(1, len(os.sched_getaffinity(0)) // local_world)


def init_process_group(backend=None):
    """Join the job torchrun started (MASTER_ADDR/MASTER_PORT from the environment).  Returns
    (rank, world_size, local_rank); a no-op for a single process."""
    rank, world_size, local_rank = world()
    if world_size == 1:
        return rank, world_size, local_rank
    if backend != "gloo":
        try:  # the ranks of a node share its cores: without this every rank would start one encoder thread per core
            from . import _lib
            _lib.set_option("host_threads", host_threads_per_rank())
            _lib.set_option("host_ranks", max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world_size))))
        except Exception:
            pass
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend)
    return rank, world_size, local_rank


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous shard [lo, hi) of n_items for this rank (sizes differ by at most one)."""
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_reads(offsets: np.ndarray, rank: int, world_size: int):
    """Contiguous range of whole reads for this rank: returns (r_lo, r_hi, base_lo, base_hi)."""
    n_reads = len(offsets) - 1
    lo, hi = shard_range(n_reads, rank, world_size)
    return lo, hi, int(offsets[lo]), int(offsets[hi])


def chunk_belongs_to_rank(chunk_index: int, rank: int, world_size: int) -> bool:
    """Round-robin chunk assignment used by the CLI (every rank scans the file, maps its own chunks)."""
    return chunk_index % world_size == rank


class Comm:
    """The job's NCCL communicator behind the C ABI (``kmb_comm_*``, include/kmer_mapper_b200.h): what sums the
    per-GPU count arrays in place of the reference's additive map-reduce (command_line_interface.py:124-130).
    The 128-byte NCCL id is made by rank 0 and handed to the other ranks through the torch.distributed
    rendezvous that torchrun already set up (control plane only: the reduction itself is one ncclAllReduce on
    the mapper's stream, issued by the library)."""

    def __init__(self, device=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        rank, world_size, local_rank = world()
        self.rank, self.world_size = rank, world_size
        self.device = local_rank if device is None else int(device)
        ident = (C.c_uint8 * _lib.COMM_ID_BYTES)()
        if rank == 0:
            _lib.check(_lib.lib().kmb_comm_unique_id(ident))
        if world_size > 1:
            if not dist.is_initialized():
                raise RuntimeError("Comm needs the torch.distributed process group (distributed.init_process_group)")
            box = [bytes(ident)]
            dist.broadcast_object_list(box, src=0)
            ident = (C.c_uint8 * _lib.COMM_ID_BYTES).from_buffer_copy(box[0])
        h = C.c_void_p()
        _lib.check(_lib.lib().kmb_comm_init_rank(self.device, world_size, rank, ident, C.byref(h)))
        self._h = h
        import weakref
        self._finalizer = weakref.finalize(self, _lib.lib().kmb_comm_destroy, h)

    def all_reduce(self, mapper):
        """Queue hit log -> counts and the in-place sum over the ranks on the mapper's stream (no host wait)."""
        from . import _lib
        _lib.check(_lib.lib().kmb_mapper_allreduce(mapper._h, self._h))

    def close(self):
        self._finalizer()


def all_reduce_counts(counts):
    """In-place sum over ranks of a uint32 count array: a torch tensor (CUDA -> NCCL, CPU -> gloo) or
    a numpy array (gloo).  Returns the reduced array (same object for tensors)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return counts
    if isinstance(counts, np.ndarray):
        assert counts.dtype == np.uint32
        t = torch.from_numpy(counts.view(np.int32))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return counts
    assert counts.element_size() == 4 and not counts.dtype.is_floating_point
    dist.all_reduce(counts.view(torch.int32), op=dist.ReduceOp.SUM)
    return counts
