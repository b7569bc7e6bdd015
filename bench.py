#!/usr/bin/env python
"""bench.py -- k-mers mapped per second on synthetic inputs of the BASELINE.json shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2] [--impl ours|reference]

One "step" = one pass of the hot path (ASCII reads -> 2-bit -> rolling k-mers -> index probe -> per-node
uint32 counts) over one batch of synthetic reads.  At N=1 the default workload is BASELINE.json's
configs[1]: k=31, 100 M-entry index (modulo 452 930 477, 80 M nodes), 50 M x 150 bp reads = 6.0 G k-mers
per step.  For N>1 (torchrun, one rank per GPU) every rank holds a replica of the index and maps its own
50 M reads (weak scaling); the private count arrays are summed by one NCCL all-reduce inside every step.

Printed JSON (rank 0, one line):
  value          whole-job G k-mers/s with the reads already resident in HBM (CUDA events, max over ranks)
  e2e            the same through the public API with HOST buffers: pinned host -> H2D -> kernels -> counts
                 D2H, every step (host wall clock between device synchronisations, max over ranks)
  roofline       fused kernel only: algorithmic bytes (L/(L-k+1) + 32 + 8h per k-mer) / its
                 CUDA-event duration, against the measured HBM copy peak of MEASURED_PEAKS.json
  cpu_baseline   the reference's CPU path (compiled mapper.pyx from oracle/_ref when present, else the C
                 port; hashing restated in C) on all host cores over a bounded sample of the same workload
  --impl reference   times only that CPU path (rank 0 alone under torchrun).

Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "kmers_mapped_per_sec"
UNIT = "GK/s"

WORKLOADS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    "config1": dict(k=31, genome=10_000_000, entries=1_000_000, nodes=800_000, modulo=2_000_003,
                    reads=100_000, read_len=150, seed=101),
    # configs[1]: human-scale short-read genotyping shape on 1 B200 -- the configuration the metric is quoted on
    "config2": dict(k=31, genome=1_000_000_000, entries=100_000_000, nodes=80_000_000, modulo=452_930_477,
                    reads=50_000_000, read_len=150, seed=201),
    # configs[2]: multi-GB table; 200 M reads sharded over the GPUs of the job (25 M per GPU at 8)
    # (5 Gbp genome, 200 M reads in total; strong scaling)
    "config3": dict(k=31, genome=5_000_000_000, entries=500_000_000, nodes=400_000_000, modulo=1_000_000_007,
                    reads=200_000_000, reads_total=200_000_000, read_len=150, seed=301),
    # configs[3]: small k, higher hit rate, Zipf nodes (atomic contention)
    "config4_k21": dict(k=21, genome=1_000_000_000, entries=100_000_000, nodes=80_000_000, modulo=452_930_477,
                        reads=50_000_000, read_len=150, seed=401, zipf=True),
    "config4_k15": dict(k=15, genome=1_000_000_000, entries=100_000_000, nodes=80_000_000, modulo=452_930_477,
                        reads=50_000_000, read_len=150, seed=402, zipf=True),
    # configs[4]: long reads, 5 % N, mixed case
    "config5": dict(k=31, genome=1_000_000_000, entries=100_000_000, nodes=80_000_000, modulo=452_930_477,
                    reads=750_000, read_len=10_000, seed=502, n_rate=0.05, lower_rate=0.5),
}


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    p.add_argument("--scale", type=float, default=1.0, help="shrink genome/index/reads (smoke runs only)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-files", action="store_true", help="skip the reads/s leg (FASTQ / FASTQ.gz files -> counts)")
    p.add_argument("--file-reads", type=int, default=8_000_000, help="reads in the files of the reads/s leg")
    p.add_argument("--host-pack", type=int, default=-1, help="e2e leg: library option host_pack (-1 default = hybrid for pinned input, "
                   "1 every chunk 2-bit packed by the CPU, 0 every chunk ASCII)")
    p.add_argument("--no-oracle", action="store_true", help="skip the oracle parity checks (tuning runs only)")
    p.add_argument("--full-oracle", action="store_true", help="N=1: also run EVERY read of the step through the oracle")
    p.add_argument("--cpu-sample-reads", type=int, default=0, help="reads in the CPU baseline sample (0 = auto)")
    p.add_argument("--opt", action="append", default=[], help="library option name=value (tuning)")
    # internal: the CPU baseline runs in a CUDA-free child process
    p.add_argument("--cpu-worker", default=None, help=argparse.SUPPRESS)
    p.add_argument("--cpu-threads", type=int, default=0, help=argparse.SUPPRESS)
    return p.parse_args(argv)


def workload(name, scale):
    w = dict(WORKLOADS[name])
    if scale != 1.0:
        for key in ("genome", "entries", "nodes", "reads", "reads_total"):
            if key in w:
                w[key] = max(int(w[key] * scale), 1000)
        w["modulo"] = max(int(w["modulo"] * scale) | 1, 1009)
    return w


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def known_traffic(name, kernel):
    """DRAM bytes per launch of the fused kernel from the committed ncu --set full capture of THAT kernel on this
    workload, if any, and where the figure comes from (it is evidence of that capture, not of this run: ncu cannot
    run inside a timed bench)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        rec = json.load(open(path)).get(name)
    except Exception:
        return None, None
    if isinstance(rec, dict) and "bytes" not in rec:
        rec = rec.get(kernel)
    if isinstance(rec, dict):
        return rec.get("bytes"), "static ncu capture: " + str(rec.get("source"))
    return None, None


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.lines = []
        self.gpu_index = gpu_index

    def start(self):
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        try:
            self.proc = subprocess.Popen([exe, "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU path (the reference arm / cpu_baseline).  Runs in a child process that never touches CUDA.
# ---------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(data_dir, k):
    from oracle import c_oracle, ref_loader
    from oracle.oracle import OracleIndex
    # copy-on-write mapping: shared page cache, but "writable" as the reference's memoryview casts demand
    ld = lambda n: np.load(os.path.join(data_dir, n + ".npy"), mmap_mode="c")  # noqa: E731
    _W["index"] = OracleIndex.__new__(OracleIndex)
    idx = _W["index"]
    idx._hashes_to_index, idx._n_kmers, idx._nodes = ld("hashes_to_index"), ld("n_kmers"), ld("nodes")
    idx._kmers, idx._frequencies = ld("kmers"), ld("frequencies")
    idx._modulo = int(np.load(os.path.join(data_dir, "modulo.npy")))
    _W["max_node"] = int(np.load(os.path.join(data_dir, "max_node.npy")))
    _W["bases"], _W["offsets"] = ld("bases"), ld("offsets")
    _W["k"] = k
    _W["ref"] = ref_loader.load_reference_mapper()
    _W["c"] = c_oracle
    _W["acc"] = None


def _cpu_chunk(rng):
    """One chunk of reads, the body of map_cpu (command_line_interface.py:32-56): N->A + hash (restated in C),
    then the lookup -- the reference's own compiled Cython function when oracle/_ref exists."""
    r0, r1 = rng
    off = np.asarray(_W["offsets"][r0:r1 + 1])
    bases = np.asarray(_W["bases"][off[0]:off[-1]])
    hashes = _W["c"].kmer_hashes(bases, off - off[0], _W["k"], n_to_a=True)
    if _W["ref"] is not None:
        res = _W["ref"].map_kmers_to_graph_index(_W["index"], _W["max_node"], hashes)   # fresh array per call (mapper.pyx:37)
    else:
        res = _W["c"].map_kmers_to_graph_index(_W["index"], _W["max_node"], hashes)
    if _W["acc"] is None:
        _W["acc"] = res
    else:
        _W["acc"] += res            # the additive reduce (command_line_interface.py:124-130), worker-local
    return hashes.shape[0]


def cpu_worker_main(args):
    """Child process: time `steps` passes over the sample after `warmup` passes; print one JSON line."""
    import multiprocessing as mp
    data_dir = args.cpu_worker
    meta = json.load(open(os.path.join(data_dir, "meta.json")))
    k, n_reads, read_len = meta["k"], meta["n_reads"], meta["read_len"]
    threads = args.cpu_threads or os.cpu_count() or 1
    n_counts = meta["n_counts"]
    try:
        ram = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
        threads = max(1, min(threads, int(0.4 * ram / (2 * 4 * n_counts + 1))))
    except Exception:
        pass
    reads_per_chunk = max(1, 2_500_000 // max(read_len, 1))   # --chunk-size default 2 500 000 bytes (cli:169)
    chunks = [(s, min(s + reads_per_chunk, n_reads)) for s in range(0, n_reads, reads_per_chunk)]
    ctx = mp.get_context("fork")
    with ctx.Pool(threads, initializer=_cpu_init, initargs=(data_dir, k)) as pool:
        times, n_kmers = [], 0
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            n_kmers = sum(pool.imap_unordered(_cpu_chunk, chunks, chunksize=1))
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    from oracle import ref_loader
    kind = "reference" if ref_loader.load_reference_mapper() is not None else "port"
    print(json.dumps({"n_kmers": n_kmers, "seconds": times, "threads": threads, "kind": kind,
                      "chunks": len(chunks)}))


def run_cpu_child(data_dir, steps, warmup, threads=0, timeout=1500):
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", data_dir, "--steps", str(steps), "--warmup", str(warmup),
           "--cpu-threads", str(threads)]
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""
    for v in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(v, None)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    if out.returncode != 0:
        raise RuntimeError("cpu worker failed: " + out.stderr[-2000:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def write_cpu_sample(data_dir, host_index, max_node, bases, offsets, n_reads, w):
    os.makedirs(data_dir, exist_ok=True)
    for name in ("hashes_to_index", "n_kmers", "nodes", "kmers", "frequencies"):
        np.save(os.path.join(data_dir, name + ".npy"), np.ascontiguousarray(getattr(host_index, "_" + name)))
    np.save(os.path.join(data_dir, "modulo.npy"), np.int64(host_index._modulo))
    np.save(os.path.join(data_dir, "max_node.npy"), np.int64(max_node))
    np.save(os.path.join(data_dir, "offsets.npy"), np.ascontiguousarray(offsets[:n_reads + 1]))
    np.save(os.path.join(data_dir, "bases.npy"), np.ascontiguousarray(bases[:int(offsets[n_reads])]))
    json.dump(dict(k=w["k"], n_reads=int(n_reads), read_len=w["read_len"], n_counts=int(max_node) + 1),
              open(os.path.join(data_dir, "meta.json"), "w"))


def cpu_baseline_record(res, sample_desc):
    sec = float(np.median(res["seconds"]))
    return {"value": res["n_kmers"] / sec / 1e9, "unit": UNIT, "cores": res["threads"], "kind": res["kind"],
            "sample": sample_desc + "; %d chunks of 2.5 MB per pass, median of %d passes (%.2f s each); lookup = %s, "
            "N->A + hashing restated in C (bionumpy absent), worker-local additive reduce"
            % (res["chunks"], len(res["seconds"]), sec,
               "reference mapper.pyx compiled unmodified (oracle/_ref)" if res["kind"] == "reference" else "C port (oracle/kmer_oracle.c)")}


# ---------------------------------------------------------------------------------------------
# data
# ---------------------------------------------------------------------------------------------
def generate(w, rank, device):
    """Index tensors (same on every rank: a replica) and this rank's reads, on `device`."""
    from kmer_mapper_b200 import synthetic as S
    genome = S.t_make_genome(w["genome"], w["seed"], device=device)
    idx = S.t_make_index(genome, w["entries"], w["k"], w["nodes"], w["modulo"], w["seed"] + 1000,
                         zipf_nodes=bool(w.get("zipf")))
    bases, offsets = S.t_make_reads(genome, w["reads"], w["read_len"], w["seed"] + 1 + 7919 * rank,
                                    n_rate=w.get("n_rate", 0.0), lower_rate=w.get("lower_rate", 0.0),
                                    slice_reads=max(1, (1 << 27) // w["read_len"]))
    del genome
    return S.TensorIndex(idx), bases, offsets



# ---------------------------------------------------------------------------------------------
# reads/s: files in, counts on the host out (the metric's third leg)
# ---------------------------------------------------------------------------------------------
def _gz_member(args):
    import zlib
    path, lo, hi = args
    with open(path, "rb") as f:
        f.seek(lo)
        blob = f.read(hi - lo)
    c = zlib.compressobj(1, zlib.DEFLATED, 31)
    return c.compress(blob) + c.flush()


def write_fastq_files(dirname, bases, n_reads, read_len):
    """<dir>/reads.fq (4-line records, constant quality) from the first n_reads device-resident reads -- assembled on the
    GPU as one [n, record_len] byte matrix -- and the same text as two multi-member .gz files (BASELINE configs[1] names
    FASTQ.gz): <dir>/reads.fq.gz and <dir>/reads_bgzf.fq.gz."""
    import multiprocessing as mp
    import torch
    dev = bases.device
    digits = 9
    rec = 1 + digits + 1 + read_len + 3 + read_len + 1
    m = torch.empty((n_reads, rec), dtype=torch.uint8, device=dev)
    m[:, 0] = ord("@")
    ids = torch.arange(n_reads, device=dev)
    for d in range(digits):
        m[:, 1 + d] = ((ids // (10 ** (digits - 1 - d))) % 10 + ord("0")).to(torch.uint8)
    m[:, 1 + digits] = 10
    s0 = 2 + digits
    m[:, s0:s0 + read_len] = bases[:n_reads * read_len].view(n_reads, read_len)
    m[:, s0 + read_len] = 10
    m[:, s0 + read_len + 1] = ord("+")
    m[:, s0 + read_len + 2] = 10
    m[:, s0 + read_len + 3:s0 + 2 * read_len + 3] = ord("I")
    m[:, rec - 1] = 10
    fq = os.path.join(dirname, "reads.fq")
    m.cpu().numpy().tofile(fq)
    del m
    size = os.path.getsize(fq)
    out = [fq]
    # members of ~4 MB of text cut at records (`cat a.gz b.gz ...`), and BGZF-like members of 65280 bytes of text cut
    # anywhere (what bgzip writes: the usual multi-member FASTQ.gz)
    for name, per in (("reads.fq.gz", max(rec, (4 << 20) // rec * rec)), ("reads_bgzf.fq.gz", 65280)):
        cuts = [(fq, lo, min(lo + per, size)) for lo in range(0, size, per)]
        with mp.get_context("fork").Pool(min(os.cpu_count() or 1, 32)) as pool, open(os.path.join(dirname, name), "wb") as f:
            for blob in pool.imap(_gz_member, cuts, chunksize=max(4, (1 << 20) // per)):
                f.write(blob)
        out.append(os.path.join(dirname, name))
    return tuple(out)


def time_file_to_counts(path, di, n_counts, k):
    """Wall clock from opening the reads file to the count array on the host, through the library's own route: text
    windows -> pinned staging -> H2D -> device-side record parsing -> fused kernel -> counts D2H.  Second of two passes."""
    from kmer_mapper_b200.device import Mapper
    from kmer_mapper_b200.reader import open_reads
    from kmer_mapper_b200.command_line_interface import map_file_text
    import torch
    from kmer_mapper_b200 import _lib
    out = torch.empty(n_counts, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)   # where the counts land
    m = Mapper(di, n_counts)           # the mapper (its staging buffers) outlives a file, like the index
    best = None
    g0 = _gz_device_members()
    for _ in range(2):
        m.reset()
        r0 = _lib.get_option("text_reads")
        t0 = time.perf_counter()
        reads = open_reads(path)
        map_file_text(m, reads, k)
        m.counts(out=out)
        dt = time.perf_counter() - t0
        n = _lib.get_option("text_reads") - r0
        reads.close()
        best = dt
    counts = out.copy()
    m.close()
    return n / best, n, best, counts, (_gz_device_members() - g0) // 2


def _gz_device_members():
    import ctypes as C
    from kmer_mapper_b200 import _lib
    a = C.c_uint64()
    _lib.check(_lib.lib().kmb_gz_device_stats(C.byref(a), None, None))
    return int(a.value)

# ---------------------------------------------------------------------------------------------
# main arms
# ---------------------------------------------------------------------------------------------
def main_reference(args):
    """The reference's CPU path on all host cores, same config/metric/unit; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    w = workload(args.workload, args.scale)
    device = "cuda" if torch.cuda.is_available() else "cpu"
    cores = os.cpu_count() or 1
    sample_reads = args.cpu_sample_reads or min(w["reads"], max(cores * 100_000 * 150 // w["read_len"], 1000))
    w = dict(w, reads=min(w["reads"], max(sample_reads, 2_000_000)))   # only the sample is needed here
    tindex, bases, offsets = generate(w, 0, device)
    host_index = tindex.to_host()
    max_node = host_index.max_node_id()
    hb = bases[:int(offsets[sample_reads].item())].cpu().numpy()
    ho = offsets[:sample_reads + 1].cpu().numpy()
    del tindex, bases, offsets
    if device == "cuda":
        torch.cuda.empty_cache()
    data_dir = tempfile.mkdtemp(prefix="kmb_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        write_cpu_sample(data_dir, host_index, max_node, hb, ho, sample_reads, w)
        del host_index
        res = run_cpu_child(data_dir, args.steps, args.warmup)
    finally:
        shutil.rmtree(data_dir, ignore_errors=True)
    sec = float(np.median(res["seconds"]))
    value = res["n_kmers"] / sec / 1e9
    desc = "first %d reads of %s (%d k-mers per step)" % (sample_reads, args.workload, res["n_kmers"])
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "impl": "reference",
            "config": dict(workload=args.workload, k=w["k"], index_entries=w["entries"], modulo=w["modulo"],
                           nodes=w["nodes"], reads_per_step=sample_reads, read_len=w["read_len"], scale=args.scale,
                           host_threads=res["threads"]),
            "cpu_baseline": cpu_baseline_record(res, desc),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def _oracle_index(host_index):
    """The host arrays as the object the oracle reads (mapper.pyx:22-29 attribute names)."""
    from oracle.oracle import OracleIndex
    oi = OracleIndex.__new__(OracleIndex)
    oi._hashes_to_index, oi._n_kmers, oi._nodes = host_index._hashes_to_index, host_index._n_kmers, host_index._nodes
    oi._kmers, oi._frequencies, oi._modulo = host_index._kmers, host_index._frequencies, int(host_index._modulo)
    return oi


def main_ours(args):
    import torch
    import torch.distributed as dist
    from kmer_mapper_b200 import _lib, distributed
    from kmer_mapper_b200.device import DeviceIndex, Mapper

    _lib.require_device()
    rank, world, local_rank = distributed.init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    for kv in args.opt:
        name, val = kv.split("=")
        _lib.set_option(name, int(val))
    _lib.set_option("time_kernels", 1)
    w = workload(args.workload, args.scale)
    k, L = w["k"], w["read_len"]
    strong = "reads_total" in w            # the job's reads are fixed and sharded over the GPUs (BASELINE configs[2])
    if strong:
        lo, hi = distributed.shard_range(w["reads_total"], rank, world)
        w["reads"] = hi - lo
    t_setup = time.perf_counter()
    tindex, bases, offsets = generate(w, rank, device)
    n_reads = w["reads"]
    n_bases = int(bases.shape[0])
    max_node = tindex.max_node_id()
    n_counts = max_node + 1

    # ---- host copy of the index for the oracle checks (rank 0) and the CPU baseline sample (N == 1 only),
    # taken before the raw index tensors are dropped
    cpu_dir = None
    do_cpu = (not args.no_cpu_baseline) and world == 1 and rank == 0
    do_oracle = rank == 0 and not args.no_oracle
    cores = os.cpu_count() or 1
    # reads per rank in the parity sample: 1.6 M in total (192 M k-mers at k = 31), at least 100 k per rank
    sample_reads = min(n_reads, args.cpu_sample_reads or max(cores * 100_000 * 150 // L // world, 100_000 * 150 // L, 1000))
    host_index = None
    if do_cpu or do_oracle:
        host_index = tindex.to_host()
    if do_cpu:
        cpu_dir = tempfile.mkdtemp(prefix="kmb_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        write_cpu_sample(cpu_dir, host_index, max_node, bases[:int(offsets[sample_reads].item())].cpu().numpy(),
                         offsets[:sample_reads + 1].cpu().numpy(), sample_reads, w)

    di = DeviceIndex.from_index(tindex, device=local_rank)
    del tindex
    torch.cuda.empty_cache()

    stream = torch.cuda.Stream(device=device)
    counts = torch.zeros(n_counts, dtype=torch.int32, device=device)
    mapper = Mapper(di, n_counts, counts_tensor=counts)
    mapper.set_stream(stream)
    comm = distributed.Comm(device=local_rank) if world > 1 else None   # NCCL behind the C ABI (kmb_comm_*)
    setup_s = time.perf_counter() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()

    def reduce_counts():
        """hit log -> node counts [-> sum over the GPUs], queued on the mapper's stream"""
        if comm is not None:
            comm.all_reduce(mapper)
        else:
            mapper.flush()

    def step_resident():
        mapper.reset()
        mapper.map_reads(bases, offsets, k)
        reduce_counts()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_resident()
        torch.cuda.synchronize()
        mapper.kernel_time()            # drop warm-up records
        mapper.apply_time()
        n_kmers_step, n_counted_step = mapper.stats()
        n_candidates_step = mapper.candidates()   # every step starts with a reset: these are per-step figures
        barrier()
        clocks = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                              os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank])
        if rank == 0:
            clocks.start()
        launches0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_resident()
        e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() - launches0
        kernel_ms, kernel_n = mapper.kernel_time()
        apply_ms, apply_n = mapper.apply_time()
    clock_rec = clocks.stop() if rank == 0 else None
    mapper.sync()                        # raises on an invalid base

    def over_ranks(value, op):
        t = torch.tensor([value], dtype=torch.float64 if isinstance(value, float) else torch.int64, device=device)
        if world > 1:
            dist.all_reduce(t, op=op)
        return t.item()

    ms_max = float(over_ranks(float(ms), dist.ReduceOp.MAX))
    total_kmers = int(over_ranks(int(n_kmers_step), dist.ReduceOp.SUM))
    total_counted = int(over_ranks(int(n_counted_step), dist.ReduceOp.SUM))
    value = total_kmers * args.steps / (ms_max / 1e3) / 1e9

    # ---- checks at full size (size-independent properties); every False fails the run (exit code 3)
    checks = {}
    full_counts = counts.clone()         # after the last timed step: the job's total on every rank
    s = int(full_counts.view(torch.int32).to(torch.int64).bitwise_and(0xFFFFFFFF).sum().item())
    checks["sum_reduced_counts_equals_entries_counted_by_all_ranks"] = (s == total_counted)
    checks["kmers_per_step"] = total_kmers
    checks["expected_kmers_per_step"] = int(over_ranks(int(n_reads * max(L - k + 1, 0)), dist.ReduceOp.SUM))
    checks["kmers_per_step_as_expected"] = checks["kmers_per_step"] == checks["expected_kmers_per_step"]

    # ---- parity of the multi-GPU path against the oracle: every rank maps the first `sample_reads` of its own
    # reads through the same device-resident path, the count arrays are summed by the same all-reduce, and rank
    # 0 compares with the oracle run over the concatenated samples (N == 1: the plain sample check)
    if not args.no_oracle:
        with torch.cuda.stream(stream):
            nb_s = int(offsets[sample_reads].item())
            mapper.reset()
            mapper.map_reads(bases[:nb_s], offsets[:sample_reads + 1], k)
            reduce_counts()
            torch.cuda.synchronize()
        got = counts.cpu().numpy().view(np.uint32)
        if world > 1:
            sizes = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([nb_s, sample_reads], dtype=torch.int64, device=device))
            sizes = [tuple(int(x) for x in t.tolist()) for t in sizes]
            cap_b, cap_r = max(b for b, _ in sizes), max(r for _, r in sizes)
            sb = torch.zeros(cap_b, dtype=torch.uint8, device=device)
            sb[:nb_s] = bases[:nb_s]
            so = torch.zeros(cap_r + 1, dtype=torch.int64, device=device)
            so[:sample_reads + 1] = offsets[:sample_reads + 1]
            gb = [torch.empty_like(sb) for _ in range(world)] if rank == 0 else None
            go = [torch.empty_like(so) for _ in range(world)] if rank == 0 else None
            dist.gather(sb, gb, dst=0)
            dist.gather(so, go, dst=0)
            if rank == 0:
                parts_b = [gb[r][:sizes[r][0]].cpu().numpy() for r in range(world)]
                parts_o, base = [np.zeros(1, np.int64)], 0
                for r in range(world):
                    o = go[r][:sizes[r][1] + 1].cpu().numpy()
                    parts_o.append(o[1:] + base)
                    base += sizes[r][0]
                samp_b, samp_o = np.concatenate(parts_b), np.concatenate(parts_o)
                del gb, go
        else:
            samp_b, samp_o = bases[:nb_s].cpu().numpy(), offsets[:sample_reads + 1].cpu().numpy()
        if rank == 0:
            from oracle import c_oracle
            want, n_want = c_oracle.map_reads(_oracle_index(host_index), max_node, samp_b, samp_o, k, n_threads=cores)
            checks["sample_counts_bit_exact_vs_oracle"] = bool(np.array_equal(got, want))
            checks["sample_kmers"] = int(n_want)
            checks["sample"] = "first %d reads of each of %d rank(s), mapped device-resident, %s" % (
                sample_reads, world, "summed by kmb_mapper_allreduce (NCCL)" if world > 1 else "one GPU")
            del samp_b, samp_o, want

    # ---- the whole step against the oracle (N == 1, --full-oracle): every read of the timed batch
    if args.full_oracle and world == 1 and not args.no_oracle:
        from oracle import c_oracle
        t0 = time.perf_counter()
        hb_all, ho_all = bases.cpu().numpy(), offsets.cpu().numpy()
        want, n_want = c_oracle.map_reads(_oracle_index(host_index), max_node, hb_all, ho_all, k, n_threads=cores)
        checks["full_step_counts_bit_exact_vs_oracle"] = bool(np.array_equal(full_counts.cpu().numpy().view(np.uint32), want))
        checks["full_step_kmers"] = int(n_want)
        checks["full_step_oracle_seconds"] = round(time.perf_counter() - t0, 1)
        del hb_all, ho_all, want

    # ---- e2e: host buffers through the public API, H2D and D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        hb = torch.empty(n_bases, dtype=torch.uint8, pin_memory=True)
        ho = torch.empty(n_reads + 1, dtype=torch.int64, pin_memory=True)
        hc = torch.empty(n_counts, dtype=torch.int32, pin_memory=True)
        hb.copy_(bases)
        ho.copy_(offsets)
        torch.cuda.synchronize()
        hb_np, ho_np, hc_np = hb.numpy(), ho.numpy(), hc.numpy().view(np.uint32)
        # the same transport policy at every N -- the library's default for a pinned source, the hybrid: a chunk goes as
        # ASCII straight from this buffer when the bus is about to run dry and is packed to 2 bits per base by this
        # rank's share of the host cores otherwise (distributed.init_process_group set host_threads = cores / ranks)
        _lib.set_option("host_pack", args.host_pack)

        def step_e2e():
            mapper.reset()
            mapper.map_reads(hb_np, ho_np, k)            # pinned host -> staged H2D (copy stream) -> kernels
            reduce_counts()
            mapper.counts(out=hc_np)                       # result D2H, synchronises

        # (a batch smaller than a few staging chunks -- config 1 -- needs more warm-up calls: the library allocates its three
        # pinned staging slots one call at a time, which costs milliseconds where a steady-state step takes 0.5 ms)
        for _ in range(8 if n_bases < (256 << 20) else max(1, min(args.warmup, 2))):
            step_e2e()
        torch.cuda.synchronize()
        barrier()
        h2d_before = _lib.get_option("h2d_bytes")
        chunks_before = _lib.get_option("host_chunks_packed"), _lib.get_option("host_chunks_ascii")
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        torch.cuda.synchronize()
        barrier()
        dt = float(over_ranks(time.perf_counter() - t0, dist.ReduceOp.MAX))
        h2d_per_step = (_lib.get_option("h2d_bytes") - h2d_before) // args.steps
        n_packed = _lib.get_option("host_chunks_packed") - chunks_before[0]
        n_ascii = _lib.get_option("host_chunks_ascii") - chunks_before[1]
        host_threads = _lib.get_option("host_threads") or len(os.sched_getaffinity(0))
        e2e = {"value": total_kmers * args.steps / dt / 1e9, "unit": UNIT,
               # bytes the library actually put on the bus (counted where it issues the copies): with the packed
               # transport the bases cross as 2 bits each, encoded on the host inside the timed region
               "h2d_bytes_per_step": int(h2d_per_step), "d2h_bytes_per_step": int(4 * n_counts),
               "host_input_bytes_per_step": int(n_bases + 8 * (n_reads + 1)),
               "host_transport": "%d of %d chunks 2-bit packed on %d CPU threads, the others ASCII over the bus from the caller's "
               "pinned buffer (host_pack=%d)" % (n_packed, n_packed + n_ascii, host_threads, args.host_pack),
               "chunks_packed_fraction": n_packed / max(n_packed + n_ascii, 1),
               "ms_per_step": dt * 1e3 / args.steps, "timing": "host wall clock between device synchronisations",
               "bytes_are": "per rank"}
        checks["e2e_counts_equal_resident_counts"] = bool(np.array_equal(hc_np, full_counts.cpu().numpy().view(np.uint32)))
        _lib.set_option("host_pack", -1)
        # what bounds it: the host's memory system (every base is read from host DRAM at least once -- by the cores that
        # pack it or by the GPU's DMA engine -- and all ranks of the node share it) and this GPU's PCIe link
        import ctypes as C
        gbs = C.c_double()
        probe = min(n_bases, 4 << 30)
        _lib.check(_lib.lib().kmb_host_read_bandwidth(hb_np.ctypes.data, probe, 0, C.byref(gbs)))      # page in / warm up
        barrier()                                                                                        # all ranks at once
        _lib.check(_lib.lib().kmb_host_read_bandwidth(hb_np.ctypes.data, probe, 0, C.byref(gbs)))
        e2e["host_read_GBps_per_rank_all_ranks_at_once"] = round(gbs.value, 1)
        dst = torch.empty(min(n_bases, 1 << 30), dtype=torch.uint8, device=device)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dst.copy_(hb[:dst.shape[0]], non_blocking=True)
        torch.cuda.synchronize()
        barrier()
        ev0.record()
        for i in range(6):
            lo = (i * dst.shape[0]) % max(n_bases - dst.shape[0], 1)
            dst.copy_(hb[lo:lo + dst.shape[0]], non_blocking=True)
        ev1.record()
        torch.cuda.synchronize()
        e2e["pcie_h2d_GBps_per_rank_all_ranks_at_once"] = round(6 * dst.shape[0] / (ev0.elapsed_time(ev1) / 1e3) / 1e9, 1)
        e2e["bound"] = ("host memory bandwidth and PCIe: every base is read from host DRAM once (1 byte: DMA of an ASCII chunk) or "
                        "~1.5 times (read by the cores, written packed, read by the DMA engine), and an ASCII chunk needs 4x the bus time")
        del dst
        del hb, ho, hc

    # ---- reads/s from files (N == 1): FASTQ and multi-member FASTQ.gz of the batch's first reads, file open -> counts on
    # the host, records parsed on the device; the counts must equal those of the same reads mapped device-resident
    reads_per_s = None
    if world == 1 and not args.no_files and L * n_reads >= 1000:
        n_file_reads = min(n_reads, args.file_reads)
        fdir = tempfile.mkdtemp(prefix="kmb_files_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            fq, fqgz, fqbgzf = write_fastq_files(fdir, bases, n_file_reads, L)
            with torch.cuda.stream(stream):
                mapper.reset()
                mapper.map_reads(bases[:n_file_reads * L], offsets[:n_file_reads + 1], k)
                mapper.flush()
                torch.cuda.synchronize()
            want_counts = counts.cpu().numpy().view(np.uint32)
            reads_per_s = {"n_reads": int(n_file_reads), "route": "file -> pinned staging -> H2D -> [gzip members inflated on the "
                           "device, one per warp: kmb_mapper_map_gz, when they average <= 512 KB of text (BGZF); by the host "
                           "decoders on all cores otherwise] -> record parsing on the device (kmb_mapper_map_text) -> fused "
                           "kernel -> counts D2H; wall clock from opening the file, index already on the device, file in "
                           "the page cache (tmpfs)", "host_cores": cores,
                           "fastq_gz_bgzf_members": "65280 bytes of text each (what bgzip writes)",
                           "fastq_gz_multi_member_members": "~4 MB of text each"}
            for name, path in (("fastq", fq), ("fastq_gz_bgzf", fqbgzf), ("fastq_gz_multi_member", fqgz)):
                rate, n_seen, secs, got_counts, dev_members = time_file_to_counts(path, di, n_counts, k)
                reads_per_s[name] = rate
                if path.endswith(".gz"):
                    reads_per_s[name + "_members_inflated_on_device"] = dev_members
                reads_per_s[name + "_file_bytes"] = os.path.getsize(path)
                reads_per_s[name + "_seconds"] = secs
                checks["file_%s_counts_equal_resident_counts" % name] = bool(n_seen == n_file_reads and np.array_equal(got_counts, want_counts))
        finally:
            shutil.rmtree(fdir, ignore_errors=True)

    # ---- roofline.  SURVEY.md 8(d): A = L/(L-k+1) + 32 + 8h bytes per k-mer.  The fused kernel streams the bases
    # and gathers the sectors (L/(L-k+1) + 32); the 8h counter bytes belong to the apply pass, which is what
    # touches the counters.  `frac` = the dominant (fused) kernel, `frac_step` = all of A over the whole step.
    h = n_counted_step / max(n_kmers_step, 1)
    kernel_bytes_per_kmer = L / max(L - k + 1, 1) + 32.0
    bytes_per_kmer = kernel_bytes_per_kmer + 8.0 * h
    peak, peak_src = measured_peak()
    kernel_avg_ms = kernel_ms / max(kernel_n, 1)
    kernels_per_step = kernel_n / max(args.steps, 1)
    achieved = kernel_bytes_per_kmer * n_kmers_step / max(kernels_per_step, 1) / (kernel_avg_ms / 1e3) / 1e9 if kernel_n else None
    step_ms_rank = ms / args.steps
    achieved_step = bytes_per_kmer * n_kmers_step / (step_ms_rank / 1e3) / 1e9
    apply_avg_ms = apply_ms / max(apply_n, 1)
    kernel_name = "kmb_map_reads_mz_kernel (read-path table)" if _lib.get_option("last_reads_kernel") else "kmb_map_reads_kernel"
    traffic, traffic_src = known_traffic(args.workload, kernel_name.split(" ")[0])
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None,
                "achieved_step": achieved_step, "frac_step": achieved_step / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": kernel_name,
                "kernel_ms": kernel_avg_ms, "kernel_share_of_step": kernel_ms / ms if ms else None,
                "kernel_bytes_per_kmer": kernel_bytes_per_kmer, "bytes_per_kmer": bytes_per_kmer, "counted_entries_per_kmer": h,
                "apply": {"kernel": "kmb_log_apply_kernel", "ms": apply_avg_ms, "share_of_step": apply_ms / ms if ms else None,
                          "bytes_per_kmer": 8.0 * h,
                          "achieved": (8.0 * h * n_kmers_step / (apply_avg_ms / 1e3) / 1e9) if apply_n else None,
                          "reductions_per_s": (n_counted_step / (apply_avg_ms / 1e3)) if apply_n else None},
                "non_kernel_share_of_step": 1.0 - (kernel_ms + apply_ms) / ms if ms else None,
                "sector_fetches_per_kmer": n_candidates_step / max(n_kmers_step, 1), "peak_source": peak_src,
                "filter_bytes": di.filter_bytes}

    # ---- the random-gather micro-roofline of SURVEY.md 8(d), measured live on this GPU: uniform random 8-byte
    # loads over a table of the sector table's footprint (capped at 8 GB); the north_star's denominator
    if rank == 0:
        try:
            import ctypes
            ms_g = ctypes.c_float(0)
            table_bytes = int(min(di.device_bytes, 8 << 30))
            n_loads = 1 << 28
            _lib.check(_lib.lib().kmb_bench_gather(local_rank, table_bytes, n_loads, 8, 8, 256, 8, ctypes.byref(ms_g)))
            gathers_per_s = n_loads / (ms_g.value / 1e3)
            kernel_kmers_per_s = n_kmers_step / max(kernels_per_step, 1) / (kernel_avg_ms / 1e3)
            roofline["gather_roofline"] = {
                "random_8B_loads_per_s": gathers_per_s, "table_bytes": table_bytes,
                "kernel_kmers_per_s": kernel_kmers_per_s,
                "fraction_of_gather_roofline": kernel_kmers_per_s / gathers_per_s,
                "note": "k-mers mapped per second by the fused kernel / random HBM gathers per second the chip sustains; "
                        "above 1 because the L2-resident filter answers most k-mers without a gather"}
        except Exception as e:  # measurement helper only
            roofline["gather_roofline"] = {"error": str(e)}

    # ---- CPU baseline (the reference's path on the host cores, bounded sample)
    cpu_rec = None
    if do_cpu:
        try:
            del host_index
            res = run_cpu_child(cpu_dir, steps=2, warmup=1)
            desc = "first %d reads of %s (%d k-mers per pass)" % (sample_reads, args.workload, res["n_kmers"])
            cpu_rec = cpu_baseline_record(res, desc)
        finally:
            shutil.rmtree(cpu_dir, ignore_errors=True)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": dict(workload=args.workload, k=k, index_entries=w["entries"], modulo=w["modulo"], nodes=w["nodes"],
                               genome_bases=w["genome"], reads_per_gpu_per_step=n_reads,
                               reads_per_step=int(w.get("reads_total", n_reads * world)), read_len=L,
                               kmers_per_step=total_kmers, scale=args.scale,
                               parallelism="reads sharded over %d GPU(s), index replicated, one uint32 all-reduce per step "
                                           "(kmb_mapper_allreduce: NCCL through the C ABI)" % world,
                               l2="inputs larger than L2 (reads %.1f GB, index %.1f GB per step); no flush"
                                  % (n_bases / 1e9, di.device_bytes / 1e9),
                               index_device_bytes=di.device_bytes, setup_seconds=round(setup_s, 1),
                               index_layout=dict(main_sectors=di.n_main_lines, overflow_sectors=di.n_overflow_lines,
                                                 filter_bytes=di.filter_bytes),
                               options={n: _lib.get_option(n) for n in ("gathers_in_flight", "use_filter", "apply_window_log2",
                                                                        "map_reads_blocks_per_sm")}),
                "clocks": clock_rec, "e2e": e2e, "reads_per_s": reads_per_s, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu_rec, "checks": checks}
        print(json.dumps(line))
    bad = [k_ for k_, v in checks.items() if v is False]
    bad_any = int(over_ranks(len(bad), dist.ReduceOp.SUM))
    if comm is not None:
        comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if bad:
        print("PARITY CHECK FAILED: %s" % bad, file=sys.stderr)
    if bad_any:
        sys.exit(3)


def main():
    args = parse_args()
    if args.cpu_worker:
        return cpu_worker_main(args)
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


if __name__ == "__main__":
    main()
