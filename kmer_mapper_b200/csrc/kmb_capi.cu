// kmb_capi.cu -- the C ABI of include/kmer_mapper_b200.h: handles, streams, staging and launches.
//
// Nothing in this file computes on the CPU.  Host code here only validates arguments, moves bytes
// (pinned/pageable host memory -> device staging on a copy stream, double buffered against the
// compute stream) and launches the kernels of kmb_kernels.cuh.  Without a CUDA device every compute
// entry point returns KMB_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "../../include/kmer_mapper_b200.h"
#include "kmb_host.h"
#include "kmb_kernels.cuh"
#include "kmb_textparse.cuh"
#include "kmb_gzdev.cuh"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int kmb_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define KMB_CUDA(call)                                                                                       \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess) {                                                                            \
            cudaGetLastError();                                                                              \
            return kmb_fail(e__ == cudaErrorMemoryAllocation ? KMB_ERR_NOMEM : KMB_ERR_CUDA, "%s: %s (%s:%d)", \
                            #call, cudaGetErrorString(e__), __FILE__, __LINE__);                             \
        }                                                                                                    \
    } while (0)

#define KMB_TRY(call)              \
    do {                           \
        int rc__ = (call);         \
        if (rc__ != KMB_OK) return rc__; \
    } while (0)

// ------------------------------------------------------------------------------------------------
// options and launch accounting
// ------------------------------------------------------------------------------------------------
struct KmbOptions {
    int64_t map_reads_blocks_per_sm = 0;  // 0 = whatever the occupancy calculator allows
    int64_t map_carveout = -1;            // shared-memory carve-out of the mapping kernels in percent of the maximum, -1 = driver's choice
    int64_t map_kmers_blocks_per_sm = 0;
    int64_t probe_variant = 1;            // map_kmers: 0 = one query per thread, 1 = staged probe with warp stack
    int64_t chunk_bytes = 64ll << 20;     // staging slot size for host input
    int64_t gathers_in_flight = 4;        // U: 2 or 4 independent filter loads per thread
    int64_t log_max_entries = 2048ll << 20;  // upper bound of the hit log (x 4 bytes)
    int64_t sectors_per_100_entries = 250; // main sectors per 100 index entries (mean occupancy 0.4 of 2 slots)
    int64_t use_filter = -1;              // -1 auto (filter fits the L2 budget), 0 off, 1 on
    int64_t filter_l2_budget_bytes = 60ll << 20;  // the L2 keeps ~72 MB of randomly accessed data (profiles/README.md)
    int64_t policy_filter = 2;            // L2 priority hints: 0 normal, 1 evict-first, 2 evict-last
    int64_t policy_line = 0;
    int64_t ablate = 0;                   // measurement only (results become wrong): 1 no RED, 2 no line loads, 4 no filter loads, 8 no key loads
    int64_t l2_persist = 1;               // set a persisting-L2 access window over the filter
    int64_t l2_fetch_granularity = 0;     // 0 = leave the device default; else 32/64/128 (cudaLimitMaxL2FetchGranularity)
    int64_t bench_load_mode = 0;          // kmb_bench_gather, 8-byte loads: 0 .nc, 1-3 L2::64B/128B/256B, 4 plain, 5 .cv
    int64_t bench_grid_blocks = 0;        // kmb_bench_gather: total CTAs (0 = SMs x blocks_per_sm)
    int64_t time_kernels = 0;             // bracket every mapping kernel with CUDA events (kmb_mapper_kernel_time)
    // Host input of map_reads travels as 2 bits per base, encoded on the CPU (kmb_hostpack.cpp): 1 always, 0 never,
    // -1 auto = when that beats sending the ASCII bytes.  Measured on the 16-core B200 boxes
    // (profiles/r01_v7_host_pack.jsonl): a pinned source crosses PCIe at ~48 GB/s as ASCII and the encoder makes
    // 4.6 GB/s per thread (DRAM-bound at ~78 GB/s), so packing wins from ~10 threads up; a pageable source only
    // reaches ~10 GB/s through the driver's staging copy, so packing wins from 2 threads up.
    // Fused reads kernel over the minimizer-bucketed read-path table (kmb_core.cuh), k = 31 without reverse
    // complements only: 1 = always, 0 = never, -1 (default) = when the key filter is too thin to help.  Measured
    // (profiles/README.md): with 5 filter bits per key (config 2) the key-addressed kernel needs 0.185 fetches per
    // k-mer and wins, 46 ms against 54 ms per 6.0 G k-mers -- the minimizer arithmetic costs more issue slots than
    // the saved look-ups give back; with 1 bit per key (config 3: 500 M entries) it needs 0.68 fetches per k-mer
    // and loses, 65.9 ms against 50.7 ms per 3.0 G k-mers.  Auto takes the table when the filter has less than 2.5
    // bits per key (one probe bit, or no filter at all) and the index is too big for the L2 anyway.
    int64_t read_table = -1;
    int64_t read_table_min_entries = 8ll << 20;  // auto: smaller indexes sit in the L2 anyway
    // buckets (two sectors = 64 bytes each) of the read-path table per 100 live entries: sparser = fewer buckets
    // that spill into secondary and pool sectors (config 2 kernel: 51.8 / 50.0 / 48.6 ms at 100 / 150 / 200)
    int64_t read_table_buckets_per_100_entries = 150;
    int64_t filter_probes = 0;            // filter bits per key: 0 = by filter density (filter_probes()), else 1..3
    // The encoder is bound by the host's DRAM bandwidth, which the ranks of a multi-GPU node share, while every GPU
    // has its own PCIe link: with 2 ranks on one host a pinned source went 46.9 GK/s packed against 74.3 as ASCII
    // (profiles/README.md), so auto packs a pinned source only when this process has the host to itself.
    // 1 = every chunk packed, 0 = every chunk as ASCII, 2 = hybrid: the bus and the cores work SIDE BY SIDE -- a chunk
    // goes as ASCII straight from the caller's pinned buffer whenever less than host_hybrid_backlog_bytes are waiting
    // for the bus (costs no CPU time), and is packed by the cores otherwise (costs a quarter of the bus time), so
    // bases reach the GPU at about the SUM of the two rates; -1 (default) = packed for a pageable source (which has to
    // be staged by the CPU anyway); for a pinned one hybrid when the process has the host to itself (host_ranks = 1),
    // ASCII when it shares it.
    int64_t host_pack = -1;
    int64_t host_hybrid_backlog_bytes = 0;   // 0 = one chunk (chunk_bytes)
    int64_t apply_slabs_per_sm = 8;          // apply pass: CTAs per SM and node window (each walks one slab of the log)
    // 1 = the encoder writes the packed words with streaming (non-temporal) stores: written once, read next by the DMA
    // engine, and a write-allocating store would first read the line it overwrites (config 2 end to end, 16 cores:
    // packed 51.7 -> 55.2 GK/s, hybrid 68.3 -> 70.3)
    int64_t host_pack_streaming = 1;
    int64_t host_threads = 0;             // CPU threads of the host-side encoder: 0 = every CPU of the affinity mask
    int64_t host_ranks = 1;               // processes sharing this host's memory system (set by distributed.py)
    // The apply pass reduces into one window of 2^apply_window_log2 nodes at a time (x 4 bytes: 23 = 32 MB), reading the
    // groups of the window's node range once per window of the range.  0 = auto: 2^23 when that means at most two
    // passes over a range's groups, else 2^24.  Measured per 0.51 G reductions with the stream evict-first and the
    // counters evict-last: 80 M nodes (ranges of 2^24): 3.3 ms at 2^23, 4.0 at 2^24; 400 M nodes (ranges of 2^26):
    // 12.1 ms at 2^23 (eight passes), 6.6 at 2^24, 10.8 at 2^25, 16.0 at 2^26 (one pass, window far larger than the L2).
    int64_t apply_window_log2 = 0;
    // Count arrays of at most this many nodes (4 Mi = 16 MB: L2-resident whatever else is going on) are reduced onto
    // directly by the mapping kernels (warp-aggregated RED, kmb_count_direct): no hit log, no apply pass, no log reset --
    // three launches less per step, which is what a small job consists of (config 1: 100 k reads, 0.18 ms per step).
    int64_t direct_counts_max_nodes = 4ll << 20;
    // gzip members inflated on the device (kmb_mapper_map_gz): a member that announces more text than this is left to
    // the host decoders (one warp decodes one member: a plain single-member .gz has no parallelism to offer); text per
    // batch; whether every member's CRC-32 is recomputed on the device and compared with its trailer
    int64_t gz_device_max_member_bytes = 16ll << 20;
    int64_t gz_device_batch_bytes = 256ll << 20;
    // one warp inflates ~15 MB/s of text, the host decoders ~0.3 GB/s per core: the device wins when a batch holds
    // hundreds of members (bgzip: 64 KB each -> 4000 per batch, 14.5 GB/s measured against 4.8 on 16 cores), and loses
    // when it holds a few dozen large ones (4 MB members: 0.9 GB/s), so files whose members average more text than
    // this are left to the host decoders
    int64_t gz_device_max_mean_member_bytes = 512ll << 10;
    int64_t gz_device_crc = 1;
};
static KmbOptions g_opt;
static std::atomic<unsigned long long> g_launches{0};
static std::atomic<unsigned long long> g_host_chunks_packed{0}, g_host_chunks_ascii{0};
static std::atomic<unsigned long long> g_h2d_bytes{0};  // bytes the mapping calls sent host -> device (bench.py's e2e)
static std::atomic<unsigned long long> g_text_reads{0}, g_text_bases{0};  // parsed by kmb_mapper_map_text
// where the host side of kmb_mapper_map_text spends its time (microseconds, cumulative): waiting for a free slot,
// copying pageable text into pinned staging, waiting for the H2D copy + parse kernels, allocating
static std::atomic<unsigned long long> g_text_us_slot{0}, g_text_us_stage{0}, g_text_us_wait{0}, g_text_us_alloc{0};
static std::atomic<int> g_last_reads_kernel{0};  // 0 = key-addressed fused kernel, 1 = read-path table kernel

extern "C" int kmb_set_option(const char *name, int64_t value) {
    if (!name) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_set_option: null name");
#define OPT(n)                    \
    if (!strcmp(name, #n)) {      \
        g_opt.n = value;          \
        return KMB_OK;            \
    }
    OPT(map_reads_blocks_per_sm)
    OPT(map_carveout)
    OPT(map_kmers_blocks_per_sm)
    OPT(probe_variant)
    OPT(gathers_in_flight)
    OPT(log_max_entries)
    OPT(use_filter)
    OPT(sectors_per_100_entries)
    OPT(filter_l2_budget_bytes)
    OPT(l2_persist)
    OPT(ablate)
    OPT(policy_filter)
    OPT(policy_line)
    OPT(time_kernels)
    OPT(l2_fetch_granularity)
    OPT(bench_grid_blocks)
    OPT(bench_load_mode)
    OPT(host_pack)
    OPT(host_threads)
    OPT(host_ranks)
    OPT(read_table)
    OPT(read_table_min_entries)
    OPT(read_table_buckets_per_100_entries)
    OPT(filter_probes)
    OPT(apply_window_log2)
    OPT(direct_counts_max_nodes)
    OPT(host_hybrid_backlog_bytes)
    OPT(apply_slabs_per_sm)
    OPT(host_pack_streaming)
    OPT(gz_device_max_member_bytes)
    OPT(gz_device_batch_bytes)
    OPT(gz_device_max_mean_member_bytes)
    OPT(gz_device_crc)
#undef OPT
    if (!strcmp(name, "chunk_bytes")) {
        if (value < (1 << 16)) return kmb_fail(KMB_ERR_BAD_ARG, "chunk_bytes must be >= 65536");
        g_opt.chunk_bytes = (value + 255) & ~255ll;
        return KMB_OK;
    }
    return kmb_fail(KMB_ERR_BAD_ARG, "kmb_set_option: unknown option '%s'", name);
}

extern "C" int kmb_get_option(const char *name, int64_t *value) {
    if (!name || !value) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_get_option: null argument");
#define OPT(n)                    \
    if (!strcmp(name, #n)) {      \
        *value = g_opt.n;         \
        return KMB_OK;            \
    }
    OPT(map_reads_blocks_per_sm)
    OPT(map_carveout)
    OPT(map_kmers_blocks_per_sm)
    OPT(probe_variant)
    OPT(gathers_in_flight)
    OPT(log_max_entries)
    OPT(use_filter)
    OPT(sectors_per_100_entries)
    OPT(filter_l2_budget_bytes)
    OPT(l2_persist)
    OPT(ablate)
    OPT(policy_filter)
    OPT(policy_line)
    OPT(time_kernels)
    OPT(l2_fetch_granularity)
    OPT(bench_grid_blocks)
    OPT(bench_load_mode)
    OPT(host_pack)
    OPT(host_threads)
    OPT(host_ranks)
    OPT(read_table)
    OPT(read_table_min_entries)
    OPT(read_table_buckets_per_100_entries)
    OPT(filter_probes)
    OPT(apply_window_log2)
    OPT(direct_counts_max_nodes)
    OPT(host_hybrid_backlog_bytes)
    OPT(apply_slabs_per_sm)
    OPT(host_pack_streaming)
    OPT(gz_device_max_member_bytes)
    OPT(gz_device_batch_bytes)
    OPT(gz_device_max_mean_member_bytes)
    OPT(gz_device_crc)
    OPT(chunk_bytes)
#undef OPT
    if (!strcmp(name, "last_reads_kernel")) {  // read-only: which fused kernel the last map_reads launch used
        *value = g_last_reads_kernel.load();
        return KMB_OK;
    }
    if (!strcmp(name, "host_chunks_packed")) { *value = (int64_t)g_host_chunks_packed.load(); return KMB_OK; }   // read-only
    if (!strcmp(name, "host_chunks_ascii")) { *value = (int64_t)g_host_chunks_ascii.load(); return KMB_OK; }     // read-only
    if (!strcmp(name, "text_us_slot")) { *value = (int64_t)g_text_us_slot.load(); return KMB_OK; }
    if (!strcmp(name, "text_us_stage")) { *value = (int64_t)g_text_us_stage.load(); return KMB_OK; }
    if (!strcmp(name, "text_us_wait")) { *value = (int64_t)g_text_us_wait.load(); return KMB_OK; }
    if (!strcmp(name, "text_us_alloc")) { *value = (int64_t)g_text_us_alloc.load(); return KMB_OK; }
    if (!strcmp(name, "text_reads")) {  // read-only: reads parsed by kmb_mapper_map_text so far
        *value = (int64_t)g_text_reads.load();
        return KMB_OK;
    }
    if (!strcmp(name, "h2d_bytes")) {  // read-only
        *value = (int64_t)g_h2d_bytes.load();
        return KMB_OK;
    }
    if (!strcmp(name, "bounds_failures")) {  // read-only: -1 unless this is the bounds-checked build
#ifdef KMB_BOUNDS_CHECKS
        unsigned long long h[KMB_BOUND_SITES];
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaMemcpyFromSymbol(h, g_kmb_bound_failures, sizeof(h));
        if (e != cudaSuccess) {
            cudaGetLastError();
            return kmb_fail(KMB_ERR_CUDA, "bounds_failures: %s", cudaGetErrorString(e));
        }
        unsigned long long total = 0;
        for (int i = 0; i < KMB_BOUND_SITES; i++) {
            if (h[i]) fprintf(stderr, "kmer_mapper_b200: bounds check site %d failed %llu times\n", i, h[i]);
            total += h[i];
        }
        *value = (int64_t)total;
#else
        *value = -1;
#endif
        return KMB_OK;
    }
    return kmb_fail(KMB_ERR_BAD_ARG, "kmb_get_option: unknown option '%s'", name);
}

extern "C" int kmb_launch_count(uint64_t *n) {
    if (!n) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_launch_count: null");
    *n = g_launches.load();
    return KMB_OK;
}

extern "C" const char *kmb_last_error(void) { return g_err; }
extern "C" const char *kmb_version(void) { return "kmer_mapper_b200 0.1.0 (sm_100a)"; }

extern "C" int kmb_device_count(int *n_devices) {
    if (!n_devices) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_device_count: null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *n_devices = 0;
        return kmb_fail(KMB_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *n_devices = n;
    return KMB_OK;
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
struct DevInfo {
    int sms = 0;
    int l2_bytes = 0;
    int max_persist_l2 = 0;
};
static int dev_info(int device, DevInfo *d) {
    KMB_CUDA(cudaDeviceGetAttribute(&d->sms, cudaDevAttrMultiProcessorCount, device));
    KMB_CUDA(cudaDeviceGetAttribute(&d->l2_bytes, cudaDevAttrL2CacheSize, device));
    KMB_CUDA(cudaDeviceGetAttribute(&d->max_persist_l2, cudaDevAttrMaxPersistingL2CacheSize, device));
    return KMB_OK;
}

// Is p device memory (usable by kernels of `device`)?  Plain malloc / numpy memory is "unregistered".
static int ptr_on_device(const void *p, int device, bool *on_device, bool *pinned = nullptr) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (pinned) *pinned = e == cudaSuccess && a.type == cudaMemoryTypeHost;
    if (e != cudaSuccess) {
        cudaGetLastError();
        *on_device = false;
        return KMB_OK;
    }
    if (a.type == cudaMemoryTypeDevice) {
        if (a.device != device)
            return kmb_fail(KMB_ERR_BAD_ARG, "device buffer lives on GPU %d, handle is on GPU %d", a.device, device);
        *on_device = true;
    } else if (a.type == cudaMemoryTypeManaged) {
        *on_device = true;
    } else {
        *on_device = false;
    }
    return KMB_OK;
}

struct DeviceGuard {  // every entry point runs on the handle's GPU and restores the caller's
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            cudaGetLastError();
            prev = -1;
        }
        ok = cudaSetDevice(device) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define KMB_ON_DEVICE(dev)  \
    DeviceGuard guard__(dev); \
    if (!guard__.ok) return kmb_fail(KMB_ERR_CUDA, "cudaSetDevice(%d) failed: no such CUDA device", dev)

template <class T>
struct DevBuf {  // RAII scratch allocation
    T *p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t n) {
        KMB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
        return KMB_OK;
    }
    T *release() {
        T *r = p;
        p = nullptr;
        return r;
    }
};

// Bring `n` elements to the device if they are on the host; `out` = usable device pointer.
template <class T>
static int to_device(const T *src, size_t n, int device, DevBuf<T> &tmp, const T **out, cudaStream_t s) {
    bool dev;
    KMB_TRY(ptr_on_device(src, device, &dev));
    if (dev) {
        *out = src;
        return KMB_OK;
    }
    KMB_TRY(tmp.alloc(n));
    if (n) KMB_CUDA(cudaMemcpyAsync(tmp.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
    *out = tmp.p;
    return KMB_OK;
}

// Bits set per key in the word-blocked filter: the classic optimum is ln 2 x bits per key; inside one 32-bit word
// more than three bits collide too often to pay, and below 2.5 bits per key a second bit only fills the filter up.
static uint32_t filter_probes(double bits_per_key) {
    if (g_opt.filter_probes >= 1 && g_opt.filter_probes <= 3) return (uint32_t)g_opt.filter_probes;
    return bits_per_key >= 4.0 ? 3u : (bits_per_key >= 2.5 ? 2u : 1u);
}

static int grid_for(size_t work_items, int block, int sms, int per_sm = 8) {
    size_t need = (work_items + block - 1) / block;
    size_t cap = (size_t)sms * per_sm;
    return (int)std::max<size_t>(1, std::min(need, cap));
}

// ------------------------------------------------------------------------------------------------
// index
// ------------------------------------------------------------------------------------------------
struct kmb_index {
    int device = 0;
    uint64_t modulo = 0, n_entries = 0, n_live = 0;
    KmbAddr addr;                       // key -> sector / filter word / filter bits
    uint64_t n_main = 0, n_lines = 0;   // main lines, main + overflow lines
    uint32_t *lines = nullptr;          // 32-byte sectors (header, frequencies, keys, nodes), read-only after the build
    uint32_t *filter = nullptr;
    size_t filter_bytes = 0;
    bool filter_on = false;
    int64_t max_node = -1;
    uint64_t device_bytes = 0;
    KmbMod mod;
    DevInfo info;
    // read-path table: the same live entries filed under the minimizer of their key (built on first use)
    std::mutex mz_mu;
    int mz_k = 0;                       // 0 = not built
    KmbAddr mz_addr;                    // n_main = buckets (two sectors each)
    uint64_t mz_n_lines = 0;
    uint32_t *mz_lines = nullptr;
    uint32_t *mz_filter = nullptr;
    size_t mz_filter_bytes = 0;
    uint64_t mz_entries = 0;
};

extern "C" int kmb_index_destroy(kmb_index *ix) {
    if (!ix) return KMB_OK;
    DeviceGuard g(ix->device);
    cudaFree(ix->lines);
    cudaFree(ix->filter);
    cudaFree(ix->mz_lines);
    cudaFree(ix->mz_filter);
    cudaGetLastError();
    delete ix;
    return KMB_OK;
}

extern "C" int kmb_index_create(int device, const int32_t *hashes_to_index, const int32_t *n_kmers, uint64_t modulo,
                                const int32_t *nodes, const uint64_t *kmers, const uint16_t *frequencies,
                                uint64_t n_entries, kmb_index **out) {
    if (!out) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_index_create: out is null");
    *out = nullptr;
    if (!hashes_to_index || !n_kmers) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_index_create: null bucket arrays");
    if (n_entries && (!nodes || !kmers || !frequencies))
        return kmb_fail(KMB_ERR_BAD_ARG, "kmb_index_create: null entry arrays");
    if (modulo == 0 || modulo >= (1ull << 32))
        return kmb_fail(KMB_ERR_BAD_ARG, "kmb_index_create: modulo %llu outside [1, 2^32)", (unsigned long long)modulo);
    if (n_entries >= (1ull << 31))
        return kmb_fail(KMB_ERR_BAD_ARG, "kmb_index_create: n_entries %llu does not fit the int32 bucket positions",
                        (unsigned long long)n_entries);
    KMB_ON_DEVICE(device);
    kmb_index *ix = new (std::nothrow) kmb_index;
    if (!ix) return kmb_fail(KMB_ERR_NOMEM, "out of host memory");
    struct Cleanup {
        kmb_index *ix;
        ~Cleanup() {
            if (ix) kmb_index_destroy(ix);
        }
    } cleanup{ix};
    ix->device = device;
    ix->modulo = modulo;
    ix->n_entries = n_entries;
    ix->mod = kmb_mod_make(modulo);
    // main sectors: 2.5 per entry (mean occupancy 0.4 of 2 slots: < 1 % of the sectors need an overflow chain)
    ix->n_main = std::min<uint64_t>(std::max<uint64_t>(n_entries * (uint64_t)std::max<int64_t>(g_opt.sectors_per_100_entries, 50) / 100, 1), 1600000000ull);
    memset(&ix->addr, 0, sizeof(ix->addr));
    ix->addr.n_main = (uint32_t)ix->n_main;
    KMB_TRY(dev_info(device, &ix->info));

    cudaStream_t s = 0;  // index construction is a one-off: legacy default stream, synchronous
    DevBuf<int32_t> t_h2i, t_nk, t_nodes;
    DevBuf<uint64_t> t_kmers;
    DevBuf<uint16_t> t_freq;
    const int32_t *d_h2i, *d_nk, *d_nodes;
    const uint64_t *d_kmers;
    const uint16_t *d_freq;
    KMB_TRY(to_device(hashes_to_index, (size_t)modulo, device, t_h2i, &d_h2i, s));
    KMB_TRY(to_device(n_kmers, (size_t)modulo, device, t_nk, &d_nk, s));
    KMB_TRY(to_device(nodes, (size_t)n_entries, device, t_nodes, &d_nodes, s));
    KMB_TRY(to_device(kmers, (size_t)n_entries, device, t_kmers, &d_kmers, s));
    KMB_TRY(to_device(frequencies, (size_t)n_entries, device, t_freq, &d_freq, s));

    // Filter geometry: as many bits as the L2 budget allows, at most 16 per key; a second probe bit only with
    // >= 2.5 bits per key; no filter at all below 0.62 bits per key (> 80 % of the absent k-mers would pass).
    const uint64_t keys_for_filter = std::max<uint64_t>(n_entries, 1);
    uint64_t filter_words = std::min<uint64_t>((uint64_t)std::max<int64_t>(g_opt.filter_l2_budget_bytes, 4) / 4,
                                               std::max<uint64_t>(keys_for_filter / 2, 1));
    const double bits_per_key = 32.0 * (double)filter_words / (double)keys_for_filter;
    bool want_filter = g_opt.use_filter == 1 || (g_opt.use_filter < 0 && bits_per_key >= 0.62);
    if (!want_filter) filter_words = 0;
    ix->addr.n_filter_words = (uint32_t)filter_words;
    ix->addr.n_probes = filter_probes(bits_per_key);
    ix->filter_bytes = (size_t)std::max<uint64_t>(filter_words, 1) * 4;
    KMB_CUDA(cudaMalloc(&ix->filter, ix->filter_bytes));
    KMB_CUDA(cudaMemsetAsync(ix->filter, 0, ix->filter_bytes, s));
    DevBuf<uint32_t> line_fill;
    KMB_TRY(line_fill.alloc((size_t)ix->n_main));
    KMB_CUDA(cudaMemsetAsync(line_fill.p, 0, (size_t)ix->n_main * 4, s));
    DevBuf<KmbStatus> d_status;
    KMB_TRY(d_status.alloc(1));
    KmbStatus hs;
    memset(&hs, 0, sizeof(hs));
    hs.first_bad_offset = ~0ull;
    hs.max_node = -1;
    KMB_CUDA(cudaMemcpyAsync(d_status.p, &hs, sizeof(hs), cudaMemcpyHostToDevice, s));

    const int sms = ix->info.sms;
    // 1. the directory must stay inside the entry arrays (the reference runs with boundscheck off)
    kmb_build_check_buckets<<<grid_for(modulo, 256, sms), 256, 0, s>>>(d_h2i, d_nk, modulo, n_entries, d_status.p);
    g_launches++;
    KMB_CUDA(cudaGetLastError());
    KMB_CUDA(cudaMemcpyAsync(&hs, d_status.p, sizeof(hs), cudaMemcpyDeviceToHost, s));
    KMB_CUDA(cudaStreamSynchronize(s));
    if (hs.index_flags & 1u)
        return kmb_fail(KMB_ERR_BAD_INDEX,
                        "index: a bucket (hashes_to_index[h], n_kmers[h]) lies outside [0, n_entries=%llu] or has a "
                        "negative size (the reference would read out of bounds, mapper.pyx:15-18)",
                        (unsigned long long)n_entries);
    // 2. per-line counts + filter bits, 3. overflow lines needed
    if (n_entries) {
        kmb_build_count<<<grid_for(n_entries, 256, sms), 256, 0, s>>>(d_kmers, d_nodes, d_h2i, d_nk, n_entries, ix->mod,
                                                                 ix->addr, line_fill.p, ix->filter, d_status.p);
        g_launches++;
    }
    kmb_build_plan<false><<<grid_for(ix->n_main, 256, sms), 256, 0, s>>>(line_fill.p, ix->n_main, nullptr, 0, d_status.p);
    g_launches++;
    KMB_CUDA(cudaGetLastError());
    KMB_CUDA(cudaMemcpyAsync(&hs, d_status.p, sizeof(hs), cudaMemcpyDeviceToHost, s));
    KMB_CUDA(cudaStreamSynchronize(s));
    if (hs.index_flags & 2u)
        return kmb_fail(KMB_ERR_BAD_INDEX, "index: negative node id (the reference would write out of bounds)");
    ix->max_node = hs.max_node;
    ix->n_live = hs.n_live_entries;
    ix->n_lines = ix->n_main + hs.pool_lines;
    if (ix->n_lines >= (1ull << 31)) return kmb_fail(KMB_ERR_BAD_INDEX, "index: too many sectors (%llu) for the 31-bit chain links", (unsigned long long)ix->n_lines);
    // 4. sectors: headers of every chain, then placement
    KMB_CUDA(cudaMalloc(&ix->lines, (size_t)ix->n_lines * KMB_LINE_BYTES));
    KMB_CUDA(cudaMemsetAsync(ix->lines, 0, (size_t)ix->n_lines * KMB_LINE_BYTES, s));
    KMB_CUDA(cudaMemsetAsync(&d_status.p->pool_lines, 0, sizeof(unsigned int), s));
    kmb_build_plan<true><<<grid_for(ix->n_main, 256, sms), 256, 0, s>>>(line_fill.p, ix->n_main, ix->lines, ix->n_lines, d_status.p);
    g_launches++;
    if (n_entries) {
        kmb_build_scatter<<<grid_for(n_entries, 256, sms), 256, 0, s>>>(d_kmers, d_nodes, d_freq, d_h2i, d_nk, n_entries, ix->mod,
                                                                   ix->addr, line_fill.p, ix->lines, ix->n_lines);
        g_launches++;
    }
    KMB_CUDA(cudaGetLastError());
    KMB_CUDA(cudaStreamSynchronize(s));

    ix->filter_on = want_filter;
    if (!want_filter) {
        cudaFree(ix->filter);
        ix->filter = nullptr;
        ix->filter_bytes = 0;
    }
    ix->device_bytes = ix->n_lines * KMB_LINE_BYTES + ix->filter_bytes;
    cleanup.ix = nullptr;
    *out = ix;
    return KMB_OK;
}

extern "C" int kmb_index_info(const kmb_index *ix, int64_t *max_node_id, uint64_t *n_entries, uint64_t *modulo,
                              uint64_t *device_bytes) {
    if (!ix) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_index_info: null index");
    if (max_node_id) *max_node_id = ix->max_node;
    if (n_entries) *n_entries = ix->n_entries;
    if (modulo) *modulo = ix->modulo;
    if (device_bytes) *device_bytes = ix->device_bytes;
    return KMB_OK;
}

extern "C" int kmb_index_filter_bytes(const kmb_index *ix, uint64_t *bytes) {
    if (!ix || !bytes) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_index_filter_bytes: null argument");
    *bytes = ix->filter_on ? ix->filter_bytes : 0;
    return KMB_OK;
}

// Geometry of the line table: buckets per line, main lines, overflow lines, live entries.
extern "C" int kmb_index_layout(const kmb_index *ix, uint64_t *n_main_lines, uint64_t *n_overflow_lines,
                                uint64_t *n_live_entries) {
    if (!ix) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_index_layout: null index");
    if (n_main_lines) *n_main_lines = ix->n_main;
    if (n_overflow_lines) *n_overflow_lines = ix->n_lines - ix->n_main;
    if (n_live_entries) *n_live_entries = ix->n_live;
    return KMB_OK;
}

// ------------------------------------------------------------------------------------------------
// mapper
// ------------------------------------------------------------------------------------------------
struct StageSlot {
    uint8_t *data = nullptr;  // bases (ASCII or packed) or k-mers of one chunk
    int64_t *offsets = nullptr;
    uint32_t *tiles = nullptr;  // per 1024-position tile: the read it begins in (kmb_tile_reads_kernel)
    size_t data_cap = 0, off_cap = 0, tiles_cap = 0;
    uint32_t *h_words = nullptr;  // pinned: the chunk's packed bases, written by the host encoder, read by the DMA engine
    uint32_t *h_off = nullptr;    // pinned: its chunk-relative read offsets
    size_t h_words_cap = 0, h_off_cap = 0;
    cudaEvent_t copied = nullptr, consumed = nullptr;
    bool used = false;
    // device-side text parsing (kmb_mapper_map_text): the chunk's raw text and what the parse kernels make of it
    uint8_t *h_text = nullptr;          // pinned staging for text that arrives in pageable memory
    size_t h_text_cap = 0;
    uint8_t *text = nullptr;            // the text on the device
    uint8_t *tbases = nullptr;          // bases of the reads, back to back
    int64_t *toffsets = nullptr;        // read offsets
    uint32_t *nl_pos = nullptr;         // position of every newline
    KmbTextCopy *copies = nullptr;      // one per line
    uint32_t *blk_nl = nullptr, *blk_reads = nullptr;
    unsigned long long *blk_bases = nullptr, *scalars = nullptr;  // scalars: [0] newlines, [1] bases, [2] reads
    KmbTextResult *d_result = nullptr, *h_result = nullptr;       // h_result pinned
    size_t text_cap = 0, line_cap = 0;
    // gzip members inflated on the device (kmb_mapper_map_gz): compressed bytes, member table, results
    uint8_t *h_gz = nullptr, *d_gz = nullptr;      // h_gz pinned
    size_t h_gz_cap = 0, d_gz_cap = 0;
    KmbGzMember *h_members = nullptr, *d_members = nullptr;   // h_members pinned
    KmbGzResult *h_results = nullptr, *d_results = nullptr;   // h_results pinned
    uint32_t *h_crc = nullptr, *d_crc = nullptr;              // h_crc pinned
    uint8_t *h_windows = nullptr;                             // pinned: the first bytes of the batch's text and of its extra member
    size_t members_cap = 0;
    uint64_t bus_bytes = 0;             // host chunks: what this slot put on the bus (0 once seen across), hybrid transport
    const uint8_t *text_ptr = nullptr;  // the text the parse kernels read: s.text, or a window of it (gz route)
    uint64_t text_n = 0;                // the chunk in flight: its size, format and the mapping arguments
    int text_format = 0, text_k = 0;
    uint32_t text_flags = 0;
};
#define KMB_SLOTS 6       // host chunks: being encoded / queued for the bus (a few, see the hybrid transport) / under the kernel
#define KMB_TEXT_SLOTS 3  // text and gz batches (large): one being staged, one on the bus + parsed, one under the kernel

struct kmb_mapper {
    kmb_index *index = nullptr;
    uint64_t n_counts = 0;
    uint32_t *counts = nullptr;
    bool own_counts = false;
    KmbLog log = {nullptr, nullptr, nullptr, 0, 0, 1, 0, KMB_LOG_BINS};  // hit log (grown on demand), its group tags and cursor
    int log_windows = 1;           // apply windows: ((n_counts - 1) >> win_shift) + 1
    bool dirty = false;            // the logs may hold hits that are not yet in the node counts
    uint64_t queries_since_flush = 0;
    int32_t max_freq = 1000;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    cudaStream_t parse_stream = nullptr;   // record parsing of chunk i runs beside the H2D copy / inflate of chunk i + 1
    cudaEvent_t text_on_device = nullptr;
    KmbStatus *d_status = nullptr;
    KmbStatus *h_status = nullptr;  // pinned: [0] what fetch_status reads back, [1] the constant initial state
    bool log_clean = true;          // nothing has been logged since the log was last emptied: no need to empty it again
    StageSlot slot[KMB_SLOTS];
    uint64_t host_bad = ~0ull;  // first invalid byte met by the host-side encoder since the last reset
    uint32_t *dtiles = nullptr;  // tile -> read table for in-place device input
    size_t dtiles_cap = 0;
    int next_slot = 0;
    int text_pending = -1;      // slot of the text chunk whose mapping kernel has not been launched yet
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed[2];  // event pairs around [0] mapping kernels, [1] apply passes
    size_t timed_used[2] = {0, 0};
};

static int text_finish_pending(kmb_mapper *m);
static int timed_begin(kmb_mapper *m, int klass = 0);
static int timed_end(kmb_mapper *m, int klass = 0);

static void slot_free_text(StageSlot &s) {
    cudaFree(s.text);
    cudaFree(s.tbases);
    cudaFree(s.toffsets);
    cudaFree(s.nl_pos);
    cudaFree(s.copies);
    cudaFree(s.blk_nl);
    cudaFree(s.blk_reads);
    cudaFree(s.blk_bases);
    cudaFree(s.scalars);
    cudaFree(s.d_result);
    if (s.h_result) cudaFreeHost(s.h_result);
    s.text = s.tbases = nullptr;
    s.toffsets = nullptr;
    s.nl_pos = s.blk_nl = s.blk_reads = nullptr;
    s.copies = nullptr;
    s.blk_bases = s.scalars = nullptr;
    s.d_result = s.h_result = nullptr;
    s.text_cap = s.line_cap = 0;
}

static void slot_free_gz(StageSlot &s) {
    if (s.h_gz) cudaFreeHost(s.h_gz);
    cudaFree(s.d_gz);
    if (s.h_members) cudaFreeHost(s.h_members);
    cudaFree(s.d_members);
    if (s.h_results) cudaFreeHost(s.h_results);
    cudaFree(s.d_results);
    if (s.h_crc) cudaFreeHost(s.h_crc);
    cudaFree(s.d_crc);
    if (s.h_windows) cudaFreeHost(s.h_windows);
    s.h_gz = s.d_gz = s.h_windows = nullptr;
    s.h_members = s.d_members = nullptr;
    s.h_results = s.d_results = nullptr;
    s.h_crc = s.d_crc = nullptr;
    s.h_gz_cap = s.d_gz_cap = s.members_cap = 0;
}

static void slot_free(StageSlot &s) {
    slot_free_text(s);
    slot_free_gz(s);
    if (s.h_text) cudaFreeHost(s.h_text);
    cudaFree(s.data);
    cudaFree(s.offsets);
    cudaFree(s.tiles);
    if (s.h_words) cudaFreeHost(s.h_words);
    if (s.h_off) cudaFreeHost(s.h_off);
    if (s.copied) cudaEventDestroy(s.copied);
    if (s.consumed) cudaEventDestroy(s.consumed);
    s = StageSlot();
}

extern "C" int kmb_mapper_destroy(kmb_mapper *m) {
    if (!m) return KMB_OK;
    DeviceGuard g(m->index->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->copy_stream) cudaStreamSynchronize(m->copy_stream);
    if (m->parse_stream) cudaStreamSynchronize(m->parse_stream);
    for (int i = 0; i < KMB_SLOTS; i++) slot_free(m->slot[i]);
    for (auto &v : m->timed)
        for (auto &pr : v) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
    cudaFree(m->dtiles);
    cudaFree(m->log.entries);
    cudaFree(m->log.tags);
    cudaFree(m->log.cursor);
    if (m->own_counts) cudaFree(m->counts);
    cudaFree(m->d_status);
    if (m->h_status) cudaFreeHost(m->h_status);
    if (m->own_stream) cudaStreamDestroy(m->own_stream);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    if (m->parse_stream) cudaStreamDestroy(m->parse_stream);
    if (m->text_on_device) cudaEventDestroy(m->text_on_device);
    cudaGetLastError();
    delete m;
    return KMB_OK;
}

// Asynchronous: the source is a pinned block that never changes after mapper creation, so no host wait is needed
// (a reset per chunk is the reference's calling pattern, command_line_interface.py:51: keep it off the host's clock).
static int status_reset(kmb_mapper *m) {
    m->host_bad = ~0ull;
    KMB_CUDA(cudaMemcpyAsync(m->d_status, &m->h_status[1], sizeof(KmbStatus), cudaMemcpyHostToDevice, m->stream));
    return KMB_OK;
}

static int launch_log_reset(kmb_mapper *m) {
    if (!m->log.entries || m->log_clean) return KMB_OK;
    const uint64_t groups = m->log.cap >> 5;
    const int grid = (int)std::min<uint64_t>((groups + 255) / 256, (uint64_t)m->index->info.sms * 4);
    kmb_log_reset_kernel<<<std::max(grid, 1), 256, 0, m->stream>>>(m->log);
    kmb_log_rewind_kernel<<<1, 1, 0, m->stream>>>(m->log);
    g_launches += 2;
    KMB_CUDA(cudaGetLastError());
    m->log_clean = true;
    return KMB_OK;
}

// Play the hit log into the node counts (asynchronous on the mapper's stream), one node range after the
// other, and empty it.
static int launch_flush(kmb_mapper *m) {
    const kmb_index *ix = m->index;
    if (m->log.entries) {
        // one launch, blockIdx.y = node window: the windows are worked off in dispatch order
        const uint64_t last = m->n_counts ? m->n_counts - 1 : 0;
        const uint32_t want = g_opt.apply_window_log2 > 0 ? (uint32_t)std::min<int64_t>(std::max<int64_t>(g_opt.apply_window_log2, 10), 31)
                                                         : (m->log.bin_shift <= 24u ? 23u : 24u);
        m->log.win_shift = std::min(m->log.bin_shift, want);
        m->log_windows = (int)((last >> m->log.win_shift) + 1);
        KMB_TRY(timed_begin(m, 1));
        kmb_log_apply_kernel<<<dim3((unsigned)ix->info.sms * (unsigned)std::min<int64_t>(std::max<int64_t>(g_opt.apply_slabs_per_sm, 1), 256), (unsigned)m->log_windows), 256, 0, m->stream>>>(m->log, m->counts);
        g_launches++;
        KMB_CUDA(cudaGetLastError());
        KMB_TRY(timed_end(m, 1));
        KMB_TRY(launch_log_reset(m));
    }
    m->dirty = false;
    m->queries_since_flush = 0;
    return KMB_OK;
}

// Size the log for a launch of n_queries look-ups: room for one hit per four queries (twice the hit rate of
// the benchmark shapes), between 2^23 and log_max_entries ids.  A log that runs full is not an error: the
// kernels then reduce directly onto the counts.
static bool may_use_read_table(const kmb_index *ix);
static int ensure_log(kmb_mapper *m, uint64_t n_queries, bool for_read_table) {
    if (!m->log.entries && m->n_counts <= (uint64_t)std::max<int64_t>(g_opt.direct_counts_max_nodes, 0)) {
        m->log.cap = 0;   // small count array: the kernels reduce onto it directly (kmb_emit)
        return KMB_OK;
    }
    uint64_t want = std::max<uint64_t>(n_queries / 4, 1ull << 23);
    want = std::min<uint64_t>(want, (uint64_t)std::max<int64_t>(g_opt.log_max_entries, 1 << 10));
    want = (want + 4095) & ~4095ull;  // whole 128-group blocks: the apply pass reads four tags per 4-byte load
    if (!m->log.cursor) {
        KMB_CUDA(cudaMalloc(&m->log.cursor, sizeof(unsigned long long)));
        KMB_CUDA(cudaMemsetAsync(m->log.cursor, 0, sizeof(unsigned long long), m->stream));
        // The log is laid out for the kernel that asks for it first: KMB_MZ_LOG_BINS node ranges for the read-path
        // kernel (12 stacks of 35 ids fit its shared memory: at 400 M nodes a range is then two apply windows wide
        // instead of four and the apply pass reads the log twice instead of four times, 6.66 -> 4.67 ms per 6 G k-mers
        // of config 3), KMB_LOG_BINS for the key-addressed kernels.  A kernel that meets a log with more ranges than
        // it has stacks reduces the ids of the ranges beyond them directly (kmb_emit).
        const uint32_t n_bins = for_read_table ? KMB_MZ_LOG_BINS : KMB_LOG_BINS;
        uint32_t shift = 0;
        while (((m->n_counts ? m->n_counts - 1 : 0) >> shift) >= n_bins) shift++;
        m->log.bin_shift = shift;
        m->log.win_shift = shift;
        m->log.n_bins = n_bins;
    }
    if (want > m->log.cap) {
        if (m->dirty) KMB_TRY(launch_flush(m));
        KMB_CUDA(cudaStreamSynchronize(m->stream));
        cudaFree(m->log.entries);
        cudaFree(m->log.tags);
        m->log.entries = nullptr;
        m->log.tags = nullptr;
        m->log.cap = 0;
        KMB_CUDA(cudaMalloc(&m->log.entries, (size_t)want * sizeof(uint32_t)));
        KMB_CUDA(cudaMalloc(&m->log.tags, (size_t)(want / 32)));
        KMB_CUDA(cudaMemsetAsync(m->log.tags, KMB_LOG_NO_BIN, (size_t)(want / 32), m->stream));
        KMB_CUDA(cudaMemsetAsync(m->log.cursor, 0, sizeof(unsigned long long), m->stream));
        m->log.cap = want;
    } else if (m->dirty && m->queries_since_flush + n_queries > 4 * m->log.cap) {
        KMB_TRY(launch_flush(m));  // make room before the log overflows into direct reductions
    }
    // big launches reserve log space eight 128-byte groups at a time (one atomic per 256 hits); small ones group by
    // group, so that the unused tail of the reservations stays small next to the hits
    m->log.chunk_groups = n_queries >= (256ull << 20) ? 8u : 1u;
    m->queries_since_flush += n_queries;
    return KMB_OK;
}

static int set_l2_window(kmb_mapper *m) {
    if (g_opt.l2_fetch_granularity > 0 &&
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g_opt.l2_fetch_granularity) != cudaSuccess)
        cudaGetLastError();
    // Hint: keep the filter in the persisting part of L2 for kernels on this stream.  Purely a
    // performance hint; failure is not an error.
    kmb_index *ix = m->index;
    if (!g_opt.l2_persist) {
        if (cudaCtxResetPersistingL2Cache() != cudaSuccess) cudaGetLastError();
        return KMB_OK;
    }
    if (!ix->filter_on || ix->info.max_persist_l2 <= 0) return KMB_OK;
    size_t want = std::min<size_t>(ix->filter_bytes, (size_t)ix->info.max_persist_l2);
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) {
        cudaGetLastError();
        return KMB_OK;
    }
    int max_win = 0;
    cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, ix->device);
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    v.accessPolicyWindow.base_ptr = ix->filter;
    v.accessPolicyWindow.num_bytes = std::min<size_t>(ix->filter_bytes, (size_t)std::max(max_win, 0));
    v.accessPolicyWindow.hitRatio = std::min(1.0f, (float)want / (float)std::max<size_t>(v.accessPolicyWindow.num_bytes, 1));
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    if (cudaStreamSetAttribute(m->stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
    return KMB_OK;
}

extern "C" int kmb_mapper_create(kmb_index *index, uint64_t n_counts, uint32_t *counts_device,
                                 int max_index_lookup_frequency, kmb_mapper **out) {
    if (!out) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_create: out is null");
    *out = nullptr;
    if (!index) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_create: null index");
    if ((int64_t)n_counts <= index->max_node)
        return kmb_fail(KMB_ERR_BAD_ARG,
                        "kmb_mapper_create: n_counts=%llu but the index holds node id %lld (max_node_id+1 counters are "
                        "needed; the reference would write out of bounds, mapper.pyx:37,68)",
                        (unsigned long long)n_counts, (long long)index->max_node);
    KMB_ON_DEVICE(index->device);
    kmb_mapper *m = new (std::nothrow) kmb_mapper;
    if (!m) return kmb_fail(KMB_ERR_NOMEM, "out of host memory");
    m->index = index;
    struct Cleanup {
        kmb_mapper *m;
        ~Cleanup() {
            if (m) kmb_mapper_destroy(m);
        }
    } cleanup{m};
    m->n_counts = n_counts;
    m->max_freq = max_index_lookup_frequency;  // a C int like the reference's (mapper.pyx:19,64)
    KMB_CUDA(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
    KMB_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    KMB_CUDA(cudaStreamCreateWithFlags(&m->parse_stream, cudaStreamNonBlocking));
    KMB_CUDA(cudaEventCreateWithFlags(&m->text_on_device, cudaEventDisableTiming));
    m->stream = m->own_stream;
    if (counts_device) {
        bool dev;
        KMB_TRY(ptr_on_device(counts_device, index->device, &dev));
        if (!dev) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_create: counts_device is not device memory");
        m->counts = counts_device;
    } else {
        KMB_CUDA(cudaMalloc(&m->counts, std::max<uint64_t>(n_counts, 1) * 4));
        m->own_counts = true;
        KMB_CUDA(cudaMemsetAsync(m->counts, 0, n_counts * 4, m->stream));
    }
    KMB_CUDA(cudaMalloc(&m->d_status, sizeof(KmbStatus)));
    KMB_CUDA(cudaMallocHost(&m->h_status, 2 * sizeof(KmbStatus)));
    memset(m->h_status, 0, 2 * sizeof(KmbStatus));
    m->h_status[0].first_bad_offset = m->h_status[1].first_bad_offset = ~0ull;
    m->h_status[0].max_node = m->h_status[1].max_node = -1;
    KMB_TRY(status_reset(m));
    for (int i = 0; i < KMB_SLOTS; i++) {
        KMB_CUDA(cudaEventCreateWithFlags(&m->slot[i].copied, cudaEventDisableTiming));
        KMB_CUDA(cudaEventCreateWithFlags(&m->slot[i].consumed, cudaEventDisableTiming));
    }
    set_l2_window(m);
    cleanup.m = nullptr;
    *out = m;
    return KMB_OK;
}

extern "C" int kmb_mapper_set_stream(kmb_mapper *m, void *cuda_stream) {
    if (!m) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_set_stream: null mapper");
    KMB_ON_DEVICE(m->index->device);
    KMB_TRY(text_finish_pending(m));
    KMB_CUDA(cudaStreamSynchronize(m->stream));
    m->stream = cuda_stream ? (cudaStream_t)cuda_stream : m->own_stream;
    set_l2_window(m);
    return KMB_OK;
}

static KmbProbe make_probe(const kmb_mapper *m) {
    const kmb_index *ix = m->index;
    KmbProbe P;
    P.lines = ix->lines;
    P.filter = ix->filter_on ? ix->filter : nullptr;
    P.addr = ix->addr;
    P.policies = (uint32_t)((g_opt.policy_filter & 3) | ((g_opt.policy_line & 3) << 2) | ((g_opt.ablate & 15) << 8));
    P.max_freq = m->max_freq;
    P.counts = m->counts;
    P.log = m->log;
    P.n_lines = ix->n_lines;
    P.n_counts = m->n_counts;
    return P;
}

static int pick_u() {
    int64_t u = g_opt.gathers_in_flight;
    return u <= 2 ? 2 : 4;
}

// ---- kernel dispatch (template instantiation table) ------------------------------------------------
typedef void (*MapReadsFn)(const uint8_t *, uint64_t, uint64_t, KmbReads, int, uint32_t, KmbProbe, KmbStatus *);
typedef void (*MapKmersFn)(const uint64_t *, uint64_t, int, KmbProbe, KmbStatus *);

template <int U>
static MapReadsFn map_reads_fn_u(bool filt, bool rc) {
    if (filt) return rc ? kmb_map_reads_kernel<U, true, true> : kmb_map_reads_kernel<U, true, false>;
    return rc ? kmb_map_reads_kernel<U, false, true> : kmb_map_reads_kernel<U, false, false>;
}
template <int U>
static MapKmersFn map_kmers_fn_u(bool filt, bool rc) {
    if (filt) return rc ? kmb_map_kmers_kernel<U, true, true> : kmb_map_kmers_kernel<U, true, false>;
    return rc ? kmb_map_kmers_kernel<U, false, true> : kmb_map_kmers_kernel<U, false, false>;
}
struct MapKernel {
    const void *fn;
    int u;
    size_t smem;
};
static MapKernel pick_map_kernel(bool reads, bool filt, bool rc) {
    MapKernel mk;
    mk.u = pick_u();
    if (mk.u == 2) {
        mk.fn = reads ? (const void *)map_reads_fn_u<2>(filt, rc) : (const void *)map_kmers_fn_u<2>(filt, rc);
        mk.smem = KMB_MAP_SMEM_BYTES(2);
    } else {
        mk.fn = reads ? (const void *)map_reads_fn_u<4>(filt, rc) : (const void *)map_kmers_fn_u<4>(filt, rc);
        mk.smem = KMB_MAP_SMEM_BYTES(4);
    }
    return mk;
}

static int resident_blocks(const MapKernel &mk, int64_t opt, int *out) {
    int b = 0;
    KMB_CUDA(cudaFuncSetAttribute(mk.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mk.smem));
    if (g_opt.map_carveout >= 0) KMB_CUDA(cudaFuncSetAttribute(mk.fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)std::min<int64_t>(g_opt.map_carveout, 100)));
    KMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, mk.fn, KMB_MAP_THREADS, mk.smem));
    if (b < 1) b = 1;
    if (opt > 0) b = (int)std::min<int64_t>(opt, 32);
    *out = b;
    return KMB_OK;
}

static int timed_begin(kmb_mapper *m, int klass) {
    if (!g_opt.time_kernels) return KMB_OK;
    auto &v = m->timed[klass];
    if (m->timed_used[klass] == v.size()) {
        cudaEvent_t a, b;
        KMB_CUDA(cudaEventCreate(&a));
        KMB_CUDA(cudaEventCreate(&b));
        v.emplace_back(a, b);
    }
    KMB_CUDA(cudaEventRecord(v[m->timed_used[klass]].first, m->stream));
    return KMB_OK;
}
static int timed_end(kmb_mapper *m, int klass) {
    if (!g_opt.time_kernels) return KMB_OK;
    KMB_CUDA(cudaEventRecord(m->timed[klass][m->timed_used[klass]].second, m->stream));
    m->timed_used[klass]++;
    return KMB_OK;
}

// Build the read-path table of an index for window length k (once; mappers on any stream share it).
static int ensure_read_table(kmb_index *ix, int k) {
    std::lock_guard<std::mutex> lock(ix->mz_mu);
    if (ix->mz_k == k || ix->mz_k < 0) return KMB_OK;
    if (ix->mz_k != 0) return kmb_fail(KMB_ERR_BAD_ARG, "read-path table already built for k=%d", ix->mz_k);
    cudaStream_t s = 0;  // one-off, synchronous
    const uint64_t per100 = (uint64_t)std::min<int64_t>(std::max<int64_t>(g_opt.read_table_buckets_per_100_entries, 25), 800);
    const uint64_t n_buckets = std::min<uint64_t>(std::max<uint64_t>(ix->n_live * per100 / 100, 1024), (1ull << 30) - 1);
    KmbAddr addr;
    memset(&addr, 0, sizeof(addr));
    addr.n_main = (uint32_t)n_buckets;
    // filter over the distinct minimizers (at most one per entry): same sizing rule as the key filter
    const uint64_t keys = std::max<uint64_t>(ix->n_live, 1);
    uint64_t filter_words = std::min<uint64_t>((uint64_t)std::max<int64_t>(g_opt.filter_l2_budget_bytes, 4) / 4,
                                               std::max<uint64_t>(keys / 2, 1));
    const double bits_per_key = 32.0 * (double)filter_words / (double)keys;
    const bool want_filter = g_opt.use_filter == 1 || (g_opt.use_filter < 0 && bits_per_key >= 0.62);
    if (!want_filter) filter_words = 0;
    addr.n_filter_words = (uint32_t)filter_words;
    addr.n_probes = filter_probes(bits_per_key);
    DevBuf<uint32_t> filter, fill, lines;
    // the table is an accelerator, not a requirement: if the device has no room for it the key-addressed path serves
    if (filter.alloc((size_t)std::max<uint64_t>(filter_words, 1)) != KMB_OK || fill.alloc((size_t)n_buckets) != KMB_OK) {
        cudaGetLastError();
        ix->mz_k = -1;
        return KMB_OK;
    }
    KMB_CUDA(cudaMemsetAsync(filter.p, 0, (size_t)std::max<uint64_t>(filter_words, 1) * 4, s));
    KMB_CUDA(cudaMemsetAsync(fill.p, 0, (size_t)n_buckets * 4, s));
    DevBuf<KmbStatus> d_status;
    KMB_TRY(d_status.alloc(1));
    KmbStatus hs;
    memset(&hs, 0, sizeof(hs));
    KMB_CUDA(cudaMemcpyAsync(d_status.p, &hs, sizeof(hs), cudaMemcpyHostToDevice, s));
    const int sms = ix->info.sms;
    kmb_mz_build_count<<<grid_for(ix->n_lines, 256, sms), 256, 0, s>>>(ix->lines, ix->n_lines, k, addr, fill.p, filter.p, d_status.p);
    kmb_mz_build_plan<false><<<grid_for(n_buckets, 256, sms), 256, 0, s>>>(fill.p, n_buckets, nullptr, 0, d_status.p);
    g_launches += 2;
    KMB_CUDA(cudaGetLastError());
    KMB_CUDA(cudaMemcpyAsync(&hs, d_status.p, sizeof(hs), cudaMemcpyDeviceToHost, s));
    KMB_CUDA(cudaStreamSynchronize(s));
    const uint64_t n_lines = 2 * n_buckets + hs.pool_lines;
    if (n_lines >= (1ull << 31) || (hs.index_flags & 4u)) {
        ix->mz_k = -1;  // cannot be built for this index: the key-addressed path serves it
        return KMB_OK;
    }
    if (lines.alloc((size_t)n_lines * KMB_LINE_WORDS) != KMB_OK) {
        cudaGetLastError();
        ix->mz_k = -1;
        return KMB_OK;
    }
    KMB_CUDA(cudaMemsetAsync(lines.p, 0, (size_t)n_lines * KMB_LINE_BYTES, s));
    KMB_CUDA(cudaMemsetAsync(&d_status.p->pool_lines, 0, sizeof(unsigned int), s));
    kmb_mz_build_plan<true><<<grid_for(n_buckets, 256, sms), 256, 0, s>>>(fill.p, n_buckets, lines.p, n_lines, d_status.p);
    kmb_mz_build_scatter<<<grid_for(ix->n_lines, 256, sms), 256, 0, s>>>(ix->lines, ix->n_lines, k, addr, fill.p, lines.p, n_lines);
    g_launches += 2;
    KMB_CUDA(cudaGetLastError());
    KMB_CUDA(cudaStreamSynchronize(s));
    ix->mz_addr = addr;
    ix->mz_n_lines = n_lines;
    ix->mz_entries = hs.n_live_entries;
    ix->mz_filter_bytes = want_filter ? (size_t)filter_words * 4 : 0;
    ix->mz_lines = lines.release();
    if (want_filter) ix->mz_filter = filter.release();
    ix->device_bytes += n_lines * KMB_LINE_BYTES + ix->mz_filter_bytes;
    ix->mz_k = k;
    return KMB_OK;
}

// whether some launch on this index could take the read-path kernel (decides the shape of a mapper's hit log)
static bool may_use_read_table(const kmb_index *ix) {
    if (g_opt.read_table == 0 || ix->mz_k < 0) return false;
    if (g_opt.read_table > 0) return true;
    // Automatic: every index that is not small.  Until the end of round 2 the table was only chosen behind a thin key
    // filter (< 2.5 bits per key: config 3); with 16-base minimizers, the one-multiplication order and the cheaper
    // encode the read-path kernel also wins where the filter is good: config 2 48.7 ms per step against 50.3 (123.3 /
    // 119.3 GK/s), config 5 39.4 against 48.6 (190 / 154).  Small indexes (config 1) fit the L2 and stay key-addressed.
    return ix->n_live >= (uint64_t)std::max<int64_t>(g_opt.read_table_min_entries, 0);
}
static bool use_read_table(const kmb_index *ix, int k, uint32_t flags) {
    if (k != KMB_MZ_K || (flags & KMB_FLAG_REVCOMP)) return false;
    return may_use_read_table(ix);
}

// launch the tile -> read table + the fused kernel over one device-resident batch.  packed: d_bases is the 2-bit
// stream of kmb_hostpack.cpp and d_offsets are uint32 offsets relative to the batch (else int64, relative to base0)
static int launch_map_reads(kmb_mapper *m, const uint8_t *d_bases, uint64_t n_bases, uint64_t base0, const void *d_offsets,
                            uint64_t n_reads, uint32_t *d_tiles, int k, uint32_t flags, bool packed) {
    if (n_bases == 0) return KMB_OK;
    const kmb_index *ix = m->index;
    const uint64_t n_wtiles = (n_bases + KMB_WTILE_POS - 1) / KMB_WTILE_POS;
    KmbReads R;
    R.offsets = d_offsets;
    R.tile_read = d_tiles;
    R.n_reads = n_reads;
    R.base0 = packed ? 0 : (int64_t)base0;
    R.off32 = packed ? 1u : 0u;
    kmb_tile_reads_kernel<<<grid_for(n_wtiles, 256, ix->info.sms), 256, 0, m->stream>>>(R, n_bases, n_wtiles, d_tiles, m->d_status);
    g_launches++;
    // every window, both strands when asked: the number of look-ups this launch can make
    KMB_TRY(ensure_log(m, ((flags & KMB_FLAG_REVCOMP) ? 2 : 1) * n_bases, use_read_table(ix, k, flags)));
    KmbProbe P = make_probe(m);
    const uint32_t in_mode = ((flags & KMB_FLAG_NO_N_TO_A) ? 0u : KMB_IN_N_TO_A) | (packed ? KMB_IN_PACKED : 0u);
    // (a log shaped for the key-addressed kernels -- the option was switched on after it was made -- keeps them)
    const bool want_mz = use_read_table(ix, k, flags) && m->log.n_bins <= KMB_MZ_LOG_BINS;
    if (want_mz) KMB_TRY(ensure_read_table(m->index, k));
    if (want_mz && ix->mz_k == k) {
        g_last_reads_kernel = 1;
        const KmbProbe Pkey = P;  // the key-addressed sectors: where the rare tile that overflows the run table goes
        P.lines = ix->mz_lines;
        P.filter = ix->mz_filter;
        P.addr = ix->mz_addr;
        P.n_lines = ix->mz_n_lines;
        typedef void (*MzFn)(const uint8_t *, uint64_t, uint64_t, KmbReads, uint32_t, KmbProbe, KmbProbe, KmbStatus *);
        MzFn fn = P.filter ? kmb_map_reads_mz_kernel<true> : kmb_map_reads_mz_kernel<false>;
        KMB_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KMB_MZ_SMEM_BYTES));
        int per_sm = 0;
        KMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)fn, KMB_MZ_THREADS, KMB_MZ_SMEM_BYTES));
        if (per_sm < 1) per_sm = 1;
        if (g_opt.map_reads_blocks_per_sm > 0) per_sm = (int)std::min<int64_t>(g_opt.map_reads_blocks_per_sm, per_sm);
        const uint64_t n_cta_tiles = (n_bases + (uint64_t)KMB_WTILE_POS * (KMB_MZ_THREADS / 32) - 1) / ((uint64_t)KMB_WTILE_POS * (KMB_MZ_THREADS / 32));
        int grid = (int)std::min<uint64_t>(n_cta_tiles, (uint64_t)ix->info.sms * per_sm);
        m->dirty = true;
        m->log_clean = false;
        KMB_TRY(timed_begin(m));
        fn<<<grid, KMB_MZ_THREADS, KMB_MZ_SMEM_BYTES, m->stream>>>(d_bases, n_bases, base0, R, in_mode, P, Pkey, m->d_status);
        g_launches++;
        KMB_CUDA(cudaGetLastError());
        KMB_TRY(timed_end(m));
        return KMB_OK;
    }
    g_last_reads_kernel = 0;
    const MapKernel mk = pick_map_kernel(true, P.filter != nullptr, (flags & KMB_FLAG_REVCOMP) != 0);
    int per_sm;
    KMB_TRY(resident_blocks(mk, g_opt.map_reads_blocks_per_sm, &per_sm));
    const uint64_t n_cta_tiles = (n_wtiles + (KMB_MAP_THREADS / 32) - 1) / (KMB_MAP_THREADS / 32);  // one 1024-window tile per warp
    int grid = (int)std::min<uint64_t>(n_cta_tiles, (uint64_t)ix->info.sms * per_sm);
    m->dirty = true;
    m->log_clean = false;
    KMB_TRY(timed_begin(m));
    ((MapReadsFn)mk.fn)<<<grid, KMB_MAP_THREADS, mk.smem, m->stream>>>(d_bases, n_bases, base0, R, k, in_mode, P, m->d_status);
    g_launches++;
    KMB_CUDA(cudaGetLastError());
    KMB_TRY(timed_end(m));
    return KMB_OK;
}

static int launch_map_kmers(kmb_mapper *m, const uint64_t *d_kmers, uint64_t n, uint32_t flags, int k) {
    if (n == 0) return KMB_OK;
    const kmb_index *ix = m->index;
    bool rc = (flags & KMB_FLAG_REVCOMP) != 0;
    KMB_TRY(ensure_log(m, (rc ? 2 : 1) * n, false));
    KmbProbe P = make_probe(m);
    m->dirty = true;
    m->log_clean = false;
    KMB_TRY(timed_begin(m));
    if (g_opt.probe_variant == 0) {
        int grid = grid_for(n, 256, ix->info.sms, 16);
        if (rc) kmb_map_kmers_simple_kernel<true><<<grid, 256, 0, m->stream>>>(d_kmers, n, k, P, m->d_status);
        else kmb_map_kmers_simple_kernel<false><<<grid, 256, 0, m->stream>>>(d_kmers, n, k, P, m->d_status);
    } else {
        const MapKernel mk = pick_map_kernel(false, P.filter != nullptr, rc);
        int per_sm;
        KMB_TRY(resident_blocks(mk, g_opt.map_kmers_blocks_per_sm, &per_sm));
        uint64_t n_blocks = (n + (uint64_t)KMB_MAP_THREADS * mk.u - 1) / ((uint64_t)KMB_MAP_THREADS * mk.u);
        int grid = (int)std::min<uint64_t>(n_blocks, (uint64_t)ix->info.sms * per_sm);
        ((MapKmersFn)mk.fn)<<<grid, KMB_MAP_THREADS, mk.smem, m->stream>>>(d_kmers, n, k, P, m->d_status);
    }
    g_launches++;
    KMB_CUDA(cudaGetLastError());
    KMB_TRY(timed_end(m));
    return KMB_OK;
}

static int slot_reserve(StageSlot &s, size_t data_bytes, size_t n_offsets, size_t n_tiles) {
    if (data_bytes > s.data_cap) {
        cudaFree(s.data);
        s.data = nullptr;
        s.data_cap = 0;
        size_t cap = (data_bytes + 255) & ~(size_t)255;
        KMB_CUDA(cudaMalloc(&s.data, cap));
        s.data_cap = cap;
    }
    if (n_offsets > s.off_cap) {
        cudaFree(s.offsets);
        s.offsets = nullptr;
        s.off_cap = 0;
        size_t cap = n_offsets + n_offsets / 4 + 64;
        KMB_CUDA(cudaMalloc(&s.offsets, cap * 8));
        s.off_cap = cap;
    }
    if (n_tiles > s.tiles_cap) {
        cudaFree(s.tiles);
        s.tiles = nullptr;
        s.tiles_cap = 0;
        size_t cap = n_tiles + n_tiles / 4 + 64;
        KMB_CUDA(cudaMalloc(&s.tiles, cap * 4));
        s.tiles_cap = cap;
    }
    return KMB_OK;
}

static int slot_reserve_host(StageSlot &s, size_t n_words, size_t n_offsets) {
    if (n_words > s.h_words_cap) {
        if (s.h_words) cudaFreeHost(s.h_words);
        s.h_words = nullptr;
        s.h_words_cap = 0;
        size_t cap = n_words + n_words / 8 + 64;
        KMB_CUDA(cudaMallocHost(&s.h_words, cap * 4));
        s.h_words_cap = cap;
    }
    if (n_offsets > s.h_off_cap) {
        if (s.h_off) cudaFreeHost(s.h_off);
        s.h_off = nullptr;
        s.h_off_cap = 0;
        size_t cap = n_offsets + n_offsets / 4 + 64;
        KMB_CUDA(cudaMallocHost(&s.h_off, cap * 4));
        s.h_off_cap = cap;
    }
    return KMB_OK;
}

static int check_k(int k) {
    if (k <= 0 || k >= 32) return kmb_fail(KMB_ERR_BAD_ARG, "k=%d outside 1..31", k);
    return KMB_OK;
}

extern "C" int kmb_mapper_map_reads(kmb_mapper *m, const uint8_t *bases, uint64_t n_bases, const int64_t *offsets,
                                    uint64_t n_reads, int k, uint32_t flags) {
    if (!m) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_reads: null mapper");
    KMB_TRY(check_k(k));
    if (n_bases == 0 || n_reads == 0) return KMB_OK;
    if (!bases || !offsets) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_reads: null buffer");
    KMB_ON_DEVICE(m->index->device);
    KMB_TRY(text_finish_pending(m));
    bool dev_b, dev_o, pinned_b;
    KMB_TRY(ptr_on_device(bases, m->index->device, &dev_b, &pinned_b));
    KMB_TRY(ptr_on_device(offsets, m->index->device, &dev_o));
    if (dev_b != dev_o)
        return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_reads: bases and offsets must both be host or both be device buffers");
    if (dev_b) {
        if (((uintptr_t)bases & 15u) != 0)
            return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_reads: device bases buffer must be 16-byte aligned");
        size_t words = (size_t)(n_bases / KMB_WTILE_POS + 1);
        if (words > m->dtiles_cap) {
            KMB_CUDA(cudaStreamSynchronize(m->stream));
            cudaFree(m->dtiles);
            m->dtiles = nullptr;
            m->dtiles_cap = 0;
            KMB_CUDA(cudaMalloc(&m->dtiles, (words + words / 8 + 64) * 4));
            m->dtiles_cap = words + words / 8 + 64;
        }
        return launch_map_reads(m, bases, n_bases, 0, offsets, n_reads, m->dtiles, k, flags, false);
    }
    // ---- host input: whole reads per chunk, double-buffered H2D on the copy stream
    if (offsets[0] != 0 || (uint64_t)offsets[n_reads] != n_bases)
        return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_reads: offsets[0] must be 0 and offsets[n_reads] must equal n_bases");
    const uint64_t chunk = (uint64_t)g_opt.chunk_bytes;
    const int pack_threads = g_opt.host_threads > 0 ? (int)g_opt.host_threads : kmb_host_cpus();
    // transport of a host chunk (see KmbOptions::host_pack): 0 ASCII, 1 packed, 2 hybrid
    int mode;
    if (g_opt.host_pack >= 0) mode = (int)std::min<int64_t>(g_opt.host_pack, 2);
    else if (!pinned_b) mode = pack_threads >= 2 ? 1 : 0;
    // a pinned source: hybrid when this process has the host to itself.  When several ranks share the host's memory
    // system the cores' streaming reads slow every rank's DMA down (measured, 8 ranks x 4 threads: 557 ms per step
    // hybrid against 380 ms ASCII; 4 ranks: 300 / 300; 2 ranks: 162 / 156), so there the bases go as they are.
    else mode = (pack_threads >= 4 && g_opt.host_ranks <= 1) ? 2 : 0;
    if (mode == 2 && !pinned_b) mode = 1;
    const uint64_t backlog_limit = g_opt.host_hybrid_backlog_bytes > 0 ? (uint64_t)g_opt.host_hybrid_backlog_bytes : chunk;
    uint64_t r0 = 0;
    while (r0 < n_reads) {
        // largest r1 with offsets[r1] - offsets[r0] <= chunk (at least one read)
        const int64_t *lo = offsets + r0 + 1, *hi = offsets + n_reads + 1;
        const int64_t *it = std::upper_bound(lo, hi, (int64_t)(offsets[r0] + chunk));
        uint64_t r1 = (uint64_t)(it - offsets) - 1;
        if (r1 <= r0) r1 = r0 + 1;
        if (offsets[r1] < offsets[r0]) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_reads: offsets must be non-decreasing");
        const uint64_t b0 = (uint64_t)offsets[r0], nb = (uint64_t)offsets[r1] - b0, nr = r1 - r0;
        if (nb) {
            StageSlot &s = m->slot[m->next_slot];
            m->next_slot = (m->next_slot + 1) % KMB_SLOTS;
            if (s.used) KMB_CUDA(cudaEventSynchronize(s.consumed));  // also makes re-allocation safe
            s.bus_bytes = 0;
            bool packed = mode == 1;
            if (mode == 2) {
                // bytes queued for the bus and not yet across: the bus must never run dry, the cores take the rest
                uint64_t waiting = 0;
                for (int i = 0; i < KMB_SLOTS; i++)
                    if (m->slot[i].bus_bytes) {
                        if (cudaEventQuery(m->slot[i].copied) == cudaErrorNotReady) waiting += m->slot[i].bus_bytes;
                        else m->slot[i].bus_bytes = 0;
                    }
                (void)cudaGetLastError();
                packed = waiting >= backlog_limit;
            }
            packed = packed && nb < (1ull << 32);
            if (packed) {
                // encode on the CPU into pinned staging while the previous chunks are on the bus / under the kernel
                const size_t n_words = (size_t)kmb_packed_words(nb);
                KMB_TRY(slot_reserve(s, n_words * 4, (nr + 2) / 2, nb / KMB_WTILE_POS + 1));
                KMB_TRY(slot_reserve_host(s, n_words, nr + 1));
                kmb_host_pack_streaming(g_opt.host_pack_streaming != 0);
                const uint64_t bad = kmb_host_pack(bases + b0, nb, !(flags & KMB_FLAG_NO_N_TO_A), pack_threads, s.h_words);
                if (bad != ~0ull) m->host_bad = std::min(m->host_bad, b0 + bad);
                kmb_host_rel_offsets(offsets + r0, nr + 1, (int64_t)b0, pack_threads, s.h_off);
                KMB_CUDA(cudaMemcpyAsync(s.data, s.h_words, n_words * 4, cudaMemcpyHostToDevice, m->copy_stream));
                KMB_CUDA(cudaMemcpyAsync(s.offsets, s.h_off, (nr + 1) * 4, cudaMemcpyHostToDevice, m->copy_stream));
                s.bus_bytes = n_words * 4 + (nr + 1) * 4;
                g_host_chunks_packed++;
            } else {
                KMB_TRY(slot_reserve(s, nb + 16, nr + 1, nb / KMB_WTILE_POS + 1));
                KMB_CUDA(cudaMemcpyAsync(s.data, bases + b0, nb, cudaMemcpyHostToDevice, m->copy_stream));
                KMB_CUDA(cudaMemcpyAsync(s.offsets, offsets + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, m->copy_stream));
                s.bus_bytes = nb + (nr + 1) * 8;
                g_host_chunks_ascii++;
            }
            g_h2d_bytes += s.bus_bytes;
            KMB_CUDA(cudaEventRecord(s.copied, m->copy_stream));
            KMB_CUDA(cudaStreamWaitEvent(m->stream, s.copied, 0));
            KMB_TRY(launch_map_reads(m, s.data, nb, b0, s.offsets, nr, s.tiles, k, flags, packed));
            KMB_CUDA(cudaEventRecord(s.consumed, m->stream));
            s.used = true;
        }
        r0 = r1;
    }
    // the caller may reuse its host buffers as soon as we return
    KMB_CUDA(cudaStreamSynchronize(m->copy_stream));
    return KMB_OK;
}

// ------------------------------------------------------------------------------------------------
// raw FASTA / FASTQ text in, counts out: record parsing on the device (kmb_textparse.cuh)
// ------------------------------------------------------------------------------------------------

static int slot_reserve_text(StageSlot &s, size_t n_text, size_t n_lines_cap) {
    if (n_text > s.text_cap || n_lines_cap > s.line_cap) {
        slot_free_text(s);
        const size_t tcap = ((std::max(n_text, s.text_cap) + (n_text >> 3)) + 4095) & ~(size_t)4095;
        const size_t lcap = std::max(n_lines_cap, s.line_cap) + 1024;
        KMB_CUDA(cudaMalloc(&s.text, tcap + 16));
        KMB_CUDA(cudaMalloc(&s.tbases, tcap + 32));
        KMB_CUDA(cudaMalloc(&s.toffsets, (lcap + 2) * sizeof(int64_t)));
        KMB_CUDA(cudaMalloc(&s.nl_pos, (lcap + 2) * sizeof(uint32_t)));
        KMB_CUDA(cudaMalloc(&s.copies, (lcap + 2) * sizeof(KmbTextCopy)));
        KMB_CUDA(cudaMalloc(&s.blk_nl, (tcap / KMB_TP_BLOCK_BYTES + 2) * sizeof(uint32_t)));
        KMB_CUDA(cudaMalloc(&s.blk_reads, (lcap / KMB_TP_LINES_PER_BLOCK + 2) * sizeof(uint32_t)));
        KMB_CUDA(cudaMalloc(&s.blk_bases, (lcap / KMB_TP_LINES_PER_BLOCK + 2) * sizeof(unsigned long long)));
        KMB_CUDA(cudaMalloc(&s.scalars, 4 * sizeof(unsigned long long)));
        KMB_CUDA(cudaMalloc(&s.d_result, sizeof(KmbTextResult)));
        KMB_CUDA(cudaMallocHost(&s.h_result, sizeof(KmbTextResult)));
        s.text_cap = tcap;
        s.line_cap = lcap;
    }
    return KMB_OK;
}

// the six parse kernels over s.text[0, n_text) on `st`; result copied to s.h_result (asynchronously)
static int launch_text_parse(int sms, StageSlot &s, uint64_t n_text, int format, cudaStream_t st, const uint8_t *text = nullptr) {
    if (!text) text = s.text;
    const unsigned n_blocks = (unsigned)((n_text + KMB_TP_BLOCK_BYTES - 1) / KMB_TP_BLOCK_BYTES);
    const uint64_t line_cap = s.line_cap;
    const unsigned n_line_blocks = (unsigned)((line_cap + 1 + KMB_TP_LINES_PER_BLOCK - 1) / KMB_TP_LINES_PER_BLOCK);
    KMB_CUDA(cudaMemsetAsync(s.d_result, 0, sizeof(KmbTextResult), st));
    kmb_tp_count_newlines<<<n_blocks, 256, 0, st>>>(text, n_text, s.blk_nl);
    kmb_tp_scan<<<1, 1024, 0, st>>>(s.blk_nl, n_blocks, s.scalars + 0);
    kmb_tp_line_ends<<<n_blocks, 256, 0, st>>>(text, n_text, s.blk_nl, s.nl_pos, line_cap, s.d_result);
    kmb_tp_classify<<<n_line_blocks, 256, 0, st>>>(text, n_text, format, s.nl_pos, s.scalars + 0, line_cap, s.blk_bases, s.blk_reads, s.d_result);
    kmb_tp_scan64<<<1, 1024, 0, st>>>(s.blk_bases, n_line_blocks, s.scalars + 1);
    kmb_tp_scan<<<1, 1024, 0, st>>>(s.blk_reads, n_line_blocks, s.scalars + 2);
    kmb_tp_emit<<<n_line_blocks, 256, 0, st>>>(text, n_text, format, s.nl_pos, s.scalars + 0, line_cap, s.blk_bases, s.blk_reads,
                                               s.scalars + 1, s.scalars + 2, s.toffsets, line_cap + 2, s.copies, s.d_result);
    kmb_tp_copy<<<std::max(1, std::min<int>((int)n_line_blocks * 4, sms * 8)), 256, 0, st>>>(text, s.copies, s.d_result, s.tbases);
    g_launches += 8;
    KMB_CUDA(cudaGetLastError());
    KMB_CUDA(cudaMemcpyAsync(s.h_result, s.d_result, sizeof(KmbTextResult), cudaMemcpyDeviceToHost, st));
    return KMB_OK;
}

struct TextCopyJob {
    const uint8_t *src;   // memory source, or
    int fd;               // file source (src == nullptr): bytes [offset, offset + n) of fd
    uint64_t offset;
    uint8_t *dst;
    uint64_t n;
    int parts;
    std::atomic<int> failed;
};
static void text_copy_part(void *ctx, int part) {
    TextCopyJob *j = (TextCopyJob *)ctx;
    const uint64_t per = (j->n + j->parts - 1) / j->parts;
    const uint64_t lo = std::min<uint64_t>(j->n, per * part), hi = std::min<uint64_t>(j->n, lo + per);
    if (hi <= lo) return;
    if (j->src) {
        memcpy(j->dst + lo, j->src + lo, (size_t)(hi - lo));
        return;
    }
    // pread: the kernel copies from the page cache straight into the pinned buffer -- no mapping, no page faults
    uint64_t done = lo;
    while (done < hi) {
        const ssize_t r = pread(j->fd, j->dst + done, (size_t)(hi - done), (off_t)(j->offset + done));
        if (r <= 0) {
            j->failed = 1;
            return;
        }
        done += (uint64_t)r;
    }
}

static unsigned long long text_now_us() {
    return (unsigned long long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// The chunk whose copy + parse was queued by the last kmb_mapper_map_text: wait for its parse result, launch its
// mapping kernel.  Called by the next map_text (after it has queued ITS copy, so that the bus stays busy) and by
// everything that needs the mapper's work to be complete or in stream order.
static int text_finish_pending(kmb_mapper *m) {
    if (m->text_pending < 0) return KMB_OK;
    StageSlot &s = m->slot[m->text_pending];
    m->text_pending = -1;
    unsigned long long t0 = text_now_us();
    KMB_CUDA(cudaEventSynchronize(s.copied));
    g_text_us_wait += text_now_us() - t0;
    if (s.h_result->error & KMB_TP_ERR_TOO_MANY_LINES) {
        // more lines than one per 8 bytes: parse again (the text is still on the device) with room for the worst case
        DevBuf<uint8_t> keep;
        KMB_TRY(keep.alloc((size_t)s.text_n + 16));
        KMB_CUDA(cudaMemcpyAsync(keep.p, s.text_ptr, s.text_n, cudaMemcpyDeviceToDevice, m->parse_stream));
        KMB_CUDA(cudaStreamSynchronize(m->parse_stream));
        KMB_TRY(slot_reserve_text(s, (size_t)s.text_n, (size_t)s.text_n + 16));
        KMB_CUDA(cudaMemcpyAsync(s.text, keep.p, s.text_n, cudaMemcpyDeviceToDevice, m->parse_stream));
        s.text_ptr = s.text;
        KMB_TRY(launch_text_parse(m->index->info.sms, s, s.text_n, s.text_format, m->parse_stream));
        KMB_CUDA(cudaStreamSynchronize(m->parse_stream));
    }
    if (s.h_result->error)
        return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_text: malformed %s text (a FASTQ record must be '@' line, bases, '+' line, "
                        "qualities; FASTA data must begin with a '>' line; the text must hold whole records)", s.text_format ? "FASTQ" : "FASTA");
    const uint64_t n_reads = s.h_result->n_reads, n_bases = s.h_result->n_bases;
    g_text_reads += n_reads;
    g_text_bases += n_bases;
    if (n_reads && n_bases) {
        KMB_TRY(slot_reserve(s, 0, 0, (size_t)(n_bases / KMB_WTILE_POS + 1)));
        KMB_TRY(launch_map_reads(m, s.tbases, n_bases, 0, s.toffsets, n_reads, s.tiles, s.text_k, s.text_flags, false));
    }
    KMB_CUDA(cudaEventRecord(s.consumed, m->stream));
    s.used = true;
    return KMB_OK;
}

// Pipeline over three slots: while this call copies chunk i into pinned staging (all cores) and queues its H2D copy
// and parse kernels on the copy stream, chunk i-1 is on the bus / being parsed and chunk i-2 is under the mapping
// kernel on the compute stream.  The mapping kernel of chunk i is launched by the NEXT call (or by sync / flush /
// read_counts ...), once its parse result -- how many reads and bases there are -- has come back.
static int map_text_impl(kmb_mapper *m, const uint8_t *text, int fd, uint64_t fd_offset, uint64_t n_text, int format, int k,
                         uint32_t flags) {
    if (!m) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_text: null mapper");
    KMB_TRY(check_k(k));
    if (format != 0 && format != 1) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_text: format must be 0 (FASTA) or 1 (FASTQ)");
    if (n_text == 0) return KMB_OK;
    if (!text && fd < 0) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_text: null text");
    if (n_text >= (1ull << 32) - 64) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_text: at most 4 GiB of text per call");
    KMB_ON_DEVICE(m->index->device);
    bool dev = false, pinned = false;
    if (text) KMB_TRY(ptr_on_device(text, m->index->device, &dev, &pinned));
    const int slot_index = m->next_slot % KMB_TEXT_SLOTS;
    if (m->text_pending == slot_index) KMB_TRY(text_finish_pending(m));
    unsigned long long t0 = text_now_us();
    StageSlot &s = m->slot[slot_index];
    m->next_slot = (slot_index + 1) % KMB_TEXT_SLOTS;
    if (s.used) KMB_CUDA(cudaEventSynchronize(s.consumed));
    g_text_us_slot += text_now_us() - t0;
    t0 = text_now_us();
    // one line per 8 bytes of text is the first guess (FASTQ of 150-base reads has one per 85)
    KMB_TRY(slot_reserve_text(s, (size_t)n_text, (size_t)(n_text / 8 + 1024)));
    g_text_us_alloc += text_now_us() - t0;
    const uint8_t *src = text;
    if (!dev && !pinned) {
        t0 = text_now_us();
        // pageable source (a mapping of the page cache, an inflate buffer): every core copies its share into pinned staging,
        // which the DMA engine then reads at full PCIe speed (a pageable cudaMemcpy goes through a ~10 GB/s bounce buffer)
        if (n_text > s.h_text_cap) {
            if (s.h_text) cudaFreeHost(s.h_text);
            s.h_text = nullptr;
            s.h_text_cap = 0;
            const size_t cap = (size_t)n_text + (size_t)(n_text >> 3) + 4096;
            KMB_CUDA(cudaMallocHost(&s.h_text, cap));
            s.h_text_cap = cap;
        }
        const int threads = g_opt.host_threads > 0 ? (int)g_opt.host_threads : kmb_host_cpus();
        TextCopyJob job;
        job.src = text;
        job.fd = fd;
        job.offset = fd_offset;
        job.dst = s.h_text;
        job.n = n_text;
        job.parts = std::max(1, std::min(threads, (int)(n_text >> 20) + 1));
        job.failed = 0;
        kmb_host_parallel(threads, job.parts, text_copy_part, &job);
        if (job.failed.load()) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_text_fd: cannot read %llu bytes at offset %llu of the file",
                                                (unsigned long long)n_text, (unsigned long long)fd_offset);
        src = s.h_text;
        g_text_us_stage += text_now_us() - t0;
    }
    KMB_CUDA(cudaMemcpyAsync(s.text, src, n_text, cudaMemcpyDefault, m->copy_stream));
    if (!dev) g_h2d_bytes += n_text;
    cudaEvent_t caller_buffer_read = nullptr;
    if (!dev && pinned) {  // the DMA engine reads the caller's own buffer: it must not be reused before the copy is done
        KMB_CUDA(cudaEventCreateWithFlags(&caller_buffer_read, cudaEventDisableTiming));
        KMB_CUDA(cudaEventRecord(caller_buffer_read, m->copy_stream));
    }
    s.text_ptr = s.text;
    s.text_n = n_text;
    s.text_format = format;
    s.text_k = k;
    s.text_flags = flags;
    // the parse kernels run on their own stream, so that the next chunk's copy does not queue up behind them
    KMB_CUDA(cudaEventRecord(m->text_on_device, m->copy_stream));
    KMB_CUDA(cudaStreamWaitEvent(m->parse_stream, m->text_on_device, 0));
    KMB_TRY(launch_text_parse(m->index->info.sms, s, n_text, format, m->parse_stream));
    KMB_CUDA(cudaEventRecord(s.copied, m->parse_stream));
    // now that this chunk is on its way, finish the previous one: its mapping kernel goes onto the compute stream
    KMB_TRY(text_finish_pending(m));
    m->text_pending = slot_index;
    if (caller_buffer_read) {
        cudaEventSynchronize(caller_buffer_read);
        cudaEventDestroy(caller_buffer_read);
    } else if (dev) {
        KMB_CUDA(cudaStreamSynchronize(m->copy_stream));  // a device buffer of the caller: read before we return
    }
    return KMB_OK;
}


// ------------------------------------------------------------------------------------------------
// multi-member .gz in, counts out: members inflated by the device (kmb_gzdev.cuh), then parsed and mapped there
// ------------------------------------------------------------------------------------------------
static std::atomic<unsigned long long> g_gz_members{0}, g_gz_text_bytes{0}, g_gz_host_batches{0};
#define KMB_GZ_WINDOW (1u << 20)   // text looked at for a record start on either side of a batch boundary

struct GzScanJob {
    const uint8_t *gz;
    uint64_t n, per;
    std::vector<std::vector<uint64_t>> *found;
};
// A gzip member header (RFC 1952): ID1 ID2 CM=8, no reserved flag bits, XFL 0/2/4, a known OS byte.  The last two make a
// chance match inside compressed data 1500x rarer than the three magic bytes alone (~3 per GB -> ~0.002 per GB).
static inline bool gz_header_at(const uint8_t *p) {
    return p[0] == 0x1f && p[1] == 0x8b && p[2] == 8 && (p[3] & 0xE0) == 0 && (p[8] == 0 || p[8] == 2 || p[8] == 4) &&
           (p[9] <= 13 || p[9] == 255);
}
static void gz_scan_part(void *ctx, int part) {
    GzScanJob *j = (GzScanJob *)ctx;
    const uint64_t lo = j->per * (uint64_t)part, hi = std::min<uint64_t>(j->n, lo + j->per);
    std::vector<uint64_t> &out = (*j->found)[(size_t)part];
    uint64_t p = lo;
    while (p < hi && p + 18 <= j->n) {
        const uint8_t *q = (const uint8_t *)memchr(j->gz + p, 0x1f, (size_t)(hi - p));
        if (!q) break;
        p = (uint64_t)(q - j->gz);
        if (p + 18 <= j->n && gz_header_at(q)) out.push_back(p);
        p++;
    }
}

static int slot_reserve_gz(StageSlot &s, size_t gz_bytes, size_t n_members) {
    if (gz_bytes + 64 > s.h_gz_cap) {
        if (s.h_gz) cudaFreeHost(s.h_gz);
        cudaFree(s.d_gz);
        s.h_gz = s.d_gz = nullptr;
        s.h_gz_cap = s.d_gz_cap = 0;
        const size_t cap = gz_bytes + (gz_bytes >> 2) + 4096;
        KMB_CUDA(cudaMallocHost(&s.h_gz, cap));
        KMB_CUDA(cudaMalloc(&s.d_gz, cap));
        s.h_gz_cap = s.d_gz_cap = cap;
    }
    if (n_members > s.members_cap) {
        if (s.h_members) cudaFreeHost(s.h_members);
        if (s.h_results) cudaFreeHost(s.h_results);
        if (s.h_crc) cudaFreeHost(s.h_crc);
        cudaFree(s.d_members);
        cudaFree(s.d_results);
        cudaFree(s.d_crc);
        s.h_members = s.d_members = nullptr;
        s.h_results = s.d_results = nullptr;
        s.h_crc = s.d_crc = nullptr;
        s.members_cap = 0;
        const size_t cap = n_members + n_members / 4 + 64;
        KMB_CUDA(cudaMallocHost(&s.h_members, cap * sizeof(KmbGzMember)));
        KMB_CUDA(cudaMalloc(&s.d_members, cap * sizeof(KmbGzMember)));
        KMB_CUDA(cudaMallocHost(&s.h_results, cap * sizeof(KmbGzResult)));
        KMB_CUDA(cudaMalloc(&s.d_results, cap * sizeof(KmbGzResult)));
        KMB_CUDA(cudaMallocHost(&s.h_crc, cap * sizeof(uint32_t)));
        KMB_CUDA(cudaMalloc(&s.d_crc, cap * sizeof(uint32_t)));
        s.members_cap = cap;
    }
    if (!s.h_windows) KMB_CUDA(cudaMallocHost(&s.h_windows, 2 * KMB_GZ_WINDOW));
    return KMB_OK;
}

struct GzBatch {      // what a slot is working on
    size_t m0 = 0, m1 = 0, x1 = 0;   // members [m0, m1) are the batch; [m1, x1) complete its last record ("overlap")
    bool first = false, overlap_is_tail = false;   // overlap_is_tail: the overlap runs to the end of the data
    uint64_t text_len = 0, extra_len = 0;   // text of [m0, m1) / of [m1, x1)
    uint64_t gz_offset = 0, gz_len = 0;     // compressed bytes of [m0, x1)
    int slot = -1;
};

// Verify the inflated members of a batch (a batch the device got wrong is inflated again by the host decoders, so a
// decoder fault costs time, not correctness; data that neither can decode is an error), cut its text at record
// starts, parse and map it.
static int gz_finish_batch(kmb_mapper *m, const GzBatch &b, int format, int k, uint32_t flags, bool check_crc, int threads) {
    StageSlot &s = m->slot[b.slot];
    KMB_CUDA(cudaEventSynchronize(s.copied));
    const size_t n_mem = b.x1 - b.m0;
    const uint64_t total = b.text_len + b.extra_len;
    bool verified = true;
    for (size_t i = 0; i < n_mem && verified; i++) {
        const KmbGzResult &r = s.h_results[i];
        const KmbGzMember &mem = s.h_members[i];
        if (r.status != KMB_GZ_OK || r.out_len != mem.out_len || r.in_used != mem.in_len) verified = false;
        if (check_crc && s.h_crc[i] != r.crc) verified = false;
    }
    if (!verified) {
        std::vector<uint8_t> text((size_t)total + 64);
        uint64_t consumed = 0, produced = 0;
        int flag = 0;
        const int rc = kmb_gunzip_members(s.h_gz, b.gz_len, threads, text.data(), total, 0, &consumed, &produced, &flag);
        if (rc != KMB_OK || consumed != b.gz_len || produced != total)
            return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_gz: corrupt gzip data in the members at bytes [%llu, %llu) (or a false member "
                            "start: map the file through the host decoders instead)", (unsigned long long)b.gz_offset,
                            (unsigned long long)(b.gz_offset + b.gz_len));
        KMB_CUDA(cudaMemcpyAsync(s.text, text.data(), (size_t)total, cudaMemcpyHostToDevice, m->copy_stream));
        KMB_CUDA(cudaStreamSynchronize(m->copy_stream));
        memcpy(s.h_windows, text.data(), (size_t)std::min<uint64_t>(KMB_GZ_WINDOW, total));
        if (b.extra_len) memcpy(s.h_windows + KMB_GZ_WINDOW, text.data() + b.text_len, (size_t)std::min<uint64_t>(KMB_GZ_WINDOW, b.extra_len));
        g_gz_host_batches++;
        g_h2d_bytes += total;
    }
    // Where the batch's records begin and end in its text.  Both neighbours apply the same rule to the same bytes
    // (kmb_find_record_start from the first byte of member m's text): the record that straddles a batch boundary
    // belongs to the batch it begins in.
    uint64_t start = 0, end = b.text_len;
    if (!b.first) {
        uint64_t off = 0;
        const uint64_t w = std::min<uint64_t>(KMB_GZ_WINDOW, total);
        KMB_TRY(kmb_find_record_start(s.h_windows, w, format, &off));
        if (off >= w && w < total) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_gz: a record longer than %u bytes: map the file through the host decoders", KMB_GZ_WINDOW);
        start = std::min<uint64_t>(off, b.text_len);
    }
    if (b.x1 > b.m1) {
        uint64_t off = 0;
        const uint64_t w = std::min<uint64_t>(KMB_GZ_WINDOW, b.extra_len);
        KMB_TRY(kmb_find_record_start(s.h_windows + KMB_GZ_WINDOW, w, format, &off));
        if (off >= w) {
            // no record start in the overlap: fine if the data ends with it (the rest IS the last record) ...
            if (!b.overlap_is_tail)
                return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_gz: a record longer than %u bytes: map the file through the host decoders", KMB_GZ_WINDOW);
            off = b.extra_len;
        }
        end = b.text_len + off;
    }
    g_gz_members += b.m1 - b.m0;
    g_gz_text_bytes += b.text_len;
    if (end <= start) {
        KMB_CUDA(cudaEventRecord(s.consumed, m->stream));
        s.used = true;
        return KMB_OK;
    }
    s.text_ptr = s.text + start;
    s.text_n = end - start;
    s.text_format = format;
    s.text_k = k;
    s.text_flags = flags;
    // (the batch's text is complete: s.copied was waited for above) -- beside the next batch's copy and inflate
    KMB_TRY(launch_text_parse(m->index->info.sms, s, s.text_n, format, m->parse_stream, s.text_ptr));
    KMB_CUDA(cudaEventRecord(s.copied, m->parse_stream));
    m->text_pending = b.slot;
    return text_finish_pending(m);   // waits for the parse result, launches the mapping kernel, records `consumed`
}

extern "C" int kmb_mapper_map_gz(kmb_mapper *m, const uint8_t *gz, uint64_t n_gz, int format, int k, uint32_t flags,
                                 int shard_index, int shard_count, uint64_t *resume_offset) {
    if (!m || !resume_offset) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_gz: null argument");
    *resume_offset = 0;
    KMB_TRY(check_k(k));
    if (format != 0 && format != 1) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_gz: format must be 0 (FASTA) or 1 (FASTQ)");
    if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_gz: bad shard");
    if (n_gz == 0) return KMB_OK;
    if (!gz) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_gz: null data");
    if (n_gz < 18 || !gz_header_at(gz)) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_gz: the data does not begin with a gzip member");
    KMB_ON_DEVICE(m->index->device);
    KMB_TRY(text_finish_pending(m));
    const int threads = g_opt.host_threads > 0 ? (int)g_opt.host_threads : kmb_host_cpus();
    // ---- 1. member starts (host, all cores) and the text length each trailer announces
    std::vector<uint64_t> starts;
    {
        const int parts = std::max(1, std::min(threads * 4, (int)(n_gz >> 20) + 1));
        std::vector<std::vector<uint64_t>> found((size_t)parts);
        GzScanJob job = {gz, n_gz, (n_gz + parts - 1) / parts, &found};
        kmb_host_parallel(threads, parts, gz_scan_part, &job);
        for (auto &v : found) starts.insert(starts.end(), v.begin(), v.end());
    }
    const size_t n_members = starts.size();
    starts.push_back(n_gz);   // starts[i + 1] = end of member i
    const uint64_t max_member = (uint64_t)std::max<int64_t>(g_opt.gz_device_max_member_bytes, 1 << 16);
    std::vector<uint32_t> isize(n_members);
    size_t n_ok = 0;          // members [0, n_ok) can be given to a warp each
    for (; n_ok < n_members; n_ok++) {
        const uint64_t len = starts[n_ok + 1] - starts[n_ok];
        const uint8_t *t = gz + starts[n_ok + 1] - 4;
        isize[n_ok] = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        if (len < 18 || len >= (1ull << 32) || isize[n_ok] > max_member) break;   // a plain single-member .gz, or a false start
    }
    // The device takes members [0, n_dev).  When something further on is not for it (a function of the file alone, so
    // all ranks agree), the last members before it -- KMB_GZ_WINDOW of text -- serve only as the overlap that completes
    // the last record, and the host decoders resume AT the first of them, skipping its first partial record by the same
    // rule: every record is mapped exactly once.
    size_t n_dev = n_members;
    if (n_ok < n_members) {
        uint64_t text = 0;
        n_dev = n_ok;
        while (n_dev > 0 && text < KMB_GZ_WINDOW) text += isize[--n_dev];
    }
    if (n_dev == 0) return KMB_OK;   // *resume_offset == 0: everything through the host decoders
    {
        uint64_t text = 0;
        for (size_t i = 0; i < n_dev; i++) text += isize[i];
        if (text / n_dev > (uint64_t)std::max<int64_t>(g_opt.gz_device_max_mean_member_bytes, 1)) return KMB_OK;   // too few, too large
    }
    const uint64_t batch_cap = (uint64_t)std::max<int64_t>(g_opt.gz_device_batch_bytes, 1 << 20);
    const bool check_crc = g_opt.gz_device_crc != 0;
    // ---- 2. batches of whole members, each followed by its overlap
    std::vector<GzBatch> batches;
    for (size_t i = 0; i < n_dev;) {
        GzBatch b;
        b.m0 = i;
        uint64_t text = 0;
        while (i < n_dev && (i == b.m0 || (text + isize[i] <= batch_cap && i - b.m0 < 60000))) text += isize[i++];
        b.m1 = i;
        b.text_len = text;
        size_t x = i;
        uint64_t extra = 0;
        while (x < n_ok && extra < KMB_GZ_WINDOW) extra += isize[x++];
        b.x1 = x;
        b.extra_len = extra;
        b.overlap_is_tail = x == n_members;
        b.first = b.m0 == 0;
        b.gz_offset = starts[b.m0];
        b.gz_len = starts[b.x1] - starts[b.m0];
        batches.push_back(b);
    }
    // ---- 3. pipeline: while batch b is on the bus and being inflated, batch b-1 is verified, parsed and mapped
    GzBatch pending;
    bool have_pending = false;
    size_t bi = 0;
    for (GzBatch &b : batches) {
        if ((int)(bi++ % (size_t)shard_count) != shard_index) continue;
        b.slot = m->next_slot % KMB_TEXT_SLOTS;
        m->next_slot = (b.slot + 1) % KMB_TEXT_SLOTS;
        StageSlot &s = m->slot[b.slot];
        if (s.used) KMB_CUDA(cudaEventSynchronize(s.consumed));
        const size_t n_mem = b.x1 - b.m0;
        const uint64_t total_text = b.text_len + b.extra_len;
        KMB_TRY(slot_reserve_gz(s, (size_t)b.gz_len, n_mem));
        KMB_TRY(slot_reserve_text(s, (size_t)total_text + 64, (size_t)(total_text / 8 + 1024)));
        {
            TextCopyJob job;
            job.src = gz + b.gz_offset;
            job.fd = -1;
            job.offset = 0;
            job.dst = s.h_gz;
            job.n = b.gz_len;
            job.parts = std::max(1, std::min(threads, (int)(b.gz_len >> 20) + 1));
            job.failed = 0;
            kmb_host_parallel(threads, job.parts, text_copy_part, &job);
            memset(s.h_gz + b.gz_len, 0, 48);
        }
        uint64_t out = 0;
        for (size_t i = 0; i < n_mem; i++) {
            KmbGzMember &mem = s.h_members[i];
            mem.in_off = starts[b.m0 + i] - b.gz_offset;
            mem.in_len = (uint32_t)(starts[b.m0 + i + 1] - starts[b.m0 + i]);
            mem.out_off = out;
            mem.out_len = isize[b.m0 + i];
            out += mem.out_len;
        }
        cudaStream_t st = m->copy_stream;
        KMB_CUDA(cudaMemcpyAsync(s.d_gz, s.h_gz, (size_t)b.gz_len + 48, cudaMemcpyHostToDevice, st));
        KMB_CUDA(cudaMemcpyAsync(s.d_members, s.h_members, n_mem * sizeof(KmbGzMember), cudaMemcpyHostToDevice, st));
        g_h2d_bytes += b.gz_len + n_mem * sizeof(KmbGzMember);
        const size_t smem = KMB_GZ_WARPS * sizeof(KmbGzShared);
        KMB_CUDA(cudaFuncSetAttribute((const void *)kmb_gz_inflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)std::min<size_t>((n_mem + KMB_GZ_WARPS - 1) / KMB_GZ_WARPS, (size_t)m->index->info.sms * 8);
        kmb_gz_inflate_kernel<<<grid, KMB_GZ_WARPS * 32, smem, st>>>(s.d_gz, s.d_members, (uint32_t)n_mem, s.text, s.d_results);
        g_launches++;
        if (check_crc) {
            kmb_gz_crc_kernel<<<(unsigned)((n_mem + 3) / 4), 128, 0, st>>>(s.text, s.d_members, (uint32_t)n_mem, s.d_crc);
            g_launches++;
            KMB_CUDA(cudaMemcpyAsync(s.h_crc, s.d_crc, n_mem * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        }
        KMB_CUDA(cudaGetLastError());
        KMB_CUDA(cudaMemcpyAsync(s.h_results, s.d_results, n_mem * sizeof(KmbGzResult), cudaMemcpyDeviceToHost, st));
        KMB_CUDA(cudaMemcpyAsync(s.h_windows, s.text, (size_t)std::min<uint64_t>(KMB_GZ_WINDOW, total_text), cudaMemcpyDeviceToHost, st));
        if (b.extra_len)
            KMB_CUDA(cudaMemcpyAsync(s.h_windows + KMB_GZ_WINDOW, s.text + b.text_len, (size_t)std::min<uint64_t>(KMB_GZ_WINDOW, b.extra_len),
                                     cudaMemcpyDeviceToHost, st));
        KMB_CUDA(cudaEventRecord(s.copied, st));
        if (have_pending) KMB_TRY(gz_finish_batch(m, pending, format, k, flags, check_crc, threads));
        pending = b;
        have_pending = true;
    }
    if (have_pending) KMB_TRY(gz_finish_batch(m, pending, format, k, flags, check_crc, threads));
    *resume_offset = n_dev == n_members ? n_gz : starts[n_dev];
    return KMB_OK;
}

// members and text bytes inflated by the device so far in this process; batches the host decoders had to redo
extern "C" int kmb_gz_device_stats(uint64_t *n_members, uint64_t *text_bytes, uint64_t *host_batches) {
    if (n_members) *n_members = g_gz_members.load();
    if (text_bytes) *text_bytes = g_gz_text_bytes.load();
    if (host_batches) *host_batches = g_gz_host_batches.load();
    return KMB_OK;
}

extern "C" int kmb_mapper_map_text(kmb_mapper *m, const uint8_t *text, uint64_t n_text, int format, int k, uint32_t flags) {
    if (!text && n_text) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_text: null text");
    return map_text_impl(m, text, -1, 0, n_text, format, k, flags);
}
// The same with the text taken from bytes [offset, offset + n_text) of an open file: every core preads its share
// straight into the pinned staging buffer (no mapping of the file, no page faults, no intermediate copy).
extern "C" int kmb_mapper_map_text_fd(kmb_mapper *m, int fd, uint64_t offset, uint64_t n_text, int format, int k, uint32_t flags) {
    if (fd < 0) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_text_fd: bad file descriptor");
    return map_text_impl(m, nullptr, fd, offset, n_text, format, k, flags);
}

// The device parser on its own: text in (host or device), bases + offsets out (host or device buffers).  The device-side
// counterpart of kmb_parse_reads; what the tests compare with the host parser record by record.
extern "C" int kmb_parse_text_device(int device, const uint8_t *text, uint64_t n_text, int format, uint8_t *bases,
                                     uint64_t bases_capacity, int64_t *offsets, uint64_t offsets_capacity, uint64_t *n_reads,
                                     uint64_t *n_bases) {
    if (!n_reads || !n_bases) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_parse_text_device: null output");
    *n_reads = *n_bases = 0;
    if (format != 0 && format != 1) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_parse_text_device: format must be 0 (FASTA) or 1 (FASTQ)");
    if (n_text == 0) {
        if (offsets && offsets_capacity) {
            const int64_t zero = 0;
            KMB_CUDA(cudaMemcpy(offsets, &zero, 8, cudaMemcpyDefault));
        }
        return KMB_OK;
    }
    if (!text || n_text >= (1ull << 32) - 64) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_parse_text_device: bad text");
    KMB_ON_DEVICE(device);
    DevInfo info;
    KMB_TRY(dev_info(device, &info));
    StageSlot s;
    struct Cleanup {
        StageSlot &s;
        ~Cleanup() { slot_free_text(s); }
    } cleanup{s};
    uint64_t line_cap = n_text / 8 + 1024;
    for (int attempt = 0;; attempt++) {
        KMB_TRY(slot_reserve_text(s, (size_t)n_text, (size_t)line_cap));
        KMB_CUDA(cudaMemcpy(s.text, text, n_text, cudaMemcpyDefault));
        KMB_TRY(launch_text_parse(info.sms, s, n_text, format, 0));
        KMB_CUDA(cudaStreamSynchronize(0));
        if ((s.h_result->error & KMB_TP_ERR_TOO_MANY_LINES) && attempt == 0) {
            line_cap = n_text + 16;
            continue;
        }
        break;
    }
    if (s.h_result->error) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_parse_text_device: malformed %s text", format ? "FASTQ" : "FASTA");
    *n_reads = s.h_result->n_reads;
    *n_bases = s.h_result->n_bases;
    if ((bases && *n_bases > bases_capacity) || (offsets && *n_reads + 1 > offsets_capacity))
        return kmb_fail(KMB_ERR_NOMEM, "kmb_parse_text_device: output capacity too small (%llu bases, %llu reads)",
                        (unsigned long long)*n_bases, (unsigned long long)*n_reads);
    if (bases && *n_bases) KMB_CUDA(cudaMemcpy(bases, s.tbases, *n_bases, cudaMemcpyDefault));
    if (offsets) KMB_CUDA(cudaMemcpy(offsets, s.toffsets, (*n_reads + 1) * 8, cudaMemcpyDefault));
    return KMB_OK;
}

// reads and bases that kmb_mapper_map_text has parsed in this process so far
extern "C" int kmb_text_parsed(uint64_t *n_reads, uint64_t *n_bases) {
    if (n_reads) *n_reads = g_text_reads.load();
    if (n_bases) *n_bases = g_text_bases.load();
    return KMB_OK;
}

extern "C" int kmb_mapper_map_kmers(kmb_mapper *m, const uint64_t *kmers, uint64_t n, uint32_t flags, int k) {
    if (!m) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_kmers: null mapper");
    if (flags & KMB_FLAG_REVCOMP) KMB_TRY(check_k(k));
    if (n == 0) return KMB_OK;
    if (!kmers) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_kmers: null buffer");
    KMB_ON_DEVICE(m->index->device);
    KMB_TRY(text_finish_pending(m));
    bool dev;
    KMB_TRY(ptr_on_device(kmers, m->index->device, &dev));
    if (dev) {
        if (((uintptr_t)kmers & 7u) != 0) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_map_kmers: misaligned device buffer");
        return launch_map_kmers(m, kmers, n, flags, k);
    }
    const uint64_t per = std::max<uint64_t>((uint64_t)g_opt.chunk_bytes / 8, 1);
    for (uint64_t i0 = 0; i0 < n; i0 += per) {
        uint64_t cnt = std::min(per, n - i0);
        StageSlot &s = m->slot[m->next_slot];
        m->next_slot = (m->next_slot + 1) % KMB_SLOTS;
        if (s.used) KMB_CUDA(cudaEventSynchronize(s.consumed));
        KMB_TRY(slot_reserve(s, cnt * 8, 0, 0));
        KMB_CUDA(cudaMemcpyAsync(s.data, kmers + i0, cnt * 8, cudaMemcpyHostToDevice, m->copy_stream));
        g_h2d_bytes += cnt * 8;
        KMB_CUDA(cudaEventRecord(s.copied, m->copy_stream));
        KMB_CUDA(cudaStreamWaitEvent(m->stream, s.copied, 0));
        KMB_TRY(launch_map_kmers(m, reinterpret_cast<const uint64_t *>(s.data), cnt, flags, k));
        KMB_CUDA(cudaEventRecord(s.consumed, m->stream));
        s.used = true;
    }
    KMB_CUDA(cudaStreamSynchronize(m->copy_stream));
    return KMB_OK;
}

static int fetch_status(kmb_mapper *m) {
    KMB_TRY(text_finish_pending(m));
    if (m->dirty) KMB_TRY(launch_flush(m));
    KMB_CUDA(cudaMemcpyAsync(m->h_status, m->d_status, sizeof(KmbStatus), cudaMemcpyDeviceToHost, m->stream));
    KMB_CUDA(cudaStreamSynchronize(m->stream));
    if (m->host_bad < m->h_status->first_bad_offset) m->h_status->first_bad_offset = m->host_bad;  // met by the host encoder
    return KMB_OK;
}

extern "C" int kmb_mapper_sync(kmb_mapper *m) {
    if (!m) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_sync: null mapper");
    KMB_ON_DEVICE(m->index->device);
    KMB_TRY(fetch_status(m));
    if (m->h_status->index_flags & KMB_FLAG_BAD_OFFSETS)
        return kmb_fail(KMB_ERR_BAD_ARG, "read offsets are not non-decreasing from 0 to n_bases (a map_reads call since the last reset)");
    if (m->h_status->first_bad_offset != ~0ull)
        return kmb_fail(KMB_ERR_INVALID_BASE, "invalid base byte at flat offset %llu (only ACGTacgt and N are accepted)",
                        m->h_status->first_bad_offset);
    return KMB_OK;
}

extern "C" int kmb_mapper_bad_offset(kmb_mapper *m, int64_t *offset) {
    if (!m || !offset) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_bad_offset: null argument");
    KMB_ON_DEVICE(m->index->device);
    KMB_TRY(fetch_status(m));
    *offset = m->h_status->first_bad_offset == ~0ull ? -1 : (int64_t)m->h_status->first_bad_offset;
    return KMB_OK;
}

extern "C" int kmb_mapper_read_counts(kmb_mapper *m, uint32_t *out, uint64_t n_counts) {
    if (!m || (!out && n_counts)) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_read_counts: null argument");
    if (n_counts > m->n_counts) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_read_counts: n_counts larger than the mapper's");
    int rc = kmb_mapper_sync(m);
    if (rc != KMB_OK) return rc;
    KMB_ON_DEVICE(m->index->device);
    if (n_counts) {
        KMB_CUDA(cudaMemcpyAsync(out, m->counts, n_counts * 4, cudaMemcpyDefault, m->stream));
        KMB_CUDA(cudaStreamSynchronize(m->stream));
    }
    return KMB_OK;
}

extern "C" int kmb_mapper_write_counts(kmb_mapper *m, const uint32_t *values, uint64_t n_counts) {
    if (!m || (!values && n_counts)) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_write_counts: null argument");
    if (n_counts != m->n_counts) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_write_counts: n_counts must equal the mapper's (%llu)", (unsigned long long)m->n_counts);
    KMB_ON_DEVICE(m->index->device);
    KMB_TRY(text_finish_pending(m));
    KMB_TRY(launch_log_reset(m));   // hits that were logged but not applied belong to the counts being replaced
    m->dirty = false;
    m->queries_since_flush = 0;
    if (n_counts) {
        KMB_CUDA(cudaMemcpyAsync(m->counts, values, n_counts * 4, cudaMemcpyDefault, m->stream));
        KMB_CUDA(cudaStreamSynchronize(m->stream));  // the caller may reuse `values` at once
    }
    return KMB_OK;
}

extern "C" int kmb_mapper_reset(kmb_mapper *m) {
    if (!m) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_reset: null mapper");
    KMB_ON_DEVICE(m->index->device);
    KMB_TRY(text_finish_pending(m));
    KMB_TRY(launch_log_reset(m));
    m->dirty = false;
    m->queries_since_flush = 0;
    KMB_CUDA(cudaMemsetAsync(m->counts, 0, m->n_counts * 4, m->stream));
    return status_reset(m);
}

// Queue the flush of the slot counters onto the node counts (asynchronous, on the mapper's stream).
extern "C" int kmb_mapper_flush(kmb_mapper *m) {
    if (!m) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_flush: null mapper");
    KMB_ON_DEVICE(m->index->device);
    KMB_TRY(text_finish_pending(m));
    if (m->dirty) KMB_TRY(launch_flush(m));
    return KMB_OK;
}

extern "C" int kmb_mapper_counts_device(kmb_mapper *m, uint32_t **counts_device, uint64_t *n_counts) {
    if (!m) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_counts_device: null mapper");
    if (counts_device) *counts_device = m->counts;
    if (n_counts) *n_counts = m->n_counts;
    return KMB_OK;
}

extern "C" int kmb_mapper_stats(kmb_mapper *m, uint64_t *n_kmers_mapped, uint64_t *n_entries_counted) {
    if (!m) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_stats: null mapper");
    KMB_ON_DEVICE(m->index->device);
    KMB_TRY(fetch_status(m));
    if (n_kmers_mapped) *n_kmers_mapped = m->h_status->n_kmers_mapped;
    if (n_entries_counted) *n_entries_counted = m->h_status->n_entries_counted;
    return KMB_OK;
}

// Sum of the device durations of the mapping kernels (the fused reads kernel / the k-mer kernel, not
// the mask or memset launches) recorded since the last call, with option "time_kernels" = 1.
// Look-ups that passed the filter and fetched a sector since the last reset (implies a sync).
extern "C" int kmb_mapper_candidates(kmb_mapper *m, uint64_t *n_candidates) {
    if (!m || !n_candidates) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_candidates: null argument");
    KMB_ON_DEVICE(m->index->device);
    KMB_TRY(fetch_status(m));
    *n_candidates = m->h_status->n_candidates;
    return KMB_OK;
}

static int timed_total(kmb_mapper *m, int klass, double *ms_total, uint64_t *n_kernels) {
    if (!m || !ms_total) return kmb_fail(KMB_ERR_BAD_ARG, "kernel time: null argument");
    KMB_ON_DEVICE(m->index->device);
    KMB_CUDA(cudaStreamSynchronize(m->stream));
    double total = 0;
    for (size_t i = 0; i < m->timed_used[klass]; i++) {
        float ms = 0;
        KMB_CUDA(cudaEventElapsedTime(&ms, m->timed[klass][i].first, m->timed[klass][i].second));
        total += ms;
    }
    *ms_total = total;
    if (n_kernels) *n_kernels = m->timed_used[klass];
    m->timed_used[klass] = 0;
    return KMB_OK;
}
extern "C" int kmb_mapper_kernel_time(kmb_mapper *m, double *ms_total, uint64_t *n_kernels) {
    return timed_total(m, 0, ms_total, n_kernels);
}
// The same for the apply passes (hit log -> node counts).
extern "C" int kmb_mapper_apply_time(kmb_mapper *m, double *ms_total, uint64_t *n_kernels) {
    return timed_total(m, 1, ms_total, n_kernels);
}

// ------------------------------------------------------------------------------------------------
// multi-GPU reduction of the node counts over NCCL (command_line_interface.py:124-130)
// ------------------------------------------------------------------------------------------------
// NCCL is bound at run time, by name, so that the library loads (and the single-GPU path works) where NCCL is
// not installed.  The five prototypes below restate nccl.h (NCCL 2.x ABI): ncclUniqueId is a 128-byte struct
// passed by value, ncclComm_t an opaque pointer, ncclUint32 = 3, ncclSum = 0.
struct KmbNcclId {
    char internal[KMB_COMM_ID_BYTES];
};
struct KmbNccl {
    void *lib = nullptr;
    int (*GetUniqueId)(KmbNcclId *) = nullptr;
    int (*CommInitRank)(void **, int, KmbNcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*CommRegister)(void *, void *, size_t, void **) = nullptr;  // optional (NCCL >= 2.19)
    int (*CommDeregister)(void *, void *) = nullptr;
};
static KmbNccl g_nccl;
static std::mutex g_nccl_mu;

static int nccl_load() {
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_nccl.lib) return KMB_OK;
    const char *env = getenv("KMB_NCCL_LIB");
    const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return kmb_fail(KMB_ERR_NCCL, "NCCL not found: dlopen(libnccl.so.2) failed (%s); set KMB_NCCL_LIB to its path", dlerror());
#define NCCL_SYM(field, name) *(void **)(&g_nccl.field) = dlsym(h, name)
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    NCCL_SYM(CommInitRank, "ncclCommInitRank");
    NCCL_SYM(CommDestroy, "ncclCommDestroy");
    NCCL_SYM(AllReduce, "ncclAllReduce");
    NCCL_SYM(GetErrorString, "ncclGetErrorString");
    NCCL_SYM(CommRegister, "ncclCommRegister");
    NCCL_SYM(CommDeregister, "ncclCommDeregister");
#undef NCCL_SYM
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.GetErrorString) {
        dlclose(h);
        return kmb_fail(KMB_ERR_NCCL, "the NCCL library found lacks a required symbol");
    }
    g_nccl.lib = h;
    return KMB_OK;
}
#define KMB_NCCL(call)                                                                                    \
    do {                                                                                                  \
        int r__ = (call);                                                                                 \
        if (r__ != 0) return kmb_fail(KMB_ERR_NCCL, "%s: %s (%s:%d)", #call, g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
    } while (0)

struct kmb_comm {
    void *comm = nullptr;
    int device = 0, n_ranks = 1, rank = 0;
    std::vector<std::pair<void *, void *>> registered;  // (buffer, NCCL registration handle)
};

extern "C" int kmb_comm_unique_id(uint8_t id[KMB_COMM_ID_BYTES]) {
    if (!id) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_comm_unique_id: null id");
    KMB_TRY(nccl_load());
    KmbNcclId nid;
    KMB_NCCL(g_nccl.GetUniqueId(&nid));
    memcpy(id, nid.internal, KMB_COMM_ID_BYTES);
    return KMB_OK;
}

extern "C" int kmb_comm_init_rank(int device, int n_ranks, int rank, const uint8_t id[KMB_COMM_ID_BYTES], kmb_comm **out) {
    if (!out) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_comm_init_rank: out is null");
    *out = nullptr;
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_comm_init_rank: bad rank %d of %d", rank, n_ranks);
    KMB_TRY(nccl_load());
    KMB_ON_DEVICE(device);
    kmb_comm *c = new (std::nothrow) kmb_comm;
    if (!c) return kmb_fail(KMB_ERR_NOMEM, "out of host memory");
    c->device = device;
    c->n_ranks = n_ranks;
    c->rank = rank;
    KmbNcclId nid;
    memcpy(nid.internal, id, KMB_COMM_ID_BYTES);
    int r = g_nccl.CommInitRank(&c->comm, n_ranks, nid, rank);
    if (r != 0) {
        delete c;
        return kmb_fail(KMB_ERR_NCCL, "ncclCommInitRank(rank %d of %d): %s", rank, n_ranks, g_nccl.GetErrorString(r));
    }
    *out = c;
    return KMB_OK;
}

extern "C" int kmb_comm_destroy(kmb_comm *c) {
    if (!c) return KMB_OK;
    DeviceGuard g(c->device);
    cudaDeviceSynchronize();
    if (g_nccl.CommDeregister)
        for (auto &pr : c->registered)
            if (pr.second) g_nccl.CommDeregister(c->comm, pr.second);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    cudaGetLastError();
    delete c;
    return KMB_OK;
}

// Hit log -> counts, then the in-place sum over the ranks, both queued on the mapper's stream.
extern "C" int kmb_mapper_allreduce(kmb_mapper *m, kmb_comm *c) {
    if (!m || !c) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_allreduce: null argument");
    if (c->device != m->index->device)
        return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_allreduce: communicator is on GPU %d, mapper on GPU %d", c->device, m->index->device);
    KMB_ON_DEVICE(m->index->device);
    KMB_TRY(text_finish_pending(m));
    if (m->dirty) KMB_TRY(launch_flush(m));
    if (c->n_ranks == 1 || m->n_counts == 0) return KMB_OK;
    if (g_nccl.CommRegister) {
        bool have = false;
        for (auto &pr : c->registered) have |= pr.first == (void *)m->counts;
        if (!have) {
            void *handle = nullptr;
            if (g_nccl.CommRegister(c->comm, m->counts, (size_t)m->n_counts * 4, &handle) == 0 && handle)
                c->registered.emplace_back((void *)m->counts, handle);
            else
                c->registered.emplace_back((void *)m->counts, nullptr);  // registration is an optimisation: do not retry
        }
    }
    KMB_NCCL(g_nccl.AllReduce(m->counts, m->counts, (size_t)m->n_counts, /*ncclUint32*/ 3, /*ncclSum*/ 0, c->comm, m->stream));
    return KMB_OK;
}

// ------------------------------------------------------------------------------------------------
// membership and per-key lookup
// ------------------------------------------------------------------------------------------------
template <int MODE, class OutT>
static int run_lookup(kmb_index *ix, uint32_t *counts, uint64_t n_counts, cudaStream_t s, const uint64_t *keys, uint64_t n,
                      OutT *out) {
    if (n == 0) return KMB_OK;
    if (!keys || !out) return kmb_fail(KMB_ERR_BAD_ARG, "lookup: null buffer");
    DevBuf<uint64_t> t_keys;
    const uint64_t *d_keys;
    KMB_TRY(to_device(keys, (size_t)n, ix->device, t_keys, &d_keys, s));
    bool out_dev;
    KMB_TRY(ptr_on_device(out, ix->device, &out_dev));
    DevBuf<OutT> t_out;
    OutT *d_out = out;
    if (!out_dev) {
        KMB_TRY(t_out.alloc((size_t)n));
        d_out = t_out.p;
    }
    KmbProbe P;
    memset(&P, 0, sizeof(P));
    P.lines = ix->lines;
    P.filter = ix->filter_on ? ix->filter : nullptr;
    P.addr = ix->addr;
    P.policies = 2u;
    P.counts = counts;
    P.n_lines = ix->n_lines;
    P.n_counts = n_counts;
    kmb_in_graph_kernel<MODE><<<grid_for(n, 256, ix->info.sms, 16), 256, 0, s>>>(
        d_keys, n, P, (uint8_t *)(MODE == 0 ? (void *)d_out : nullptr), (uint32_t *)(MODE == 1 ? (void *)d_out : nullptr));
    g_launches++;
    KMB_CUDA(cudaGetLastError());
    if (!out_dev) KMB_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(OutT), cudaMemcpyDeviceToHost, s));
    KMB_CUDA(cudaStreamSynchronize(s));
    return KMB_OK;
}

extern "C" int kmb_pack_bases(const uint8_t *bases, uint64_t n_bases, uint32_t flags, int n_threads, uint32_t *words,
                              uint64_t words_capacity, int64_t *first_bad_offset) {
    if ((!bases && n_bases) || !words) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_pack_bases: null buffer");
    if (words_capacity < kmb_packed_words(n_bases))
        return kmb_fail(KMB_ERR_BAD_ARG, "kmb_pack_bases: words_capacity %llu < %llu", (unsigned long long)words_capacity,
                        (unsigned long long)kmb_packed_words(n_bases));
    kmb_host_pack_streaming(g_opt.host_pack_streaming != 0);
    const uint64_t bad = kmb_host_pack(bases, n_bases, !(flags & KMB_FLAG_NO_N_TO_A), n_threads, words);
    if (first_bad_offset) *first_bad_offset = bad == ~0ull ? -1 : (int64_t)bad;
    if (bad != ~0ull)
        return kmb_fail(KMB_ERR_INVALID_BASE, "invalid base byte at flat offset %llu (only ACGTacgt%s are accepted)",
                        (unsigned long long)bad, (flags & KMB_FLAG_NO_N_TO_A) ? "" : " and N");
    return KMB_OK;
}

// Host memory read bandwidth: n_threads workers of the library's pool sum buf[0, n_bytes) once (64-bit loads);
// what the packed transport and the DMA engines of every GPU on the host have to share (bench.py reports it next to
// the end-to-end number it bounds).
struct HostReadJob {
    const uint64_t *p;
    uint64_t n_words, per;
    std::atomic<uint64_t> sum;
};
static void host_read_part(void *ctx, int part) {
    HostReadJob *j = (HostReadJob *)ctx;
    const uint64_t lo = std::min(j->n_words, j->per * (uint64_t)part), hi = std::min(j->n_words, lo + j->per);
    uint64_t a = 0, b = 0, c = 0, d = 0;
    uint64_t i = lo;
    for (; i + 4 <= hi; i += 4) {
        a += j->p[i];
        b += j->p[i + 1];
        c += j->p[i + 2];
        d += j->p[i + 3];
    }
    for (; i < hi; i++) a += j->p[i];
    j->sum += a + b + c + d;
}
extern "C" int kmb_host_read_bandwidth(const void *buf, uint64_t n_bytes, int n_threads, double *gb_per_s) {
    if (!buf || !gb_per_s || n_bytes < 8) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_host_read_bandwidth: bad argument");
    const int threads = n_threads > 0 ? n_threads : kmb_host_cpus();
    HostReadJob job;
    job.p = (const uint64_t *)(((uintptr_t)buf + 7) & ~(uintptr_t)7);
    job.n_words = (n_bytes - 8) / 8;
    const int parts = threads * 8;
    job.per = (job.n_words + parts - 1) / parts;
    job.sum = 0;
    const auto t0 = std::chrono::steady_clock::now();
    kmb_host_parallel(threads, parts, host_read_part, &job);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *gb_per_s = (double)job.n_words * 8.0 / dt / 1e9 + (job.sum.load() == 0x123456789ull ? 1e-30 : 0.0);
    return KMB_OK;
}

extern "C" int kmb_in_graph_index(kmb_index *ix, const uint64_t *kmers, uint64_t n, uint8_t *out) {
    if (!ix) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_in_graph_index: null index");
    KMB_ON_DEVICE(ix->device);
    return run_lookup<0, uint8_t>(ix, nullptr, 0, 0, kmers, n, out);
}

extern "C" int kmb_mapper_lookup_counts(kmb_mapper *m, const uint64_t *keys, uint64_t n, uint32_t *out) {
    if (!m) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_mapper_lookup_counts: null mapper");
    KMB_ON_DEVICE(m->index->device);
    KMB_TRY(text_finish_pending(m));
    if (m->dirty) KMB_TRY(launch_flush(m));
    return run_lookup<1, uint32_t>(m->index, m->counts, m->n_counts, m->stream, keys, n, out);
}

// ------------------------------------------------------------------------------------------------
// hashing (util.py:71-75)
// ------------------------------------------------------------------------------------------------
// Scratch of the hashing entry point (read-boundary mask, per-tile counts, status), kept per device and grown
// on demand: the call is made once per chunk, and a cudaMalloc/cudaFree pair per call cost more than its kernels.
struct HashScratch {
    uint32_t *mask = nullptr;
    size_t mask_cap = 0;
    unsigned long long *tiles = nullptr;
    size_t tiles_cap = 0;
    KmbStatus *status = nullptr;
};
static HashScratch g_hash_scratch[64];
static std::mutex g_hash_mutex;

static int hash_scratch(int device, size_t mask_words, size_t n_tiles, HashScratch **out) {
    if (device < 0 || device >= 64) return kmb_fail(KMB_ERR_BAD_ARG, "device %d out of range", device);
    HashScratch &h = g_hash_scratch[device];
    if (mask_words > h.mask_cap) {
        cudaFree(h.mask);
        h.mask = nullptr;
        h.mask_cap = 0;
        KMB_CUDA(cudaMalloc(&h.mask, (mask_words + mask_words / 4) * 4));
        h.mask_cap = mask_words + mask_words / 4;
    }
    if (n_tiles + 1 > h.tiles_cap) {
        cudaFree(h.tiles);
        h.tiles = nullptr;
        h.tiles_cap = 0;
        KMB_CUDA(cudaMalloc(&h.tiles, (n_tiles + 1 + n_tiles / 4) * 8));
        h.tiles_cap = n_tiles + 1 + n_tiles / 4;
    }
    if (!h.status) KMB_CUDA(cudaMalloc(&h.status, sizeof(KmbStatus)));
    *out = &h;
    return KMB_OK;
}

extern "C" int kmb_hash_reads(int device, const uint8_t *bases, uint64_t n_bases, const int64_t *offsets, uint64_t n_reads,
                              int k, uint32_t flags, uint64_t *out, uint64_t out_capacity, uint64_t *n_out,
                              int64_t *bad_offset) {
    KMB_TRY(check_k(k));
    if (n_out) *n_out = 0;
    if (bad_offset) *bad_offset = -1;
    if (n_bases == 0 || n_reads == 0) return KMB_OK;
    if (!bases || !offsets) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_hash_reads: null buffer");
    if (out_capacity && !out) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_hash_reads: null output with non-zero capacity");
    KMB_ON_DEVICE(device);
    DevInfo info;
    KMB_TRY(dev_info(device, &info));
    cudaStream_t s = 0;
    DevBuf<uint8_t> t_bases;
    DevBuf<int64_t> t_off;
    const uint8_t *d_bases;
    const int64_t *d_off;
    bool dev_b;
    KMB_TRY(ptr_on_device(bases, device, &dev_b));
    if (dev_b) {
        if (((uintptr_t)bases & 15u) != 0) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_hash_reads: device bases buffer must be 16-byte aligned");
        d_bases = bases;
    } else {
        KMB_TRY(t_bases.alloc((size_t)n_bases + 16));
        KMB_CUDA(cudaMemcpyAsync(t_bases.p, bases, (size_t)n_bases, cudaMemcpyHostToDevice, s));
        d_bases = t_bases.p;
    }
    KMB_TRY(to_device(offsets, (size_t)n_reads + 1, device, t_off, &d_off, s));
    const size_t mask_words = (size_t)(n_bases / 32 + 1);
    const uint64_t n_tiles = (n_bases + KMB_TILE_POS - 1) / KMB_TILE_POS;
    std::lock_guard<std::mutex> lock(g_hash_mutex);  // one hashing call per process at a time (shared scratch)
    HashScratch *hsb = nullptr;
    KMB_TRY(hash_scratch(device, mask_words, (size_t)n_tiles, &hsb));
    struct {
        uint32_t *p;
    } d_mask{hsb->mask};
    struct {
        unsigned long long *p;
    } d_tiles{hsb->tiles};
    struct {
        KmbStatus *p;
    } d_status{hsb->status};
    KmbStatus hs;
    memset(&hs, 0, sizeof(hs));
    hs.first_bad_offset = ~0ull;
    KMB_CUDA(cudaMemcpyAsync(d_status.p, &hs, sizeof(hs), cudaMemcpyHostToDevice, s));
    KMB_CUDA(cudaMemsetAsync(d_mask.p, 0, mask_words * 4, s));
    kmb_mark_read_ends<int64_t><<<grid_for(n_reads, 256, info.sms), 256, 0, s>>>(d_off, n_reads, 0, k, d_mask.p);
    kmb_hash_count_kernel<<<(unsigned)n_tiles, KMB_TILE_THREADS, 0, s>>>(d_mask.p, n_bases, k, d_tiles.p);
    kmb_hash_scan_kernel<<<1, 1024, 0, s>>>(d_tiles.p, n_tiles, d_tiles.p + n_tiles);
    g_launches += 3;
    KMB_CUDA(cudaGetLastError());
    unsigned long long total = 0;
    KMB_CUDA(cudaMemcpyAsync(&total, d_tiles.p + n_tiles, 8, cudaMemcpyDeviceToHost, s));
    KMB_CUDA(cudaStreamSynchronize(s));
    if (n_out) *n_out = total;
    // emit (also the pass that validates the bytes); with out_capacity == 0 this only validates
    bool out_dev = true;
    if (out_capacity) KMB_TRY(ptr_on_device(out, device, &out_dev));
    uint64_t cap = std::min<uint64_t>(out_capacity, total);
    DevBuf<uint64_t> t_out;
    uint64_t *d_out = out;
    if (!out_dev) {
        KMB_TRY(t_out.alloc((size_t)cap));
        d_out = t_out.p;
    }
    kmb_hash_emit_kernel<<<(unsigned)n_tiles, KMB_TILE_THREADS, 0, s>>>(d_bases, n_bases, d_mask.p, k, !(flags & KMB_FLAG_NO_N_TO_A),
                                                                       d_tiles.p, d_out, cap, d_status.p);
    g_launches++;
    KMB_CUDA(cudaGetLastError());
    if (!out_dev && cap) KMB_CUDA(cudaMemcpyAsync(out, d_out, (size_t)cap * 8, cudaMemcpyDeviceToHost, s));
    KMB_CUDA(cudaMemcpyAsync(&hs, d_status.p, sizeof(hs), cudaMemcpyDeviceToHost, s));
    KMB_CUDA(cudaStreamSynchronize(s));
    if (hs.first_bad_offset != ~0ull) {
        if (bad_offset) *bad_offset = (int64_t)hs.first_bad_offset;
        return kmb_fail(KMB_ERR_INVALID_BASE, "invalid base byte at flat offset %llu (only ACGTacgt%s are accepted)",
                        hs.first_bad_offset, (flags & KMB_FLAG_NO_N_TO_A) ? "" : " and N");
    }
    if (out_capacity < total)
        return kmb_fail(KMB_ERR_BAD_ARG, "kmb_hash_reads: output capacity %llu < %llu hashes", (unsigned long long)out_capacity, total);
    return KMB_OK;
}

// ------------------------------------------------------------------------------------------------
// legacy codec (encodings.py)
// ------------------------------------------------------------------------------------------------
template <class Launch>
static int run_codec(int device, const void *in, size_t in_bytes, void *out, size_t out_bytes, Launch launch) {
    if (in_bytes == 0 || out_bytes == 0) return KMB_OK;
    if (!in || !out) return kmb_fail(KMB_ERR_BAD_ARG, "codec: null buffer");
    KMB_ON_DEVICE(device);
    DevInfo info;
    KMB_TRY(dev_info(device, &info));
    cudaStream_t s = 0;
    DevBuf<uint8_t> t_in, t_out;
    const uint8_t *d_in;
    KMB_TRY(to_device((const uint8_t *)in, in_bytes, device, t_in, &d_in, s));
    bool out_dev;
    KMB_TRY(ptr_on_device(out, device, &out_dev));
    uint8_t *d_out = (uint8_t *)out;
    if (!out_dev) {
        KMB_TRY(t_out.alloc(out_bytes));
        d_out = t_out.p;
    }
    if ((((uintptr_t)d_in) | ((uintptr_t)d_out)) & 7u) return kmb_fail(KMB_ERR_BAD_ARG, "codec: device buffers must be 8-byte aligned");
    launch(d_in, d_out, info.sms, s);
    g_launches++;
    KMB_CUDA(cudaGetLastError());
    if (!out_dev) KMB_CUDA(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, s));
    KMB_CUDA(cudaStreamSynchronize(s));
    return KMB_OK;
}

extern "C" int kmb_codec_actg_from_bytes(int device, const uint8_t *seq, uint64_t n, uint8_t *out) {
    if (n % 4) return kmb_fail(KMB_ERR_BAD_ARG, "from_bytes: sequence length %llu is not a multiple of 4 (encodings.py:53)", (unsigned long long)n);
    return run_codec(device, seq, n, out, n / 4, [&](const uint8_t *i, uint8_t *o, int sms, cudaStream_t s) {
        kmb_codec_actg_from_bytes_kernel<<<grid_for(n / 4, 256, sms), 256, 0, s>>>(i, n / 4, o);
    });
}
extern "C" int kmb_codec_simple_from_bytes(int device, const uint8_t *seq, uint64_t n, uint8_t *out) {
    if (n % 4) return kmb_fail(KMB_ERR_BAD_ARG, "from_bytes: sequence length %llu is not a multiple of 4 (encodings.py:99)", (unsigned long long)n);
    return run_codec(device, seq, n, out, n / 4, [&](const uint8_t *i, uint8_t *o, int sms, cudaStream_t s) {
        kmb_codec_simple_from_bytes_kernel<<<grid_for(n / 4, 256, sms), 256, 0, s>>>(i, n / 4, o);
    });
}
extern "C" int kmb_codec_to_bytes(int device, const uint8_t *packed, uint64_t n, uint8_t *out) {
    return run_codec(device, packed, n, out, n * 4, [&](const uint8_t *i, uint8_t *o, int sms, cudaStream_t s) {
        kmb_codec_to_bytes_kernel<<<grid_for(n, 256, sms), 256, 0, s>>>(i, n, o);
    });
}
extern "C" int kmb_codec_complement(int device, const uint8_t *in, uint64_t n_bytes, uint8_t *out) {
    return run_codec(device, in, n_bytes, out, n_bytes, [&](const uint8_t *i, uint8_t *o, int sms, cudaStream_t s) {
        kmb_codec_complement_kernel<<<grid_for(n_bytes, 256, sms), 256, 0, s>>>(i, n_bytes, o);
    });
}
extern "C" int kmb_codec_twobit_swap(int device, const void *in, uint64_t n_words, int word_bytes, void *out) {
    if (word_bytes != 1 && word_bytes != 2 && word_bytes != 4 && word_bytes != 8)
        return kmb_fail(KMB_ERR_BAD_ARG, "twobit_swap: word size %d not in {1,2,4,8}", word_bytes);
    size_t bytes = (size_t)n_words * word_bytes;
    return run_codec(device, in, bytes, out, bytes, [&](const uint8_t *i, uint8_t *o, int sms, cudaStream_t s) {
        int grid = grid_for(n_words, 256, sms);
        switch (word_bytes) {
            case 1: kmb_codec_twobit_swap_kernel<uint8_t><<<grid, 256, 0, s>>>((const uint8_t *)i, n_words, (uint8_t *)o); break;
            case 2: kmb_codec_twobit_swap_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t *)i, n_words, (uint16_t *)o); break;
            case 4: kmb_codec_twobit_swap_kernel<uint32_t><<<grid, 256, 0, s>>>((const uint32_t *)i, n_words, (uint32_t *)o); break;
            default: kmb_codec_twobit_swap_kernel<uint64_t><<<grid, 256, 0, s>>>((const uint64_t *)i, n_words, (uint64_t *)o); break;
        }
    });
}

// ------------------------------------------------------------------------------------------------
// pinned host memory
// ------------------------------------------------------------------------------------------------
extern "C" int kmb_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_host_alloc: null");
    *ptr = nullptr;
    KMB_CUDA(cudaMallocHost(ptr, std::max<size_t>(bytes, 1)));
    return KMB_OK;
}
extern "C" int kmb_host_free(void *ptr) {
    if (ptr) KMB_CUDA(cudaFreeHost(ptr));
    return KMB_OK;
}

// ------------------------------------------------------------------------------------------------
// gather micro-roofline
// ------------------------------------------------------------------------------------------------
typedef void (*GatherFn)(const uint8_t *, uint64_t, uint64_t, uint64_t, uint64_t *, int);
template <int W>
static GatherFn gather_fn_w(int unroll) {
    switch (unroll) {
        case 1: return kmb_gather_bench_kernel<W, 1>;
        case 2: return kmb_gather_bench_kernel<W, 2>;
        case 4: return kmb_gather_bench_kernel<W, 4>;
        case 8: return kmb_gather_bench_kernel<W, 8>;
        case 16: return kmb_gather_bench_kernel<W, 16>;
        default: return nullptr;
    }
}

extern "C" int kmb_bench_gather(int device, uint64_t table_bytes, uint64_t n_loads, int load_bytes, int unroll,
                                int threads_per_block, int blocks_per_sm, float *ms) {
    if (!ms) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_bench_gather: null ms");
    GatherFn fn = load_bytes == 8 ? gather_fn_w<8>(unroll) : load_bytes == 16 ? gather_fn_w<16>(unroll) : load_bytes == 32 ? gather_fn_w<32>(unroll) : nullptr;
    if (!fn) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_bench_gather: load_bytes in {8,16,32}, unroll in {1,2,4,8,16}");
    if (table_bytes < 32 || threads_per_block < 32 || threads_per_block > 1024 || blocks_per_sm < 1)
        return kmb_fail(KMB_ERR_BAD_ARG, "kmb_bench_gather: bad launch shape");
    KMB_ON_DEVICE(device);
    DevInfo info;
    KMB_TRY(dev_info(device, &info));
    if (g_opt.l2_fetch_granularity > 0) KMB_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g_opt.l2_fetch_granularity));
    DevBuf<uint8_t> table;
    DevBuf<uint64_t> sink;
    KMB_TRY(table.alloc((size_t)table_bytes));
    KMB_TRY(sink.alloc(1));
    KMB_CUDA(cudaMemset(table.p, 1, (size_t)table_bytes));
    cudaEvent_t e0, e1;
    KMB_CUDA(cudaEventCreate(&e0));
    KMB_CUDA(cudaEventCreate(&e1));
    int grid = g_opt.bench_grid_blocks > 0 ? (int)g_opt.bench_grid_blocks : info.sms * blocks_per_sm;
    size_t smem = 0;
    if (g_opt.bench_load_mode >= 6) {  // the cp.async variant: 32-byte sectors into per-thread shared-memory slots
        if (load_bytes != 32) return kmb_fail(KMB_ERR_BAD_ARG, "kmb_bench_gather: the cp.async modes copy 32-byte sectors");
        switch (unroll) {
            case 1: fn = kmb_gather_async_bench_kernel<1>; break;
            case 2: fn = kmb_gather_async_bench_kernel<2>; break;
            case 4: fn = kmb_gather_async_bench_kernel<4>; break;
            default: return kmb_fail(KMB_ERR_BAD_ARG, "kmb_bench_gather: cp.async modes take unroll 1, 2 or 4");
        }
        smem = (size_t)threads_per_block * unroll * 32;
        KMB_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    fn<<<grid, threads_per_block, smem>>>(table.p, table_bytes / 32, n_loads / 8 + 1, 1, sink.p, (int)g_opt.bench_load_mode);  // warm-up
    KMB_CUDA(cudaEventRecord(e0));
    fn<<<grid, threads_per_block, smem>>>(table.p, table_bytes / 32, n_loads, 2, sink.p, (int)g_opt.bench_load_mode);
    KMB_CUDA(cudaEventRecord(e1));
    g_launches += 2;
    KMB_CUDA(cudaEventSynchronize(e1));
    KMB_CUDA(cudaGetLastError());
    KMB_CUDA(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return KMB_OK;
}
