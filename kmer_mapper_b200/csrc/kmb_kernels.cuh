// kmb_kernels.cuh -- hand-written sm_100a kernels of the k-mer mapping path.
//
//   K5  kmb_build_check_buckets / _count / _plan / _scatter   index re-layout into 32-byte sectors (once per index)
//   K0  kmb_tile_reads_kernel     first read of every 1024-base tile (the fused kernels derive "which windows exist"
//                                 from it); kmb_mark_read_ends: the one-bit-per-base mask the hashing kernels use
//   K1-4 kmb_map_reads_kernel     fused encode + window + filter + sector probe + hit log  (any k; small indexes)
//   K1-4 kmb_map_reads_mz_kernel  the same over the minimizer-bucketed read-path table (k = 31, indexes of >= 8 M
//                                 entries: the default there), kmb_mz_build_* build that table
//   K3-4 kmb_map_kmers_kernel     probe + hit log on ready-made uint64 k-mers (mapper.pyx:19 drop-in)
//   K4b kmb_log_apply_kernel      hit log -> per-node counts, one L2-sized window of nodes at a time, hot nodes
//                                 aggregated in shared memory first
//   K6  kmb_in_graph_kernel       membership mask (mapper.pyx:81)
//   K2  kmb_hash_count/_scan/_emit  flat hash array (util.py:71-75 drop-in)
//   E1  kmb_codec_* kernels       legacy 2-bit codec (encodings.py)
//   B   kmb_gather_bench_kernel   random-gather micro-roofline
//
// Nothing here is a dense contraction, so no tensor-core / TMEM / TMA-tile machinery is used: the
// path is bound by the rate of random DRAM transactions (38.2 G/s measured) plus one L2 hit per
// k-mer (the filter) and a thin coalesced stream of bases.  What matters: as few
// transactions per k-mer as possible -- an L2-resident filter in front, one read-only 32-byte sector
// that answers a query completely, and no read-modify-write of random memory on the hot path: hits
// are appended to coalesced logs and applied later to node windows that fit in L2 -- the fetch
// latency overlapped with the next batch's arithmetic, persistent grid sized to the SMs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmb_core.cuh"

#define KMB_FULL_MASK 0xFFFFFFFFu
#define KMB_TILE_THREADS 256
// Three resident CTAs per SM = up to 85 registers per thread.  At four (64 registers) ptxas spills the sector
// that is in flight between issue and consume, and the spill store waits for the load: the overlap is gone
// (30 % of all stall samples sat on three STL instructions, profiles/README.md).
#define KMB_MAP_MIN_BLOCKS 3
// The mapping kernels have no CTA-level synchronisation (everything is per warp), so their CTA size only sets the
// granularity at which registers and shared memory are handed out.  Measured on config 2: 256 threads x 3 CTAs
// (24 warps per SM) 47.7 ms, 128 x 7 (28 warps, 70 registers) 55.9 ms: more warps only thrash the L2 the filter
// lives in.  (KMB_MAP_THREADS / KMB_MAP_BLOCKS; A/B builds via -D.)
#ifndef KMB_MAP_THREADS
#define KMB_MAP_THREADS 256
#endif
#ifndef KMB_MAP_BLOCKS
#define KMB_MAP_BLOCKS 3
#endif
#define KMB_POS_PER_THREAD 32
#define KMB_TILE_POS (KMB_TILE_THREADS * KMB_POS_PER_THREAD)  // 8192 window starts per tile
#define KMB_WTILE_POS (32 * KMB_POS_PER_THREAD)                // 1024 window starts per warp tile
#define KMB_QUEUE_SLOTS(U) (32 * ((U) + 2))  // per-warp candidate stack: < 64 waiting + 32 U new
#define KMB_IN_N_TO_A 1u  // kmb_map_reads_kernel in_mode: 'N' reads as 'A'
#define KMB_IN_PACKED 2u  //   the input is the packed 2-bit stream, not ASCII

struct KmbStatus {
    unsigned long long first_bad_offset;   // min flat offset of an invalid byte, ~0 if none
    unsigned long long n_kmers_mapped;     // windows looked up
    unsigned long long n_entries_counted;  // +1s destined for node counts (mapper.pyx:68)
    unsigned long long n_live_entries;     // index build: entries reachable through their own bucket
    unsigned long long n_candidates;       // look-ups that passed the filter and fetched a sector
    unsigned int index_flags;              // bit0 bucket out of range, bit1 negative node
    unsigned int pool_lines;               // index build: overflow sectors needed / handed out
    int max_node;
};

struct KmbLog {  // hit log of one mapper: `cap` node ids in groups of 32, one node-range tag per group
    uint32_t *entries;
    uint8_t *tags;               // [cap / 32] bin of the group, KMB_LOG_NO_BIN = reserved but never written
    unsigned long long *cursor;  // [0] ids reserved so far (may run past cap: those went straight onto the counts)
    uint64_t cap;                // multiple of 32
    uint32_t bin_shift;          // bin = node >> bin_shift
    uint32_t chunk_groups;       // groups reserved per atomic
    uint32_t win_shift;          // apply pass: one window of 2^win_shift nodes at a time (<= bin_shift)
    uint32_t n_bins;             // node ranges in use: (n_counts - 1) >> bin_shift < n_bins <= KMB_LOG_BINS
};

struct KmbProbe {  // everything a probe needs, passed by value to the kernels
    const uint32_t *__restrict__ lines;      // 32-byte sectors, read-only
    const uint32_t *__restrict__ filter;     // word-blocked Bloom filter over the keys, or nullptr
    KmbAddr addr;                            // key -> sector, filter word, filter bits (kmb_locate)
    uint32_t policies;                       // L2 priority of: filter (bits 0-1), sector loads (2-3); bits 8+: ablation
    int32_t max_freq;                        // C int like the reference's cut-off (mapper.pyx:19,64)
    uint32_t *counts;                        // node counts: target of the log and of the rare direct reductions
    KmbLog log;
    uint64_t n_lines;                        // sizes of lines[] (in sectors) and counts[]: only the bounds-checked
    uint64_t n_counts;                       // build reads them
};

// Bounds-checked build (-DKMB_BOUNDS_CHECKS: `python -m kmer_mapper_b200._build --bounds`).  Every index the
// kernels compute is compared with the size of what it indexes; a violation bumps the counter of its site
// (kmb_get_option("bounds_failures") adds them up) and the access still happens.  Compiled out otherwise.
#ifdef KMB_BOUNDS_CHECKS
#define KMB_BOUND_SITES 16
__device__ unsigned long long g_kmb_bound_failures[KMB_BOUND_SITES];
#define KMB_BOUND(site, idx, limit)                                                                  \
    do {                                                                                             \
        if (!((unsigned long long)(idx) < (unsigned long long)(limit))) atomicAdd(&g_kmb_bound_failures[site], 1ull); \
    } while (0)
#else
#define KMB_BOUND(site, idx, limit) \
    do {                            \
    } while (0)
#endif
// sites: 0 filter word, 1 main sector, 2 chain sector, 3 staging slot, 4 log entry, 5 log tag, 6 node count,
//        7 candidate stack, 8 base vector, 9 boundary-mask word, 10 apply: tag word, 11 apply: entry, 12 packed tile word

// ------------------------------------------------------------------------------------------------
// cache-hinted loads.  Gathers are use-once: keep them out of L1 (L1::no_allocate) and mark their
// L2 lines evict-first; the read stream likewise.  The bucket-occupancy filter is the one
// structure meant to stay in L2: evict-last.  (On sm_100a the direct .L2::evict_* qualifiers exist
// only for 256-bit loads, so narrower loads carry a createpolicy descriptor as L2::cache_hint.)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t kmb_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t kmb_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t kmb_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// 0 = normal, 1 = evict-first, 2 = evict-last
__device__ __forceinline__ uint64_t kmb_policy_select(uint32_t which) {
    return which == 1u ? kmb_policy_evict_first() : (which == 2u ? kmb_policy_evict_last() : kmb_policy_evict_normal());
}
__device__ __forceinline__ uint4 kmb_ldg_v4_nc(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ uint64_t kmb_ldg_u64_hint(const uint64_t *p, uint64_t pol) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t kmb_ldg_u32_hint(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint4 kmb_ldg_v4_hint(const void *p, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
// node_counts[node] += v as a reduction whose line carries an L2 priority (the apply pass keeps its window evict-last)
__device__ __forceinline__ void kmb_red_add_hint(uint32_t *p, uint32_t v, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
// One 32-byte sector of the index.
// L2::64B: a miss then fetches 64 bytes from HBM instead of the default 128 (measured: 1.98 vs 3.91
// DRAM sectors per random load, profiles/README.md).
__device__ __forceinline__ void kmb_ld_sector(const uint32_t *p, uint32_t (&r)[8], uint64_t pol) {
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p), "l"(pol));
}

// ================================================================================================
// K5: index re-layout.  The reference's structure is "scan n_kmers[h] entries from
// hashes_to_index[h], compare keys" (mapper.pyx:55-62).  Entry l can only ever match a query of
// bucket hl = kmers[l] % modulo, and only if l lies inside that bucket's range: such entries are
// "live", and the set of live entries defines the lookup result for ANY directory, well-formed or
// not (overlapping ranges, entries filed under a foreign bucket).  Live entries are scattered into
// the line of their own bucket; order inside a line is irrelevant for counting.
// ================================================================================================
__global__ void kmb_build_check_buckets(const int32_t *__restrict__ hashes_to_index, const int32_t *__restrict__ n_kmers,
                                     uint64_t modulo, uint64_t n_entries, KmbStatus *status) {
    unsigned flags = 0;
    for (uint64_t h = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; h < modulo; h += (uint64_t)gridDim.x * blockDim.x) {
        int n = n_kmers[h];
        int pos = hashes_to_index[h];
        if (n < 0 || (n > 0 && (pos < 0 || (uint64_t)pos + (uint64_t)n > n_entries))) flags |= 1u;
    }
    if (flags) atomicOr(&status->index_flags, flags);
}

__device__ __forceinline__ bool kmb_entry_live(const int32_t *__restrict__ hashes_to_index,
                                               const int32_t *__restrict__ n_kmers, uint64_t l, uint64_t h) {
    int64_t pos = hashes_to_index[h], n = n_kmers[h];
    return n > 0 && (int64_t)l >= pos && (int64_t)l < pos + n;
}

// pass 1: per-line entry counts, filter bits, node statistics
__global__ void kmb_build_count(const uint64_t *__restrict__ kmers, const int32_t *__restrict__ nodes,
                             const int32_t *__restrict__ hashes_to_index, const int32_t *__restrict__ n_kmers,
                             uint64_t n_entries, KmbMod mod, KmbAddr addr, uint32_t *__restrict__ line_fill,
                             uint32_t *__restrict__ filter, KmbStatus *status) {
    int local_max = -1;
    bool neg = false;
    unsigned long long live_n = 0;
    for (uint64_t l = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; l < n_entries; l += (uint64_t)gridDim.x * blockDim.x) {
        int node = nodes[l];
        neg |= node < 0;
        local_max = max(local_max, node);
        uint64_t q, h;
        const uint64_t key = kmers[l];
        kmb_divmod(key, mod, q, h);  // the reference's bucket (mapper.pyx:54) decides liveness, nothing else
        if (!kmb_entry_live(hashes_to_index, n_kmers, l, h)) continue;
        live_n++;
        const KmbLoc loc = kmb_locate(key, addr);
        KMB_BOUND(1, loc.sector, addr.n_main);
        if (addr.n_filter_words) KMB_BOUND(0, loc.fword, addr.n_filter_words);
        atomicAdd(&line_fill[loc.sector], 1u);
        if (addr.n_filter_words) atomicOr(&filter[loc.fword], loc.fmask);
    }
    for (int o = 16; o > 0; o >>= 1) {
        local_max = max(local_max, __shfl_xor_sync(KMB_FULL_MASK, local_max, o));
        live_n += __shfl_xor_sync(KMB_FULL_MASK, live_n, o);
    }
    unsigned any_neg = __ballot_sync(KMB_FULL_MASK, neg);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&status->max_node, local_max);
        if (live_n) atomicAdd(&status->n_live_entries, live_n);
        if (any_neg) atomicOr(&status->index_flags, 2u);
    }
}

// pass 2 (ASSIGN = false): total overflow sectors needed; (ASSIGN = true): hand them out, write the
// headers of the whole chain and reset line_fill for the scatter
template <bool ASSIGN>
__global__ void kmb_build_plan(uint32_t *__restrict__ line_fill, uint64_t n_main, uint32_t *__restrict__ lines,
                            uint64_t n_lines, KmbStatus *status) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_main; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t c = line_fill[i];
        uint32_t extra = kmb_chain_extra_lines(c);
        if (!ASSIGN) {
            if (extra) atomicAdd(&status->pool_lines, extra);
        } else {
            uint32_t base = 0;
            if (extra) base = (uint32_t)n_main + atomicAdd(&status->pool_lines, extra);
            lines[i * KMB_LINE_WORDS] = kmb_sector_header(c, base);
            for (uint32_t t = 0; t < extra; t++) KMB_BOUND(2, base + t, n_lines);
            for (uint32_t t = 0; t < extra; t++)
                lines[(uint64_t)(base + t) * KMB_LINE_WORDS] = kmb_sector_header(c - KMB_LINE_SLOTS * (t + 1), base + t + 1);
            line_fill[i] = 0;
        }
    }
}

// pass 3: place every live entry; slot order inside a chain is whatever the atomics give
__global__ void kmb_build_scatter(const uint64_t *__restrict__ kmers, const int32_t *__restrict__ nodes,
                               const uint16_t *__restrict__ freqs, const int32_t *__restrict__ hashes_to_index,
                               const int32_t *__restrict__ n_kmers, uint64_t n_entries, KmbMod mod, KmbAddr addr,
                               uint32_t *__restrict__ line_fill, uint32_t *__restrict__ lines, uint64_t n_lines) {
    for (uint64_t l = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; l < n_entries; l += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t key = kmers[l];
        uint64_t q, h;
        kmb_divmod(key, mod, q, h);
        if (!kmb_entry_live(hashes_to_index, n_kmers, l, h)) continue;
        uint64_t main_line = kmb_locate(key, addr).sector;
        uint32_t s = atomicAdd(&line_fill[main_line], 1u);
        uint32_t ovf_base = lines[main_line * KMB_LINE_WORDS] & ~KMB_HDR_CHAIN;  // only meaningful (and used) for s >= 2
        KMB_BOUND(1, main_line, addr.n_main);
        KMB_BOUND(2, kmb_chain_line(main_line, ovf_base, s), n_lines);
        uint32_t *lp = lines + kmb_chain_line(main_line, ovf_base, s) * KMB_LINE_WORDS;
        uint32_t j = kmb_chain_slot(s);
        *reinterpret_cast<uint2 *>(lp + KMB_LINE_KEY_WORD0 + 2 * j) = make_uint2((uint32_t)key, (uint32_t)(key >> 32));
        lp[KMB_LINE_NODE_WORD0 + j] = (uint32_t)nodes[l];
        reinterpret_cast<uint16_t *>(lp + KMB_LINE_FREQ_WORD)[j] = freqs[l];
    }
}

// ================================================================================================
// K0: read-boundary mask.  Bit p of mask set <=> no window may start at flat base p, i.e. p lies in
// the last k-1 bases of its read (or the read is shorter than k).  One thread per read; a read
// touches at most k-1 <= 30 bits = at most two 32-bit words.
// ================================================================================================
template <class OffT>  // int64 offsets of the caller (base0 = the chunk's first base) or uint32 chunk-relative ones
__global__ void kmb_mark_read_ends(const OffT *__restrict__ offsets, uint64_t n_reads, int64_t base0, int k,
                                   uint32_t *mask) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_reads; r += (uint64_t)gridDim.x * blockDim.x) {
        int64_t s = (int64_t)offsets[r] - base0;
        int64_t e = (int64_t)offsets[r + 1] - base0;
        int64_t lo = e - (k - 1);
        if (lo < s) lo = s;
        if (lo >= e) continue;
        uint64_t w0 = (uint64_t)lo >> 5, w1 = (uint64_t)(e - 1) >> 5;
        uint32_t b0 = (uint32_t)lo & 31u, b1 = (uint32_t)(e - 1) & 31u;
        if (w0 == w1) {
            uint32_t bits = (b1 == 31u ? 0xFFFFFFFFu : ((1u << (b1 + 1u)) - 1u)) & ~((1u << b0) - 1u);
            atomicOr(&mask[w0], bits);
        } else {
            atomicOr(&mask[w0], ~((1u << b0) - 1u));
            atomicOr(&mask[w1], b1 == 31u ? 0xFFFFFFFFu : ((1u << (b1 + 1u)) - 1u));
        }
    }
}

// ================================================================================================
// K0 for the fused kernels: where reads begin and end, without a pass over the bases.
//
// The fused kernels work on tiles of 1024 window starts per warp.  A small table -- for every tile the last read
// that starts at or before the tile's first position (one binary search per tile, kmb_tile_reads_kernel: 4 bytes
// per 1024 bases) -- lets a warp load just the offsets of the handful of reads that touch its tile and mark the
// last k-1 positions of each in a 1024-bit mask in shared memory.  This replaces the global read-boundary bitmask
// of the first version (one bit per base: a 0.94 GB memset + an atomicOr kernel + a mask load per tile, for the
// 7.5 G bases of the benchmark batch), and no malformed offset can make anyone write outside the tile's 32 words.
// ================================================================================================
struct KmbReads {
    const void *offsets;        // n_reads + 1 values: int64 of the caller (base0 is subtracted) or uint32 chunk-relative
    const uint32_t *tile_read;  // [n_tiles] see above
    uint64_t n_reads;
    int64_t base0;
    uint32_t off32;             // offsets are uint32
};
__device__ __forceinline__ int64_t kmb_read_offset(const KmbReads &R, uint64_t r) {
    return R.off32 ? (int64_t)reinterpret_cast<const uint32_t *>(R.offsets)[r]
                   : reinterpret_cast<const int64_t *>(R.offsets)[r] - R.base0;
}
#define KMB_FLAG_BAD_OFFSETS 8u  // KmbStatus::index_flags: offsets not monotonic, or not covering [0, n_bases]
__global__ void kmb_tile_reads_kernel(KmbReads R, uint64_t n_bases, uint64_t n_tiles, uint32_t *__restrict__ tile_read,
                                      KmbStatus *status) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && R.n_reads &&
        (kmb_read_offset(R, 0) != 0 || kmb_read_offset(R, R.n_reads) != (int64_t)n_bases))
        atomicOr(&status->index_flags, KMB_FLAG_BAD_OFFSETS);
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n_tiles; t += (uint64_t)gridDim.x * blockDim.x) {
        const int64_t t0 = (int64_t)(t * KMB_WTILE_POS);
        uint64_t lo = 0, hi = R.n_reads;  // first read in [lo, hi) that starts after t0
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (kmb_read_offset(R, mid) <= t0) lo = mid + 1;
            else hi = mid;
        }
        tile_read[t] = (uint32_t)(lo ? lo - 1 : 0);
    }
}
// Where read r begins and ends (relative to the launch's first base); INT64_MAX beyond the last read.
__device__ __forceinline__ void kmb_read_span(const KmbReads &R, uint64_t r, int64_t &s, int64_t &e) {
    s = e = INT64_MAX;
    if (r < R.n_reads) {
        s = kmb_read_offset(R, r);
        e = kmb_read_offset(R, r + 1);
    }
}
// Valid window starts among this lane's 32 positions [t0 + 32 lane, +32) of tile `tile`.  Called by all 32 lanes;
// tmask = 32 words of the warp's shared memory; r0 = tile_read[tile] and (s, e) = span of read r0 + lane, loaded by
// the caller ahead of time (kmb_read_span) so that their latency hides behind the tile's base loads.
__device__ __forceinline__ uint32_t kmb_tile_valid_starts(const KmbReads &R, uint64_t tile, uint64_t n_bases, int k,
                                                          uint32_t *tmask, int lane, KmbStatus *status, uint64_t r0,
                                                          int64_t s, int64_t e) {
    const int64_t t0 = (int64_t)(tile * KMB_WTILE_POS), t1 = t0 + KMB_WTILE_POS;
    tmask[lane] = 0u;
    __syncwarp();
    uint64_t r = r0;
#pragma unroll 1
    for (;;) {
        if (r + (uint64_t)lane < R.n_reads) {
            if (e < s) atomicOr(&status->index_flags, KMB_FLAG_BAD_OFFSETS);
            // the last k-1 bases of the read (all of it if shorter) cannot start a window: [lo, hi) within the tile
            int64_t lo = e - (int64_t)(k - 1);
            if (lo < s) lo = s;
            if (lo < t0) lo = t0;
            const int64_t hi = e < t1 ? e : t1;
            if (lo < hi) {
                const uint32_t a = (uint32_t)(lo - t0), b = (uint32_t)(hi - 1 - t0);  // first and last bit, < 1024
                const uint32_t wa = a >> 5, wb = b >> 5;
                const uint32_t from_a = ~0u << (a & 31u), upto_b = ~0u >> (31u - (b & 31u));
                if (wa == wb) {
                    atomicOr(&tmask[wa], from_a & upto_b);
                } else {  // at most k-1 <= 30 bits: two words
                    atomicOr(&tmask[wa], from_a);
                    for (uint32_t w = wa + 1; w < wb; w++) atomicOr(&tmask[w], ~0u);
                    atomicOr(&tmask[wb], upto_b);
                }
            }
        }
        if (__shfl_sync(KMB_FULL_MASK, s, 31) >= t1) break;  // the next read starts beyond the tile (or there is none)
        r += 32;
        kmb_read_span(R, r + (uint64_t)lane, s, e);             // more than 32 reads touch the tile (short reads)
    }
    __syncwarp();
    const uint64_t p0 = (uint64_t)t0 + (uint64_t)lane * KMB_POS_PER_THREAD;
    if (p0 >= n_bases) return 0u;
    uint32_t valid = ~tmask[lane];
    if (p0 + 32 + (uint64_t)k > n_bases + 1) {  // window must end inside the buffer: p + k <= n_bases
        const int64_t last = (int64_t)n_bases - (int64_t)k - (int64_t)p0;  // last valid i
        valid &= last < 0 ? 0u : (last >= 31 ? 0xFFFFFFFFu : ((1u << (last + 1)) - 1u));
    }
    return valid;
}

// ================================================================================================
// The probe, shared by every mapping kernel.
//
// Level 0 (FILT): one or two bits of one filter word per query (kmb_filter_mask).  modulo/8 bytes
//   -- 57 MB for the reference's default modulo 452 930 477 -- so it stays resident in the L2.  At
//   the reference's load factor (~0.22 entries per bucket) ~87 % of the absent k-mers end here
//   without touching HBM.
// Level 1: survivors are compacted onto a per-warp stack in shared memory and drained 32 at a time,
//   one per lane (32 independent DRAM transactions in flight per warp): the candidate's 32-byte
//   sector carries keys, nodes and frequencies, so every equal key (mapper.pyx:58-62, no break)
//   whose frequency passes the cut-off (:64) yields its node id at once.
// Level 2: the hit.  `node_counts[node] += 1` (:68) on a 320 MB array would be a DRAM read plus a
//   write-back per hit; instead the node id goes to a per-warp staging area in shared memory, binned
//   by node range, and leaves in full 128-byte groups, each tagged with its range.  kmb_log_apply_kernel
//   later plays the groups of one range at a time into their window of node counts, which is small
//   enough to stay in L2.  uint32 addition wraps, so any order gives the reference's bits.  A full
//   log falls back to the direct reduction, as do the (rare) hits found in overflow chains.
// ================================================================================================
struct KmbPol {
    uint64_t first;   // L2 evict-first: the read stream (use once)
    uint64_t filter;  // the filter words (default evict-last: the structure meant to live in L2)
    uint64_t line;    // index sectors (default normal)
};
__device__ __forceinline__ KmbPol kmb_make_policies(uint32_t bits) {
    KmbPol p;
    p.first = kmb_policy_evict_first();
    p.filter = kmb_policy_select(bits & 3u);
    p.line = kmb_policy_select((bits >> 2) & 3u);
    return p;
}

// Compare the (up to two) keys of a loaded sector with km; on_match(node, frequency) for every equal
// key until it returns true.
template <class F>
__device__ __forceinline__ bool kmb_match_sector(const uint32_t (&r)[8], uint64_t km, F on_match) {
    const uint32_t klo = (uint32_t)km, khi = (uint32_t)(km >> 32);
    const uint32_t n_here = kmb_header_count(r[0]);
    if (n_here >= 1u && r[KMB_LINE_KEY_WORD0] == klo && r[KMB_LINE_KEY_WORD0 + 1] == khi)
        if (on_match(r[KMB_LINE_NODE_WORD0], r[KMB_LINE_FREQ_WORD] & 0xFFFFu)) return true;
    if (n_here >= 2u && r[KMB_LINE_KEY_WORD0 + 2] == klo && r[KMB_LINE_KEY_WORD0 + 3] == khi)
        if (on_match(r[KMB_LINE_NODE_WORD0 + 1], r[KMB_LINE_FREQ_WORD] >> 16)) return true;
    return false;
}

// Follow the overflow sectors behind a sector whose header has the chain bit, synchronously.  Rare.
template <class F>
__device__ __forceinline__ void kmb_walk_chain(const KmbProbe &P, const KmbPol &pol, uint64_t km, uint32_t hdr, F on_match) {
    while (hdr & KMB_HDR_CHAIN) {
        uint32_t r[8];
        KMB_BOUND(2, hdr & ~KMB_HDR_CHAIN, P.n_lines);
        kmb_ld_sector(P.lines + (uint64_t)(hdr & ~KMB_HDR_CHAIN) * KMB_LINE_WORDS, r, pol.line);
        if (kmb_match_sector(r, km, on_match)) return;
        hdr = r[0];
    }
}

// Synchronous probe of the whole chain behind a main sector (cross-check variant, membership, lookup).
template <class F>
__device__ __forceinline__ void kmb_probe_line(const KmbProbe &P, const KmbPol &pol, uint64_t km, uint32_t sector, F on_match) {
    uint32_t r[8];
    KMB_BOUND(1, sector, P.addr.n_main);
    kmb_ld_sector(P.lines + (uint64_t)sector * KMB_LINE_WORDS, r, pol.line);
    if (kmb_match_sector(r, km, on_match)) return;
    kmb_walk_chain(P, pol, km, r[0], on_match);
}

// ------------------------------------------------------------------------------------------------
// Hit staging and logs.  Per warp: KMB_LOG_BINS stacks of node ids in shared memory (31 left over + at
// most 2 x 32 new ones per drain = KMB_STAGE_SLOTS; kernels that are short of shared memory use fewer
// and let the overflow valve of kmb_emit work).  A stack with >= 32 ids sends its top 32 as one
// coalesced 128-byte store to the log; the space is reserved with one atomic per 32 hits or more.
// ------------------------------------------------------------------------------------------------
#ifndef KMB_STAGE_SLOTS
#define KMB_STAGE_SLOTS 96
#endif
#define KMB_LOG_HOLE 0xFFFFFFFFu
#define KMB_LOG_NO_BIN 0xFFu
#define KMB_RES_FULL 0xFFFFFFFFu
struct KmbStage {
    uint32_t *cnt;                 // [KMB_LOG_BINS] ids staged per bin
    uint32_t *buf;                 // [KMB_LOG_BINS][slots]
    unsigned long long *res_base;  // [KMB_LOG_BINS] next free position of this warp's reservation for the bin
    uint32_t *res_left;            // [KMB_LOG_BINS] groups left in that reservation, or KMB_RES_FULL
    uint32_t slots;                // capacity of one bin's stack
    uint32_t bins;                 // stacks there are room for (>= the log's n_bins)
};
// `node_counts[node] += 1` (mapper.pyx:68) straight onto the count array: the fallback of every path that cannot
// use the log (log full, hit found in an overflow chain, staging stack full, apply-table collision).  Warp-
// aggregated: the lanes that arrive here together group equal node ids (match.any) and the lowest lane of
// each group adds the group's size, so a hot node costs one atomic per warp instead of one per hit.
__device__ __forceinline__ void kmb_count_direct(uint32_t *counts, uint32_t node, uint32_t weight = 1u) {
    const unsigned here = __activemask();
    const unsigned same = __match_any_sync(here, node);
    if (weight == 1u) {
        if ((int)(threadIdx.x & 31u) == __ffs(same) - 1) atomicAdd(counts + node, (uint32_t)__popc(same));
    } else {
        atomicAdd(counts + node, weight);
    }
}
__device__ __forceinline__ void kmb_emit(const KmbProbe &P, const KmbStage &st, uint32_t node) {
    const uint32_t b = node >> P.log.bin_shift;
    KMB_BOUND(6, node, P.n_counts);
    // (a log may have more ranges than this kernel has stacks: a mapper whose log was laid out for the read-path
    // kernel's KMB_MZ_LOG_BINS and that is then used with another k or with reverse complements; those ids go direct)
    const uint32_t pos = b < st.bins ? atomicAdd(&st.cnt[b], 1u) : 0xFFFFFFFFu;
    if (pos < st.slots) st.buf[b * st.slots + pos] = node;
    else kmb_count_direct(P.counts, node);  // valve: the stack is full (the flush clamps the count)
}
// One group of 32 ids of bin b (lanes >= n write holes) to the log -- or, if the log is full, straight onto
// the counts.  All bins share one pool: space is reserved chunk_groups groups at a time from a single cursor
// (one atomic, and one wait for its result, per chunk), and a group is tagged with its bin when it is written,
// so a skewed node distribution cannot overflow "its" log while the others stay empty.
__device__ __forceinline__ void kmb_log_write(const KmbLog &log, uint32_t *counts, const KmbStage &st, uint32_t b,
                                              const uint32_t *src, uint32_t n, int lane) {
    unsigned long long base = ~0ull;
    if (lane == 0) {
        uint32_t left = st.res_left[b];
        if (left != KMB_RES_FULL) {
            base = st.res_base[b];
            if (left == 0) {
                left = log.chunk_groups;
                base = atomicAdd(&log.cursor[0], 32ull * left);
            }
            if (base + 32 <= log.cap) {
                st.res_left[b] = left - 1;
                st.res_base[b] = base + 32;
                KMB_BOUND(5, base >> 5, log.cap >> 5);
                log.tags[base >> 5] = (uint8_t)b;
            } else {
                st.res_left[b] = KMB_RES_FULL;  // stop reserving: every further atomic would hit the same address
                base = ~0ull;
            }
        }
    }
    base = __shfl_sync(KMB_FULL_MASK, base, 0);
    const uint32_t id = (uint32_t)lane < n ? src[lane] : KMB_LOG_HOLE;
    if (base != ~0ull) {
        KMB_BOUND(4, base + lane, log.cap);
        log.entries[base + lane] = id;
    } else if (id != KMB_LOG_HOLE) {
        kmb_count_direct(counts, id);
    }
}
// The bins named in `ready` have something to send: full groups of 32, or (all = true, end of kernel) everything.
__device__ __forceinline__ void kmb_stage_send(const KmbLog &log, uint32_t *counts, const KmbStage &st, unsigned ready, uint32_t c, int lane, bool all) {
    while (ready) {
        const int b = __ffs(ready) - 1;
        ready &= ready - 1u;
        uint32_t cb = __shfl_sync(KMB_FULL_MASK, c, b);
        while (cb >= 32u || (all && cb > 0u)) {
            const uint32_t n = min(cb, 32u);
            kmb_log_write(log, counts, st, (uint32_t)b, st.buf + b * st.slots + (cb - n), n, lane);
            cb -= n;
        }
        if (lane == 0) st.cnt[b] = cb;
        __syncwarp();
    }
}
// Called by all 32 lanes.  all = false: send full groups of 32; all = true (end of kernel): everything.
// Groups that were reserved but never written keep the tag KMB_LOG_NO_BIN and are skipped by the apply pass.
__device__ __forceinline__ void kmb_stage_flush(const KmbProbe &P, const KmbStage &st, int lane, bool all) {
    __syncwarp();
    const uint32_t c = (uint32_t)lane < st.bins ? min(st.cnt[lane], st.slots) : 0u;  // clamp: the valve of kmb_emit
    const unsigned ready = __ballot_sync(KMB_FULL_MASK, all ? c > 0u : c >= 32u);
    if (ready) kmb_stage_send(P.log, P.counts, st, ready, c, lane, all);
    __syncwarp();
}
// (Measured and rejected for the read-path kernel, whose 96 KB of SASS are half this send path inlined at four sites
// and whose stall samples were 20 % `no_instruction`: the send path and the valve of kmb_emit as __noinline__
// functions shrink the kernel by 42 % and make it 1.2 % SLOWER -- 29.25 ms against 28.88 per 3 G k-mers of config 3.)
// no_log: the mapper has no hit log (small, L2-resident count array: KmbOptions::direct_counts_max_nodes) -- every
// group then leaves through the "log is full" exit of kmb_log_write, straight onto the counts, warp-aggregated.  (A
// test inside kmb_emit instead made the read-path kernel 10 % slower: 355 ms against 321 on config 3.)
__device__ __forceinline__ void kmb_stage_init(const KmbStage &st, int lane, bool no_log) {
    if ((uint32_t)lane < st.bins) {
        st.cnt[lane] = 0;
        st.res_left[lane] = no_log ? KMB_RES_FULL : 0u;
        st.res_base[lane] = 0;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// The sector fetch pipeline.  A warp that has 32 candidates requests their 32 sectors -- ONE 256-bit load per
// lane, one DRAM transaction each -- at the top of a batch iteration, runs the batch's arithmetic and filter
// loads, and compares the keys at the bottom of the same iteration.
//
// Two things were measured on the way here (profiles/README.md, round 2):
//  * With the consume step at the top of the NEXT iteration (round 1) ptxas made the loop header wait for every
//    outstanding load (control code `wait=01245` on its first instruction: 19 % of all stall samples), because the
//    load was in flight across the loop's back edge.  Requesting and comparing inside one iteration avoids that.
//  * cp.async (LDGSTS) into shared memory instead of registers -- which would let a batch stay in flight across
//    iterations -- is NOT an option for scattered sectors: per-lane scattered 16-byte cp.async cost ~3.5 LSU cycles
//    per lane and block the other loads meanwhile; the kernel took 74 ms instead of 47 ms whether the copy was
//    consumed one iteration later or in the same one, at 2, 3 or 4 CTAs per SM alike.
// A sector that is full and chained does not make anyone wait either: (k-mer, next sector) goes back onto the
// candidate stack (overflow sectors live in the same array, same layout) and is fetched with a later batch.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void kmb_cp_async_sector(uint32_t *smem_dst, const uint32_t *gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
    asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(d + 16u), "l"(gmem_src + 4) : "memory");
}
__device__ __forceinline__ void kmb_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

struct KmbPipe {
    uint32_t r[8];  // the candidate's sector, in flight between issue and consume
    uint64_t km;
    unsigned issued;  // sector fetches of this warp so far (warp-uniform)
    bool valid;
};
__device__ __forceinline__ void kmb_pipe_init(KmbPipe &pp) {
    pp.valid = false;
    pp.issued = 0;
    pp.km = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) pp.r[i] = 0;
}
// Request the sectors of candidates [base, base + cnt) of the warp's stack, one per lane.  Called by all lanes.
__device__ __forceinline__ void kmb_pipe_issue(const KmbProbe &P, const KmbPol &pol, KmbPipe &pp, const uint64_t *q_kmer,
                                               const uint32_t *q_h, int base, int cnt, int lane) {
    pp.issued += (unsigned)cnt;
    if (lane < cnt && !(P.policies & 0x200u)) {
        pp.km = q_kmer[base + lane];
        KMB_BOUND(1, q_h[base + lane], P.n_lines);
        kmb_ld_sector(P.lines + (uint64_t)q_h[base + lane] * KMB_LINE_WORDS, pp.r, pol.line);
        pp.valid = true;
    }
}
// Compare a sector that has arrived with its candidate, stage the hits.  Returns the next sector of the chain
// when this one is full and chained (0.8 % of the main sectors, but 3 % of the candidates: a sector that holds
// the key one is looking for is more likely to be a crowded one), else 0.
__device__ __forceinline__ uint32_t kmb_pipe_match(const KmbProbe &P, const KmbStage &st, const uint32_t (&r)[8],
                                                   uint64_t km, unsigned &counted) {
    const int32_t max_freq = P.max_freq;
    const bool no_emit = (P.policies & 0x100u) != 0u;
    kmb_match_sector(r, km, [&](uint32_t node, uint32_t freq) {
        if ((int32_t)freq <= max_freq) {
            if (!no_emit) kmb_emit(P, st, node);
            counted++;
        }
        return false;
    });
    return (r[0] & KMB_HDR_CHAIN) ? (r[0] & ~KMB_HDR_CHAIN) : 0u;
}
// Called by all 32 lanes: compare the batch requested by the last kmb_pipe_issue; chain continuations go back
// onto the stack.
__device__ __forceinline__ void kmb_pipe_consume(const KmbProbe &P, KmbPipe &pp, const KmbStage &st, unsigned &counted,
                                                 uint64_t *q_kmer, uint32_t *q_h, int &qcount, int lane) {
    if (!__any_sync(KMB_FULL_MASK, pp.valid)) return;
    uint32_t next = 0;
    if (pp.valid) {
        pp.valid = false;
        next = kmb_pipe_match(P, st, pp.r, pp.km, counted);
    }
    const unsigned m = __ballot_sync(KMB_FULL_MASK, next != 0u);
    if (m) {
        if (next != 0u) {
            const int slot = qcount + __popc(m & ((1u << lane) - 1u));
            q_kmer[slot] = pp.km;
            q_h[slot] = next;
        }
        qcount += __popc(m);
        __syncwarp();
    }
    kmb_stage_flush(P, st, lane, false);
}

// Level 0 for U queries of this lane (all filter loads in flight together).  kf(u) yields query u; bit u of
// vbits says whether query u exists.  Order of one batch:
//   0. request the sectors of 32 candidates that earlier batches left on the stack;
//   1. kmb_locate + the U filter loads of this batch;
//   2. test the filter words; survivors go onto the warp's stack (one inclusive scan of the per-lane
//      survivor counts gives every lane its first slot);
//   3. compare the sectors requested in step 0 (they have had steps 1-2 to arrive);
//   4. if two or more batches of candidates are still waiting (thin or no filter), fetch them now.
// The stack holds < 64 entries on entry and on exit, so KMB_QUEUE_SLOTS >= 32 * (U + 2).
template <int U, bool FILT, class KF>
__device__ __forceinline__ void kmb_probe_batch(const KmbProbe &P, const KmbPol &pol, KmbPipe &pp, const KmbStage &st,
                                                unsigned &counted, const KF &kf, uint32_t vbits, uint64_t *q_kmer,
                                                uint32_t *q_h, int &qcount, int lane) {
    if (qcount >= 32) {
        qcount -= 32;
        kmb_pipe_issue(P, pol, pp, q_kmer, q_h, qcount, 32, lane);
    }
    uint64_t km[U];    // the queries (kept: recomputing them for the push cost more than the registers)
    uint32_t hh[U];    // main sector of the query
    uint32_t need[U];  // filter bits the query needs; 0 = no query
    uint32_t fw[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        km[u] = kf(u);
        const KmbLoc loc = kmb_locate(km[u], P.addr);
        hh[u] = loc.sector;
        const bool valid = (vbits >> u) & 1u;
        if (FILT) {
            need[u] = valid ? loc.fmask : 0u;
            if (valid) KMB_BOUND(0, loc.fword, P.addr.n_filter_words);
            fw[u] = (valid && !(P.policies & 0x400u)) ? kmb_ldg_u32_hint(P.filter + loc.fword, pol.filter) : 0u;
        } else {
            need[u] = valid ? 1u : 0u;
            fw[u] = 1u;
        }
    }
    uint32_t cmask = 0;
#pragma unroll
    for (int u = 0; u < U; u++) cmask |= (need[u] != 0u && (fw[u] & need[u]) == need[u]) ? (1u << u) : 0u;
    const uint32_t mine = __popc(cmask);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(KMB_FULL_MASK, incl, o);
        if (lane >= o) incl += v;
    }
    uint32_t slot = (uint32_t)qcount + incl - mine;
    qcount += (int)__shfl_sync(KMB_FULL_MASK, incl, 31);
#pragma unroll
    for (int u = 0; u < U; u++) {
        if ((cmask >> u) & 1u) {
            KMB_BOUND(7, slot, KMB_QUEUE_SLOTS(U));
            q_kmer[slot] = km[u];
            q_h[slot] = hh[u];
            slot++;
        }
    }
    __syncwarp();
    kmb_pipe_consume(P, pp, st, counted, q_kmer, q_h, qcount, lane);
#pragma unroll 1
    while (qcount >= 64) {
        qcount -= 32;
        kmb_pipe_issue(P, pol, pp, q_kmer, q_h, qcount, 32, lane);
        kmb_pipe_consume(P, pp, st, counted, q_kmer, q_h, qcount, lane);
    }
    __syncwarp();
}

// End of kernel: what is left on the stack (a consume step may put chain continuations back: go on until it is
// empty), then the staged hits and the statistics.
__device__ __forceinline__ void kmb_pipe_finish(const KmbProbe &P, const KmbPol &pol, KmbPipe &pp, const KmbStage &st,
                                                unsigned &counted, uint64_t *q_kmer, uint32_t *q_h, int qcount,
                                                int lane, KmbStatus *status) {
    __syncwarp();
#pragma unroll 1
    while (qcount > 0) {
        const int n = min(qcount, 32);
        qcount -= n;
        kmb_pipe_issue(P, pol, pp, q_kmer, q_h, qcount, n, lane);
        kmb_pipe_consume(P, pp, st, counted, q_kmer, q_h, qcount, lane);
    }
    kmb_stage_flush(P, st, lane, true);
    for (int o = 16; o > 0; o >>= 1) counted += __shfl_xor_sync(KMB_FULL_MASK, counted, o);
    if (lane == 0 && counted) atomicAdd(&status->n_entries_counted, (unsigned long long)counted);
    if (lane == 0 && pp.issued) atomicAdd(&status->n_candidates, (unsigned long long)pp.issued);
}

// Forward windows b0 .. b0+U-1 of a lane's 64 bases (four packed words w0..w3), b0 a multiple of U <= 4.  All
// windows of a batch start inside the same word, so the three words they can touch are picked once per batch and
// each window is two 32-bit funnel shifts -- not two 64-bit variable shifts (which were 14 % of the kernel's
// instructions).  Same value as kmb_window(lo, hi, b0 + u, kmask).
struct KmbWindowFn {
    uint32_t a, b, c;    // the words holding bases 16 i .. 16 i + 47, i = b0 / 16
    uint32_t sh;         // bit offset of window b0 inside a: (2 b0) mod 32 <= 28 and 2 u <= 6 with b0 a multiple of U, so sh + 2 u < 32
    uint32_t mlo, mhi;   // the k-mer mask
    __device__ __forceinline__ KmbWindowFn(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int b0, uint64_t kmask) {
        const bool up = b0 >= 16;
        a = up ? w1 : w0;
        b = up ? w2 : w1;
        c = up ? w3 : w2;
        sh = (2u * (uint32_t)b0) & 31u;
        mlo = (uint32_t)kmask;
        mhi = (uint32_t)(kmask >> 32);
    }
    __device__ __forceinline__ uint64_t operator()(int u) const {
        const uint32_t s = sh + 2u * (uint32_t)u;
        return (uint64_t)(__funnelshift_r(a, b, s) & mlo) | ((uint64_t)(__funnelshift_r(b, c, s) & mhi) << 32);
    }
};
struct KmbRcWindowFn {  // their reverse complements
    KmbWindowFn fwd;
    int k;
    __device__ __forceinline__ uint64_t operator()(int u) const { return kmb_revcomp(fwd(u), k); }
};
struct KmbArrayFn {
    const uint64_t *km;
    __device__ __forceinline__ uint64_t operator()(int u) const { return km[u]; }
};

// ================================================================================================
// K1-4 fused: raw ASCII bases in, slot counters out.  Each base is read from HBM exactly once.
//
// Persistent CTAs of 256 threads walk tiles of 8192 window-start positions.  Per tile:
//   1. 16-byte vector loads of 8192+32 bases (coalesced, evict-first), SWAR-encoded in registers to
//      2 bits/base, stored to shared memory as a packed little-endian bit stream (2 KB + halo);
//   2. each thread owns 32 consecutive positions: two 64-bit shared loads give it every window
//      (window i = bits [2i, 2i+2k) -- the reference's first-base-lowest hash, util.py:71-75);
//      one 32-bit word of the read-boundary mask says which of its 32 starts are real windows;
//   3. in batches of U positions: kmb_locate (sector + filter bits of the k-mer), then the probe levels above.
// base0 = flat offset of bases[0] inside the caller's buffer (chunked host input), only used to
// report the position of an invalid byte.
// ================================================================================================
__device__ __forceinline__ uint4 kmb_load_bases16(const uint8_t *__restrict__ bases, uint64_t v, uint64_t n_vec_full,
                                                  uint64_t n_bases, uint64_t pol) {
    if (v < n_vec_full) {
        KMB_BOUND(8, v * 16 + 15, n_bases);
        return kmb_ldg_v4_hint(bases + v * 16, pol);
    }
    // tail of the buffer: byte loads, 'A' beyond the end (window starts there are masked out)
    uint32_t t[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        uint32_t x = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            uint64_t p = v * 16 + (uint64_t)(a * 4 + b);
            uint32_t c = p < n_bases ? (uint32_t)bases[p] : 65u;
            x |= c << (8 * b);
        }
        t[a] = x;
    }
    return make_uint4(t[0], t[1], t[2], t[3]);
}

__device__ __forceinline__ uint32_t kmb_valid_starts(const uint32_t *__restrict__ mask, uint64_t p0, uint64_t n_bases,
                                                     int k) {
    if (p0 >= n_bases) return 0u;
    KMB_BOUND(9, p0 >> 5, n_bases / 32 + 1);
    uint32_t valid = ~mask[p0 >> 5];
    if (p0 + 32 + (uint64_t)k > n_bases + 1) {  // window must end inside the buffer: p + k <= n_bases
        int64_t last = (int64_t)n_bases - (int64_t)k - (int64_t)p0;  // last valid i
        valid &= last < 0 ? 0u : (last >= 31 ? 0xFFFFFFFFu : ((1u << (last + 1)) - 1u));
    }
    return valid;
}

// Bulk-async copy (TMA, 1-D): `bytes` (a multiple of 16) from global to shared memory, completion counted in bytes on
// an mbarrier.  One instruction of one lane moves a whole tile of bases without passing through the LSU / L1TEX, the
// unit this kernel keeps busiest, and without holding registers while the bytes are under way.
__device__ __forceinline__ void kmb_mbar_init(unsigned long long *mbar, uint32_t arrivals) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void kmb_bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes, unsigned long long *mbar, uint64_t pol) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst), a = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(d), "l"(gmem_src), "r"(bytes), "r"(a), "l"(pol) : "memory");
}
__device__ __forceinline__ void kmb_mbar_wait(unsigned long long *mbar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(mbar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}

// Per-warp shared memory of the key-addressed mapping kernels: 5.9 KB per warp = 47 KB per CTA, three CTAs per SM.
#define KMB_TILE_VECS (KMB_WTILE_POS / 16 + 2)  // 16-byte vectors of bases per warp tile, halo included
template <int U>
struct alignas(16) KmbWarpShared {
    static constexpr int kStageSlots = KMB_STAGE_SLOTS;
#ifdef KMB_MAP_BULK_BASES
    uint4 raw[KMB_TILE_VECS];                             // the NEXT tile's ASCII bases, landed by a bulk-async copy
    unsigned long long raw_mbar;                          // its mbarrier (complete_tx)
    unsigned long long raw_pad;
#endif
    uint64_t qk[KMB_QUEUE_SLOTS(U)];                      // candidate stack: k-mer
    unsigned long long stage_res[KMB_LOG_BINS];
    uint32_t qh[KMB_QUEUE_SLOTS(U)];                      //                  its sector
    uint32_t pack[KMB_WTILE_POS / 16 + 4];                // the tile's packed 2-bit stream: 64 words + halo
    uint32_t tmask[32];                                   // the tile's 1024 "no window starts here" bits
    uint32_t stage[KMB_LOG_BINS * kStageSlots];           // staged hits per node range
    uint32_t stage_cnt[2 * KMB_LOG_BINS];
};
#define KMB_MAP_SMEM_BYTES(U) ((KMB_MAP_THREADS / 32) * sizeof(KmbWarpShared<U>))

template <int U, bool FILT, bool REVCOMP>
__global__ void __launch_bounds__(KMB_MAP_THREADS, KMB_MAP_BLOCKS)
kmb_map_reads_kernel(const uint8_t *__restrict__ bases, uint64_t n_bases, uint64_t base0,
                     KmbReads R, int k, uint32_t in_mode, KmbProbe P, KmbStatus *status) {
    const bool n_to_a = (in_mode & KMB_IN_N_TO_A) != 0u;
    // packed transport (host input, kmb_hostpack.cpp): `bases` holds the 2-bit stream already, validated on the host
    const bool packed = (in_mode & KMB_IN_PACKED) != 0u;
    const uint32_t *__restrict__ words = reinterpret_cast<const uint32_t *>(bases);
    const uint64_t n_words = (n_bases + 15) / 16 + 4;  // kmb_packed_words
    // Everything is per warp (tile of 1024 window starts, packed stream, candidate stack): no CTA barrier,
    // so a warp that is walking a long chain never holds the other seven back.
    extern __shared__ __align__(16) unsigned char kmb_map_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    KmbWarpShared<U> &S = reinterpret_cast<KmbWarpShared<U> *>(kmb_map_smem)[warp];
    uint32_t *pack = S.pack;
    uint64_t *q_kmer = S.qk;
    uint32_t *q_h = S.qh;
    const KmbStage st = {S.stage_cnt, S.stage, S.stage_res, S.stage_cnt + KMB_LOG_BINS, (uint32_t)KmbWarpShared<U>::kStageSlots, (uint32_t)KMB_LOG_BINS};
    kmb_stage_init(st, lane, P.log.cap == 0);
    unsigned counted = 0;
    int qcount = 0;
    KmbPipe pp;
    kmb_pipe_init(pp);
    const KmbPol pol = kmb_make_policies(P.policies);
    const uint64_t kmask = kmb_kmer_mask(k);
    const uint64_t n_tiles = (n_bases + KMB_WTILE_POS - 1) / KMB_WTILE_POS;
    const uint64_t n_vec_full = n_bases / 16;  // vectors entirely inside the buffer
    const uint64_t warp_stride = (uint64_t)gridDim.x * (KMB_MAP_THREADS / 32);
    unsigned long long mapped = 0;

    const uint64_t first_tile = (uint64_t)blockIdx.x * (KMB_MAP_THREADS / 32) + warp;
    uint32_t r_next = first_tile < n_tiles ? R.tile_read[first_tile] : 0u;  // one tile ahead: its latency is never waited for
#ifdef KMB_MAP_BULK_BASES
    // Opt-in (-DKMB_MAP_BULK_BASES), measured and NOT the default: the ASCII bases of a tile (1056 bytes, halo
    // included) fetched ONE TILE AHEAD by a bulk-async copy (UBLKCP) into S.raw -- requested by lane 0 as soon as the
    // previous tile has been encoded, waited for on the warp's mbarrier; tiles that are not wholly inside the buffer,
    // unaligned buffers and packed input take the loads below.  Waiting for the bases is 14 % of this kernel's stall
    // samples, and the copy does remove it, but the kernel gets SLOWER: config 2 50.0 ms against 46.9, config 5 52.9 / 47.0,
    // config 4 k=21 54.3 / 51.0 (same box, parity green on both).  The 8.6 KB per CTA come out of the L1 (3 CTAs: 167 KB
    // of shared memory instead of 141), which this kernel needs for the sectors of its loads in flight.
    const bool bulk_ok = !packed && (reinterpret_cast<uintptr_t>(bases) & 15u) == 0;
    auto bulk_tile = [&](uint64_t t) { return bulk_ok && t < n_tiles && t * KMB_WTILE_POS + 16ull * KMB_TILE_VECS <= n_bases; };
    uint32_t raw_parity = 0;
    bool raw_pending = false;  // warp-uniform
    if (lane == 0) kmb_mbar_init(&S.raw_mbar, 1);
    __syncwarp();
    if (bulk_tile(first_tile)) {
        if (lane == 0) kmb_bulk_load(S.raw, bases + first_tile * KMB_WTILE_POS, 16u * KMB_TILE_VECS, &S.raw_mbar, pol.first);
        raw_pending = true;
    }
#endif
    for (uint64_t tile = first_tile; tile < n_tiles; tile += warp_stride) {
        const uint64_t t0 = tile * KMB_WTILE_POS;
        const uint64_t r0 = r_next;
        if (tile + warp_stride < n_tiles) r_next = R.tile_read[tile + warp_stride];
        int64_t rs, re;  // span of read r0 + lane: in flight while the bases are loaded and encoded
        kmb_read_span(R, r0 + (uint64_t)lane, rs, re);
        __syncwarp();  // the previous tile's readers are done with pack
        // ---- 1. load + encode: vectors v = t0/16 + i, i in [0, 66)
#ifdef KMB_MAP_BULK_BASES
        const bool from_raw = raw_pending;
        if (from_raw) {
            kmb_mbar_wait(&S.raw_mbar, raw_parity);
            raw_parity ^= 1u;
#pragma unroll
            for (int i = lane; i < KMB_TILE_VECS; i += 32) {
                const uint4 w = S.raw[i];
                uint32_t inv;
                pack[i] = kmb_encode16(w.x, w.y, w.z, w.w, n_to_a, inv);
                if (inv) atomicMin(&status->first_bad_offset, (unsigned long long)(base0 + t0 + 16ull * (uint64_t)i + (uint64_t)(__ffs(inv) - 1)));
            }
            __syncwarp();  // every lane has read its share of raw: the next tile may land there
        }
        raw_pending = bulk_tile(tile + warp_stride);
        if (raw_pending && lane == 0)
            kmb_bulk_load(S.raw, bases + (tile + warp_stride) * KMB_WTILE_POS, 16u * KMB_TILE_VECS, &S.raw_mbar, pol.first);
        if (!from_raw)
#endif
        {
#pragma unroll
        for (int i = lane; i < KMB_WTILE_POS / 16 + 2; i += 32) {
            uint64_t v = t0 / 16 + (uint64_t)i;
            KMB_BOUND(12, i, KMB_WTILE_POS / 16 + 4);
            if (packed) {
                pack[i] = v < n_words ? kmb_ldg_u32_hint(words + v, pol.first) : 0u;
                continue;
            }
            uint4 w = kmb_load_bases16(bases, v, n_vec_full, n_bases, pol.first);
            uint32_t inv;
            pack[i] = kmb_encode16(w.x, w.y, w.z, w.w, n_to_a, inv);
            if (inv) atomicMin(&status->first_bad_offset, (unsigned long long)(base0 + v * 16 + (uint64_t)(__ffs(inv) - 1)));
        }
        }
        __syncwarp();
        // ---- 2. this lane's 32 positions
        const uint32_t valid = kmb_tile_valid_starts(R, tile, n_bases, k, S.tmask, lane, status, r0, rs, re);
        mapped += __popc(valid);
        const uint2 a = *reinterpret_cast<const uint2 *>(&pack[2 * lane]);
        const uint2 b = *reinterpret_cast<const uint2 *>(&pack[2 * lane + 2]);
        // ---- 3. probe in batches of U
#pragma unroll 1
        for (int b0 = 0; b0 < KMB_POS_PER_THREAD; b0 += U) {
            const uint32_t vb = (valid >> b0) & ((U == 32) ? 0xFFFFFFFFu : ((1u << U) - 1u));
            if (!__any_sync(KMB_FULL_MASK, vb != 0u)) continue;
            const KmbWindowFn fw(a.x, a.y, b.x, b.y, b0, kmask);
            kmb_probe_batch<U, FILT>(P, pol, pp, st, counted, fw, vb, q_kmer, q_h, qcount, lane);
            if (REVCOMP) {
                const KmbRcWindowFn rc = {fw, k};
                kmb_probe_batch<U, FILT>(P, pol, pp, st, counted, rc, vb, q_kmer, q_h, qcount, lane);
            }
        }
    }
    kmb_pipe_finish(P, pol, pp, st, counted, q_kmer, q_h, qcount, lane, status);
    // statistics: one atomic per warp
    for (int o = 16; o > 0; o >>= 1) mapped += __shfl_xor_sync(KMB_FULL_MASK, mapped, o);
    if (lane == 0 && mapped) atomicAdd(&status->n_kmers_mapped, REVCOMP ? 2ull * mapped : mapped);
}

// ================================================================================================
// K1-4 over the read-path table (kmb_core.cuh, "Read-path table"): the fused kernel for k = 31 reads.
// Same tiles, same encode, same hit log as kmb_map_reads_kernel; what differs is the unit of a
// look-up: a RUN of consecutive windows with the same minimizer costs one filter word, one 64-byte
// bucket fetch, and one key comparison per entry of the bucket (the entry's minimizer offset says
// which window of the run it could equal).  Tile-synchronous, per warp:
//   1. load + encode the tile (as in kmb_map_reads_kernel);
//   2. per lane: 47 m-mer ordering keys (26 hash bits | position) and a log-step sliding minimum
//      (window 16 = k - 16 + 1) give minimizer and position for its 32 windows; a bit mask marks
//      where runs start;
//   then, in two passes of 16 lanes' runs each (so that everything below is sized for half a tile):
//   2b. the pass's run list (lanes append the runs that hold an existing window);
//   3. one filter word per run, 32 runs per round; the runs that pass get a staging slot;
//   4. all their primary sectors are fetched with cp.async in one burst (~50 independent 64-byte
//      DRAM fetches in flight per warp, no registers held), then the secondary sectors of the fuller
//      buckets (L2 hits: same 64 bytes);
//   5. 32 runs per round, one per lane; per entry of the run's bucket the one window that could match
//      is extracted and compared; buckets that continue in the pool go onto a per-warp list that is
//      retired 32 runs at a time.
// No cross-tile state except the staged hits and that list.
// Also measured and rejected (config 3, 3 G k-mers, same box): an L2 prefetch of the warp's next tile of bases 29.3 ms
// against 28.9 without; the filter words of three rounds requested before the first is tested 30.4 ms.
// ================================================================================================
#define KMB_MZ_K 31
#define KMB_MZ_THREADS 128  // 4 warps per CTA, seven CTAs per SM
#ifdef KMB_BOUNDS_CHECKS  // the checked build doubles as a stress build: tiny limits, so every overflow path runs in the tests
#define KMB_MZ_SLOTS 4
#define KMB_MZ_SLOTS2 2
#define KMB_MZ_RUNS_KEPT 8
#else
// A tile is worked off in two passes of 16 lanes' runs each, so that everything below is sized for half a tile:
// 7.7 KB of shared memory per warp = 28 warps per SM instead of 20.
#define KMB_MZ_SLOTS 64     // primary sectors staged per pass (half a tile has ~73 runs, ~52 pass the filter)
#ifndef KMB_MZ_SLOTS2
#define KMB_MZ_SLOTS2 24    // secondary sectors staged per pass (buckets with more than two entries)
#endif
#define KMB_MZ_RUNS_KEPT 96  // passing runs remembered per pass; beyond KMB_MZ_SLOTS they load their bucket late
#endif
#define KMB_MZ_NONE 0xFFu   // secondary table: the bucket has no secondary sector to look at
#define KMB_MZ_LATE 0xFEu   // secondary table: no staging slot left, load from global memory instead
#define KMB_MZ_PACK_WORDS (KMB_WTILE_POS / 16 + 4)
#define KMB_MZ_LATE_CAP 40  // <= 8 left over + at most 32 new per round
#ifndef KMB_MZ_STAGE_SLOTS
#define KMB_MZ_STAGE_SLOTS 35  // 12 ranges x 35 ids (was 8 x 48): a full group of 32 + what one round can add before the flush
#endif
// a run of one lane: lane | first window << 5 | last window << 10 | minimizer position (0..47) << 15
#define KMB_MZ_RUN(lane, s, e, j) ((uint32_t)(lane) | ((uint32_t)(s) << 5) | ((uint32_t)((e) - 1) << 10) | ((uint32_t)(j) << 15))
struct alignas(16) KmbMzShared {  // per warp
    unsigned long long stage_res[KMB_MZ_LOG_BINS];
    unsigned long long late_lo[KMB_MZ_LATE_CAP], late_hi[KMB_MZ_LATE_CAP];  // the 64 bases of the run's lane
    uint32_t pack[KMB_MZ_PACK_WORDS];
    union {
        uint32_t runs[16 * 32];                          // every run of the pass, until the filter has been asked
        uint32_t slots[KMB_MZ_SLOTS][KMB_LINE_WORDS];    // then: the primary sectors of the runs that passed
    } a;
    uint8_t mzpos[32][32];                               // [window of the lane][lane]: where its minimizer sits
    uint32_t slots2[KMB_MZ_SLOTS2][KMB_LINE_WORDS];      // secondary sectors
    uint32_t kept_run[KMB_MZ_RUNS_KEPT];     // the runs that passed the filter, in slot order
    uint32_t kept_sector[KMB_MZ_RUNS_KEPT];  // and the primary sector of their bucket
    uint32_t valid[32];                      // per lane: which of its 32 windows exist
    uint8_t slot2_of[KMB_MZ_SLOTS];          // secondary slot of a primary slot, KMB_MZ_NONE, or KMB_MZ_LATE
    uint32_t late_valid[KMB_MZ_LATE_CAP], late_sector[KMB_MZ_LATE_CAP], late_sej[KMB_MZ_LATE_CAP];  // s | e << 8 | jpos << 16
    uint32_t stage[KMB_MZ_LOG_BINS * KMB_MZ_STAGE_SLOTS];
    uint32_t stage_cnt[2 * KMB_MZ_LOG_BINS];
};
#define KMB_MZ_SMEM_BYTES ((KMB_MZ_THREADS / 32) * sizeof(KmbMzShared))
static_assert(sizeof(KmbMzShared) % 16 == 0, "per-warp shared block must keep the 16-byte alignment of the staged sectors");
#ifndef KMB_BOUNDS_CHECKS
static_assert(7 * (KMB_MZ_SMEM_BYTES + 1024) <= 233472, "seven CTAs of the read-path kernel must fit one SM");
#endif

// ---- index side: file every live entry of the key-addressed sectors under its minimizer -------------
__global__ void kmb_mz_build_count(const uint32_t *__restrict__ lines, uint64_t n_lines, int k, KmbAddr addr,
                                   uint32_t *__restrict__ fill, uint32_t *__restrict__ filter, KmbStatus *status) {
    unsigned long long n_here_total = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_lines; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t *lp = lines + i * KMB_LINE_WORDS;
        const uint32_t n = kmb_header_count(lp[0]);
        for (uint32_t j = 0; j < n; j++) {
            const uint64_t key = (uint64_t)lp[KMB_LINE_KEY_WORD0 + 2 * j] | ((uint64_t)lp[KMB_LINE_KEY_WORD0 + 2 * j + 1] << 32);
            if (key >> (2 * k)) continue;  // cannot equal a k-base window
            uint32_t off;
            const KmbLoc loc = kmb_locate((uint64_t)kmb_minimizer(key, k, &off), addr);
            KMB_BOUND(1, loc.sector, addr.n_main);
            atomicAdd(&fill[loc.sector], 1u);
            if (addr.n_filter_words) atomicOr(&filter[loc.fword], loc.fmask);
            n_here_total++;
        }
    }
    if (n_here_total) atomicAdd(&status->n_live_entries, n_here_total);
}

template <bool ASSIGN>
__global__ void kmb_mz_build_plan(uint32_t *__restrict__ fill, uint64_t n_buckets, uint32_t *__restrict__ lines,
                                  uint64_t n_lines, KmbStatus *status) {
    for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < n_buckets; b += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = fill[b];
        const uint32_t extra = kmb_mz_pool_sectors(c);
        if (!ASSIGN) {
            if (extra) atomicAdd(&status->pool_lines, extra);
            if (extra && c - 4u > KMB_MZ_POOL_MAX_ENTRIES) atomicOr(&status->index_flags, 4u);  // does not fit the pool header
            continue;
        }
        uint32_t base = 0;
        if (extra) base = (uint32_t)(2 * n_buckets) + atomicAdd(&status->pool_lines, extra);
        lines[(2 * b) * KMB_LINE_WORDS] = c < 5u ? c : 5u;  // the scatter ORs the minimizer offsets in
        lines[(2 * b + 1) * KMB_LINE_WORDS] = extra ? (KMB_HDR_CHAIN | base) : 0u;
        for (uint32_t t = 0; t < extra; t++) {
            KMB_BOUND(2, base + t, n_lines);
            lines[(uint64_t)(base + t) * KMB_LINE_WORDS] = (c - 4u - 2u * t) & KMB_MZ_POOL_MAX_ENTRIES;  // offsets: scatter
        }
        fill[b] = 0;
    }
}

__global__ void kmb_mz_build_scatter(const uint32_t *__restrict__ src, uint64_t n_src_lines, int k, KmbAddr addr,
                                     uint32_t *__restrict__ fill, uint32_t *__restrict__ lines, uint64_t n_lines) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_src_lines; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t *sp = src + i * KMB_LINE_WORDS;
        const uint32_t n = kmb_header_count(sp[0]);
        for (uint32_t j = 0; j < n; j++) {
            const uint32_t klo = sp[KMB_LINE_KEY_WORD0 + 2 * j], khi = sp[KMB_LINE_KEY_WORD0 + 2 * j + 1];
            const uint64_t key = (uint64_t)klo | ((uint64_t)khi << 32);
            if (key >> (2 * k)) continue;
            uint32_t off;
            const uint64_t b = kmb_locate((uint64_t)kmb_minimizer(key, k, &off), addr).sector;
            const uint32_t s = atomicAdd(&fill[b], 1u);
            const uint32_t pool_base = lines[(2 * b + 1) * KMB_LINE_WORDS] & ~KMB_HDR_CHAIN;  // used for s >= 4 only
            const uint64_t sec = kmb_mz_sector(b, pool_base, s);
            KMB_BOUND(2, sec, n_lines);
            uint32_t *lp = lines + sec * KMB_LINE_WORDS;
            const uint32_t slot = s & 1u;
            *reinterpret_cast<uint2 *>(lp + KMB_LINE_KEY_WORD0 + 2 * slot) = make_uint2(klo, khi);
            lp[KMB_LINE_NODE_WORD0 + slot] = sp[KMB_LINE_NODE_WORD0 + j];
            reinterpret_cast<uint16_t *>(lp + KMB_LINE_FREQ_WORD)[slot] = reinterpret_cast<const uint16_t *>(sp + KMB_LINE_FREQ_WORD)[j];
            if (s < 4u) atomicOr(&lines[(2 * b) * KMB_LINE_WORDS], off << (3u + 5u * s));
            else atomicOr(&lp[0], off << (22u + 5u * slot));
        }
    }
}

// ---- query side ---------------------------------------------------------------------------------------
// Minimizers of windows 16 H .. 16 H + 15 of this lane: w = its 64 bases (4 packed words).  Per window the base
// position (relative to the lane's first base) of the winning m-mer is stored; a bit mask marks where the
// (ordering key, position) pair changes = where a new run starts.
// (h is a run-time value and the two halves share one copy of the code: the kernel is long and straight-line, and
// instruction fetch is what its warps wait for most after memory.)
__device__ __forceinline__ void kmb_mz_half(const uint32_t (&w4)[4], int h, uint8_t *pos_col, uint32_t &prev, uint32_t &startbits) {
    uint32_t a[32];
    const uint32_t w[3] = {h ? w4[1] : w4[0], h ? w4[2] : w4[1], h ? w4[3] : w4[2]};  // bases 16 h .. 16 h + 47
    const uint32_t jbase = 16u * (uint32_t)h;
#pragma unroll
    for (int t = 0; t < 32; t++) {
        // m-mer starting at base 16 h + t: bits [2t, 2t + 30) of the 96-bit stream w
        a[t] = kmb_mmer_order(__funnelshift_r(w[t >> 4], w[(t >> 4) + 1], (2 * t) & 31) & KMB_MZ_MASK) | (jbase + (uint32_t)t);
    }
    // sliding minimum over the k - m + 1 m-mers of a window by doubling: after the four rounds a[t] = min of m-mers t .. t+15
#pragma unroll
    for (int t = 0; t < 31; t++) a[t] = min(a[t], a[t + 1]);
#pragma unroll
    for (int t = 0; t < 29; t++) a[t] = min(a[t], a[t + 2]);
#pragma unroll
    for (int t = 0; t < 25; t++) a[t] = min(a[t], a[t + 4]);
#pragma unroll
    for (int t = 0; t < 17; t++) a[t] = min(a[t], a[t + 8]);
    static_assert(KMB_MZ_K - KMB_MZ_M + 1 == 16 || KMB_MZ_K - KMB_MZ_M + 1 == 17, "window of 16 or 17 m-mers");
    uint32_t sb = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t v = (KMB_MZ_K - KMB_MZ_M + 1 == 17) ? min(a[i], a[i + 1]) : a[i];
        pos_col[(jbase + i) * 32] = (uint8_t)(v & 63u);
        sb |= (v != prev ? 1u : 0u) << i;
        prev = v;
    }
    startbits |= sb << jbase;
}

// the k-mer that starts at base p of the packed tile
__device__ __forceinline__ uint64_t kmb_mz_window_at(const uint32_t *pack, uint32_t p, uint64_t kmask) {
    const uint32_t wi = p >> 4, sh = (p & 15u) * 2u;
    KMB_BOUND(12, wi + 2, KMB_MZ_PACK_WORDS);
    const uint32_t w0 = pack[wi], w1 = pack[wi + 1], w2 = pack[wi + 2];
    return ((uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32)) & kmask;
}
// the k-mer that starts at base i (0..31) of a lane's 64 bases, given as its four packed words
__device__ __forceinline__ uint64_t kmb_mz_window_of(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int i, uint64_t kmask) {
    const bool up = i >= 16;
    const uint32_t a = up ? w1 : w0, b = up ? w2 : w1, c = up ? w3 : w2;
    const uint32_t sh = (2u * (uint32_t)i) & 31u;
    return ((uint64_t)__funnelshift_r(a, b, sh) | ((uint64_t)__funnelshift_r(b, c, sh) << 32)) & kmask;
}
// The two entries of one sector (words r[1..7]: frequencies, keys, nodes) against one run.  off0/off1 = minimizer
// offsets of the two entries, n = how many of them exist; jp = tile position of the run's minimizer; [ps, pe) =
// tile positions of the run's windows; vrow = valid bits of the run's lane (bit = position - lane_base).
__device__ __forceinline__ void kmb_mz_match_pair(const KmbProbe &P, const KmbStage &st, const uint32_t (&r)[8], uint32_t off0,
                                                  uint32_t off1, uint32_t n, int jp, int ps, int pe, int lane_base, uint32_t vrow,
                                                  const uint32_t *pack, uint64_t kmask, unsigned &counted) {
#pragma unroll
    for (uint32_t t = 0; t < 2u; t++) {
        if (t >= n) break;
        const int p = jp - (int)(t ? off1 : off0);  // the only window that has the minimizer at that offset
        if (p < ps || p >= pe || !((vrow >> (p - lane_base)) & 1u)) continue;
        const uint64_t km = kmb_mz_window_at(pack, (uint32_t)p, kmask);
        if (r[KMB_LINE_KEY_WORD0 + 2 * t] != (uint32_t)km || r[KMB_LINE_KEY_WORD0 + 2 * t + 1] != (uint32_t)(km >> 32)) continue;
        const uint32_t freq = t ? (r[KMB_LINE_FREQ_WORD] >> 16) : (r[KMB_LINE_FREQ_WORD] & 0xFFFFu);
        if ((int32_t)freq > P.max_freq) continue;
        kmb_emit(P, st, r[KMB_LINE_NODE_WORD0 + t]);
        counted++;
    }
}

// Retire `n` (<= 32) runs from the top of the late list, one per lane: walk the contiguous pool sectors of the run's
// bucket; every entry names, through its minimizer offset, the one window it could equal.  Called by all lanes.
__device__ __forceinline__ void kmb_mz_late_drain(const KmbProbe &P, const KmbPol &pol, KmbMzShared &S, const KmbStage &st,
                                                  int top, int n, uint64_t kmask, unsigned &counted, unsigned &fetched, int lane) {
    __syncwarp();
    bool more = lane < n;
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;  // the 64 bases of the run's lane
    uint32_t valid = 0, sector = 0;
    int s = 0, e = 0, jpos = 0;
    if (more) {
        const int idx = top - 1 - lane;
        w0 = (uint32_t)S.late_lo[idx], w1 = (uint32_t)(S.late_lo[idx] >> 32);
        w2 = (uint32_t)S.late_hi[idx], w3 = (uint32_t)(S.late_hi[idx] >> 32);
        valid = S.late_valid[idx];
        sector = S.late_sector[idx];
        const uint32_t sej = S.late_sej[idx];
        s = (int)(sej & 255u), e = (int)((sej >> 8) & 255u), jpos = (int)(sej >> 16);
    }
    fetched += (unsigned)n;
    // the two entries of one pool sector against the run
    auto match2 = [&](const uint32_t (&r)[8], uint32_t left) {
#pragma unroll
        for (uint32_t t = 0; t < 2u; t++) {
            if (t >= left) break;
            const int i = jpos - (int)KMB_MZ_POOL_OFFSET(r[0], t);
            if (i < s || i >= e || !((valid >> i) & 1u)) continue;
            const uint64_t km = kmb_mz_window_of(w0, w1, w2, w3, i, kmask);
            if (r[KMB_LINE_KEY_WORD0 + 2 * t] != (uint32_t)km || r[KMB_LINE_KEY_WORD0 + 2 * t + 1] != (uint32_t)(km >> 32)) continue;
            const uint32_t freq = t ? (r[KMB_LINE_FREQ_WORD] >> 16) : (r[KMB_LINE_FREQ_WORD] & 0xFFFFu);
            if ((int32_t)freq > P.max_freq) continue;
            kmb_emit(P, st, r[KMB_LINE_NODE_WORD0 + t]);
            counted++;
        }
    };
    // The first sector says how many entries the chain holds; the sectors are contiguous, so from then on two are
    // fetched per round, independently of each other.  Hits are staged as usual; a bin that fills up before the
    // flush at the end sends its hits straight to the counts (kmb_emit).
    uint32_t left = 0;
    if (more) {
        uint32_t r[8];
        KMB_BOUND(2, sector, P.n_lines);
        kmb_ld_sector(P.lines + (uint64_t)sector * KMB_LINE_WORDS, r, pol.line);
        left = KMB_MZ_POOL_LEFT(r[0]);
        match2(r, left);
        left = left > 2u ? left - 2u : 0u;
        sector++;
    }
    while (__any_sync(KMB_FULL_MASK, left != 0u)) {
        if (left) {
            uint32_t r[8], q[8];
            KMB_BOUND(2, sector + (left > 2u ? 1u : 0u), P.n_lines);
            kmb_ld_sector(P.lines + (uint64_t)sector * KMB_LINE_WORDS, r, pol.line);
            if (left > 2u) kmb_ld_sector(P.lines + (uint64_t)(sector + 1u) * KMB_LINE_WORDS, q, pol.line);
            match2(r, left);
            if (left > 2u) match2(q, left - 2u);
            left = left > 4u ? left - 4u : 0u;
            sector += 2u;
        }
    }
    kmb_stage_flush(P, st, lane, false);
}

template <bool FILT>
__global__ void __launch_bounds__(KMB_MZ_THREADS, 7)
kmb_map_reads_mz_kernel(const uint8_t *__restrict__ bases, uint64_t n_bases, uint64_t base0,
                        KmbReads R, uint32_t in_mode, KmbProbe P, KmbProbe Pkey, KmbStatus *status) {
    extern __shared__ __align__(16) unsigned char kmb_mz_smem[];
    KmbMzShared &S = reinterpret_cast<KmbMzShared *>(kmb_mz_smem)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool n_to_a = (in_mode & KMB_IN_N_TO_A) != 0u;
    const bool packed = (in_mode & KMB_IN_PACKED) != 0u;
    const uint32_t *__restrict__ words = reinterpret_cast<const uint32_t *>(bases);
    const uint64_t n_words = (n_bases + 15) / 16 + 4;
    const KmbStage st = {S.stage_cnt, S.stage, S.stage_res, S.stage_cnt + KMB_MZ_LOG_BINS, KMB_MZ_STAGE_SLOTS, KMB_MZ_LOG_BINS};
    kmb_stage_init(st, lane, P.log.cap == 0);
    unsigned counted = 0, fetched = 0;
    const KmbPol pol = kmb_make_policies(P.policies);
    const int k = KMB_MZ_K;
    const uint64_t kmask = kmb_kmer_mask(k);
    const uint64_t n_tiles = (n_bases + KMB_WTILE_POS - 1) / KMB_WTILE_POS;
    const uint64_t n_vec_full = n_bases / 16;
    const uint64_t warp_stride = (uint64_t)gridDim.x * (KMB_MZ_THREADS / 32);
    unsigned long long mapped = 0;
    uint32_t *pack = S.pack;
    int late_n = 0;  // warp-uniform
    const uint32_t lanemask_lt = (1u << lane) - 1u;

    const uint64_t first_tile = (uint64_t)blockIdx.x * (KMB_MZ_THREADS / 32) + warp;
    uint32_t r_next = first_tile < n_tiles ? R.tile_read[first_tile] : 0u;
    for (uint64_t tile = first_tile; tile < n_tiles; tile += warp_stride) {
        const uint64_t t0 = tile * KMB_WTILE_POS;
        const uint64_t r0 = r_next;
        if (tile + warp_stride < n_tiles) r_next = R.tile_read[tile + warp_stride];
        int64_t rs, re;
        kmb_read_span(R, r0 + (uint64_t)lane, rs, re);
        __syncwarp();
        // ---- 1. load + encode
#pragma unroll 1
        for (int i = lane; i < KMB_WTILE_POS / 16 + 2; i += 32) {
            uint64_t v = t0 / 16 + (uint64_t)i;
            if (packed) {
                pack[i] = v < n_words ? kmb_ldg_u32_hint(words + v, pol.first) : 0u;
                continue;
            }
            uint4 w = kmb_load_bases16(bases, v, n_vec_full, n_bases, pol.first);
            uint32_t inv;
            pack[i] = kmb_encode16(w.x, w.y, w.z, w.w, n_to_a, inv);
            if (inv) atomicMin(&status->first_bad_offset, (unsigned long long)(base0 + v * 16 + (uint64_t)(__ffs(inv) - 1)));
        }
        if (lane < 2) pack[KMB_WTILE_POS / 16 + 2 + lane] = 0u;  // read (not used) by the window extraction of the last positions
        __syncwarp();
        // ---- 2. this lane's 32 windows: which exist, where their minimizers sit, where runs start
        const uint32_t valid = kmb_tile_valid_starts(R, tile, n_bases, k, S.valid, lane, status, r0, rs, re);
        mapped += __popc(valid);
        if (!__any_sync(KMB_FULL_MASK, valid != 0u)) continue;
        __syncwarp();
        S.valid[lane] = valid;
        uint32_t startbits = 0;
        {
            const uint2 qa = *reinterpret_cast<const uint2 *>(&pack[2 * lane]);
            const uint2 qb = *reinterpret_cast<const uint2 *>(&pack[2 * lane + 2]);
            const uint32_t w[4] = {qa.x, qa.y, qb.x, qb.y};
            uint32_t prev = 0;
#pragma unroll 1
            for (int h = 0; h < 2; h++) kmb_mz_half(w, h, &S.mzpos[0][lane], prev, startbits);
            startbits |= 1u;
        }
#pragma unroll 1
        for (int pass_no = 0; pass_no < 2; pass_no++) {
            __syncwarp();
            // ---- 2b. the tile's run list: lanes append their runs (those with at least one existing window)
            unsigned n_runs;
            {
                // Runs of this lane that contain an existing window, without a loop: smear the valid bits down to the
                // start of their run (a bit moves d places down unless a run starts in between), then keep the start bits.
                uint32_t keep = 0;  // bit s set <=> the run starting at s has an existing window
                if ((lane >> 4) == pass_no) {
                    uint32_t d = valid, m = ~(startbits >> 1);  // m: bit i may receive from bit i+1 (no run starts at i+1)
                    d |= (d >> 1) & m;
                    m &= m >> 1;
                    d |= (d >> 2) & m;
                    m &= m >> 2;
                    d |= (d >> 4) & m;
                    m &= m >> 4;
                    d |= (d >> 8) & m;
                    m &= m >> 8;
                    d |= (d >> 16) & m;
                    keep = d & startbits;
                }
                const uint32_t mine = __popc(keep);
                uint32_t incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(KMB_FULL_MASK, incl, o);
                    if (lane >= o) incl += v;
                }
                n_runs = __shfl_sync(KMB_FULL_MASK, incl, 31);
                unsigned at = incl - mine;
                for (uint32_t bits = startbits, kb = keep; kb;) {
                    const int s = __ffs(kb) - 1;
                    kb &= kb - 1u;
                    const uint32_t after = bits & ~((2u << s) - 1u);  // start bits above s
                    const int e = after ? __ffs(after) - 1 : 32;
                    S.a.runs[at++] = KMB_MZ_RUN(lane, s, e, S.mzpos[s][lane]);
                }
            }
            __syncwarp();
            // ---- 3. one filter word per run, 32 runs per round; the runs that pass are kept in slot order
            unsigned n_kept = 0;  // warp-uniform; a tile that exceeds KMB_MZ_RUNS_KEPT takes the slow exit below
#pragma unroll 1
            for (unsigned r0 = 0; r0 < n_runs; r0 += 32u) {
                const unsigned ri = r0 + (unsigned)lane;
                uint32_t run = 0, sector = 0;
                bool pass = false;
                if (ri < n_runs) {
                    run = S.a.runs[ri];
                    const uint32_t p = (run & 31u) * KMB_POS_PER_THREAD + (run >> 15);  // tile position of the minimizer m-mer
                    const uint32_t mmer = __funnelshift_r(pack[p >> 4], pack[(p >> 4) + 1], (p & 15u) * 2u) & KMB_MZ_MASK;
                    const KmbLoc loc = kmb_locate((uint64_t)mmer, P.addr);
                    sector = 2u * loc.sector;
                    if (FILT) {
                        KMB_BOUND(0, loc.fword, P.addr.n_filter_words);
                        const uint32_t fw = kmb_ldg_u32_hint(P.filter + loc.fword, pol.filter);
                        pass = (fw & loc.fmask) == loc.fmask;
                    } else {
                        pass = true;
                    }
                }
                const unsigned m = __ballot_sync(KMB_FULL_MASK, pass);
                if (pass) {
                    const unsigned slot = n_kept + __popc(m & lanemask_lt);
                    if (slot < KMB_MZ_RUNS_KEPT) {
                        S.kept_run[slot] = run;
                        S.kept_sector[slot] = sector;
                    }
                }
                n_kept += __popc(m);
            }
            if (n_kept > KMB_MZ_RUNS_KEPT) {
                // A pass with more passing runs than can be remembered (a minimizer change at almost every window: never
                // seen on real or synthetic reads; the checked build forces it).  Forget the runs: every existing window
                // of the pass's 16 lanes goes through the key-addressed sectors, one at a time.
                const uint2 qa = *reinterpret_cast<const uint2 *>(&pack[2 * lane]);
                const uint2 qb = *reinterpret_cast<const uint2 *>(&pack[2 * lane + 2]);
                const uint64_t lo = (uint64_t)qa.x | ((uint64_t)qa.y << 32), hi = (uint64_t)qb.x | ((uint64_t)qb.y << 32);
#pragma unroll 1
                for (uint32_t vb = (lane >> 4) == pass_no ? valid : 0u; vb; vb &= vb - 1u) {
                    kmb_walk_one(Pkey, pol, kmb_window(lo, hi, __ffs(vb) - 1, kmask), [&](uint32_t node, uint32_t freq) {
                        if ((int32_t)freq <= Pkey.max_freq) {
                            KMB_BOUND(6, node, Pkey.n_counts);
                            kmb_count_direct(Pkey.counts, node);
                            counted++;
                        }
                        return false;
                    });
                    fetched++;
                }
                continue;
            }
            __syncwarp();  // the run list is no longer needed: its space becomes the staging slots
            // ---- 4. fetch: primaries in one burst, then the secondaries of the buckets with more than two entries
            const unsigned n_have = n_kept;
            const unsigned n_staged = min(n_have, (unsigned)KMB_MZ_SLOTS);
            fetched += n_have;
#ifndef KMB_MZ_LDG_STAGING
            for (unsigned sl = (unsigned)lane; sl < n_staged; sl += 32u) {
                KMB_BOUND(1, S.kept_sector[sl], 2ull * P.addr.n_main);
                kmb_cp_async_sector(S.a.slots[sl], P.lines + (uint64_t)S.kept_sector[sl] * KMB_LINE_WORDS);
            }
            kmb_cp_async_wait_all();
#else
            // Alternative (-DKMB_MZ_LDG_STAGING): one 256-bit load per sector into registers, two sectors per lane in
            // flight, then into the staging slots.  Measured on config 3 (24 G k-mers): 336 ms against 317 ms for the
            // cp.async burst above, which keeps ~52 fetches per warp in flight without holding registers.
#pragma unroll 1
            for (unsigned base = 0; base < n_staged; base += 64u) {
                const unsigned s0 = base + (unsigned)lane, s1 = s0 + 32u;
                uint32_t r0[8], r1[8];
                if (s0 < n_staged) {
                    KMB_BOUND(1, S.kept_sector[s0], 2ull * P.addr.n_main);
                    kmb_ld_sector(P.lines + (uint64_t)S.kept_sector[s0] * KMB_LINE_WORDS, r0, pol.line);
                }
                if (s1 < n_staged) {
                    KMB_BOUND(1, S.kept_sector[s1], 2ull * P.addr.n_main);
                    kmb_ld_sector(P.lines + (uint64_t)S.kept_sector[s1] * KMB_LINE_WORDS, r1, pol.line);
                }
                if (s0 < n_staged) {
                    *reinterpret_cast<uint4 *>(&S.a.slots[s0][0]) = make_uint4(r0[0], r0[1], r0[2], r0[3]);
                    *reinterpret_cast<uint4 *>(&S.a.slots[s0][4]) = make_uint4(r0[4], r0[5], r0[6], r0[7]);
                }
                if (s1 < n_staged) {
                    *reinterpret_cast<uint4 *>(&S.a.slots[s1][0]) = make_uint4(r1[0], r1[1], r1[2], r1[3]);
                    *reinterpret_cast<uint4 *>(&S.a.slots[s1][4]) = make_uint4(r1[4], r1[5], r1[6], r1[7]);
                }
            }
#endif
            __syncwarp();
            {
                unsigned n2 = 0;
#pragma unroll 1
                for (unsigned base = 0; base < n_staged; base += 32u) {
                    const unsigned sl = base + (unsigned)lane;
                    const bool more = sl < n_staged && KMB_MZ_HDR_COUNT(S.a.slots[sl][0]) > 2u;
                    const unsigned m = __ballot_sync(KMB_FULL_MASK, more);
                    if (sl < n_staged) {
                        uint8_t tag = KMB_MZ_NONE;
                        if (more) {
                            const unsigned j = n2 + __popc(m & lanemask_lt);
                            if (j < KMB_MZ_SLOTS2) {
#ifndef KMB_MZ_LDG_STAGING
                                kmb_cp_async_sector(S.slots2[j], P.lines + (uint64_t)(S.kept_sector[sl] + 1u) * KMB_LINE_WORDS);
#else
                                uint32_t r2[8];  // the other half of the 64 bytes the primary's fetch brought into the L2
                                kmb_ld_sector(P.lines + (uint64_t)(S.kept_sector[sl] + 1u) * KMB_LINE_WORDS, r2, pol.line);
                                *reinterpret_cast<uint4 *>(&S.slots2[j][0]) = make_uint4(r2[0], r2[1], r2[2], r2[3]);
                                *reinterpret_cast<uint4 *>(&S.slots2[j][4]) = make_uint4(r2[4], r2[5], r2[6], r2[7]);
#endif
                                tag = (uint8_t)j;
                            } else {
                                tag = KMB_MZ_LATE;
                            }
                        }
                        S.slot2_of[sl] = tag;
                    }
                    n2 += __popc(m);
                }
#ifndef KMB_MZ_LDG_STAGING
                if (n2) kmb_cp_async_wait_all();
#endif
            }
            __syncwarp();
            // ---- 5. 32 kept runs per round, one per lane; per entry of the run's bucket the one window that could match
#pragma unroll 1
            for (unsigned s0 = 0; s0 < n_have; s0 += 32u) {
                const unsigned sl = s0 + (unsigned)lane;
                uint32_t pool_sector = 0, pool_run = 0;
                if (sl < n_have) {
                    const uint32_t run = S.kept_run[sl];
                    const int lane_base = (int)(run & 31u) * KMB_POS_PER_THREAD;
                    const int ps = lane_base + (int)((run >> 5) & 31u), pe = lane_base + (int)((run >> 10) & 31u) + 1;
                    const int jp = lane_base + (int)(run >> 15);
                    const uint32_t vrow = S.valid[run & 31u];
                    uint32_t r[8];
                    uint32_t t2;
                    if (sl < n_staged) {
                        const uint4 x = *reinterpret_cast<const uint4 *>(&S.a.slots[sl][0]);
                        const uint4 y = *reinterpret_cast<const uint4 *>(&S.a.slots[sl][4]);
                        r[0] = x.x, r[1] = x.y, r[2] = x.z, r[3] = x.w, r[4] = y.x, r[5] = y.y, r[6] = y.z, r[7] = y.w;
                        t2 = S.slot2_of[sl];
                    } else {  // more runs passed than there are staging slots: load the bucket now
                        kmb_ld_sector(P.lines + (uint64_t)S.kept_sector[sl] * KMB_LINE_WORDS, r, pol.line);
                        t2 = KMB_MZ_LATE;
                    }
                    const uint32_t hdr = r[0];
                    const uint32_t n_entries = KMB_MZ_HDR_COUNT(hdr);
                    kmb_mz_match_pair(P, st, r, KMB_MZ_HDR_OFFSET(hdr, 0), KMB_MZ_HDR_OFFSET(hdr, 1), n_entries, jp, ps, pe, lane_base,
                                      vrow, pack, kmask, counted);
                    if (n_entries > 2u) {
                        if (t2 < KMB_MZ_SLOTS2) {
                            const uint4 x = *reinterpret_cast<const uint4 *>(&S.slots2[t2][0]);
                            const uint4 y = *reinterpret_cast<const uint4 *>(&S.slots2[t2][4]);
                            r[0] = x.x, r[1] = x.y, r[2] = x.z, r[3] = x.w, r[4] = y.x, r[5] = y.y, r[6] = y.z, r[7] = y.w;
                        } else {
                            kmb_ld_sector(P.lines + (uint64_t)(S.kept_sector[sl] + 1u) * KMB_LINE_WORDS, r, pol.line);
                        }
                        kmb_mz_match_pair(P, st, r, KMB_MZ_HDR_OFFSET(hdr, 2), KMB_MZ_HDR_OFFSET(hdr, 3), n_entries - 2u, jp, ps, pe,
                                          lane_base, vrow, pack, kmask, counted);
                        if (n_entries > 4u) {
                            pool_sector = r[0] & ~KMB_HDR_CHAIN;
                            pool_run = run;
                        }
                    }
                }
                kmb_stage_flush(P, st, lane, false);
                const unsigned pm = __ballot_sync(KMB_FULL_MASK, pool_sector != 0u);
                if (pm) {
                    if (pool_sector) {
                        const int idx = late_n + __popc(pm & lanemask_lt);
                        KMB_BOUND(7, idx, KMB_MZ_LATE_CAP);
                        const uint32_t L = pool_run & 31u;
                        const uint2 qa = *reinterpret_cast<const uint2 *>(&pack[2 * L]);
                        const uint2 qb = *reinterpret_cast<const uint2 *>(&pack[2 * L + 2]);
                        S.late_lo[idx] = (unsigned long long)qa.x | ((unsigned long long)qa.y << 32);
                        S.late_hi[idx] = (unsigned long long)qb.x | ((unsigned long long)qb.y << 32);
                        S.late_valid[idx] = S.valid[L];
                        S.late_sector[idx] = pool_sector;
                        S.late_sej[idx] = ((pool_run >> 5) & 31u) | ((((pool_run >> 10) & 31u) + 1u) << 8) | ((pool_run >> 15) << 16);
                    }
                    late_n += __popc(pm);
                    if (late_n > 8) {
                        const int n = min(late_n, 32);
                        kmb_mz_late_drain(P, pol, S, st, late_n, n, kmask, counted, fetched, lane);
                        late_n -= n;
                    }
                    __syncwarp();
                }
            }
    
        }
    }
    // ---- the rest
    if (late_n) kmb_mz_late_drain(P, pol, S, st, late_n, late_n, kmask, counted, fetched, lane);
    kmb_stage_flush(P, st, lane, true);
    for (int o = 16; o > 0; o >>= 1) {
        counted += __shfl_xor_sync(KMB_FULL_MASK, counted, o);
        mapped += __shfl_xor_sync(KMB_FULL_MASK, mapped, o);
    }
    if (lane == 0 && counted) atomicAdd(&status->n_entries_counted, (unsigned long long)counted);
    if (lane == 0 && fetched) atomicAdd(&status->n_candidates, (unsigned long long)fetched);  // approximate: bucket and pool fetches
    if (lane == 0 && mapped) atomicAdd(&status->n_kmers_mapped, mapped);
}

// ================================================================================================
// K3-4 on ready-made k-mers (drop-in for map_kmers_to_graph_index, mapper.pyx:19-72).
// Coalesced 8-byte loads, U filter loads in flight per thread, same probe.
// ================================================================================================
template <int U, bool FILT, bool REVCOMP>
__global__ void __launch_bounds__(KMB_MAP_THREADS, KMB_MAP_BLOCKS)
kmb_map_kmers_kernel(const uint64_t *__restrict__ kmers, uint64_t n, int k, KmbProbe P, KmbStatus *status) {
    extern __shared__ __align__(16) unsigned char kmb_map_smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    KmbWarpShared<U> &S = reinterpret_cast<KmbWarpShared<U> *>(kmb_map_smem)[warp];
    uint64_t *q_kmer = S.qk;
    uint32_t *q_h = S.qh;
    const KmbStage st = {S.stage_cnt, S.stage, S.stage_res, S.stage_cnt + KMB_LOG_BINS, (uint32_t)KmbWarpShared<U>::kStageSlots, (uint32_t)KMB_LOG_BINS};
    kmb_stage_init(st, lane, P.log.cap == 0);
    unsigned counted = 0;
    int qcount = 0;
    KmbPipe pp;
    kmb_pipe_init(pp);
    const KmbPol pol = kmb_make_policies(P.policies);
    const uint64_t per_block = (uint64_t)KMB_MAP_THREADS * U;
    const uint64_t n_blocks = (n + per_block - 1) / per_block;
    for (uint64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        // warp w of the CTA owns 32*U consecutive k-mers; lane l takes elements l, l+32, ... (coalesced)
        uint64_t base = blk * per_block + (uint64_t)warp * (32 * U) + (uint64_t)lane;
        uint64_t km[U];
        uint32_t vb = 0;
#pragma unroll
        for (int u = 0; u < U; u++) {
            uint64_t i = base + (uint64_t)u * 32;
            bool in = i < n;
            km[u] = in ? kmb_ldg_u64_hint(kmers + i, pol.first) : 0ull;
            vb |= in ? (1u << u) : 0u;
        }
        KmbArrayFn fa = {km};
        kmb_probe_batch<U, FILT>(P, pol, pp, st, counted, fa, vb, q_kmer, q_h, qcount, lane);
        if (REVCOMP) {
#pragma unroll
            for (int u = 0; u < U; u++) km[u] = kmb_revcomp(km[u], k);
            kmb_probe_batch<U, FILT>(P, pol, pp, st, counted, fa, vb, q_kmer, q_h, qcount, lane);
        }
    }
    kmb_pipe_finish(P, pol, pp, st, counted, q_kmer, q_h, qcount, lane, status);
    if (blockIdx.x == 0 && tid == 0) atomicAdd(&status->n_kmers_mapped, REVCOMP ? 2ull * n : (unsigned long long)n);
}

// ------------------------------------------------------------------------------------------------
// One query per thread without the warp stack: the cross-check variant of the mapping kernels
// (kmb_set_option("probe_variant", 0)), the membership kernel and the per-key lookup.
// ------------------------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ void kmb_walk_one(const KmbProbe &P, const KmbPol &pol, uint64_t km, F on_match) {
    const KmbLoc loc = kmb_locate(km, P.addr);
    if (P.filter != nullptr) {
        KMB_BOUND(0, loc.fword, P.addr.n_filter_words);
        uint32_t w = kmb_ldg_u32_hint(P.filter + loc.fword, pol.filter);
        if ((w & loc.fmask) != loc.fmask) return;
    }
    kmb_probe_line(P, pol, km, loc.sector, on_match);
}

template <bool REVCOMP>
__global__ void kmb_map_kmers_simple_kernel(const uint64_t *__restrict__ kmers, uint64_t n, int k, KmbProbe P,
                                            KmbStatus *status) {
    const KmbPol pol = kmb_make_policies(P.policies);
    unsigned long long counted = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t km = kmers[i];
#pragma unroll
        for (int strand = 0; strand < (REVCOMP ? 2 : 1); strand++) {
            if (strand == 1) km = kmb_revcomp(km, k);
            kmb_walk_one(P, pol, km, [&](uint32_t node, uint32_t freq) {
                if ((int32_t)freq <= P.max_freq) {
                    KMB_BOUND(6, node, P.n_counts);
                    kmb_count_direct(P.counts, node);
                    counted++;
                }
                return false;
            });
        }
    }
    if (counted) atomicAdd(&status->n_entries_counted, counted);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&status->n_kmers_mapped, REVCOMP ? 2ull * n : (unsigned long long)n);
}

// ================================================================================================
// K4b: play the logged hits into the node counts (mapper.pyx:68), one WINDOW of 2^win_shift nodes at a
// time, so that the window being reduced into stays in L2; the log itself is a coalesced stream.
// One launch: blockIdx.y = window.  CTAs are dispatched x-fastest, so the windows are worked off in
// order (the last CTAs of one overlap the first of the next, which fills the tails).  A group is
// tagged with its node range (bin = node >> bin_shift, one of KMB_LOG_BINS); a range wider than a
// window (large count arrays: 400 M nodes = 6 ranges of 268 MB) is played in 2^(bin_shift - win_shift)
// windows, each re-reading the range's groups and keeping the ids that fall inside it -- a few extra
// coalesced passes over the log instead of reductions that miss the L2.
//
// Count accumulation with aggregated atomics: when a few nodes are very hot (config 4: Zipf), their
// reductions serialise on one L2 atomic unit.  Every CTA therefore walks a contiguous slab of the log
// and first tries to count an id in a small shared-memory table (first id to claim a slot keeps it;
// later ids that collide go straight to global memory, grouped by match.any); the table is flushed
// with ONE reduction per id at the end.  A warp that finds less than 1 in 16 of its first 256 ids
// already present switches the table off, so uniformly distributed nodes (config 2) pay almost nothing.
// ================================================================================================
#define KMB_APPLY_TABLE 2048
#ifndef KMB_APPLY_UNROLL
#define KMB_APPLY_UNROLL 8   // groups of 32 ids a warp has in flight
#endif
__global__ void __launch_bounds__(256) kmb_log_apply_kernel(KmbLog log, uint32_t *__restrict__ counts) {
    __shared__ uint32_t s_id[KMB_APPLY_TABLE];
    __shared__ uint32_t s_cnt[KMB_APPLY_TABLE];
    for (int i = threadIdx.x; i < KMB_APPLY_TABLE; i += blockDim.x) {
        s_id[i] = KMB_LOG_HOLE;
        s_cnt[i] = 0;
    }
    __syncthreads();
    // The log is a stream that is read once per window and must not push the window's counters out of the L2 (with
    // default priorities 52 % of the reductions missed and 7.2 GB of dirty sectors were written back per 0.5 G
    // reductions): the stream is evict-first, the counters evict-last.
    const uint64_t pol_stream = kmb_policy_evict_first(), pol_keep = kmb_policy_evict_last();
    const uint32_t window = blockIdx.y;
    const uint32_t sub_shift = log.bin_shift - log.win_shift;
    const uint32_t bin = min(window >> sub_shift, log.n_bins - 1u);
    // the last range takes every node beyond it (kmb_log_bin clamps), so only a window that is not the last one of
    // its range has to look at the ids; with one window per range nothing does
    const bool filter_ids = sub_shift != 0u;
    const bool last_window = window + 1u == gridDim.y;
    unsigned long long n = log.cursor[0];
    if (n > log.cap) n = log.cap;
    const uint64_t n_groups = n >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    // slab of this CTA: a multiple of 128 groups, so that a warp reads 128 tags with one coalesced 4-byte load per lane
    const uint64_t per_cta = (((n_groups + gridDim.x - 1) / gridDim.x) + 127) & ~127ull;
    const uint64_t g_lo = (uint64_t)blockIdx.x * per_cta;
    const uint64_t g_hi = min(g_lo + per_cta, n_groups);
    const uint32_t *__restrict__ tags4 = reinterpret_cast<const uint32_t *>(log.tags);  // capacity is a multiple of 4096 ids
    bool use_table = true, decided = false;
    uint32_t seen = 0, present = 0;
    const uint64_t g_first = g_lo + (uint64_t)warp * 128, g_step = (uint64_t)warps * 128;
    // the tags of the NEXT 128 groups are requested before this block's groups are played, and the groups of a block
    // that belong to the window are fetched KMB_APPLY_UNROLL at a time: their DRAM latencies overlap instead of adding up
    // (one group per warp in flight made the pass latency-bound at 0.4 TB/s)
    uint32_t t4_next = (g_first < g_hi && g_first + 4ull * lane < g_hi) ? kmb_ldg_u32_hint(tags4 + ((g_first + 4ull * lane) >> 2), pol_stream) : 0xFFFFFFFFu;
    for (uint64_t g0 = g_first; g0 < g_hi; g0 += g_step) {
        const uint64_t gl = g0 + 4ull * lane;  // this lane's four groups
        if (gl < g_hi) KMB_BOUND(10, gl >> 2, log.cap >> 7);
        const uint32_t t4 = t4_next;
        {
            const uint64_t gn = g0 + g_step + 4ull * lane;
            t4_next = gn < g_hi ? kmb_ldg_u32_hint(tags4 + (gn >> 2), pol_stream) : 0xFFFFFFFFu;
        }
        // bit p of (hi:lo) <=> group g0 + 4 (p % 32) + p / 32 belongs to this window's range
        unsigned long long lo = (unsigned long long)__ballot_sync(KMB_FULL_MASK, gl + 0 < g_hi && ((t4 >> 0) & 0xFFu) == bin) |
                                ((unsigned long long)__ballot_sync(KMB_FULL_MASK, gl + 1 < g_hi && ((t4 >> 8) & 0xFFu) == bin) << 32);
        unsigned long long hi = (unsigned long long)__ballot_sync(KMB_FULL_MASK, gl + 2 < g_hi && ((t4 >> 16) & 0xFFu) == bin) |
                                ((unsigned long long)__ballot_sync(KMB_FULL_MASK, gl + 3 < g_hi && ((t4 >> 24) & 0xFFu) == bin) << 32);
        while (lo | hi) {
            uint32_t ids[KMB_APPLY_UNROLL];
#pragma unroll
            for (int q = 0; q < KMB_APPLY_UNROLL; q++) {
                int p = -1;
                if (lo) {
                    p = __ffsll((long long)lo) - 1;
                    lo &= lo - 1ull;
                } else if (hi) {
                    p = 64 + __ffsll((long long)hi) - 1;
                    hi &= hi - 1ull;
                }
                ids[q] = KMB_LOG_HOLE;
                if (p >= 0) {
                    const uint64_t g = g0 + 4ull * (uint64_t)(p & 31) + (uint64_t)(p >> 5);
                    KMB_BOUND(11, (g << 5) + lane, log.cap);
                    ids[q] = kmb_ldg_u32_hint(log.entries + ((g << 5) + lane), pol_stream);
                }
            }
#pragma unroll
            for (int q = 0; q < KMB_APPLY_UNROLL; q++) {
                uint32_t id = ids[q];
                if (filter_ids && id != KMB_LOG_HOLE) {
                    const uint32_t w = id >> log.win_shift;
                    if (!(w == window || (last_window && w > window))) id = KMB_LOG_HOLE;
                }
                if (id != KMB_LOG_HOLE) {
                    if (use_table) {
#ifdef KMB_APPLY_CAS_FIRST
                        const uint32_t s = (id * 0x9E3779B1u) >> 21;  // 11 bits
                        const uint32_t old = atomicCAS(&s_id[s], KMB_LOG_HOLE, id);
                        if (old == KMB_LOG_HOLE || old == id) atomicAdd(&s_cnt[s], 1u);
                        else kmb_count_direct(counts, id);
                        present += old == id ? 1u : 0u;
#else
                        // Two candidate slots (s, s ^ 1), looked at with plain loads first: a claimed slot never changes, so
                        // the compare-and-swap is only needed while a slot is still free -- the hot ids, which are what the
                        // table is for, cost one shared load and one shared add.  (With the CAS in front of every id the CAS
                        // and the wait for its result were a third of this kernel's stall samples on the Zipf nodes of
                        // config 4: 12.8 -> 9.3 ms per 3.4 G k-mers at k = 15, 6.0 -> 4.4 at k = 21; uniform nodes switch the
                        // table off and do not notice.)  Two slots instead of one: two hot ids that hash alike both find room.
                        uint32_t s = (id * 0x9E3779B1u) >> 21;  // 11 bits
                        uint32_t old = s_id[s];
                        if (old != id) {
                            const uint32_t old1 = s_id[s ^ 1u];
                            if (old1 == id) {
                                s ^= 1u, old = id;
                            } else if (old == KMB_LOG_HOLE) {
                                old = atomicCAS(&s_id[s], KMB_LOG_HOLE, id);
                            } else if (old1 == KMB_LOG_HOLE) {
                                s ^= 1u;
                                old = atomicCAS(&s_id[s], KMB_LOG_HOLE, id);
                            }
                        }
                        if (old == KMB_LOG_HOLE || old == id) atomicAdd(&s_cnt[s], 1u);
                        else kmb_red_add_hint(counts + id, 1u, pol_keep);
                        present += old == id ? 1u : 0u;
#endif
                        seen++;
                    } else {
                        kmb_red_add_hint(counts + id, 1u, pol_keep);
                    }
                }
                if (use_table && !decided && __any_sync(KMB_FULL_MASK, seen >= 8u)) {  // ~256 ids per warp looked at: decide once
                    uint32_t pp = present, t = seen;
                    for (int o = 16; o > 0; o >>= 1) {
                        pp += __shfl_xor_sync(KMB_FULL_MASK, pp, o);
                        t += __shfl_xor_sync(KMB_FULL_MASK, t, o);
                    }
                    if (pp * 16u < t) use_table = false;
                    decided = true;
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < KMB_APPLY_TABLE; i += blockDim.x)
        if (s_cnt[i]) kmb_red_add_hint(counts + s_id[i], s_cnt[i], pol_keep);
}
// Empty the log: untag the groups that were used, rewind the cursor.
__global__ void kmb_log_reset_kernel(KmbLog log) {
    unsigned long long n = log.cursor[0];
    if (n > log.cap) n = log.cap;
    const uint64_t n_groups = (n + 31) >> 5;
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n_groups; g += (uint64_t)gridDim.x * blockDim.x)
        log.tags[g] = KMB_LOG_NO_BIN;
}
__global__ void kmb_log_rewind_kernel(KmbLog log) { log.cursor[0] = 0ull; }

// ================================================================================================
// K6 membership (mapper.pyx:81-130): any key match, frequency ignored.
// MODE 0: out_u8[i] = hit.  MODE 1: out_u32[i] = counts[node of a matching slot] (Counter.__getitem__;
// the counter's keys are unique, so "a" matching slot is "the" slot).
// ================================================================================================
template <int MODE>
__global__ void kmb_in_graph_kernel(const uint64_t *__restrict__ kmers, uint64_t n, KmbProbe P, uint8_t *out8,
                                    uint32_t *out32) {
    const KmbPol pol = kmb_make_policies(P.policies);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        bool hit = false;
        uint32_t node = 0;
        kmb_walk_one(P, pol, kmers[i], [&](uint32_t nd, uint32_t) {
            hit = true;
            node = nd;
            return true;
        });
        if (MODE == 0) out8[i] = hit ? 1 : 0;
        else {
            if (hit) KMB_BOUND(6, node, P.n_counts);
            out32[i] = hit ? P.counts[node] : 0u;
        }
    }
}

// ================================================================================================
// K2 standalone: flat hash array (get_kmer_hashes_from_chunk_sequence, util.py:71-75).
// count -> scan -> emit; tiles and per-thread ownership exactly as in the fused kernel.
// ================================================================================================
__global__ void __launch_bounds__(KMB_TILE_THREADS)
kmb_hash_count_kernel(const uint32_t *__restrict__ mask, uint64_t n_bases, int k, unsigned long long *tile_counts) {
    __shared__ uint32_t s_w[KMB_TILE_THREADS / 32];
    uint64_t tile = blockIdx.x;
    uint64_t p0 = tile * KMB_TILE_POS + (uint64_t)threadIdx.x * KMB_POS_PER_THREAD;
    uint32_t c = __popc(kmb_valid_starts(mask, p0, n_bases, k));
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(KMB_FULL_MASK, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < KMB_TILE_THREADS / 32; w++) t += s_w[w];
        tile_counts[tile] = t;
    }
}

// single-CTA exclusive scan of tile_counts (n_tiles <= a few million); total written to *total
__global__ void __launch_bounds__(1024) kmb_hash_scan_kernel(unsigned long long *tile_counts, uint64_t n_tiles,
                                                             unsigned long long *total) {
    __shared__ unsigned long long s_part[1024];
    const int t = threadIdx.x;
    const uint64_t per = (n_tiles + 1023) / 1024;
    const uint64_t lo = (uint64_t)t * per;
    const uint64_t hi = lo + per < n_tiles ? lo + per : n_tiles;
    unsigned long long sum = 0;
    for (uint64_t i = lo; i < hi; i++) sum += tile_counts[i];
    s_part[t] = sum;
    __syncthreads();
    if (t == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; i++) {
            unsigned long long v = s_part[i];
            s_part[i] = run;
            run += v;
        }
        *total = run;
    }
    __syncthreads();
    unsigned long long run = s_part[t];
    for (uint64_t i = lo; i < hi; i++) {
        unsigned long long v = tile_counts[i];
        tile_counts[i] = run;
        run += v;
    }
}

// Emit: thread t first owns the 32 window starts [32t, 32t+32) of the tile (one mask word) to build the
// per-word output offsets, then the warps walk the tile 32 consecutive positions at a time: the 32 lanes
// share two 64-bit words of the packed stream (shared-memory broadcast), compact their valid windows with one
// popcount of the mask word and store them to consecutive addresses -- the output, 8 bytes per k-mer, is the
// traffic that bounds this kernel, so it leaves as full coalesced lines.
__global__ void __launch_bounds__(KMB_TILE_THREADS)
kmb_hash_emit_kernel(const uint8_t *__restrict__ bases, uint64_t n_bases, const uint32_t *__restrict__ mask, int k,
                     bool n_to_a, const unsigned long long *__restrict__ tile_offsets, uint64_t *__restrict__ out,
                     uint64_t out_capacity, KmbStatus *status) {
    __shared__ __align__(16) uint32_t s_pack[KMB_TILE_POS / 16 + 8];
    __shared__ uint32_t s_valid[KMB_TILE_THREADS];  // valid-start bits of word t
    __shared__ uint32_t s_off[KMB_TILE_THREADS];    // windows of the tile before word t
    __shared__ uint32_t s_w[KMB_TILE_THREADS / 32];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const uint64_t tile = blockIdx.x;
    const uint64_t t0 = tile * KMB_TILE_POS;
    const uint64_t n_vec_full = n_bases / 16;
    const uint64_t kmask = kmb_kmer_mask(k);
    const uint64_t pol_first = kmb_policy_evict_first();
    for (int i = tid; i < KMB_TILE_POS / 16 + 2; i += KMB_TILE_THREADS) {
        uint64_t v = t0 / 16 + (uint64_t)i;
        uint4 w = kmb_load_bases16(bases, v, n_vec_full, n_bases, pol_first);
        uint32_t inv;
        s_pack[i] = kmb_encode16(w.x, w.y, w.z, w.w, n_to_a, inv);
        if (inv) atomicMin(&status->first_bad_offset, (unsigned long long)(v * 16 + (uint64_t)(__ffs(inv) - 1)));
    }
    const uint32_t valid = kmb_valid_starts(mask, t0 + (uint64_t)tid * KMB_POS_PER_THREAD, n_bases, k);
    // exclusive prefix of popc(valid) over the CTA
    const uint32_t c = __popc(valid);
    uint32_t incl = c;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(KMB_FULL_MASK, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t warp_base = 0;
    for (int w = 0; w < warp; w++) warp_base += s_w[w];
    s_valid[tid] = valid;
    s_off[tid] = warp_base + (incl - c);
    __syncthreads();
    const uint64_t tile_base = tile_offsets[tile];
    for (int word = warp; word < KMB_TILE_THREADS; word += KMB_TILE_THREADS / 32) {
        const uint32_t vw = s_valid[word];
        if (vw == 0u) continue;
        const uint2 a = *reinterpret_cast<const uint2 *>(&s_pack[2 * word]);
        const uint2 b = *reinterpret_cast<const uint2 *>(&s_pack[2 * word + 2]);
        const uint64_t lo = (uint64_t)a.x | ((uint64_t)a.y << 32);
        const uint64_t hi = (uint64_t)b.x | ((uint64_t)b.y << 32);
        if ((vw >> lane) & 1u) {
            const uint64_t o = tile_base + s_off[word] + __popc(vw & ((1u << lane) - 1u));
            if (o < out_capacity) out[o] = kmb_window(lo, hi, lane, kmask);
        }
    }
}

// ================================================================================================
// E1 legacy codec (encodings.py).  Element-wise, vector width chosen so one thread writes >= 4 B.
// ================================================================================================
// ACTGTwoBitEncoding.from_bytes (encodings.py:51-59): `& 31`, aligned PAIRS through the 64Ki LUT
// (:30-31): low-5-bit values 1,3,20,7 -> 0,1,2,3 (A,C,T,G); a pair with any other value -> 0000.
__device__ __forceinline__ uint32_t kmb_actg_code5(uint32_t v, bool &ok) {
    ok = (v == 1u) | (v == 3u) | (v == 20u) | (v == 7u);
    return v == 3u ? 1u : (v == 20u ? 2u : (v == 7u ? 3u : 0u));
}
__global__ void kmb_codec_actg_from_bytes_kernel(const uint8_t *__restrict__ seq, uint64_t n_out, uint8_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_out; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t w = *reinterpret_cast<const uint32_t *>(seq + 4 * i);
        bool o0, o1, o2, o3;
        uint32_t c0 = kmb_actg_code5(w & 31u, o0), c1 = kmb_actg_code5((w >> 8) & 31u, o1);
        uint32_t c2 = kmb_actg_code5((w >> 16) & 31u, o2), c3 = kmb_actg_code5((w >> 24) & 31u, o3);
        uint32_t lo = (o0 && o1) ? (c0 | (c1 << 2)) : 0u;
        uint32_t hi = (o2 && o3) ? (c2 | (c3 << 2)) : 0u;
        out[i] = (uint8_t)(lo | (hi << 4));
    }
}
// SimpleEncoding.from_bytes (encodings.py:96-102): per-byte LUT a/A,c/C,t/T,g/G -> 0,1,2,3, else 0.
__device__ __forceinline__ uint32_t kmb_simple_code(uint32_t b) {
    uint32_t f = b & 0xDFu;
    return f == 67u ? 1u : (f == 84u ? 2u : (f == 71u ? 3u : 0u));
}
__global__ void kmb_codec_simple_from_bytes_kernel(const uint8_t *__restrict__ seq, uint64_t n_out, uint8_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_out; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t w = *reinterpret_cast<const uint32_t *>(seq + 4 * i);
        out[i] = (uint8_t)(kmb_simple_code(w & 255u) | (kmb_simple_code((w >> 8) & 255u) << 2) |
                           (kmb_simple_code((w >> 16) & 255u) << 4) | (kmb_simple_code(w >> 24) << 6));
    }
}
// to_bytes (encodings.py:70-75): 4 lower-case letters per packed byte, reverse[code] + 96.
__global__ void kmb_codec_to_bytes_kernel(const uint8_t *__restrict__ packed, uint64_t n, uint8_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t b = packed[i];
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t c = (b >> (2 * j)) & 3u;
            uint32_t ch = 96u + (c == 0u ? 1u : (c == 1u ? 3u : (c == 2u ? 20u : 7u)));
            w |= ch << (8 * j);
        }
        *reinterpret_cast<uint32_t *>(out + 4 * i) = w;
    }
}
// complement (encodings.py:44-48): XOR 0xAA per byte.
__global__ void kmb_codec_complement_kernel(const uint8_t *__restrict__ in, uint64_t n, uint8_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = in[i] ^ 0xAAu;
}
// twobit_swap (encodings.py:104-112): reverse all 2-bit groups of each word.
template <typename T>
__global__ void kmb_codec_twobit_swap_kernel(const T *__restrict__ in, uint64_t n, T *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = (uint64_t)in[i];
        x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
        x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
        x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
        x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
        x = (x >> 32) | (x << 32);
        out[i] = (T)(x >> (64 - 8 * sizeof(T)));
    }
}

// ================================================================================================
// B: random-gather micro-roofline.  Each thread issues UNROLL independent loads of W bytes from
// uniformly random W-aligned... (32-byte-sector-aligned) addresses; the XOR of everything loaded is
// written once so the loads cannot be elided.
// ================================================================================================
// 8-byte load variants for the fetch-granularity experiment: mode 0 plain .nc, 1/2/3 = L2::64B/128B/256B
// prefetch-size qualifier, 4 = plain coherent ld.global, 5 = ld.global.cv (volatile-like, "don't cache")
__device__ __forceinline__ uint64_t kmb_ldg_u64_mode(const uint64_t *p, int mode) {
    uint64_t v;
    switch (mode) {
        case 1: asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p)); break;
        case 2: asm volatile("ld.global.nc.L1::no_allocate.L2::128B.u64 %0, [%1];" : "=l"(v) : "l"(p)); break;
        case 3: asm volatile("ld.global.nc.L1::no_allocate.L2::256B.u64 %0, [%1];" : "=l"(v) : "l"(p)); break;
        case 4: asm volatile("ld.global.u64 %0, [%1];" : "=l"(v) : "l"(p)); break;
        case 5: asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p)); break;
        default: asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p)); break;
    }
    return v;
}

// The same gather through cp.async: every thread copies UNROLL random 32-byte sectors into its own shared-memory slots
// (two 16-byte cp.async each, as the mapping kernels would), waits for the group and reads one word of each back.
// mode 6: plain; mode 7: with the L2::64B prefetch qualifier.  Measures what rate of random fetches LDGSTS sustains.
template <int UNROLL>
__global__ void kmb_gather_async_bench_kernel(const uint8_t *__restrict__ table, uint64_t n_sectors, uint64_t n_loads,
                                              uint64_t seed, uint64_t *sink, int mode) {
    extern __shared__ __align__(16) unsigned char kmb_ga_smem[];
    uint32_t *slots = reinterpret_cast<uint32_t *>(kmb_ga_smem) + (size_t)threadIdx.x * UNROLL * 8;
    uint64_t acc = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_loads; i += stride * UNROLL) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const uint64_t r = kmb_mix64((i + (uint64_t)u * stride) ^ seed);
            const uint8_t *p = table + kmb_umulhi64(r, n_sectors) * 32;
            const uint32_t d = (uint32_t)__cvta_generic_to_shared(slots + u * 8);
            if (mode == 7) {
                asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(d), "l"(p) : "memory");
                asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(d + 16u), "l"(p + 16) : "memory");
            } else {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(p) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u), "l"(p + 16) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int u = 0; u < UNROLL; u++) acc ^= slots[u * 8];
    }
    if (acc == 0x123456789ABCDEFull) *sink = acc;
}

template <int W, int UNROLL>
__global__ void kmb_gather_bench_kernel(const uint8_t *__restrict__ table, uint64_t n_sectors, uint64_t n_loads,
                                        uint64_t seed, uint64_t *sink, int mode) {
    uint64_t acc = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_loads; i += stride * UNROLL) {
        uint64_t v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            uint64_t idx = i + (uint64_t)u * stride;
            uint64_t r = kmb_mix64(idx ^ seed);
            uint64_t sector = kmb_umulhi64(r, n_sectors);
            const uint8_t *p = table + sector * 32;
            if (W == 8) {
                v[u] = kmb_ldg_u64_mode(reinterpret_cast<const uint64_t *>(p), mode);
            } else if (W == 16) {
                uint4 t = kmb_ldg_v4_nc(p);
                v[u] = (uint64_t)t.x ^ ((uint64_t)t.y << 32) ^ t.z ^ ((uint64_t)t.w << 32);
            } else if (mode == 8) {  // one 256-bit load per sector (kmb_ld_sector: what the mapping kernels issue)
                uint32_t r8[8];
                kmb_ld_sector(reinterpret_cast<const uint32_t *>(p), r8, kmb_policy_evict_normal());
                v[u] = (uint64_t)r8[0] ^ ((uint64_t)r8[1] << 32) ^ r8[2] ^ ((uint64_t)r8[3] << 32) ^ r8[4] ^ ((uint64_t)r8[5] << 32) ^ r8[6] ^ r8[7];
            } else {
                uint4 t0 = kmb_ldg_v4_nc(p), t1 = kmb_ldg_v4_nc(p + 16);
                v[u] = (uint64_t)t0.x ^ ((uint64_t)t0.y << 32) ^ t0.z ^ ((uint64_t)t0.w << 32) ^ t1.x ^
                       ((uint64_t)t1.y << 32) ^ t1.z ^ ((uint64_t)t1.w << 32);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) acc ^= v[u];
    }
    if (acc == 0x123456789ABCDEFull) *sink = acc;
}
