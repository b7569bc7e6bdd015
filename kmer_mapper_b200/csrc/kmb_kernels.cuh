// kmb_kernels.cuh -- hand-written sm_100a kernels of the k-mer mapping path.
//
//   K5  kmb_pack_entries / kmb_build_directory / kmb_scan_nodes   index re-layout (once per index)
//   K0  kmb_mark_read_ends        read-boundary bitmask (one bit per base = "no window starts here")
//   K1-4 kmb_map_reads_kernel     fused encode + window + directory probe + count  (production path)
//   K3-4 kmb_map_kmers_kernel     probe + count on ready-made uint64 k-mers (mapper.pyx:19 drop-in)
//   K6  kmb_in_graph_kernel       membership mask (mapper.pyx:81)
//   K2  kmb_hash_count/_scan/_emit  flat hash array (util.py:71-75 drop-in)
//   E1  kmb_codec_* kernels       legacy 2-bit codec (encodings.py)
//   B   kmb_gather_bench_kernel   random-sector gather micro-roofline
//
// Nothing here is a dense contraction, so no tensor-core / TMEM / TMA-tile machinery is used: the
// path is bound by random 32-byte-sector gathers from HBM (directory) plus a thin coalesced stream
// of bases.  What matters: one sector per query in the common case, many independent
// gathers in flight per thread, warp-compacted second-level work so hits do not serialise the
// warp, no-return reductions (RED) for the counters, persistent grid sized to the SM count.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmb_core.cuh"

#define KMB_FULL_MASK 0xFFFFFFFFu
#define KMB_TILE_THREADS 256
#define KMB_POS_PER_THREAD 32
#define KMB_TILE_POS (KMB_TILE_THREADS * KMB_POS_PER_THREAD)  // 8192 window starts per tile
#define KMB_QUEUE_SLOTS 64                                    // per-warp candidate stack (>= 32 + 32)

struct KmbStatus {
    unsigned long long first_bad_offset;  // min flat offset of an invalid byte, ~0 if none
    unsigned long long n_kmers_mapped;    // windows looked up
    unsigned long long n_entries_counted; // index entries that received a +1 (mapper.pyx:68)
    unsigned int index_flags;             // bit0 bucket out of range, bit1 negative node, bit2 overflow bucket seen
    int max_node;
};

struct KmbEntry {  // 16 bytes, one LDG.128
    uint64_t key;
    uint32_t node;
    uint32_t freq;
};

struct KmbProbe {  // everything a probe needs, passed by value to the kernels
    const uint64_t *__restrict__ dir;
    const KmbEntry *__restrict__ entries;
    const int32_t *__restrict__ n_overflow;  // original n_kmers[], only read for n == 31 buckets
    const uint32_t *__restrict__ filter;     // bit h set <=> bucket h is not empty (L2-resident), or nullptr
    uint32_t *counts;
    KmbMod mod;
    int32_t max_freq;  // C int like the reference's cut-off (mapper.pyx:19,64); negative = nothing counts
};

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
__device__ __forceinline__ uint64_t kmb_ldg_u64_nc(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
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
__device__ __forceinline__ KmbEntry kmb_load_entry(const KmbEntry *p, uint64_t pol) {
    uint4 v = kmb_ldg_v4_hint(p, pol);
    KmbEntry e;
    e.key = (uint64_t)v.x | ((uint64_t)v.y << 32);
    e.node = v.z;
    e.freq = v.w;
    return e;
}

// ================================================================================================
// K5: index re-layout
// ================================================================================================
__global__ void kmb_pack_entries(const uint64_t *__restrict__ kmers, const int32_t *__restrict__ nodes,
                                 const uint16_t *__restrict__ freqs, uint64_t n, KmbEntry *__restrict__ out,
                                 KmbStatus *status) {
    int local_max = -1;
    bool neg = false;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        int node = nodes[i];
        uint4 v;
        uint64_t key = kmers[i];
        v.x = (uint32_t)key;
        v.y = (uint32_t)(key >> 32);
        v.z = (uint32_t)node;
        v.w = (uint32_t)freqs[i];
        reinterpret_cast<uint4 *>(out)[i] = v;
        neg |= node < 0;
        local_max = max(local_max, node);
    }
    for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(KMB_FULL_MASK, local_max, o));
    unsigned any_neg = __ballot_sync(KMB_FULL_MASK, neg);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&status->max_node, local_max);
        if (any_neg) atomicOr(&status->index_flags, 2u);
    }
}

__global__ void kmb_build_directory(const int32_t *__restrict__ hashes_to_index, const int32_t *__restrict__ n_kmers,
                                    const uint64_t *__restrict__ kmers, uint64_t modulo, uint64_t n_entries,
                                    KmbMod mod, uint64_t *__restrict__ dir, uint32_t *__restrict__ filter,
                                    KmbStatus *status) {
    // every warp walks aligned groups of 32 consecutive buckets so that one ballot is one filter word
    unsigned flags = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (uint64_t h0 = blockIdx.x * (uint64_t)blockDim.x + (threadIdx.x - lane); h0 < modulo; h0 += stride) {
        const uint64_t h = h0 + lane;
        uint64_t w = 0;
        if (h < modulo) {
            int n = n_kmers[h];
            int pos = hashes_to_index[h];
            if (n > 0) {
                if (pos < 0 || (uint64_t)pos + (uint64_t)n > n_entries) {
                    flags |= 1u;
                } else {
                    uint64_t q, r;
                    kmb_divmod(kmers[pos], mod, q, r);
                    w = kmb_dir_pack((uint32_t)pos, (uint32_t)n, (uint32_t)q);
                    if (n >= (int)KMB_DIR_N_OVERFLOW) flags |= 4u;
                }
            } else if (n < 0) {
                flags |= 1u;
            }
            dir[h] = w;
        }
        // NB: a non-empty bucket always has a non-zero word (n field >= 1)
        unsigned occ = __ballot_sync(KMB_FULL_MASK, w != 0ull);
        if (lane == 0) filter[h0 >> 5] = occ;
    }
    if (flags) atomicOr(&status->index_flags, flags);
}

// ================================================================================================
// K0: read-boundary mask.  Bit p of mask set <=> no window may start at flat base p, i.e. p lies in
// the last k-1 bases of its read (or the read is shorter than k).  One thread per read; a read
// touches at most k-1 <= 30 bits = at most two 32-bit words.
// ================================================================================================
__global__ void kmb_mark_read_ends(const int64_t *__restrict__ offsets, uint64_t n_reads, int64_t base0, int k,
                                   uint32_t *mask) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_reads; r += (uint64_t)gridDim.x * blockDim.x) {
        int64_t s = offsets[r] - base0;
        int64_t e = offsets[r + 1] - base0;
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
// The probe, shared by every mapping kernel.
//
// Level 0 (optional, FILT): one bit per bucket, "bucket h is not empty".  modulo/8 bytes -- 57 MB for
//   the reference's default modulo 452 930 477 -- so it stays resident in the 126 MB L2 while the
//   3.6 GB directory cannot.  At the reference's load factor (~0.22 entries per bucket) four out of
//   five queries end here without touching HBM.
// Level 1: one 8-byte directory word per surviving query (one HBM sector): position, size and a
//   28-bit quotient fingerprint of the bucket's first entry -> single-entry buckets whose key
//   differs are rejected without reading the entry.
// Level 2: surviving candidates are compacted onto a per-warp stack in shared memory and drained 32
//   at a time, one candidate per lane, so the entry walk (mapper.pyx:58-69: every equal key counts,
//   no break; frequency filter :64) runs with full lanes instead of one divergent lane per hit.
//   One no-return reduction (RED) per counted entry; with AGG the lanes that hit the same node in
//   the same step are merged first (__match_any_sync) so hot nodes cost one RED per warp step.
// ================================================================================================
struct KmbPol {
    uint64_t first;  // L2 evict-first: use-once gathers and streams
    uint64_t last;   // L2 evict-last: the filter
};
__device__ __forceinline__ KmbPol kmb_make_policies() {
    KmbPol p;
    p.first = kmb_policy_evict_first();
    p.last = kmb_policy_evict_last();
    return p;
}

template <bool AGG>
__device__ __forceinline__ void kmb_drain(const KmbProbe &P, const KmbPol &pol, const uint64_t *q_kmer,
                                          const uint64_t *q_dir, int base, int cnt, int lane, unsigned &counted) {
    bool active = lane < cnt;
    uint64_t km = 0, dw = 0;
    if (active) {
        km = q_kmer[base + lane];
        dw = q_dir[base + lane];
    }
    uint32_t n = kmb_dir_n(dw);
    uint32_t pos = kmb_dir_pos(dw);
    if (n == KMB_DIR_N_OVERFLOW) {
        uint64_t q, h;
        kmb_divmod(km, P.mod, q, h);
        n = (uint32_t)P.n_overflow[h];
    }
    for (uint32_t j = 0;; ++j) {
        bool more = active && j < n;
        if (!__any_sync(KMB_FULL_MASK, more)) break;
        bool hit = false;
        uint32_t node = 0;
        if (more) {
            KmbEntry e = kmb_load_entry(P.entries + pos + j, pol.first);
            hit = (e.key == km) && ((int32_t)e.freq <= P.max_freq);
            node = e.node;
        }
        counted += hit ? 1u : 0u;
        if (AGG) {
            unsigned hm = __ballot_sync(KMB_FULL_MASK, hit);
            if (hit) {
                unsigned peers = __match_any_sync(hm, node);
                if (lane == __ffs(peers) - 1) atomicAdd(P.counts + node, (uint32_t)__popc(peers));
            }
        } else {
            if (hit) atomicAdd(P.counts + node, 1u);
        }
    }
}

// Push this lane's candidate (if any) on the warp's stack; drain when 32 are waiting.
// Must be called by all 32 lanes (cand=false for lanes without one).
template <bool AGG>
__device__ __forceinline__ void kmb_push_candidate(const KmbProbe &P, const KmbPol &pol, uint64_t *q_kmer,
                                                   uint64_t *q_dir, int &qcount, bool cand, uint64_t km, uint64_t dw,
                                                   int lane, unsigned &counted) {
    unsigned bal = __ballot_sync(KMB_FULL_MASK, cand);
    if (bal == 0u) return;
    if (cand) {
        int slot = qcount + __popc(bal & ((1u << lane) - 1u));
        q_kmer[slot] = km;
        q_dir[slot] = dw;
    }
    qcount += __popc(bal);
    if (qcount >= 32) {
        __syncwarp();
        qcount -= 32;
        kmb_drain<AGG>(P, pol, q_kmer, q_dir, qcount, 32, lane, counted);
        __syncwarp();
    }
}

// Levels 0 and 1 for U queries of this lane, all gathers of a level in flight together.
// kf(u) yields query u (cheap to recompute, so it is not kept in registers); bit u of vbits says
// whether query u exists.
template <int U, bool FILT, bool AGG, class KF>
__device__ __forceinline__ void kmb_probe_batch(const KmbProbe &P, const KmbPol &pol, const KF &kf, uint32_t vbits,
                                                uint64_t *q_kmer, uint64_t *q_dir, int &qcount, int lane,
                                                unsigned &counted) {
    uint64_t dw[U];
    uint32_t fq[U];
    if (FILT) {
        uint32_t hh[U], fw[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            uint64_t q, h;
            kmb_divmod(kf(u), P.mod, q, h);
            fq[u] = (uint32_t)q;
            hh[u] = (uint32_t)h;
            fw[u] = ((vbits >> u) & 1u) ? kmb_ldg_u32_hint(P.filter + (hh[u] >> 5), pol.last) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            dw[u] = ((fw[u] >> (hh[u] & 31u)) & 1u) ? kmb_ldg_u64_hint(P.dir + hh[u], pol.first) : 0ull;
    } else {
#pragma unroll
        for (int u = 0; u < U; u++) {
            uint64_t q, h;
            kmb_divmod(kf(u), P.mod, q, h);
            fq[u] = (uint32_t)q;
            dw[u] = ((vbits >> u) & 1u) ? kmb_ldg_u64_hint(P.dir + h, pol.first) : 0ull;
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        bool cand = !kmb_dir_rejects(dw[u], fq[u]);
        kmb_push_candidate<AGG>(P, pol, q_kmer, q_dir, qcount, cand, kf(u), dw[u], lane, counted);
    }
}

struct KmbWindowFn {  // forward window b0+u of the 64 bases in hi:lo
    uint64_t lo, hi, kmask;
    int b0;
    __device__ __forceinline__ uint64_t operator()(int u) const { return kmb_window(lo, hi, b0 + u, kmask); }
};
struct KmbRcWindowFn {  // its reverse complement
    uint64_t lo, hi, kmask;
    int b0, k;
    __device__ __forceinline__ uint64_t operator()(int u) const { return kmb_revcomp(kmb_window(lo, hi, b0 + u, kmask), k); }
};
struct KmbArrayFn {
    const uint64_t *km;
    __device__ __forceinline__ uint64_t operator()(int u) const { return km[u]; }
};

// ================================================================================================
// K1-4 fused: raw ASCII bases in, node counts out.  Each base is read from HBM exactly once.
//
// Persistent CTAs of 256 threads walk tiles of 8192 window-start positions.  Per tile:
//   1. 16-byte vector loads of 8192+32 bases (coalesced, evict-first), SWAR-encoded in registers to
//      2 bits/base, stored to shared memory as a packed little-endian bit stream (2 KB + halo);
//   2. each thread owns 32 consecutive positions: two 64-bit shared loads give it every window
//      (window i = bits [2i, 2i+2k) -- the reference's first-base-lowest hash, util.py:71-75);
//      one 32-bit word of the read-boundary mask says which of its 32 starts are real windows;
//   3. in batches of U positions: exact kmer % modulo (Barrett), then the three probe levels above.
// base0 = flat offset of bases[0] inside the caller's buffer (chunked host input), only used to
// report the position of an invalid byte.
// ================================================================================================
__device__ __forceinline__ uint4 kmb_load_bases16(const uint8_t *__restrict__ bases, uint64_t v, uint64_t n_vec_full,
                                                  uint64_t n_bases, uint64_t pol) {
    if (v < n_vec_full) return kmb_ldg_v4_hint(bases + v * 16, pol);
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
    uint32_t valid = ~mask[p0 >> 5];
    if (p0 + 32 + (uint64_t)k > n_bases + 1) {  // window must end inside the buffer: p + k <= n_bases
        int64_t last = (int64_t)n_bases - (int64_t)k - (int64_t)p0;  // last valid i
        valid &= last < 0 ? 0u : (last >= 31 ? 0xFFFFFFFFu : ((1u << (last + 1)) - 1u));
    }
    return valid;
}

template <int U, bool FILT, bool AGG, bool REVCOMP>
__global__ void __launch_bounds__(KMB_TILE_THREADS)
kmb_map_reads_kernel(const uint8_t *__restrict__ bases, uint64_t n_bases, uint64_t base0,
                     const uint32_t *__restrict__ mask, int k, bool n_to_a, KmbProbe P, KmbStatus *status) {
    __shared__ __align__(16) uint32_t s_pack[KMB_TILE_POS / 16 + 8];  // 512 words + halo
    __shared__ uint64_t s_qk[KMB_TILE_THREADS / 32][KMB_QUEUE_SLOTS];
    __shared__ uint64_t s_qd[KMB_TILE_THREADS / 32][KMB_QUEUE_SLOTS];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    uint64_t *q_kmer = s_qk[warp];
    uint64_t *q_dir = s_qd[warp];
    int qcount = 0;
    unsigned counted = 0;
    const KmbPol pol = kmb_make_policies();
    const uint64_t kmask = kmb_kmer_mask(k);
    const uint64_t n_tiles = (n_bases + KMB_TILE_POS - 1) / KMB_TILE_POS;
    const uint64_t n_vec_full = n_bases / 16;  // vectors entirely inside the buffer
    unsigned long long mapped = 0;

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t t0 = tile * KMB_TILE_POS;
        __syncthreads();  // previous tile's readers are done with s_pack
        // ---- 1. load + encode: vectors v = t0/16 + i, i in [0, 514)
        for (int i = tid; i < KMB_TILE_POS / 16 + 2; i += KMB_TILE_THREADS) {
            uint64_t v = t0 / 16 + (uint64_t)i;
            uint4 w = kmb_load_bases16(bases, v, n_vec_full, n_bases, pol.first);
            uint32_t inv;
            s_pack[i] = kmb_encode16(w.x, w.y, w.z, w.w, n_to_a, inv);
            if (inv) atomicMin(&status->first_bad_offset, (unsigned long long)(base0 + v * 16 + (uint64_t)(__ffs(inv) - 1)));
        }
        __syncthreads();
        // ---- 2. this thread's 32 positions
        const uint64_t p0 = t0 + (uint64_t)tid * KMB_POS_PER_THREAD;
        const uint32_t valid = kmb_valid_starts(mask, p0, n_bases, k);
        mapped += __popc(valid);
        const uint2 a = *reinterpret_cast<const uint2 *>(&s_pack[2 * tid]);
        const uint2 b = *reinterpret_cast<const uint2 *>(&s_pack[2 * tid + 2]);
        const uint64_t lo = (uint64_t)a.x | ((uint64_t)a.y << 32);
        const uint64_t hi = (uint64_t)b.x | ((uint64_t)b.y << 32);
        // ---- 3. probe in batches of U
#pragma unroll 1
        for (int b0 = 0; b0 < KMB_POS_PER_THREAD; b0 += U) {
            const uint32_t vb = (valid >> b0) & ((U == 32) ? 0xFFFFFFFFu : ((1u << U) - 1u));
            if (!__any_sync(KMB_FULL_MASK, vb != 0u)) continue;
            KmbWindowFn fw = {lo, hi, kmask, b0};
            kmb_probe_batch<U, FILT, AGG>(P, pol, fw, vb, q_kmer, q_dir, qcount, lane, counted);
            if (REVCOMP) {
                KmbRcWindowFn rc = {lo, hi, kmask, b0, k};
                kmb_probe_batch<U, FILT, AGG>(P, pol, rc, vb, q_kmer, q_dir, qcount, lane, counted);
            }
        }
    }
    __syncwarp();
    if (qcount > 0) kmb_drain<AGG>(P, pol, q_kmer, q_dir, 0, qcount, lane, counted);
    // statistics: one atomic per warp
    for (int o = 16; o > 0; o >>= 1) {
        mapped += __shfl_xor_sync(KMB_FULL_MASK, mapped, o);
        counted += __shfl_xor_sync(KMB_FULL_MASK, counted, o);
    }
    if (lane == 0 && mapped) atomicAdd(&status->n_kmers_mapped, REVCOMP ? 2ull * mapped : mapped);
    if (lane == 0 && counted) atomicAdd(&status->n_entries_counted, (unsigned long long)counted);
}

// ================================================================================================
// K3-4 on ready-made k-mers (drop-in for map_kmers_to_graph_index, mapper.pyx:19-72).
// Coalesced 8-byte loads, U gathers in flight per thread, same probe.
// ================================================================================================
template <int U, bool FILT, bool AGG, bool REVCOMP>
__global__ void __launch_bounds__(KMB_TILE_THREADS)
kmb_map_kmers_kernel(const uint64_t *__restrict__ kmers, uint64_t n, int k, KmbProbe P, KmbStatus *status) {
    __shared__ uint64_t s_qk[KMB_TILE_THREADS / 32][KMB_QUEUE_SLOTS];
    __shared__ uint64_t s_qd[KMB_TILE_THREADS / 32][KMB_QUEUE_SLOTS];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    uint64_t *q_kmer = s_qk[warp];
    uint64_t *q_dir = s_qd[warp];
    int qcount = 0;
    unsigned counted = 0;
    const KmbPol pol = kmb_make_policies();
    const uint64_t per_block = (uint64_t)KMB_TILE_THREADS * U;
    const uint64_t n_blocks = (n + per_block - 1) / per_block;
    for (uint64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        uint64_t base = blk * per_block + (uint64_t)tid;
        uint64_t km[U];
        uint32_t vb = 0;
#pragma unroll
        for (int u = 0; u < U; u++) {
            uint64_t i = base + (uint64_t)u * KMB_TILE_THREADS;
            bool in = i < n;
            km[u] = in ? kmb_ldg_u64_hint(kmers + i, pol.first) : 0ull;
            vb |= in ? (1u << u) : 0u;
        }
        KmbArrayFn fa = {km};
        kmb_probe_batch<U, FILT, AGG>(P, pol, fa, vb, q_kmer, q_dir, qcount, lane, counted);
        if (REVCOMP) {
#pragma unroll
            for (int u = 0; u < U; u++) km[u] = kmb_revcomp(km[u], k);
            kmb_probe_batch<U, FILT, AGG>(P, pol, fa, vb, q_kmer, q_dir, qcount, lane, counted);
        }
    }
    __syncwarp();
    if (qcount > 0) kmb_drain<AGG>(P, pol, q_kmer, q_dir, 0, qcount, lane, counted);
    for (int o = 16; o > 0; o >>= 1) counted += __shfl_xor_sync(KMB_FULL_MASK, counted, o);
    if (lane == 0 && counted) atomicAdd(&status->n_entries_counted, (unsigned long long)counted);
    if (blockIdx.x == 0 && tid == 0) atomicAdd(&status->n_kmers_mapped, REVCOMP ? 2ull * n : (unsigned long long)n);
}

// One-query-per-thread probe without the warp stack: the cross-check variant
// (kmb_set_option("probe_variant", 0)) and the baseline the staged probe is measured against.
__device__ __forceinline__ bool kmb_probe_one(const KmbProbe &P, const KmbPol &pol, uint64_t km, uint32_t &pos,
                                              uint32_t &nn) {
    uint64_t q, h;
    kmb_divmod(km, P.mod, q, h);
    if (P.filter != nullptr) {
        uint32_t w = kmb_ldg_u32_hint(P.filter + (h >> 5), pol.last);
        if (!((w >> (h & 31u)) & 1u)) return false;
    }
    uint64_t dw = kmb_ldg_u64_hint(P.dir + h, pol.first);
    if (kmb_dir_rejects(dw, (uint32_t)q)) return false;
    nn = kmb_dir_n(dw);
    pos = kmb_dir_pos(dw);
    if (nn == KMB_DIR_N_OVERFLOW) nn = (uint32_t)P.n_overflow[h];
    return true;
}

template <bool REVCOMP>
__global__ void kmb_map_kmers_simple_kernel(const uint64_t *__restrict__ kmers, uint64_t n, int k, KmbProbe P,
                                            KmbStatus *status) {
    const KmbPol pol = kmb_make_policies();
    unsigned long long counted = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t km = kmers[i];
#pragma unroll
        for (int strand = 0; strand < (REVCOMP ? 2 : 1); strand++) {
            if (strand == 1) km = kmb_revcomp(km, k);
            uint32_t pos, nn;
            if (!kmb_probe_one(P, pol, km, pos, nn)) continue;
            for (uint32_t j = 0; j < nn; j++) {
                KmbEntry e = kmb_load_entry(P.entries + pos + j, pol.first);
                if (e.key == km && (int32_t)e.freq <= P.max_freq) {
                    atomicAdd(P.counts + e.node, 1u);
                    counted++;
                }
            }
        }
    }
    if (counted) atomicAdd(&status->n_entries_counted, counted);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&status->n_kmers_mapped, REVCOMP ? 2ull * n : (unsigned long long)n);
}

// ================================================================================================
// K6 membership (mapper.pyx:81-130): first key match wins, frequency ignored.
// MODE 0: out_u8[i] = hit.  MODE 1: out_u32[i] = counts[node of first match] (Counter.__getitem__).
// ================================================================================================
template <int MODE>
__global__ void kmb_in_graph_kernel(const uint64_t *__restrict__ kmers, uint64_t n, KmbProbe P, uint8_t *out8,
                                    uint32_t *out32) {
    const KmbPol pol = kmb_make_policies();
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t km = kmers[i];
        bool hit = false;
        uint32_t node = 0, pos, nn;
        if (kmb_probe_one(P, pol, km, pos, nn)) {
            for (uint32_t j = 0; j < nn; j++) {
                KmbEntry e = kmb_load_entry(P.entries + pos + j, pol.first);
                if (e.key == km) {
                    hit = true;
                    node = e.node;
                    break;
                }
            }
        }
        if (MODE == 0) out8[i] = hit ? 1 : 0;
        else out32[i] = hit ? P.counts[node] : 0u;
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

__global__ void __launch_bounds__(KMB_TILE_THREADS)
kmb_hash_emit_kernel(const uint8_t *__restrict__ bases, uint64_t n_bases, const uint32_t *__restrict__ mask, int k,
                     bool n_to_a, const unsigned long long *__restrict__ tile_offsets, uint64_t *__restrict__ out,
                     uint64_t out_capacity, KmbStatus *status) {
    __shared__ __align__(16) uint32_t s_pack[KMB_TILE_POS / 16 + 8];
    __shared__ uint32_t s_w[KMB_TILE_THREADS / 32];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
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
    __syncthreads();
    const uint64_t p0 = t0 + (uint64_t)tid * KMB_POS_PER_THREAD;
    uint32_t valid = kmb_valid_starts(mask, p0, n_bases, k);
    // exclusive prefix of popc(valid) over the CTA
    uint32_t c = __popc(valid);
    uint32_t incl = c;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(KMB_FULL_MASK, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_w[tid >> 5] = incl;
    __syncthreads();
    uint32_t warp_base = 0;
    for (int w = 0; w < (tid >> 5); w++) warp_base += s_w[w];
    uint64_t o = tile_offsets[tile] + warp_base + (incl - c);
    const uint2 a = *reinterpret_cast<const uint2 *>(&s_pack[2 * tid]);
    const uint2 b = *reinterpret_cast<const uint2 *>(&s_pack[2 * tid + 2]);
    const uint64_t lo = (uint64_t)a.x | ((uint64_t)a.y << 32);
    const uint64_t hi = (uint64_t)b.x | ((uint64_t)b.y << 32);
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
        if ((valid >> i) & 1u) {
            if (o < out_capacity) out[o] = kmb_window(lo, hi, i, kmask);
            o++;
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
template <int W, int UNROLL>
__global__ void kmb_gather_bench_kernel(const uint8_t *__restrict__ table, uint64_t n_sectors, uint64_t n_loads,
                                        uint64_t seed, uint64_t *sink) {
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
                v[u] = kmb_ldg_u64_nc(reinterpret_cast<const uint64_t *>(p));
            } else if (W == 16) {
                uint4 t = kmb_ldg_v4_nc(p);
                v[u] = (uint64_t)t.x ^ ((uint64_t)t.y << 32) ^ t.z ^ ((uint64_t)t.w << 32);
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
