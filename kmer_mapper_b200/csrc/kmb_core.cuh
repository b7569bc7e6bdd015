// kmb_core.cuh -- bit-level primitives shared by every kernel of the k-mer mapping path.
//
// All functions here are pure and __host__ __device__ so that tests/test_core_host.py can compile
// this header with g++ and check the arithmetic (exact u64 % modulo, SWAR 2-bit encoding, window
// extraction, directory word packing) against Python integers without a GPU.  The host build is a
// unit-test harness for these primitives only; no mapping path exists on the CPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KMB_HD __host__ __device__ __forceinline__
#else
#define KMB_HD inline
#endif

// ---------------------------------------------------------------------------------------------
// Exact n / d and n % d for a runtime 64-bit divisor (reference: `kmers[i] % modulo`,
// mapper.pyx:54).  Barrett: m = floor((2^64-1)/d); q' = mulhi(n, m) is floor(n/d) or one less
// (n/d - n*m/2^64 < 1 for every n < 2^64, d >= 1), so a single correction step is exact.
// ---------------------------------------------------------------------------------------------
struct KmbMod {
    uint64_t d;
    uint64_t m;
};

KMB_HD uint64_t kmb_umulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

KMB_HD KmbMod kmb_mod_make(uint64_t d) {
    KmbMod r;
    r.d = d;
    r.m = ~0ull / d;
    return r;
}

KMB_HD void kmb_divmod(uint64_t n, const KmbMod md, uint64_t &q, uint64_t &r) {
    uint64_t qe = kmb_umulhi64(n, md.m);
    if (md.d < 0x80000000ull) {
        // n - qe*d lies in [0, 2d) and 2d < 2^32, so the low 32 bits of the difference are the difference
        uint32_t re = (uint32_t)n - (uint32_t)qe * (uint32_t)md.d;
        bool fix = re >= (uint32_t)md.d;
        r = fix ? re - (uint32_t)md.d : re;
        q = qe + (fix ? 1u : 0u);
        return;
    }
    uint64_t re = n - qe * md.d;
    if (re >= md.d) {
        re -= md.d;
        qe += 1;
    }
    q = qe;
    r = re;
}

// ---------------------------------------------------------------------------------------------
// Index layout: the 32-byte sector.
//
// Measured on B200 (profiles/README.md): (1) a random global load that misses L2 is one DRAM
// transaction whatever its width, and the chip sustains 38.2 G of them per second; (2) every dirty
// sector costs a second transaction when it is written back; (3) the L2 holds ~72 MB, of which the
// filter takes 57 MB, so a fetched line is gone again after a few microseconds.  Hence ONE
// read-only 32-byte sector answers a query completely -- keys, nodes and frequencies -- and hits
// are not counted in place but appended to a log (see kmb_kernels.cuh):
//   w0      header: n (0..2 entries, chain ends here) or KMB_HDR_CHAIN | index of the next sector
//   w1      frequency 0 | frequency 1 << 16
//   w2..w5  key 0, key 1  (lo, hi)
//   w6, w7  node 0, node 1
// G = 2^g consecutive buckets (h = key % modulo, mapper.pyx:54) share a sector, with G chosen so
// that a sector holds 0.25-0.5 entries on average (99 % of the non-empty ones need no chain);
// entries beyond two go to overflow sectors ovf_base + (s-2)/2, linked through their headers.
// HBM capacity (180 GB) is what pays for this: 7.2 GB of sectors for the 100 M-entry index.
// ---------------------------------------------------------------------------------------------
#define KMB_LINE_BYTES 32
#define KMB_LINE_WORDS 8
#define KMB_LINE_SLOTS 2
#define KMB_LINE_FREQ_WORD 1
#define KMB_LINE_KEY_WORD0 2
#define KMB_LINE_NODE_WORD0 6
#define KMB_HDR_CHAIN 0x80000000u

// chain position s (0-based among the entries of a main sector) -> (sector index, slot)
KMB_HD uint64_t kmb_chain_line(uint64_t main_line, uint32_t ovf_base, uint32_t s) {
    return s < KMB_LINE_SLOTS ? main_line : (uint64_t)ovf_base + (s - KMB_LINE_SLOTS) / KMB_LINE_SLOTS;
}
KMB_HD uint32_t kmb_chain_slot(uint32_t s) { return s % KMB_LINE_SLOTS; }
KMB_HD uint32_t kmb_chain_extra_lines(uint32_t n_total) {
    return n_total > KMB_LINE_SLOTS ? (n_total - 1) / KMB_LINE_SLOTS : 0u;
}
// header of a sector that still has `remaining` entries to hold (itself included)
KMB_HD uint32_t kmb_sector_header(uint32_t remaining, uint32_t next_sector) {
    return remaining > KMB_LINE_SLOTS ? (KMB_HDR_CHAIN | next_sector) : remaining;
}
KMB_HD uint32_t kmb_header_count(uint32_t hdr) { return (hdr & KMB_HDR_CHAIN) ? (uint32_t)KMB_LINE_SLOTS : hdr; }

// Hit log: node ids are appended in groups of 32, each group tagged with one of KMB_LOG_BINS node
// ranges, so that applying the groups of one range touches a window of the count array small
// enough to stay in L2.
// A log has n_bins <= KMB_LOG_BINS ranges.  Measured with 16 (80 M nodes: ranges of 2^23 nodes = one apply window
// each, so the log is read once): apply pass 5.0 -> 4.4 ms, but 3 KB more staging per warp push the mapping kernel's
// CTAs to a 228 KB shared-memory carve-out, the 28 KB of L1 that remain cannot hold the sectors of the loads in
// flight, and the kernel takes 88.8 ms instead of 47.0 (the same with an explicit 100 % carve-out and 8 bins).
#ifndef KMB_LOG_BINS
#define KMB_LOG_BINS 8
#endif
#ifndef KMB_MZ_LOG_BINS
#define KMB_MZ_LOG_BINS 12  // the read-path kernel's stacks are shallower (KMB_MZ_STAGE_SLOTS), so it has room for more ranges
#endif
KMB_HD uint32_t kmb_log_bin(uint32_t node, uint32_t bin_shift) {
    uint32_t b = node >> bin_shift;
    return b < KMB_LOG_BINS ? b : (uint32_t)(KMB_LOG_BINS - 1);
}

// Addressing.  The reference finds an entry through its bucket h = key % modulo (mapper.pyx:54).
// That rule decides, once, at index build time, which entries are reachable ("live"); after that
// the only requirement is that a query finds every live entry with an equal key -- so the device
// layout is free to address sectors and filter bits by ANY function of the key.  It uses a
// multiplicative hash (6 integer instructions) instead of an exact 64-bit modulo (25): sector and
// filter word by multiply-shift range reduction, so neither count has to be a power of two.
//
// Filter (probe level 0): a Bloom filter blocked into single 32-bit words, sized to stay in L2.
// Every live entry sets one, two or three bits of its word (the more filter bits per key there are, the more:
// kmb_filter_probes); a query passes iff all of its bits are set.  57-64 MB for 100 M keys: ~10 % of absent
// k-mers pass.
struct KmbAddr {
    uint32_t n_main;          // main sectors
    uint32_t n_filter_words;  // 0 = no filter
    uint32_t n_probes;        // filter bits per key: 1, 2 or 3
};
struct KmbLoc {
    uint32_t sector, fword, fmask;
};
KMB_HD KmbLoc kmb_locate(uint64_t key, const KmbAddr a) {
    const uint64_t x = key * 0x9E3779B97F4A7C15ull;
    const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
    const uint32_t f = hi * 0x85EBCA6Bu + lo;
    // The word index consumes the top ~24 bits of f, so inside one word f keeps only ~8 free bits: the bit
    // positions must come from a separate mix (taken from f they crowded onto 8 of the 32 positions of a
    // word and the false-pass rate of the 57 MB filter was 19.7 % instead of 12.6 %).
    const uint32_t g = lo * 0xC2B2AE35u + hi;
    KmbLoc l;
#if defined(__CUDA_ARCH__)
    l.sector = __umulhi(hi, a.n_main);
    l.fword = __umulhi(f, a.n_filter_words);
#else
    l.sector = (uint32_t)(((uint64_t)hi * a.n_main) >> 32);
    l.fword = (uint32_t)(((uint64_t)f * a.n_filter_words) >> 32);
#endif
    l.fmask = (1u << (g >> 27)) | (a.n_probes >= 2u ? (1u << ((g >> 22) & 31u)) : 0u) |
              (a.n_probes >= 3u ? (1u << ((g >> 17) & 31u)) : 0u);
    return l;
}

// ---------------------------------------------------------------------------------------------
// 2-bit encoding of ASCII bases, 4 bytes at a time in one 32-bit register (SWAR).
// Codes follow bionumpy's DNAEncoding as used by util.py:72-73: A/a,C/c,G/g,T/t -> 0,1,2,3.
// Upper-case 'N' -> code 0 ('A') when n_to_a (command_line_interface.py:40-41); every other byte
// (including lower-case 'n') is flagged invalid: bit 7 of its byte lane in *invalid.
// Returns the 4 codes packed into 8 bits, first base in bits 1:0 (hash = sum code[j] * 4^j,
// tests/test_hashing.py:13-26).
// ---------------------------------------------------------------------------------------------
KMB_HD uint32_t kmb_zero_bytes(uint32_t x) {  // 0x80 in every byte lane of x that is 0x00
    uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | x | 0x7F7F7F7Fu);
}

// The precise version: which bytes are invalid.  Only run when kmb_encode16 has seen that one is (rare).
KMB_HD uint32_t kmb_invalid4(uint32_t w, bool n_to_a) {
    uint32_t cf = w & 0xDFDFDFDFu;  // fold case
    uint32_t isN = n_to_a ? kmb_zero_bytes(w ^ 0x4E4E4E4Eu) : 0u;
    uint32_t ok = kmb_zero_bytes(cf ^ 0x41414141u) | kmb_zero_bytes(cf ^ 0x43434343u) |
                  kmb_zero_bytes(cf ^ 0x47474747u) | kmb_zero_bytes(cf ^ 0x54545454u) | isN;
    return ok ^ 0x80808080u;  // bit 7 of every invalid byte lane
}

// byte i of the result = byte sel_i of `table`, sel_i = bits [4i, 4i+2) of sel (PRMT)
KMB_HD uint32_t kmb_pick_bytes(uint32_t table, uint32_t sel) {
#ifdef __CUDA_ARCH__
    return __byte_perm(table, 0u, sel);
#else
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= ((table >> (8u * ((sel >> (4 * i)) & 3u))) & 0xFFu) << (8 * i);
    return r;
#endif
}

// Four bases.  Validity is checked by rebuilding the letter from the code: the code of a byte is taken from two
// of its bits, so a byte is one of A C G T (either case) iff "ACGT"[code] equals its case-folded self; `bad`
// collects the differences (nonzero <=> some byte of the word is invalid).  One byte permute and an xor instead
// of four zero-byte tests per word: the encode step was ~10 % of the read-path kernel's instructions.
KMB_HD uint32_t kmb_encode4(uint32_t w, bool n_to_a, uint32_t &bad) {
    uint32_t cf = w & 0xDFDFDFDFu;  // fold case
    uint32_t x = (cf >> 1) & 0x03030303u;  // A0 C1 G3 T2
    x ^= (x >> 1) & 0x01010101u;           // A0 C1 G2 T3
    const uint32_t n4 = (x | (x >> 4));    // byte 0: x0 | x1 << 4, byte 2: x2 | x3 << 4
    const uint32_t expect = kmb_pick_bytes(0x54474341u, (n4 & 0xFFu) | ((n4 >> 8) & 0xFF00u));
    uint32_t d = cf ^ expect;
    if (n_to_a) {
        const uint32_t isN = (kmb_zero_bytes(w ^ 0x4E4E4E4Eu) >> 7) * 0xFFu;  // 0xFF in the byte lanes that hold 'N'
        d &= ~isN;   // fine as it is
        x &= ~isN;   // N -> 0
    }
    bad |= d;
    uint32_t t = (x | (x >> 6)) & 0x000F000Fu;
    return (t | (t >> 12)) & 0xFFu;
}

// 16 bases (four little-endian 32-bit words of ASCII) -> 32 bits of 2-bit codes, base 0 lowest.
// invalid_lanes: bit j set <=> base j invalid (0 in all but the rarest case; computed only then)
KMB_HD uint32_t kmb_encode16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, bool n_to_a,
                             uint32_t &invalid_lanes) {
    uint32_t bad = 0;
    const uint32_t p = kmb_encode4(w0, n_to_a, bad) | (kmb_encode4(w1, n_to_a, bad) << 8) |
                       (kmb_encode4(w2, n_to_a, bad) << 16) | (kmb_encode4(w3, n_to_a, bad) << 24);
    invalid_lanes = 0;
    if (bad) {
        const uint32_t i0 = kmb_invalid4(w0, n_to_a), i1 = kmb_invalid4(w1, n_to_a), i2 = kmb_invalid4(w2, n_to_a),
                       i3 = kmb_invalid4(w3, n_to_a);
        uint32_t m = 0;
        m |= ((i0 >> 7) & 1u) | ((i0 >> 14) & 2u) | ((i0 >> 21) & 4u) | ((i0 >> 28) & 8u);
        m |= (((i1 >> 7) & 1u) | ((i1 >> 14) & 2u) | ((i1 >> 21) & 4u) | ((i1 >> 28) & 8u)) << 4;
        m |= (((i2 >> 7) & 1u) | ((i2 >> 14) & 2u) | ((i2 >> 21) & 4u) | ((i2 >> 28) & 8u)) << 8;
        m |= (((i3 >> 7) & 1u) | ((i3 >> 14) & 2u) | ((i3 >> 21) & 4u) | ((i3 >> 28) & 8u)) << 12;
        invalid_lanes = m;
    }
    return p;
}

// ---------------------------------------------------------------------------------------------
// Window extraction: lo/hi are two consecutive 64-bit words of the packed 2-bit stream
// (32 bases each, base 0 in the lowest bits); the k-mer starting at base i (0..31) of lo is
// bits [2i, 2i+2k) of the 128-bit value hi:lo.  This IS the reference hash: first base lowest.
// ---------------------------------------------------------------------------------------------
KMB_HD uint64_t kmb_kmer_mask(int k) { return k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull); }

KMB_HD uint64_t kmb_window(uint64_t lo, uint64_t hi, int i, uint64_t kmask) {
    int s = 2 * i;
    uint64_t v = (lo >> s) | ((hi << 1) << (63 - s));
    return v & kmask;
}

// Reverse complement of a k-mer hash in the A,C,G,T=0..3 / first-base-lowest convention:
// complement = 3 - code = bitwise NOT of the 2-bit group; order reversed.
KMB_HD uint64_t kmb_revcomp(uint64_t x, int k) {
    x = ~x;
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    x = (x >> 32) | (x << 32);
    return x >> (64 - 2 * k);
}

// ---------------------------------------------------------------------------------------------
// Read-path table (second device structure over the same live entries, used by the fused reads kernel
// only): entries are filed under the MINIMIZER of their key instead of the key itself.
//
// Consecutive windows of a read overlap in k-1 bases, and the minimizer of a k-mer -- the smallest
// hash among its k-m+1 m-mers (m = 16) -- is shared by ~(k-m+2)/2 consecutive windows on average (8.5
// at k = 31).  Filing the index entries under hash(minimizer) therefore sends a whole RUN of windows
// to the same bucket: one filter word and one sector fetch per run instead of one per window, after
// which every window of the run is compared, in registers, with the full keys the bucket holds.
// Exactness is untouched: equal keys have equal minimizers, so a query meets every live entry with
// its key, and only those can match.  (The key-addressed sectors remain what map_kmers, membership
// and per-key look-ups use: a lone key has no neighbours to share a bucket with.)
//
// Bucket b = two adjacent 32-byte sectors 2b, 2b+1 -- one 64-byte DRAM fetch (L2::64B) brings both:
// the primary holds entries 0-1, the secondary entries 2-3 and, beyond that, the link into a pool of
// ordinary chained sectors.
// ---------------------------------------------------------------------------------------------
// m = 16: the longest m-mer that is still one 32-bit word.  The table was first built with m = 15; on a 5 Gbp genome
// (config 3) every 15-mer occurs 4.7 times and the m-mers that win are the ~1/9 with the smallest ordering keys, so
// a popular minimizer collected the entries of several places of the genome: half of the entries sat in buckets of
// more than four, 23 % of the buckets ended in a pool chain, and walking those chains was 17 % of the kernel's
// instructions and its largest source of memory stalls (profiles/r02_v11_config3_*).  One base more divides the
// occurrences per m-mer by four (16 % of the entries / 6 % of the buckets beyond four in the same simulation).
#define KMB_MZ_M 16
#define KMB_MZ_MASK (KMB_MZ_M >= 16 ? 0xFFFFFFFFu : ((1u << (2 * (KMB_MZ_M & 15))) - 1u))
// Ordering key of an m-mer: 26 hash bits above 6 bits that the caller fills with the m-mer's position, so that one
// 32-bit minimum yields the minimizer AND where it sits (leftmost among equal hashes).  The bucket is addressed by
// the m-mer itself, not by this hash, so the 26 bits only decide which m-mer wins.
KMB_HD uint32_t kmb_mmer_order(uint32_t mmer) {
    // One multiplication: the upper 26 bits of the product depend on every base of the m-mer, which is all a minimizer
    // order needs.  (A second xor-shift-multiply round -- the first version -- made no difference to the runs per k-mer,
    // 0.1409 against 0.1405 bucket fetches per k-mer on config 3, and cost 2.4 % of the kernel's time.)
    return ((mmer ^ 0x2C1B3C6Du) * 0x9E3779B1u) & ~63u;  // the xor keeps poly-A (0) from always winning
}
// minimizer m-mer of a k-mer hash (first base in the lowest bits), k >= KMB_MZ_M, and its base offset inside the k-mer
KMB_HD uint32_t kmb_minimizer(uint64_t key, int k, uint32_t *offset) {
    uint32_t best = 0xFFFFFFFFu, best_mmer = 0;
    for (int j = 0; j + KMB_MZ_M <= k; j++) {
        const uint32_t x = (uint32_t)(key >> (2 * j)) & KMB_MZ_MASK;
        const uint32_t v = kmb_mmer_order(x) | (uint32_t)j;
        if (v < best) {
            best = v;
            best_mmer = x;
        }
    }
    *offset = best & 63u;
    return best_mmer;
}
// Bucket header (word 0 of the primary sector 2b): bits 0-2 = entries under this minimizer (5 = more than four),
// bits 3+5s .. 7+5s = minimizer offset of entry s (s = 0..3).  An entry can only equal the window that puts the
// run's minimizer at that offset, so a run costs one key comparison per ENTRY, not one per window.
// Secondary sector 2b+1: word 0 = 0, or KMB_HDR_CHAIN | first pool sector when there are more than four entries.
// The pool sectors of a bucket are contiguous; word 0 of each = entries left in the chain including its own
// (bits 0-21) | offset of its entry 0 (bits 22-26) | offset of its entry 1 (bits 27-31).
#define KMB_MZ_HDR_COUNT(h) ((h) & 7u)
#define KMB_MZ_HDR_OFFSET(h, s) (((h) >> (3u + 5u * (s))) & 31u)
#define KMB_MZ_POOL_MAX_ENTRIES ((1u << 22) - 1u)
#define KMB_MZ_POOL_LEFT(h) ((h) & KMB_MZ_POOL_MAX_ENTRIES)
#define KMB_MZ_POOL_OFFSET(h, t) (((h) >> (22u + 5u * (t))) & 31u)
KMB_HD uint64_t kmb_mz_sector(uint64_t bucket, uint32_t pool_base, uint32_t s) {
    return s < 4u ? 2ull * bucket + (s >> 1) : (uint64_t)pool_base + ((s - 4u) >> 1);
}
KMB_HD uint32_t kmb_mz_pool_sectors(uint32_t n_total) { return n_total > 4u ? (n_total - 3u) / 2u : 0u; }

// 64-bit mix (splitmix64 finaliser) for the synthetic-address generator of the gather benchmark.
KMB_HD uint64_t kmb_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
