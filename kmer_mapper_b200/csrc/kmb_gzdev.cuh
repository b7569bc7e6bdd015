// kmb_gzdev.cuh -- gzip members inflated on the device.
//
// The reads of BASELINE configs[1] arrive as FASTQ.gz.  The reference inflates them on the host inside bionumpy
// (bnp.open(path), command_line_interface.py:102-103); the host-side decoders of this library (kmb_gunzip.cpp,
// kmb_inflate.cpp) reach ~5 GB/s of text on 16 cores, which caps a B200 at ~14 M reads/s.  A multi-member .gz
// (bgzip / BGZF, `cat a.gz b.gz`, the benchmark's files) consists of independent deflate streams, so here the
// COMPRESSED bytes cross PCIe (a quarter of the text) and every member is decoded by one warp:
//
//   * lane 0 walks the bit stream (RFC 1951): 64-bit bit buffer refilled with two aligned loads, two-level decode
//     tables (10-bit primary for literals/lengths, 8-bit for distances) in the warp's shared memory, built per
//     block by the same lane; literals are stored as they come;
//   * a match (length, distance) is broadcast and copied by all 32 lanes -- byte i of the match is
//     out[pos - distance + i % distance], so overlapping matches (distance < length: the runs of equal quality
//     characters FASTQ is full of) need no serial copy;
//   * stored blocks are copied by the whole warp.
// Every member is then checked on the host: decoded length == the ISIZE of its trailer, the stream ended exactly
// where the next member begins, and (option gz_device_crc) its CRC-32.  Anything else -- a member too large for a
// warp's patience, a corrupt stream, a false member start -- makes the caller fall back to the host decoders from
// that member on.
//
// Measured and rejected: the last 32 KB of a member's text in a shared-memory ring (matches copied ring -> ring, the
// ring written out 16 KB at a time) with the stream words read two ahead.  A single warp gets faster (a file of 4 MB
// members: 0.93 -> 1.02 GB/s of text), but 44 KB of shared memory per warp leave 5 warps per SM instead of 16, and BGZF
// -- the files this route is for: tens of thousands of 64 KB members, all latency hidden by the other warps -- falls
// from 46 to 30 M reads/s (14.5 -> 9.3 GB/s of text).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define KMB_GZ_LIT_BITS 10
#define KMB_GZ_DIST_BITS 8
#define KMB_GZ_LIT_MAX 2048   // primary + sub-tables (zlib's bound for 286 symbols, 10 root bits, 15-bit codes is 1332)
#define KMB_GZ_DIST_MAX 768   //                                  (30 symbols, 8 root bits: 400)
#define KMB_GZ_WARPS 4        // members per CTA

#define KMB_GZ_OK 0u
#define KMB_GZ_ERR_HEADER 1u
#define KMB_GZ_ERR_STREAM 2u
#define KMB_GZ_ERR_TABLE 3u
#define KMB_GZ_ERR_OUTPUT 4u   // would write beyond the space its ISIZE announced
#define KMB_GZ_ERR_INPUT 5u    // ran past the end of its compressed bytes

struct KmbGzMember {
    unsigned long long in_off;   // first byte of the member (its gzip header) in the compressed buffer
    unsigned long long out_off;  // where its text goes
    uint32_t in_len;             // bytes up to the next member (or the end of the data)
    uint32_t out_len;            // ISIZE announced by its trailer
};
struct KmbGzResult {
    uint32_t status;
    uint32_t out_len;   // bytes produced
    uint32_t in_used;   // bytes consumed including the 8-byte trailer
    uint32_t crc;       // CRC-32 field of the trailer
};

// entry: bits 0-4 code bits consumed at this level, 5-9 extra bits, 10-12 kind, 16-31 value
#define KMB_GZ_K_LITERAL 0u
#define KMB_GZ_K_BASE 1u
#define KMB_GZ_K_END 2u
#define KMB_GZ_K_SUB 3u
#define KMB_GZ_K_INVALID 4u
__device__ __forceinline__ uint32_t kmb_gz_entry(uint32_t len, uint32_t extra, uint32_t kind, uint32_t value) {
    return len | (extra << 5) | (kind << 10) | (value << 16);
}

__constant__ uint16_t c_kmb_gz_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_kmb_gz_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_kmb_gz_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_kmb_gz_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_kmb_gz_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

__device__ __forceinline__ uint32_t kmb_gz_symbol_entry(int alphabet, int sym, uint32_t len) {
    if (alphabet == 2) return kmb_gz_entry(len, 0, KMB_GZ_K_LITERAL, (uint32_t)sym);
    if (alphabet == 1)
        return sym < 30 ? kmb_gz_entry(len, c_kmb_gz_dist_extra[sym], KMB_GZ_K_BASE, c_kmb_gz_dist_base[sym]) : kmb_gz_entry(len, 0, KMB_GZ_K_INVALID, 0);
    if (sym < 256) return kmb_gz_entry(len, 0, KMB_GZ_K_LITERAL, (uint32_t)sym);
    if (sym == 256) return kmb_gz_entry(len, 0, KMB_GZ_K_END, 0);
    if (sym < 286) return kmb_gz_entry(len, c_kmb_gz_len_extra[sym - 257], KMB_GZ_K_BASE, c_kmb_gz_len_base[sym - 257]);
    return kmb_gz_entry(len, 0, KMB_GZ_K_INVALID, 0);
}

// Canonical Huffman code (RFC 1951 3.2.2) -> two-level table indexed by the bit-reversed code.  One thread.
__device__ bool kmb_gz_build_table(const uint8_t *lens, int n_syms, int alphabet, int primary_bits, uint32_t *table, int table_max) {
    int count[16];
    for (int i = 0; i < 16; i++) count[i] = 0;
    for (int i = 0; i < n_syms; i++) count[lens[i]]++;
    count[0] = 0;
    int left = 1, max_len = 0, n_codes = 0;
    for (int l = 1; l <= 15; l++) {
        left = (left << 1) - count[l];
        if (left < 0) return false;
        if (count[l]) max_len = l;
        n_codes += count[l];
    }
    const int primary_size = 1 << primary_bits;
    for (int i = 0; i < primary_size; i++) table[i] = 0u;   // pass 1 collects the sub-table widths here
    if (n_codes == 0) {
        const uint32_t invalid = kmb_gz_entry(1, 0, KMB_GZ_K_INVALID, 0);
        for (int i = 0; i < primary_size; i++) table[i] = invalid;
        return alphabet == 1;
    }
    if (left > 0 && !(n_codes == 1 && max_len == 1)) return false;
    uint32_t first_code[16], next_code[16];
    {
        uint32_t code = 0;
        first_code[0] = 0;
        for (int l = 1; l <= 15; l++) {
            code = (code + (uint32_t)count[l - 1]) << 1;
            first_code[l] = code;
        }
    }
    // pass 1: width of the sub-table behind every primary prefix that long codes share
    if (max_len > primary_bits) {
        for (int l = 0; l < 16; l++) next_code[l] = first_code[l];
        for (int s = 0; s < n_syms; s++) {
            const int l = lens[s];
            if (!l) continue;
            const uint32_t code = next_code[l]++;
            if (l <= primary_bits) continue;
            const uint32_t rev = __brev(code) >> (32 - l);
            const uint32_t prefix = rev & (uint32_t)(primary_size - 1);
            const uint32_t w = (uint32_t)(l - primary_bits);
            if (table[prefix] < w) table[prefix] = w;
        }
    }
    int used = primary_size;
    const uint32_t invalid = kmb_gz_entry(1, 0, KMB_GZ_K_INVALID, 0);
    for (int p = 0; p < primary_size; p++) {
        const uint32_t w = table[p];
        if (!w) {
            table[p] = invalid;
            continue;
        }
        const int size = 1 << w;
        if (used + size > table_max) return false;
        for (int i = 0; i < size; i++) table[used + i] = invalid;
        table[p] = kmb_gz_entry((uint32_t)primary_bits, w, KMB_GZ_K_SUB, (uint32_t)used);
        used += size;
    }
    // pass 2: the entries
    for (int l = 0; l < 16; l++) next_code[l] = first_code[l];
    for (int s = 0; s < n_syms; s++) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t rev = __brev(next_code[l]++) >> (32 - l);
        if (l <= primary_bits) {
            const uint32_t e = kmb_gz_symbol_entry(alphabet, s, (uint32_t)l);
            for (uint32_t i = rev; i < (uint32_t)primary_size; i += 1u << l) table[i] = e;
        } else {
            const uint32_t sub = table[rev & (uint32_t)(primary_size - 1)];
            const uint32_t start = sub >> 16, sb = (sub >> 5) & 31u;
            const uint32_t e = kmb_gz_symbol_entry(alphabet, s, (uint32_t)(l - primary_bits));
            for (uint32_t i = rev >> primary_bits; i < (1u << sb); i += 1u << (l - primary_bits)) table[start + i] = e;
        }
    }
    return true;
}

struct KmbGzBits {   // lane 0's view of the compressed stream
    const uint8_t *base;   // 8-byte aligned address at or below the member's first byte
    uint64_t pos;          // next unread byte, relative to base
    uint64_t end;          // end of the member's bytes, relative to base
    uint64_t buf;
    int cnt;
};
// at least 56 bits in the buffer (zeros beyond the end: the caller checks pos against end when a block ends)
__device__ __forceinline__ void kmb_gz_refill(KmbGzBits &b) {
    const uint64_t *w = reinterpret_cast<const uint64_t *>(b.base + (b.pos & ~7ull));
    const uint32_t sh = (uint32_t)(b.pos & 7ull) * 8u;
    uint64_t v = w[0] >> sh;
    if (sh) v |= w[1] << (64u - sh);
    b.buf |= v << b.cnt;
    b.pos += (uint64_t)((63 - b.cnt) >> 3);
    b.cnt |= 56;
}
__device__ __forceinline__ uint32_t kmb_gz_take(KmbGzBits &b, int n) {
    const uint32_t v = (uint32_t)(b.buf & ((1ull << n) - 1ull));
    b.buf >>= n;
    b.cnt -= n;
    return v;
}
__device__ __forceinline__ void kmb_gz_align(KmbGzBits &b) {  // drop to the next byte boundary, give whole bytes back
    const int drop = b.cnt & 7;
    b.buf >>= drop;
    b.cnt -= drop;
    b.pos -= (uint64_t)(b.cnt >> 3);
    b.buf = 0;
    b.cnt = 0;
}

struct alignas(16) KmbGzShared {
    uint32_t lit[KMB_GZ_LIT_MAX];
    uint32_t dist[KMB_GZ_DIST_MAX];
    uint32_t pre[128];
    uint8_t lens[288 + 32 + 32];
};

// One warp per member.  `gz` must be readable 16 bytes past the last member (the host pads its buffer).
__global__ void __launch_bounds__(KMB_GZ_WARPS * 32) kmb_gz_inflate_kernel(const uint8_t *__restrict__ gz, const KmbGzMember *__restrict__ members,
                                                                            uint32_t n_members, uint8_t *__restrict__ out, KmbGzResult *__restrict__ results) {
    extern __shared__ __align__(16) unsigned char kmb_gz_smem[];
    KmbGzShared &S = reinterpret_cast<KmbGzShared *>(kmb_gz_smem)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * KMB_GZ_WARPS;
    for (uint32_t mi = blockIdx.x * KMB_GZ_WARPS + (threadIdx.x >> 5); mi < n_members; mi += warps) {
        const KmbGzMember M = members[mi];
        uint8_t *dst = out + M.out_off;
        uint32_t produced = 0;
        uint32_t status = KMB_GZ_OK;
        KmbGzBits B;
        B.base = gz + (M.in_off & ~7ull);
        B.pos = M.in_off & 7ull;
        B.end = B.pos + M.in_len;
        B.buf = 0;
        B.cnt = 0;
        // ---- gzip header (RFC 1952), lane 0
        if (lane == 0) {
            const uint8_t *p = B.base + B.pos;
            const uint8_t *e = B.base + B.end;
            if (M.in_len < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || (p[3] & 0xE0)) {
                status = KMB_GZ_ERR_HEADER;
            } else {
                const int flags = p[3];
                p += 10;
                if (flags & 4) {
                    const uint32_t xlen = (uint32_t)p[0] | ((uint32_t)p[1] << 8);
                    p += 2 + xlen;
                }
                for (int f = 8; f <= 16; f <<= 1) {
                    if (!(flags & f)) continue;
                    while (p < e && *p) p++;
                    p++;
                }
                if (flags & 2) p += 2;
                if (p + 8 > e) status = KMB_GZ_ERR_HEADER;
                B.pos = (uint64_t)(p - B.base);
            }
        }
        status = __shfl_sync(0xFFFFFFFFu, status, 0);
        bool last_block = false;
        // ---- deflate blocks
        while (status == KMB_GZ_OK && !last_block) {
            uint32_t type = 0, stored = 0;
            if (lane == 0) {
                kmb_gz_refill(B);
                last_block = kmb_gz_take(B, 1) != 0;
                type = kmb_gz_take(B, 2);
                if (type == 0) {
                    kmb_gz_align(B);
                    const uint8_t *p = B.base + B.pos;
                    const uint32_t len = (uint32_t)p[0] | ((uint32_t)p[1] << 8), nlen = (uint32_t)p[2] | ((uint32_t)p[3] << 8);
                    if (B.pos + 4 > B.end || (len ^ 0xFFFFu) != nlen) status = KMB_GZ_ERR_STREAM;
                    B.pos += 4;
                    stored = len;
                    if (B.pos + len > B.end) status = KMB_GZ_ERR_INPUT;
                } else if (type == 3) {
                    status = KMB_GZ_ERR_STREAM;
                } else {
                    int n_lit = 288, n_dist = 32;
                    uint8_t *lens = S.lens;
                    if (type == 1) {
                        for (int i = 0; i < 144; i++) lens[i] = 8;
                        for (int i = 144; i < 256; i++) lens[i] = 9;
                        for (int i = 256; i < 280; i++) lens[i] = 7;
                        for (int i = 280; i < 288; i++) lens[i] = 8;
                        for (int i = 0; i < 32; i++) lens[288 + i] = 5;
                    } else {
                        n_lit = (int)kmb_gz_take(B, 5) + 257;
                        n_dist = (int)kmb_gz_take(B, 5) + 1;
                        const int n_pre = (int)kmb_gz_take(B, 4) + 4;
                        if (n_lit > 286 || n_dist > 30) status = KMB_GZ_ERR_STREAM;
                        uint8_t *pre_lens = S.lens + 320;
                        for (int i = 0; i < 19; i++) pre_lens[i] = 0;
                        for (int i = 0; i < n_pre && status == KMB_GZ_OK; i++) {
                            if (B.cnt < 3) kmb_gz_refill(B);
                            pre_lens[c_kmb_gz_order[i]] = (uint8_t)kmb_gz_take(B, 3);
                        }
                        if (status == KMB_GZ_OK && !kmb_gz_build_table(pre_lens, 19, 2, 7, S.pre, 128)) status = KMB_GZ_ERR_TABLE;
                        int i = 0;
                        while (status == KMB_GZ_OK && i < n_lit + n_dist) {
                            if (B.cnt < 14) kmb_gz_refill(B);
                            const uint32_t e = S.pre[B.buf & 127u];
                            if (((e >> 10) & 7u) != KMB_GZ_K_LITERAL) {
                                status = KMB_GZ_ERR_STREAM;
                                break;
                            }
                            kmb_gz_take(B, (int)(e & 31u));
                            const uint32_t sym = e >> 16;
                            if (sym < 16) {
                                lens[i++] = (uint8_t)sym;
                                continue;
                            }
                            int rep;
                            uint8_t v = 0;
                            if (sym == 16) {
                                if (i == 0) {
                                    status = KMB_GZ_ERR_STREAM;
                                    break;
                                }
                                v = lens[i - 1];
                                rep = 3 + (int)kmb_gz_take(B, 2);
                            } else if (sym == 17) {
                                rep = 3 + (int)kmb_gz_take(B, 3);
                            } else {
                                rep = 11 + (int)kmb_gz_take(B, 7);
                            }
                            if (i + rep > n_lit + n_dist) {
                                status = KMB_GZ_ERR_STREAM;
                                break;
                            }
                            while (rep--) lens[i++] = v;
                        }
                        if (status == KMB_GZ_OK && lens[256] == 0) status = KMB_GZ_ERR_STREAM;
                        if (status == KMB_GZ_OK) {  // distance lengths follow the literal/length ones: to their own place
                            for (int j = n_dist - 1; j >= 0; j--) lens[288 + j] = lens[n_lit + j];
                            for (int j = n_lit; j < 288; j++) lens[j] = 0;
                            for (int j = n_dist; j < 32; j++) lens[288 + j] = 0;
                        }
                    }
                    if (status == KMB_GZ_OK && !kmb_gz_build_table(lens, type == 1 ? 288 : n_lit, 0, KMB_GZ_LIT_BITS, S.lit, KMB_GZ_LIT_MAX))
                        status = KMB_GZ_ERR_TABLE;
                    if (status == KMB_GZ_OK && !kmb_gz_build_table(lens + 288, type == 1 ? 32 : n_dist, 1, KMB_GZ_DIST_BITS, S.dist, KMB_GZ_DIST_MAX))
                        status = KMB_GZ_ERR_TABLE;
                }
            }
            status = __shfl_sync(0xFFFFFFFFu, status, 0);
            type = __shfl_sync(0xFFFFFFFFu, type, 0);
            last_block = __shfl_sync(0xFFFFFFFFu, (int)last_block, 0) != 0;
            if (status != KMB_GZ_OK) break;
            if (type == 0) {  // stored block: the whole warp copies
                stored = __shfl_sync(0xFFFFFFFFu, stored, 0);
                const uint64_t src = __shfl_sync(0xFFFFFFFFu, B.pos, 0);
                if (produced + stored > M.out_len) {
                    status = KMB_GZ_ERR_OUTPUT;
                    break;
                }
                for (uint32_t i = (uint32_t)lane; i < stored; i += 32u) dst[produced + i] = B.base[src + i];
                produced += stored;
                if (lane == 0) B.pos += stored;
                __syncwarp();
                continue;
            }
            // ---- coded block: lane 0 decodes, literals stored as they come; matches copied by the warp
            for (;;) {
                uint32_t mlen = 0, mdist = 0, lits = 0;   // what lane 0 found: `lits` literals written, then a match or the end
                uint32_t st = KMB_GZ_OK;
                bool end_of_block = false;
                if (lane == 0) {
                    uint32_t at = produced;
                    for (;;) {
                        if (B.cnt < 48) kmb_gz_refill(B);
                        uint32_t e = S.lit[B.buf & ((1u << KMB_GZ_LIT_BITS) - 1u)];
                        if (((e >> 10) & 7u) == KMB_GZ_K_SUB) {
                            B.buf >>= KMB_GZ_LIT_BITS;
                            B.cnt -= KMB_GZ_LIT_BITS;
                            e = S.lit[(e >> 16) + (uint32_t)(B.buf & ((1ull << ((e >> 5) & 31u)) - 1ull))];
                        }
                        const uint32_t kind = (e >> 10) & 7u;
                        B.buf >>= (e & 31u);
                        B.cnt -= (int)(e & 31u);
                        if (kind == KMB_GZ_K_LITERAL) {
                            if (at >= M.out_len) {
                                st = KMB_GZ_ERR_OUTPUT;
                                break;
                            }
                            dst[at++] = (uint8_t)(e >> 16);
                            continue;
                        }
                        if (kind == KMB_GZ_K_END) {
                            end_of_block = true;
                            break;
                        }
                        if (kind != KMB_GZ_K_BASE) {
                            st = KMB_GZ_ERR_STREAM;
                            break;
                        }
                        const uint32_t xl = (e >> 5) & 31u;
                        mlen = (e >> 16) + (uint32_t)(B.buf & ((1ull << xl) - 1ull));
                        B.buf >>= xl;
                        B.cnt -= (int)xl;
                        uint32_t d = S.dist[B.buf & ((1u << KMB_GZ_DIST_BITS) - 1u)];
                        if (((d >> 10) & 7u) == KMB_GZ_K_SUB) {
                            B.buf >>= KMB_GZ_DIST_BITS;
                            B.cnt -= KMB_GZ_DIST_BITS;
                            d = S.dist[(d >> 16) + (uint32_t)(B.buf & ((1ull << ((d >> 5) & 31u)) - 1ull))];
                        }
                        if (((d >> 10) & 7u) != KMB_GZ_K_BASE) {
                            st = KMB_GZ_ERR_STREAM;
                            break;
                        }
                        B.buf >>= (d & 31u);
                        B.cnt -= (int)(d & 31u);
                        const uint32_t xd = (d >> 5) & 31u;
                        if (B.cnt < (int)xd) kmb_gz_refill(B);
                        mdist = (d >> 16) + (uint32_t)(B.buf & ((1ull << xd) - 1ull));
                        B.buf >>= xd;
                        B.cnt -= (int)xd;
                        if (mdist > at || at + mlen > M.out_len) st = mdist > at ? KMB_GZ_ERR_STREAM : KMB_GZ_ERR_OUTPUT;
                        break;
                    }
                    lits = at - produced;
                    if (B.pos - (uint64_t)(B.cnt >> 3) > B.end) st = KMB_GZ_ERR_INPUT;
                }
                st = __shfl_sync(0xFFFFFFFFu, st, 0);
                lits = __shfl_sync(0xFFFFFFFFu, lits, 0);
                mlen = __shfl_sync(0xFFFFFFFFu, mlen, 0);
                mdist = __shfl_sync(0xFFFFFFFFu, mdist, 0);
                end_of_block = __shfl_sync(0xFFFFFFFFu, (int)end_of_block, 0) != 0;
                produced += lits;
                if (st != KMB_GZ_OK) {
                    status = st;
                    break;
                }
                if (end_of_block) break;
                // the match: byte i = out[produced - mdist + i % mdist]  (everything before `produced` is written and,
                // after the shuffles above, visible to the whole warp)
                __syncwarp();
                const uint8_t *src = dst + produced - mdist;
                if (mdist >= mlen) {
                    for (uint32_t i = (uint32_t)lane; i < mlen; i += 32u) dst[produced + i] = src[i];
                } else {
                    for (uint32_t i = (uint32_t)lane; i < mlen; i += 32u) dst[produced + i] = src[i % mdist];
                }
                produced += mlen;
                __syncwarp();
            }
        }
        // ---- trailer
        uint32_t in_used = 0, crc = 0;
        if (lane == 0) {
            if (status == KMB_GZ_OK) {
                kmb_gz_align(B);
                if (B.pos + 8 > B.end) {
                    status = KMB_GZ_ERR_INPUT;
                } else {
                    const uint8_t *p = B.base + B.pos;
                    crc = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
                    const uint32_t isize = (uint32_t)p[4] | ((uint32_t)p[5] << 8) | ((uint32_t)p[6] << 16) | ((uint32_t)p[7] << 24);
                    if (isize != produced) status = KMB_GZ_ERR_STREAM;
                    B.pos += 8;
                    in_used = (uint32_t)(B.pos - (M.in_off & 7ull));
                }
            }
            KmbGzResult r;
            r.status = status;
            r.out_len = produced;
            r.in_used = in_used;
            r.crc = crc;
            results[mi] = r;
        }
        __syncwarp();
    }
}

// CRC-32 (RFC 1952) of every member's text: one warp per member, every lane the CRC of one 32nd of it (byte-wise, table
// in shared memory), combined by lane 0 with the x^n-mod-P shift operator (the algebra of zlib's crc32_combine:
// crc(A || B) = crc(A) * x^(8 |B|) mod P  xor  crc(B), on final CRC values).  crc_out[m] is compared with the trailer.
__device__ __forceinline__ uint32_t kmb_crc_multmodp(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1u)) == 0u) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
__global__ void __launch_bounds__(128) kmb_gz_crc_kernel(const uint8_t *__restrict__ text, const KmbGzMember *__restrict__ members,
                                                          uint32_t n_members, uint32_t *__restrict__ crc_out) {
    __shared__ uint32_t s_table[256];
    __shared__ uint32_t s_x2n[32];   // x^(2^k) mod P
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = (uint32_t)i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        s_table[i] = c;
    }
    if (threadIdx.x == 0) {
        uint32_t p = 1u << 30;   // x^1
        s_x2n[0] = p;
        for (int k = 1; k < 32; k++) s_x2n[k] = p = kmb_crc_multmodp(p, p);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < n_members; m += warps) {
        const uint8_t *p = text + members[m].out_off;
        const uint32_t n = members[m].out_len;
        const uint32_t per = (n + 31u) / 32u;
        const uint32_t lo = min(n, per * (uint32_t)lane), hi = min(n, lo + per);
        uint32_t c = 0xFFFFFFFFu;
        for (uint32_t i = lo; i < hi; i++) c = s_table[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
        c ^= 0xFFFFFFFFu;
        uint32_t total = 0;   // lane 0 folds the 32 pieces together, left to right
        for (int j = 0; j < 32; j++) {
            const uint32_t cj = __shfl_sync(0xFFFFFFFFu, c, j);
            const uint32_t lj = __shfl_sync(0xFFFFFFFFu, hi - lo, j);
            if (lane == 0 && lj) {
                uint32_t xp = 1u << 31, nn = lj, k = 3;   // x^(8 lj) mod P
                while (nn) {
                    if (nn & 1u) xp = kmb_crc_multmodp(s_x2n[k & 31u], xp);
                    nn >>= 1;
                    k++;
                }
                total = kmb_crc_multmodp(xp, total) ^ cj;
            }
        }
        if (lane == 0) crc_out[m] = total;
    }
}
