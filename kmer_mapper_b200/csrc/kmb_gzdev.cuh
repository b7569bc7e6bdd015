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
//     block by the same lane;
//   * the last 32 KB of the member's text -- everything a match can refer to -- live in a ring in shared memory:
//     literals are stored there, matches are copied ring -> ring (a first version kept the text in global memory
//     only: every match then waited for an L2 round trip and a warp inflated 15 MB/s);
//   * whenever 16 KB are waiting the whole warp writes them to global memory in 16-byte vectors (the ring is
//     indexed by the low bits of the output ADDRESS, so ring and output are aligned alike).
// Every member is then checked on the host: decoded length == the ISIZE of its trailer, the stream ended exactly
// where the next member begins, and (option gz_device_crc) its CRC-32.  Anything else -- a member too large for a
// warp's patience, a corrupt stream, a false member start -- makes the caller fall back to the host decoders from
// that member on.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define KMB_GZ_LIT_BITS 10
#define KMB_GZ_DIST_BITS 8
#define KMB_GZ_LIT_MAX 2048   // primary + sub-tables (zlib's bound for 286 symbols, 10 root bits, 15-bit codes is 1332)
#define KMB_GZ_DIST_MAX 768   //                                  (30 symbols, 8 root bits: 400)

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

// Lane 0's view of the compressed stream.  The stream is read in aligned 8-byte words, two words AHEAD of the bits being
// decoded (`next`, partly moved into the bit buffer, and `after`, untouched), so that topping the bit buffer up never
// waits for memory: the load issued when the reader crosses into a new word is needed one word -- several symbols --
// later.  (With the two loads inside every refill a warp inflated 15 MB/s: the decode loop is one dependent chain and
// each refill put an L1 round trip on it.)
struct KmbGzBits {
    const uint64_t *words;   // 8-byte aligned address at or below the member's first byte
    uint64_t end;            // end of the member's bytes, relative to words
    uint64_t buf;            // bit buffer: the low `cnt` bits are unread stream bits (bits above them are stream bits too)
    uint64_t next, after;    // words[wi], words[wi + 1]
    uint32_t wi;             // index of `next`
    uint32_t c;              // bytes of `next` already moved into buf
    int cnt;
};
__device__ __forceinline__ void kmb_gz_bits_init(KmbGzBits &b, uint64_t byte_pos) {   // start reading at this byte
    b.wi = (uint32_t)(byte_pos >> 3);
    b.c = (uint32_t)(byte_pos & 7ull);
    b.next = b.words[b.wi];
    b.after = b.words[b.wi + 1];
    b.buf = 0;
    b.cnt = 0;
}
// byte position (relative to words) of the first byte none of whose bits has been handed out
__device__ __forceinline__ uint64_t kmb_gz_bits_pos(const KmbGzBits &b) {
    return ((uint64_t)b.wi << 3) + b.c - (uint64_t)(b.cnt >> 3);
}
// at least 56 bits in the buffer (zeros / foreign bytes beyond the member's end: the caller checks the position)
__device__ __forceinline__ void kmb_gz_refill(KmbGzBits &b) {
    uint32_t k = (uint32_t)(63 - b.cnt) >> 3;   // whole bytes that fit
    const uint32_t avail = 8u - b.c;
    if (k >= avail) {                            // everything left of `next`, then on into `after`
        b.buf |= (b.next >> (8u * b.c)) << b.cnt;
        b.cnt += (int)(8u * avail);
        k -= avail;
        b.next = b.after;
        b.wi++;
        b.after = b.words[b.wi + 1];
        b.c = 0;
    }
    b.buf |= (b.next >> (8u * b.c)) << b.cnt;    // (shift count < 64: c <= 7, cnt <= 63)
    b.cnt += (int)(8u * k);
    b.c += k;
}
__device__ __forceinline__ uint32_t kmb_gz_take(KmbGzBits &b, int n) {
    const uint32_t v = (uint32_t)(b.buf & ((1ull << n) - 1ull));
    b.buf >>= n;
    b.cnt -= n;
    return v;
}
// drop to the next byte boundary and return its position; the reader has to be re-initialised before it is used again
__device__ __forceinline__ uint64_t kmb_gz_align(KmbGzBits &b) {
    b.cnt -= b.cnt & 7;
    return kmb_gz_bits_pos(b);
}

#define KMB_GZ_RING 32768u    // DEFLATE looks back at most 32768 bytes
#define KMB_GZ_FLUSH 16384u   // the warp writes the ring out when this much is waiting (+ 258 for a match in progress < the ring)
struct alignas(16) KmbGzShared {
    uint8_t ring[KMB_GZ_RING];   // the last 32 KB of the member's text; byte p lives at (address of p in `out`) mod 32768
    uint32_t lit[KMB_GZ_LIT_MAX];
    uint32_t dist[KMB_GZ_DIST_MAX];
    uint32_t pre[128];
    uint8_t lens[288 + 32 + 32];
};

// why lane 0 came back from kmb_gz_run
#define KMB_GZ_R_FLUSH 0u   // KMB_GZ_FLUSH bytes are waiting in the ring
#define KMB_GZ_R_DONE 1u    // the member's last block has ended
#define KMB_GZ_R_ERROR 2u

struct KmbGzState {   // lane 0's registers across calls of kmb_gz_run
    KmbGzBits B;
    uint32_t pos;        // bytes of text produced
    uint32_t flushed;    // bytes of text written to global memory
    uint32_t stored;     // bytes left of a stored block
    uint64_t stored_at;  // and where they are in the input
    uint32_t status;
    bool in_block, last_block, coded;
};

// Lane 0: decode until the ring has to be written out, the member ends or the stream is bad.
__device__ __forceinline__ uint32_t kmb_gz_run(KmbGzState &T, KmbGzShared &S, const uint32_t ring0, const uint32_t out_len) {
    KmbGzBits &B = T.B;
    uint8_t *ring = S.ring;
    for (;;) {
        if (!T.in_block) {
            if (T.last_block) return KMB_GZ_R_DONE;
            // ---- block header
            kmb_gz_refill(B);
            T.last_block = kmb_gz_take(B, 1) != 0;
            const uint32_t type = kmb_gz_take(B, 2);
            if (type == 0) {
                const uint64_t at0 = kmb_gz_align(B);
                const uint8_t *p = reinterpret_cast<const uint8_t *>(B.words) + at0;
                const uint32_t len = (uint32_t)p[0] | ((uint32_t)p[1] << 8), nlen = (uint32_t)p[2] | ((uint32_t)p[3] << 8);
                if (at0 + 4 > B.end || (len ^ 0xFFFFu) != nlen) return T.status = KMB_GZ_ERR_STREAM, KMB_GZ_R_ERROR;
                if (at0 + 4 + len > B.end) return T.status = KMB_GZ_ERR_INPUT, KMB_GZ_R_ERROR;
                if (T.pos + len > out_len) return T.status = KMB_GZ_ERR_OUTPUT, KMB_GZ_R_ERROR;
                T.stored = len;
                T.stored_at = at0 + 4;
                T.coded = false;
            } else if (type == 3) {
                return T.status = KMB_GZ_ERR_STREAM, KMB_GZ_R_ERROR;
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
                    if (n_lit > 286 || n_dist > 30) return T.status = KMB_GZ_ERR_STREAM, KMB_GZ_R_ERROR;
                    uint8_t *pre_lens = S.lens + 320;
                    for (int i = 0; i < 19; i++) pre_lens[i] = 0;
                    for (int i = 0; i < n_pre; i++) {
                        if (B.cnt < 3) kmb_gz_refill(B);
                        pre_lens[c_kmb_gz_order[i]] = (uint8_t)kmb_gz_take(B, 3);
                    }
                    if (!kmb_gz_build_table(pre_lens, 19, 2, 7, S.pre, 128)) return T.status = KMB_GZ_ERR_TABLE, KMB_GZ_R_ERROR;
                    int i = 0;
                    while (i < n_lit + n_dist) {
                        if (B.cnt < 14) kmb_gz_refill(B);
                        const uint32_t e = S.pre[B.buf & 127u];
                        if (((e >> 10) & 7u) != KMB_GZ_K_LITERAL) return T.status = KMB_GZ_ERR_STREAM, KMB_GZ_R_ERROR;
                        kmb_gz_take(B, (int)(e & 31u));
                        const uint32_t sym = e >> 16;
                        if (sym < 16) {
                            lens[i++] = (uint8_t)sym;
                            continue;
                        }
                        int rep;
                        uint8_t v = 0;
                        if (sym == 16) {
                            if (i == 0) return T.status = KMB_GZ_ERR_STREAM, KMB_GZ_R_ERROR;
                            v = lens[i - 1];
                            rep = 3 + (int)kmb_gz_take(B, 2);
                        } else if (sym == 17) {
                            rep = 3 + (int)kmb_gz_take(B, 3);
                        } else {
                            rep = 11 + (int)kmb_gz_take(B, 7);
                        }
                        if (i + rep > n_lit + n_dist) return T.status = KMB_GZ_ERR_STREAM, KMB_GZ_R_ERROR;
                        while (rep--) lens[i++] = v;
                    }
                    if (lens[256] == 0) return T.status = KMB_GZ_ERR_STREAM, KMB_GZ_R_ERROR;
                    // distance lengths follow the literal/length ones: to their own place
                    for (int j = n_dist - 1; j >= 0; j--) lens[288 + j] = lens[n_lit + j];
                    for (int j = n_lit; j < 288; j++) lens[j] = 0;
                    for (int j = n_dist; j < 32; j++) lens[288 + j] = 0;
                }
                if (!kmb_gz_build_table(lens, type == 1 ? 288 : n_lit, 0, KMB_GZ_LIT_BITS, S.lit, KMB_GZ_LIT_MAX) ||
                    !kmb_gz_build_table(lens + 288, type == 1 ? 32 : n_dist, 1, KMB_GZ_DIST_BITS, S.dist, KMB_GZ_DIST_MAX))
                    return T.status = KMB_GZ_ERR_TABLE, KMB_GZ_R_ERROR;
                T.coded = true;
            }
            T.in_block = true;
        }
        if (!T.coded) {
            // ---- stored block: bytes from the input to the ring, as many as the ring has room for
            while (T.stored) {
                if (T.pos - T.flushed >= KMB_GZ_FLUSH) return KMB_GZ_R_FLUSH;
                const uint32_t n = min(T.stored, KMB_GZ_FLUSH - (T.pos - T.flushed));
                const uint8_t *src = reinterpret_cast<const uint8_t *>(B.words) + T.stored_at;
                for (uint32_t i = 0; i < n; i++) ring[(ring0 + T.pos + i) & (KMB_GZ_RING - 1u)] = src[i];
                T.pos += n;
                T.stored_at += n;
                T.stored -= n;
            }
            kmb_gz_bits_init(B, T.stored_at);   // the bit stream goes on after the stored bytes
            T.in_block = false;
            continue;
        }
        // ---- coded block: everything stays in the ring (literals stored, matches copied ring -> ring)
        uint32_t at = T.pos;
        for (;;) {
            if (at - T.flushed >= KMB_GZ_FLUSH) {
                T.pos = at;
                return KMB_GZ_R_FLUSH;
            }
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
                if (at >= out_len) return T.pos = at, T.status = KMB_GZ_ERR_OUTPUT, KMB_GZ_R_ERROR;
                ring[(ring0 + at) & (KMB_GZ_RING - 1u)] = (uint8_t)(e >> 16);
                at++;
                continue;
            }
            if (kind == KMB_GZ_K_END) {
                T.in_block = false;
                break;
            }
            if (kind != KMB_GZ_K_BASE) return T.pos = at, T.status = KMB_GZ_ERR_STREAM, KMB_GZ_R_ERROR;
            const uint32_t xl = (e >> 5) & 31u;
            const uint32_t mlen = (e >> 16) + (uint32_t)(B.buf & ((1ull << xl) - 1ull));
            B.buf >>= xl;
            B.cnt -= (int)xl;
            uint32_t d = S.dist[B.buf & ((1u << KMB_GZ_DIST_BITS) - 1u)];
            if (((d >> 10) & 7u) == KMB_GZ_K_SUB) {
                B.buf >>= KMB_GZ_DIST_BITS;
                B.cnt -= KMB_GZ_DIST_BITS;
                d = S.dist[(d >> 16) + (uint32_t)(B.buf & ((1ull << ((d >> 5) & 31u)) - 1ull))];
            }
            if (((d >> 10) & 7u) != KMB_GZ_K_BASE) return T.pos = at, T.status = KMB_GZ_ERR_STREAM, KMB_GZ_R_ERROR;
            B.buf >>= (d & 31u);
            B.cnt -= (int)(d & 31u);
            const uint32_t xd = (d >> 5) & 31u;
            if (B.cnt < (int)xd) kmb_gz_refill(B);
            const uint32_t mdist = (d >> 16) + (uint32_t)(B.buf & ((1ull << xd) - 1ull));
            B.buf >>= xd;
            B.cnt -= (int)xd;
            if (mdist > at) return T.pos = at, T.status = KMB_GZ_ERR_STREAM, KMB_GZ_R_ERROR;
            if (at + mlen > out_len) return T.pos = at, T.status = KMB_GZ_ERR_OUTPUT, KMB_GZ_R_ERROR;
            // byte i of the match = byte (at - mdist + i): in order, so an overlapping match (mdist < mlen) reads what
            // it has just written; ring indices wrap
            uint32_t w = (ring0 + at) & (KMB_GZ_RING - 1u), r = (w - mdist) & (KMB_GZ_RING - 1u);
            if (mdist >= 4u && w + mlen <= KMB_GZ_RING && r + mlen <= KMB_GZ_RING) {
                uint32_t i = 0;
                for (; i + 4u <= mlen; i += 4u) {   // four independent loads, then four stores (mdist >= 4: no overlap inside a quad)
                    const uint8_t b0 = ring[r + i], b1 = ring[r + i + 1], b2 = ring[r + i + 2], b3 = ring[r + i + 3];
                    ring[w + i] = b0;
                    ring[w + i + 1] = b1;
                    ring[w + i + 2] = b2;
                    ring[w + i + 3] = b3;
                }
                for (; i < mlen; i++) ring[w + i] = ring[r + i];
            } else {
                for (uint32_t i = 0; i < mlen; i++) ring[(w + i) & (KMB_GZ_RING - 1u)] = ring[(r + i) & (KMB_GZ_RING - 1u)];
            }
            at += mlen;
        }
        T.pos = at;
        if (kmb_gz_bits_pos(B) > B.end) return T.status = KMB_GZ_ERR_INPUT, KMB_GZ_R_ERROR;
    }
}

// One warp per member (one-warp CTAs: 45 KB of shared memory each, five per SM).  Lane 0 walks the bit stream and keeps
// the member's text in a 32 KB ring in shared memory (kmb_gz_run); whenever 16 KB are waiting the whole warp writes them
// to global memory in 16-byte vectors.  `gz` must be readable 16 bytes past the last member (the host pads its buffer).
__global__ void __launch_bounds__(32) kmb_gz_inflate_kernel(const uint8_t *__restrict__ gz, const KmbGzMember *__restrict__ members,
                                                            uint32_t n_members, uint8_t *__restrict__ out, KmbGzResult *__restrict__ results) {
    extern __shared__ __align__(16) unsigned char kmb_gz_smem[];
    KmbGzShared &S = *reinterpret_cast<KmbGzShared *>(kmb_gz_smem);
    const int lane = threadIdx.x & 31;
    for (uint32_t mi = blockIdx.x; mi < n_members; mi += gridDim.x) {
        const KmbGzMember M = members[mi];
        uint8_t *dst = out + M.out_off;
        const uint32_t ring0 = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & (KMB_GZ_RING - 1u));
        KmbGzState T;
        T.B.words = reinterpret_cast<const uint64_t *>(gz + (M.in_off & ~7ull));
        T.B.end = (M.in_off & 7ull) + M.in_len;
        T.B.buf = T.B.next = T.B.after = 0;
        T.B.wi = T.B.c = 0;
        T.B.cnt = 0;
        T.pos = T.flushed = T.stored = 0;
        T.stored_at = 0;
        T.status = KMB_GZ_OK;
        T.in_block = T.last_block = T.coded = false;
        // ---- gzip header (RFC 1952), lane 0
        if (lane == 0) {
            const uint8_t *p0 = reinterpret_cast<const uint8_t *>(T.B.words);
            const uint8_t *p = p0 + (M.in_off & 7ull);
            const uint8_t *e = p0 + T.B.end;
            if (M.in_len < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || (p[3] & 0xE0)) {
                T.status = KMB_GZ_ERR_HEADER;
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
                if (p + 8 > e) T.status = KMB_GZ_ERR_HEADER;
                else kmb_gz_bits_init(T.B, (uint64_t)(p - p0));
            }
        }
        uint32_t reason = KMB_GZ_R_ERROR;
        for (;;) {
            if (lane == 0) reason = T.status != KMB_GZ_OK ? KMB_GZ_R_ERROR : kmb_gz_run(T, S, ring0, M.out_len);
            __syncwarp();   // lane 0's writes to the ring are visible to the warp
            reason = __shfl_sync(0xFFFFFFFFu, reason, 0);
            const uint32_t pos = __shfl_sync(0xFFFFFFFFu, T.pos, 0);
            uint32_t flushed = __shfl_sync(0xFFFFFFFFu, T.flushed, 0);
            // ---- write [flushed, upto) out: single bytes up to a 16-byte boundary of the output, then whole vectors; the
            // rest waits for the next round unless this is the last one
            const bool final_round = reason != KMB_GZ_R_FLUSH;
            uintptr_t a = reinterpret_cast<uintptr_t>(dst) + flushed;
            const uint32_t head = min((uint32_t)((16u - (a & 15u)) & 15u), pos - flushed);
            if ((uint32_t)lane < head) dst[flushed + lane] = S.ring[(ring0 + flushed + lane) & (KMB_GZ_RING - 1u)];
            flushed += head;
            const uint32_t n_vec = (pos - flushed) >> 4;
            for (uint32_t v = (uint32_t)lane; v < n_vec; v += 32u) {
                const uint32_t p = flushed + (v << 4);
                *reinterpret_cast<uint4 *>(dst + p) = *reinterpret_cast<const uint4 *>(&S.ring[(ring0 + p) & (KMB_GZ_RING - 1u)]);
            }
            flushed += n_vec << 4;
            if (final_round) {
                if ((uint32_t)lane < pos - flushed) dst[flushed + lane] = S.ring[(ring0 + flushed + lane) & (KMB_GZ_RING - 1u)];
                flushed = pos;
            }
            T.flushed = flushed;   // (every lane keeps a copy; lane 0's is the one kmb_gz_run reads)
            __syncwarp();          // the ring may be overwritten from here on
            if (final_round) break;
        }
        // ---- trailer
        if (lane == 0) {
            uint32_t status = T.status, in_used = 0, crc = 0;
            if (status == KMB_GZ_OK) {
                const uint64_t at0 = kmb_gz_align(T.B);
                if (at0 + 8 > T.B.end) {
                    status = KMB_GZ_ERR_INPUT;
                } else {
                    const uint8_t *p = reinterpret_cast<const uint8_t *>(T.B.words) + at0;
                    crc = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
                    const uint32_t isize = (uint32_t)p[4] | ((uint32_t)p[5] << 8) | ((uint32_t)p[6] << 16) | ((uint32_t)p[7] << 24);
                    if (isize != T.pos) status = KMB_GZ_ERR_STREAM;
                    in_used = (uint32_t)(at0 + 8 - (M.in_off & 7ull));
                }
            }
            KmbGzResult r;
            r.status = status;
            r.out_len = T.pos;
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
