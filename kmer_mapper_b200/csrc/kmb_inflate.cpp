// kmb_inflate.cpp -- streaming gzip decoder for single-member .gz read files (include/kmer_mapper_b200.h, "reader").
//
// A plain `gzip reads.fq` file is ONE deflate stream: it cannot be inflated member-parallel (kmb_gunzip.cpp), and
// zlib's inflate, at ~0.45 GB/s of text, then caps the whole pipeline at ~1.4 M reads/s -- three orders of
// magnitude below what the mapping kernel eats (the reference reads such files through Python's gzip inside
// bionumpy, command_line_interface.py:102-103).  This is a from-scratch DEFLATE (RFC 1951) decoder built for
// throughput on one core: 64-bit bit buffer refilled with one unaligned load, 11-bit primary decode tables whose
// entries carry base value, extra-bit count and code length, up to three literals per refill, matches copied in
// 16-byte words, and an inner loop without exhaustion checks while 16 input and 274 output bytes remain.  Every member is verified against the CRC-32 and length of its gzip trailer (RFC 1952), the CRC
// computed by zlib's crc32 on the worker pool, so a decoder fault cannot go unnoticed.
//
// Streaming: the caller owns the output buffers; each call continues the stream into `out`, and the 32 KB of
// history the format may refer back to are expected directly in front of `out` (the caller copies the tail of the
// previous block there: reader.py keeps 1 MB of headroom in front of every block anyway).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <functional>
#include <new>
#include <vector>

#include "../../include/kmer_mapper_b200.h"
#include "kmb_host.h"

namespace {

// ---- decode tables -------------------------------------------------------------------------------------
// Entry layout (uint32): bits 0-7 code length consumed from the bit buffer (primary: this level's bits),
//   bits 8-12 number of extra bits, bits 13-15 kind, bits 16-31 value (literal byte / base length / base
//   distance / sub-table start).
enum : uint32_t { K_LITERAL = 0, K_BASE = 1, K_END = 2, K_SUBTABLE = 3, K_INVALID = 4 };
inline uint32_t make_entry(uint32_t len, uint32_t extra, uint32_t kind, uint32_t value) {
    return len | (extra << 8) | (kind << 13) | (value << 16);
}
inline uint32_t e_len(uint32_t e) { return e & 0xFFu; }
inline uint32_t e_extra(uint32_t e) { return (e >> 8) & 0x1Fu; }
inline uint32_t e_kind(uint32_t e) { return (e >> 13) & 7u; }
inline uint32_t e_value(uint32_t e) { return e >> 16; }

const int LIT_BITS = 11, DIST_BITS = 8, PRE_BITS = 7;
const int LIT_TABLE_MAX = (1 << LIT_BITS) + 2048;   // primary + all sub-tables (codes up to 15 bits)
const int DIST_TABLE_MAX = (1 << DIST_BITS) + 1024;

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; i++) r |= ((code >> i) & 1u) << (len - 1 - i);
    return r;
}

// what a symbol of the given alphabet decodes to (alphabet 0 = literal/length, 1 = distance, 2 = code lengths)
inline uint32_t symbol_entry(int alphabet, int sym, uint32_t len) {
    if (alphabet == 2) return make_entry(len, 0, K_LITERAL, (uint32_t)sym);
    if (alphabet == 1) return sym < 30 ? make_entry(len, DIST_EXTRA[sym], K_BASE, DIST_BASE[sym]) : make_entry(len, 0, K_INVALID, 0);
    if (sym < 256) return make_entry(len, 0, K_LITERAL, (uint32_t)sym);
    if (sym == 256) return make_entry(len, 0, K_END, 0);
    if (sym < 286) return make_entry(len, LEN_EXTRA[sym - 257], K_BASE, LEN_BASE[sym - 257]);
    return make_entry(len, 0, K_INVALID, 0);
}

// Canonical Huffman code (RFC 1951 3.2.2) -> two-level decode table indexed by the bit-reversed code.
// Returns false for an over-subscribed code, or an incomplete one that is not the single-code special case.
bool build_table(const uint8_t *lens, int n_syms, int alphabet, int primary_bits, uint32_t *table, int table_max) {
    int count[16] = {0};
    for (int i = 0; i < n_syms; i++) count[lens[i]]++;
    count[0] = 0;
    int left = 1, max_len = 0, n_codes = 0;
    for (int l = 1; l <= 15; l++) {
        left = (left << 1) - count[l];
        if (left < 0) return false;  // over-subscribed
        if (count[l]) max_len = l;
        n_codes += count[l];
    }
    const uint32_t invalid = make_entry(1, 0, K_INVALID, 0);
    const int primary_size = 1 << primary_bits;
    for (int i = 0; i < primary_size; i++) table[i] = invalid;
    if (n_codes == 0) return alphabet == 1;  // no distance codes at all is legal (a block of literals only)
    if (left > 0 && !(n_codes == 1 && max_len == 1)) return false;  // incomplete (RFC allows only the one-code case)
    uint32_t next_code[16] = {0};
    {
        uint32_t code = 0;
        for (int l = 1; l <= 15; l++) {
            code = (code + (uint32_t)count[l - 1]) << 1;
            next_code[l] = code;
        }
    }
    int used = primary_size;  // next free sub-table slot
    // sub-tables are created per distinct primary prefix; remember where each one starts and how wide it is
    std::vector<int> sub_start((size_t)primary_size, -1), sub_bits((size_t)primary_size, 0);
    if (max_len > primary_bits) {
        // width of the sub-table behind a prefix = longest code with that prefix - primary_bits
        std::vector<uint32_t> code_of((size_t)n_syms);
        uint32_t nc[16];
        memcpy(nc, next_code, sizeof(nc));
        for (int s = 0; s < n_syms; s++)
            if (lens[s]) code_of[(size_t)s] = nc[lens[s]]++;
        for (int s = 0; s < n_syms; s++) {
            const int l = lens[s];
            if (l <= primary_bits) continue;
            const uint32_t rev = reverse_bits(code_of[(size_t)s], l);
            const int prefix = (int)(rev & (uint32_t)(primary_size - 1));
            sub_bits[(size_t)prefix] = std::max(sub_bits[(size_t)prefix], l - primary_bits);
        }
        for (int p = 0; p < primary_size; p++) {
            if (!sub_bits[(size_t)p]) continue;
            const int size = 1 << sub_bits[(size_t)p];
            if (used + size > table_max) return false;
            sub_start[(size_t)p] = used;
            for (int i = 0; i < size; i++) table[used + i] = invalid;
            table[p] = make_entry((uint32_t)primary_bits, (uint32_t)sub_bits[(size_t)p], K_SUBTABLE, (uint32_t)used);
            used += size;
        }
    }
    for (int s = 0; s < n_syms; s++) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t rev = reverse_bits(next_code[l]++, l);
        if (l <= primary_bits) {
            const uint32_t e = symbol_entry(alphabet, s, (uint32_t)l);
            for (uint32_t i = rev; i < (uint32_t)primary_size; i += 1u << l) table[i] = e;
        } else {
            const int prefix = (int)(rev & (uint32_t)(primary_size - 1));
            const int sb = sub_bits[(size_t)prefix];
            const uint32_t e = symbol_entry(alphabet, s, (uint32_t)(l - primary_bits));
            for (uint32_t i = rev >> primary_bits; i < (1u << sb); i += 1u << (l - primary_bits)) table[sub_start[(size_t)prefix] + (int)i] = e;
        }
    }
    return true;
}

// ---- the stream ----------------------------------------------------------------------------------------
struct Stream {
    const uint8_t *in_begin = nullptr, *in_end = nullptr, *in = nullptr;  // the whole compressed file (mapped)
    uint64_t bitbuf = 0;
    int bitcnt = 0;
    // member / block state
    enum Phase { MEMBER_HEADER, BLOCK_HEADER, STORED, CODED, MEMBER_TRAILER, DONE, FAILED } phase = MEMBER_HEADER;
    bool last_block = false;
    uint32_t stored_left = 0;
    uint32_t lit_table[LIT_TABLE_MAX];
    uint32_t dist_table[DIST_TABLE_MAX];
    // verification of the current member
    uint32_t crc = 0;
    uint64_t member_out = 0;
    uint64_t total_out = 0;  // text bytes produced so far over all members (history available = min(total_out, 32768))
    int n_threads = 1;
    bool stop_after_member = false;  // kmb_inflate_member: one member only
    const char *error = nullptr;
};

inline uint64_t load64(const uint8_t *p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
}

// make at least `need` (<= 56) bits available; near the end of the input fall back to byte-wise refills
inline bool refill(Stream &s, int need) {
    if (s.in_end - s.in >= 8) {
        s.bitbuf |= load64(s.in) << s.bitcnt;
        s.in += (63 - s.bitcnt) >> 3;
        s.bitcnt |= 56;
        return true;
    }
    while (s.bitcnt < need) {
        if (s.in >= s.in_end) return false;
        s.bitbuf |= (uint64_t)*s.in++ << s.bitcnt;
        s.bitcnt += 8;
    }
    return true;
}
inline uint32_t take(Stream &s, int n) {
    const uint32_t v = (uint32_t)(s.bitbuf & ((1ull << n) - 1));
    s.bitbuf >>= n;
    s.bitcnt -= n;
    return v;
}
// drop the bits up to the next byte boundary and hand whole unread bytes back to the input pointer
inline void align_to_byte(Stream &s) {
    const int drop = s.bitcnt & 7;
    s.bitbuf >>= drop;
    s.bitcnt -= drop;
    s.in -= s.bitcnt >> 3;
    s.bitbuf = 0;
    s.bitcnt = 0;
}

bool fail(Stream &s, const char *why) {
    s.error = why;
    s.phase = Stream::FAILED;
    return false;
}

bool read_member_header(Stream &s) {
    align_to_byte(s);
    const uint8_t *p = s.in, *e = s.in_end;
    if (e - p < 18) return fail(s, "truncated gzip header");
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || (p[3] & 0xE0)) return fail(s, "not a gzip member");
    const int flags = p[3];
    p += 10;
    if (flags & 4) {  // FEXTRA
        if (e - p < 2) return fail(s, "truncated gzip header");
        const size_t xlen = (size_t)p[0] | ((size_t)p[1] << 8);
        p += 2;
        if ((size_t)(e - p) < xlen) return fail(s, "truncated gzip header");
        p += xlen;
    }
    for (int f = 8; f <= 16; f <<= 1) {  // FNAME, FCOMMENT: zero-terminated
        if (!(flags & f)) continue;
        const uint8_t *z = (const uint8_t *)memchr(p, 0, (size_t)(e - p));
        if (!z) return fail(s, "truncated gzip header");
        p = z + 1;
    }
    if (flags & 2) {  // FHCRC
        if (e - p < 2) return fail(s, "truncated gzip header");
        p += 2;
    }
    s.in = p;
    s.crc = (uint32_t)crc32(0L, Z_NULL, 0);
    s.member_out = 0;
    s.phase = Stream::BLOCK_HEADER;
    return true;
}

bool read_block_header(Stream &s) {
    if (!refill(s, 3)) return fail(s, "truncated deflate stream");
    s.last_block = take(s, 1) != 0;
    const uint32_t type = take(s, 2);
    if (type == 0) {
        align_to_byte(s);
        if (s.in_end - s.in < 4) return fail(s, "truncated stored block");
        const uint32_t len = (uint32_t)s.in[0] | ((uint32_t)s.in[1] << 8), nlen = (uint32_t)s.in[2] | ((uint32_t)s.in[3] << 8);
        if ((len ^ 0xFFFFu) != nlen) return fail(s, "corrupt stored block");
        s.in += 4;
        s.stored_left = len;
        s.phase = Stream::STORED;
        return true;
    }
    if (type == 3) return fail(s, "invalid deflate block type");
    uint8_t lens[288 + 32];
    int n_lit, n_dist;
    if (type == 1) {  // fixed codes (RFC 1951 3.2.6)
        for (int i = 0; i < 144; i++) lens[i] = 8;
        for (int i = 144; i < 256; i++) lens[i] = 9;
        for (int i = 256; i < 280; i++) lens[i] = 7;
        for (int i = 280; i < 288; i++) lens[i] = 8;
        for (int i = 0; i < 32; i++) lens[288 + i] = 5;
        n_lit = 288;
        n_dist = 32;
    } else {
        if (!refill(s, 14)) return fail(s, "truncated deflate stream");
        n_lit = (int)take(s, 5) + 257;
        n_dist = (int)take(s, 5) + 1;
        const int n_pre = (int)take(s, 4) + 4;
        if (n_lit > 286 || n_dist > 30) return fail(s, "corrupt dynamic block header");
        static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t pre_lens[19] = {0};
        for (int i = 0; i < n_pre; i++) {
            if (!refill(s, 3)) return fail(s, "truncated deflate stream");
            pre_lens[ORDER[i]] = (uint8_t)take(s, 3);
        }
        uint32_t pre_table[1 << PRE_BITS];
        if (!build_table(pre_lens, 19, 2, PRE_BITS, pre_table, 1 << PRE_BITS)) return fail(s, "corrupt code-length code");
        int i = 0;
        while (i < n_lit + n_dist) {
            if (!refill(s, 14)) return fail(s, "truncated deflate stream");
            const uint32_t e = pre_table[s.bitbuf & ((1u << PRE_BITS) - 1)];
            if (e_kind(e) != K_LITERAL) return fail(s, "corrupt code lengths");
            take(s, (int)e_len(e));
            const uint32_t sym = e_value(e);
            if (sym < 16) {
                lens[i++] = (uint8_t)sym;
                continue;
            }
            int rep;
            uint8_t v = 0;
            if (sym == 16) {
                if (i == 0) return fail(s, "corrupt code lengths");
                v = lens[i - 1];
                rep = 3 + (int)take(s, 2);
            } else if (sym == 17) {
                rep = 3 + (int)take(s, 3);
            } else {
                rep = 11 + (int)take(s, 7);
            }
            if (i + rep > n_lit + n_dist) return fail(s, "corrupt code lengths");
            while (rep--) lens[i++] = v;
        }
        if (lens[256] == 0) return fail(s, "dynamic block without an end-of-block code");
        // the distance lengths follow the literal/length ones directly: move them to their own array position
        memmove(lens + 288, lens + n_lit, (size_t)n_dist);
    }
    uint8_t lit_lens[288] = {0}, dist_lens[32] = {0};
    memcpy(lit_lens, lens, (size_t)n_lit);
    memcpy(dist_lens, lens + 288, (size_t)n_dist);
    if (!build_table(lit_lens, type == 1 ? 288 : n_lit, 0, LIT_BITS, s.lit_table, LIT_TABLE_MAX)) return fail(s, "corrupt literal/length code");
    if (!build_table(dist_lens, type == 1 ? 32 : n_dist, 1, DIST_BITS, s.dist_table, DIST_TABLE_MAX)) return fail(s, "corrupt distance code");
    s.phase = Stream::CODED;
    return true;
}

// Decode symbols of the current coded block into [out, out_end); `hist` = first byte the stream may refer back to.
// Stops at the end of the block, or when fewer than 258 + 16 bytes of output are left.  Returns the new out.
//
// Fast loop: while at least 16 input bytes and 274 output bytes remain, one refill guarantees 56 bits -- enough for
// a literal/length symbol with its extra bits and a distance symbol with its extra bits (15 + 5 + 15 + 13 = 48), or
// for three literals -- so nothing inside checks for exhaustion.
inline void copy_match(uint8_t *dst, const uint8_t *src, uint8_t *end, uint32_t dist) {
    if (dist >= 16) {  // whole 16-byte words; writes up to 15 bytes past `end` (slack guaranteed by the caller)
        do {
            memcpy(dst, src, 16);
            dst += 16;
            src += 16;
        } while (dst < end);
    } else if (dist >= 8) {
        do {
            memcpy(dst, src, 8);
            dst += 8;
            src += 8;
        } while (dst < end);
    } else if (dist == 1) {
        memset(dst, *src, (size_t)(end - dst));
    } else {
        do *dst++ = *src++;
        while (dst < end);
    }
}

uint8_t *decode_block_careful(Stream &s, uint8_t *out, uint8_t *out_end, const uint8_t *hist, bool *block_done);

uint8_t *decode_block(Stream &s, uint8_t *out, uint8_t *out_end, const uint8_t *hist, bool *block_done) {
    *block_done = false;
    const uint32_t lit_mask = (1u << LIT_BITS) - 1, dist_mask = (1u << DIST_BITS) - 1;
    const uint32_t *lit_table = s.lit_table, *dist_table = s.dist_table;
    const uint8_t *in = s.in, *in_end = s.in_end;
    uint64_t bitbuf = s.bitbuf;
    int bitcnt = s.bitcnt;
#define KMB_REFILL()                               \
    do {                                           \
        bitbuf |= load64(in) << bitcnt;            \
        in += (63 - bitcnt) >> 3;                  \
        bitcnt |= 56;                              \
    } while (0)
#define KMB_LIT_LOOKUP(e)                                                                                           \
    do {                                                                                                            \
        e = lit_table[bitbuf & lit_mask];                                                                           \
        if (e_kind(e) == K_SUBTABLE)                                                                                \
            e = lit_table[e_value(e) + ((bitbuf >> LIT_BITS) & ((1u << e_extra(e)) - 1))] + (uint32_t)LIT_BITS;     \
    } while (0)
    while (in_end - in >= 16 && out_end - out >= 258 + 16) {
        KMB_REFILL();
        uint32_t e;
        KMB_LIT_LOOKUP(e);
        if (e_kind(e) == K_LITERAL) {
            bitbuf >>= e_len(e);
            bitcnt -= (int)e_len(e);
            *out++ = (uint8_t)e_value(e);
            KMB_LIT_LOOKUP(e);
            if (e_kind(e) == K_LITERAL) {
                bitbuf >>= e_len(e);
                bitcnt -= (int)e_len(e);
                *out++ = (uint8_t)e_value(e);
                KMB_LIT_LOOKUP(e);
                if (e_kind(e) == K_LITERAL) {
                    bitbuf >>= e_len(e);
                    bitcnt -= (int)e_len(e);
                    *out++ = (uint8_t)e_value(e);
                    continue;
                }
            }
            KMB_REFILL();  // up to 30 bits went into literals: top up before a length/distance pair (e stays valid)
        }
        bitbuf >>= e_len(e);
        bitcnt -= (int)e_len(e);
        const uint32_t kind = e_kind(e);
        if (kind == K_END) {
            *block_done = true;
            break;
        }
        if (kind != K_BASE) {
            s.in = in, s.bitbuf = bitbuf, s.bitcnt = bitcnt;
            fail(s, "invalid literal/length code");
            return out;
        }
        const uint32_t lx = e_extra(e);
        const uint32_t len = e_value(e) + (uint32_t)(bitbuf & ((1ull << lx) - 1));
        bitbuf >>= lx;
        bitcnt -= (int)lx;
        uint32_t d = dist_table[bitbuf & dist_mask];
        if (e_kind(d) == K_SUBTABLE) d = dist_table[e_value(d) + ((bitbuf >> DIST_BITS) & ((1u << e_extra(d)) - 1))] + (uint32_t)DIST_BITS;
        if (e_kind(d) != K_BASE) {
            s.in = in, s.bitbuf = bitbuf, s.bitcnt = bitcnt;
            fail(s, "invalid distance code");
            return out;
        }
        bitbuf >>= e_len(d);
        bitcnt -= (int)e_len(d);
        const uint32_t dx = e_extra(d);
        const uint32_t dist = e_value(d) + (uint32_t)(bitbuf & ((1ull << dx) - 1));
        bitbuf >>= dx;
        bitcnt -= (int)dx;
        if ((size_t)(out - hist) < dist) {
            s.in = in, s.bitbuf = bitbuf, s.bitcnt = bitcnt;
            fail(s, "distance reaches before the start of the stream");
            return out;
        }
        copy_match(out, out - dist, out + len, dist);
        out += len;
    }
#undef KMB_REFILL
#undef KMB_LIT_LOOKUP
    s.in = in, s.bitbuf = bitbuf, s.bitcnt = bitcnt;
    if (*block_done) return out;
    return decode_block_careful(s, out, out_end, hist, block_done);
}

// The same loop with every bit-availability check: used near the end of the input (and of the output buffer).
uint8_t *decode_block_careful(Stream &s, uint8_t *out, uint8_t *out_end, const uint8_t *hist, bool *block_done) {
    *block_done = false;
    const uint32_t lit_mask = (1u << LIT_BITS) - 1, dist_mask = (1u << DIST_BITS) - 1;
    while (out_end - out >= 258 + 16) {
        if (!refill(s, 48)) {
            // fewer than 48 bits left in the whole input: still decodable if the remaining symbols are short
            if (s.bitcnt == 0) {
                fail(s, "truncated deflate stream");
                return out;
            }
        }
        uint32_t e = s.lit_table[s.bitbuf & lit_mask];
        if (e_kind(e) == K_SUBTABLE) e = s.lit_table[e_value(e) + ((s.bitbuf >> LIT_BITS) & ((1u << e_extra(e)) - 1))] + (uint32_t)LIT_BITS;
        // (a sub-table entry's length is the bits beyond the primary ones: + LIT_BITS gives the total; lengths < 256 fit)
        if ((int)e_len(e) > s.bitcnt) {
            fail(s, "truncated deflate stream");
            return out;
        }
        s.bitbuf >>= e_len(e);
        s.bitcnt -= (int)e_len(e);
        const uint32_t kind = e_kind(e);
        if (kind == K_LITERAL) {
            *out++ = (uint8_t)e_value(e);
            // a second literal from the same refill (up to 2 x 15 bits are certainly there)
            uint32_t e2 = s.lit_table[s.bitbuf & lit_mask];
            if (e_kind(e2) == K_LITERAL && (int)e_len(e2) <= s.bitcnt) {
                s.bitbuf >>= e_len(e2);
                s.bitcnt -= (int)e_len(e2);
                *out++ = (uint8_t)e_value(e2);
            }
            continue;
        }
        if (kind == K_END) {
            *block_done = true;
            return out;
        }
        if (kind != K_BASE) {
            fail(s, "invalid literal/length code");
            return out;
        }
        uint32_t len = e_value(e);
        const int lx = (int)e_extra(e);
        if (lx + 15 + 13 > s.bitcnt && !refill(s, lx + 15 + 13) && s.bitcnt < lx) {
            fail(s, "truncated deflate stream");
            return out;
        }
        len += take(s, lx);
        uint32_t d = s.dist_table[s.bitbuf & dist_mask];
        if (e_kind(d) == K_SUBTABLE) d = s.dist_table[e_value(d) + ((s.bitbuf >> DIST_BITS) & ((1u << e_extra(d)) - 1))] + (uint32_t)DIST_BITS;
        if (e_kind(d) != K_BASE || (int)(e_len(d) + e_extra(d)) > s.bitcnt) {
            fail(s, e_kind(d) != K_BASE ? "invalid distance code" : "truncated deflate stream");
            return out;
        }
        s.bitbuf >>= e_len(d);
        s.bitcnt -= (int)e_len(d);
        const uint32_t dist = e_value(d) + take(s, (int)e_extra(d));
        if ((size_t)(out - hist) < dist) {
            fail(s, "distance reaches before the start of the stream");
            return out;
        }
        copy_match(out, out - dist, out + len, dist);
        out += len;
    }
    return out;
}

// CRC-32 of a block of output, in parallel pieces combined with crc32_combine
uint32_t crc_update(uint32_t crc, const uint8_t *p, size_t n, int n_threads) {
    if (n < (4u << 20) || n_threads <= 1) {
        while (n) {
            const uInt step = (uInt)std::min<size_t>(n, 1u << 30);
            crc = (uint32_t)crc32(crc, p, step);
            p += step;
            n -= step;
        }
        return crc;
    }
    const int parts = std::min<int>(n_threads, (int)(n >> 20));
    const size_t per = (n + (size_t)parts - 1) / (size_t)parts;
    std::vector<uint32_t> c((size_t)parts);
    std::function<void(int)> fn = [&](int i) {
        const size_t lo = (size_t)i * per, hi = std::min(lo + per, n);
        uint32_t v = (uint32_t)crc32(0L, Z_NULL, 0);
        for (size_t q = lo; q < hi;) {
            const uInt step = (uInt)std::min<size_t>(hi - q, 1u << 30);
            v = (uint32_t)crc32(v, p + q, step);
            q += step;
        }
        c[(size_t)i] = v;
    };
    struct T {
        static void tramp(void *ctx, int part) { (*static_cast<std::function<void(int)> *>(ctx))(part); }
    };
    kmb_host_parallel(n_threads, parts, T::tramp, &fn);
    for (int i = 0; i < parts; i++) {
        const size_t lo = (size_t)i * per, hi = std::min(lo + per, n);
        crc = (uint32_t)crc32_combine(crc, c[(size_t)i], (z_off_t)(hi - lo));
    }
    return crc;
}

}  // namespace

struct kmb_gzstream {
    Stream s;
};

extern "C" int kmb_gzstream_open(const uint8_t *gz, uint64_t n_gz, int n_threads, kmb_gzstream **out) {
    if (!out || (!gz && n_gz)) return KMB_ERR_BAD_ARG;
    kmb_gzstream *g = new (std::nothrow) kmb_gzstream;
    if (!g) return KMB_ERR_NOMEM;
    g->s.in_begin = g->s.in = gz;
    g->s.in_end = gz + n_gz;
    g->s.n_threads = n_threads > 0 ? n_threads : kmb_host_cpus();
    if (n_gz == 0) g->s.phase = Stream::DONE;
    *out = g;
    return KMB_OK;
}

extern "C" int kmb_gzstream_close(kmb_gzstream *g) {
    delete g;
    return KMB_OK;
}

extern "C" const char *kmb_gzstream_error(const kmb_gzstream *g) { return g && g->s.error ? g->s.error : ""; }

extern "C" int kmb_gzstream_read(kmb_gzstream *g, uint8_t *out, uint64_t out_capacity, uint64_t history_bytes,
                                 uint64_t *produced, int *finished) {
    if (!g || !out || !produced || !finished) return KMB_ERR_BAD_ARG;
    Stream &s = g->s;
    *produced = 0;
    *finished = 0;
    if (s.phase == Stream::FAILED) return KMB_ERR_BAD_ARG;
    if (out_capacity < 65536) return KMB_ERR_BAD_ARG;
    // how far back a match may reach: what the caller placed in front of `out`, but never more than was produced
    const uint64_t have_hist = std::min<uint64_t>(std::min<uint64_t>(history_bytes, 32768), s.total_out);
    if (s.total_out > 0 && have_hist < std::min<uint64_t>(s.total_out, 32768)) return KMB_ERR_BAD_ARG;  // history missing
    const uint8_t *hist = out - have_hist;
    uint8_t *o = out, *out_end = out + out_capacity;
    uint8_t *verified = out;  // output before this pointer is already in the member's CRC
    auto account = [&](uint8_t *upto) {
        if (upto > verified) {
            s.crc = crc_update(s.crc, verified, (size_t)(upto - verified), s.n_threads);
            s.member_out += (uint64_t)(upto - verified);
            verified = upto;
        }
    };
    for (;;) {
        if (s.phase == Stream::DONE) {
            *finished = 1;
            break;
        }
        if (s.phase == Stream::MEMBER_HEADER) {
            if (!read_member_header(s)) break;
            // a new member starts a new history: matches must not reach into the previous member
            hist = o;
            continue;
        }
        if (s.phase == Stream::BLOCK_HEADER) {
            if (!read_block_header(s)) break;
            continue;
        }
        if (s.phase == Stream::STORED) {
            const uint64_t room = (uint64_t)(out_end - o);
            const uint64_t n = std::min<uint64_t>(std::min<uint64_t>(s.stored_left, room), (uint64_t)(s.in_end - s.in));
            memcpy(o, s.in, (size_t)n);
            o += n;
            s.in += n;
            s.stored_left -= (uint32_t)n;
            if (s.stored_left == 0) {
                s.phase = s.last_block ? Stream::MEMBER_TRAILER : Stream::BLOCK_HEADER;
                continue;
            }
            if (s.in >= s.in_end) {
                fail(s, "truncated stored block");
                break;
            }
            break;  // output full
        }
        if (s.phase == Stream::CODED) {
            bool done = false;
            o = decode_block(s, o, out_end, hist, &done);
            if (s.phase == Stream::FAILED) break;
            if (done) {
                s.phase = s.last_block ? Stream::MEMBER_TRAILER : Stream::BLOCK_HEADER;
                continue;
            }
            break;  // output (nearly) full: the caller comes back with a fresh buffer
        }
        if (s.phase == Stream::MEMBER_TRAILER) {
            account(o);
            align_to_byte(s);
            if (s.in_end - s.in < 8) {
                fail(s, "truncated gzip trailer");
                break;
            }
            const uint32_t want_crc = (uint32_t)s.in[0] | ((uint32_t)s.in[1] << 8) | ((uint32_t)s.in[2] << 16) | ((uint32_t)s.in[3] << 24);
            const uint32_t want_len = (uint32_t)s.in[4] | ((uint32_t)s.in[5] << 8) | ((uint32_t)s.in[6] << 16) | ((uint32_t)s.in[7] << 24);
            s.in += 8;
            if (want_crc != s.crc || want_len != (uint32_t)s.member_out) {
                fail(s, "gzip member fails its CRC-32 / length check");
                break;
            }
            if (s.stop_after_member) {
                s.phase = Stream::DONE;
                continue;
            }
            // more members?  (zero padding after the last one is tolerated, like gzip does)
            const uint8_t *p = s.in;
            while (p < s.in_end && *p == 0) p++;
            if (p >= s.in_end) {
                s.phase = Stream::DONE;
            } else {
                s.in = p;
                s.phase = Stream::MEMBER_HEADER;
            }
            continue;
        }
    }
    if (s.phase != Stream::FAILED) account(o);
    *produced = (uint64_t)(o - out);
    s.total_out += *produced;
    return s.phase == Stream::FAILED ? KMB_ERR_BAD_ARG : KMB_OK;
}

// One gzip member from gz[0, n_gz) into a malloc'd buffer (*buf, caller frees), for the member-parallel reader
// (kmb_gunzip.cpp).  Returns 1 = ok (*consumed compressed bytes used, *n_out text bytes), 2 = not a member /
// corrupt / truncated, 3 = the member inflates to more than `limit` bytes.
int kmb_inflate_member(const uint8_t *gz, uint64_t n_gz, uint64_t limit, uint8_t **buf, size_t *n_out, uint64_t *consumed) {
    *buf = nullptr;
    *n_out = 0;
    *consumed = 0;
    kmb_gzstream g;
    Stream &s = g.s;
    s.in_begin = s.in = gz;
    s.in_end = gz + n_gz;
    s.n_threads = 1;  // members are inflated in parallel with each other: the CRC stays on this thread
    s.stop_after_member = true;
    const size_t slack = 65536 + 512;  // kmb_gzstream_read wants 64 KB of room to make progress
    size_t cap = (size_t)std::min<uint64_t>(limit, std::max<uint64_t>(1u << 17, std::min<uint64_t>(n_gz * 5, 8u << 20))) + slack;
    uint8_t *p = (uint8_t *)malloc(cap);
    if (!p) return 2;
    size_t produced = 0;
    for (;;) {
        uint64_t got = 0;
        int finished = 0;
        const int rc = kmb_gzstream_read(&g, p + produced, cap - produced, produced, &got, &finished);
        produced += (size_t)got;
        if (rc != KMB_OK) {
            free(p);
            return 2;
        }
        if (produced > limit) {
            free(p);
            return 3;
        }
        if (finished) break;
        if (cap - produced < slack) {
            const size_t want = std::min<size_t>((size_t)limit + slack + 1, cap * 2);
            if (want <= cap) {  // already at the limit and still not finished
                free(p);
                return 3;
            }
            uint8_t *q = (uint8_t *)realloc(p, want);
            if (!q) {
                free(p);
                return 2;
            }
            p = q;
            cap = want;
        }
    }
    *buf = p;
    *n_out = produced;
    *consumed = (uint64_t)(s.in - s.in_begin);
    return 1;
}
