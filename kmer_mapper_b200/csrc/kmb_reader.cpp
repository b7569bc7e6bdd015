// kmb_reader.cpp -- host side of the chunked FASTA/FASTQ reader (include/kmer_mapper_b200.h,
// "reader" section): text chunk in, flat bases + read offsets out, written straight into the
// caller's (pinned) buffers so the chunk can be handed to kmb_mapper_map_reads without a copy.
//
// Replaces what the reference gets from bionumpy at command_line_interface.py:102-111
// (bnp.open(path).read_chunks(min_chunk_size) -> chunk.sequence): cut at the last complete record,
// strip headers / '+' lines / qualities / newlines.  This is file plumbing, not arithmetic of the
// mapped path; it is native and multi-threaded because the kernels consume ~100 G bases/s and a
// numpy parser delivers 0.05.
//
// Parallel scheme: the text is cut into T pieces at record boundaries found by resynchronisation
// (FASTQ: a line starting with '@' whose next-but-one line starts with '+' -- a quality line that
// starts with '@' is followed by a header and then by bases, never by '+'; FASTA: a line starting
// with '>'); pass 1 counts reads and bases per piece, a prefix sum gives every piece its output
// position, pass 2 copies.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <vector>

#include "../../include/kmer_mapper_b200.h"
#include "kmb_host.h"

namespace {

struct Piece {
    uint64_t begin = 0, end = 0;      // text range [begin, end): whole records only
    uint64_t n_reads = 0, n_bases = 0;
    uint64_t read0 = 0, base0 = 0;    // output positions (prefix sums)
    int error = 0;
};

inline const uint8_t *find_nl(const uint8_t *p, const uint8_t *e) {
    return (const uint8_t *)memchr(p, '\n', (size_t)(e - p));
}
inline uint64_t line_len(const uint8_t *s, const uint8_t *nl) {  // without a trailing '\r'
    uint64_t n = (uint64_t)(nl - s);
    return (n && s[n - 1] == '\r') ? n - 1 : n;
}

// ---- FASTQ --------------------------------------------------------------------------------------
// Walk records in [p, e).  COPY = false: count; COPY = true: write bases/offsets.  Returns the end
// of the last complete record; `final` lets the last record end without a newline.
template <bool COPY>
const uint8_t *fastq_walk(const uint8_t *text, const uint8_t *p, const uint8_t *e, bool final, Piece &pc,
                          uint8_t *bases, int64_t *offsets) {
    uint64_t nr = 0, nb = 0;
    uint64_t r = pc.read0, b = pc.base0;
    while (p < e) {
        const uint8_t *l0 = find_nl(p, e);
        if (!l0) break;
        const uint8_t *s1 = l0 + 1;
        const uint8_t *l1 = find_nl(s1, e);
        if (!l1) break;
        const uint8_t *s2 = l1 + 1;
        const uint8_t *l2 = find_nl(s2, e);
        if (!l2) break;
        const uint8_t *s3 = l2 + 1;
        const uint8_t *l3 = find_nl(s3, e);
        const uint8_t *next;
        if (l3) {
            next = l3 + 1;
        } else if (final) {
            next = e;  // last quality line without a newline (possibly empty)
        } else {
            break;
        }
        if (*p != '@' || *s2 != '+') {
            pc.error = 1;
            return p;
        }
        uint64_t len = line_len(s1, l1);
        if (COPY) {
            memcpy(bases + b, s1, (size_t)len);
            offsets[r + 1] = (int64_t)(b + len);
            r++;
            b += len;
        }
        nr++;
        nb += len;
        p = next;
    }
    (void)text;
    if (!COPY) {
        pc.n_reads = nr;
        pc.n_bases = nb;
    }
    return p;
}

// first record start at or after p (p itself if it is one and at_line_start)
const uint8_t *fastq_resync(const uint8_t *p, const uint8_t *e) {
    const uint8_t *nl = find_nl(p, e);
    while (nl) {
        const uint8_t *s = nl + 1;
        if (s >= e) return e;
        const uint8_t *l0 = find_nl(s, e);
        if (!l0) return e;
        const uint8_t *l1 = find_nl(l0 + 1, e);
        if (!l1) return e;
        if (*s == '@' && l1 + 1 < e && l1[1] == '+') return s;
        nl = l0;
    }
    return e;
}

// ---- FASTA --------------------------------------------------------------------------------------
template <bool COPY>
const uint8_t *fasta_walk(const uint8_t *p, const uint8_t *e, bool final, Piece &pc, uint8_t *bases, int64_t *offsets) {
    // a record is complete once the next '>' line (or, when final, the end of the data) is seen
    uint64_t nr = 0, nb = 0;
    uint64_t r = pc.read0, b = pc.base0;
    const uint8_t *rec = p;  // start of the current record's header
    if (p < e && *p != '>') {
        pc.error = 1;
        return p;
    }
    const uint8_t *last_complete = p;
    while (rec < e) {
        const uint8_t *hl = find_nl(rec, e);
        if (!hl) {
            if (final) {  // header without sequence at the very end: an empty read
                if (COPY) {
                    offsets[r + 1] = (int64_t)b;
                    r++;
                }
                nr++;
                last_complete = e;
            }
            break;
        }
        const uint8_t *s = hl + 1;
        uint64_t rec_bases = 0;
        const uint8_t *next_rec = nullptr;
        const uint8_t *q = s;
        bool closed = false;
        while (q < e) {
            if (*q == '>') {
                next_rec = q;
                closed = true;
                break;
            }
            const uint8_t *nl = find_nl(q, e);
            const uint8_t *le = nl ? nl : e;
            if (!nl && !final) break;  // partial sequence line: record not complete in this chunk
            uint64_t len = line_len(q, le);
            if (!nl && len == (uint64_t)(le - q)) { /* last line without newline */ }
            if (COPY) memcpy(bases + b + rec_bases, q, (size_t)len);
            rec_bases += len;
            q = nl ? nl + 1 : e;
        }
        if (!closed && !(final && q >= e)) break;  // ran out of data mid-record
        if (COPY) {
            offsets[r + 1] = (int64_t)(b + rec_bases);
            r++;
            b += rec_bases;
        }
        nr++;
        nb += rec_bases;
        rec = closed ? next_rec : e;
        last_complete = rec;
    }
    if (!COPY) {
        pc.n_reads = nr;
        pc.n_bases = nb;
    }
    return last_complete;
}

const uint8_t *fasta_resync(const uint8_t *p, const uint8_t *e) {
    const uint8_t *nl = find_nl(p, e);
    while (nl) {
        if (nl + 1 >= e) return e;
        if (nl[1] == '>') return nl + 1;
        nl = find_nl(nl + 1, e);
    }
    return e;
}

}  // namespace

extern "C" int kmb_find_record_start(const uint8_t *text, uint64_t n_text, int format, uint64_t *offset) {
    if (!offset || (!text && n_text) || (format != 0 && format != 1)) return KMB_ERR_BAD_ARG;
    const uint8_t *e = text + n_text;
    const uint8_t *p = format == 1 ? fastq_resync(text, e) : fasta_resync(text, e);
    *offset = (uint64_t)(p - text);
    return KMB_OK;
}

// one piece per worker of the library's thread pool (kmb_hostpack.cpp)
static void run_piece_trampoline(void *ctx, int part) { (*static_cast<std::function<void(int)> *>(ctx))(part); }
static void run_pieces(int n, std::function<void(int)> fn) { kmb_host_parallel(n, n, run_piece_trampoline, &fn); }

extern "C" int kmb_parse_reads(const uint8_t *text, uint64_t n_text, int format, int final_chunk, int n_threads,
                               uint8_t *bases, uint64_t bases_capacity, int64_t *offsets, uint64_t offsets_capacity,
                               uint64_t *n_reads, uint64_t *n_bases, uint64_t *consumed) {
    if (!n_reads || !n_bases || !consumed) return KMB_ERR_BAD_ARG;
    *n_reads = *n_bases = *consumed = 0;
    if (n_text == 0) return KMB_OK;
    if (!text || (format != 0 && format != 1)) return KMB_ERR_BAD_ARG;
    const bool fastq = format == 1, final = final_chunk != 0;
    const uint8_t *e = text + n_text;
    int T = std::max(1, std::min(n_threads, 64));
    if (n_text < (1u << 20)) T = 1;
    // ---- cut into pieces at record boundaries
    std::vector<const uint8_t *> cut(T + 1);
    cut[0] = text;
    cut[T] = e;
    for (int t = 1; t < T; t++) {
        const uint8_t *p = text + n_text / T * t;
        cut[t] = fastq ? fastq_resync(p, e) : fasta_resync(p, e);
    }
    for (int t = 1; t <= T; t++) cut[t] = std::max(cut[t], cut[t - 1]);
    std::vector<Piece> pc(T);
    std::vector<const uint8_t *> piece_end(T);
    // ---- pass 1: count.  Only the last non-empty piece may end in an incomplete record.
    auto count_piece = [&](int t) {
        const bool last = cut[t + 1] == e;
        const bool fin = last ? final : true;  // inner pieces end exactly at a record start
        piece_end[t] = fastq ? fastq_walk<false>(text, cut[t], cut[t + 1], fin, pc[t], nullptr, nullptr)
                             : fasta_walk<false>(cut[t], cut[t + 1], fin, pc[t], nullptr, nullptr);
    };
    run_pieces(T, count_piece);
    uint64_t r = 0, b = 0;
    const uint8_t *done = text;
    for (int t = 0; t < T; t++) {
        if (pc[t].error) return KMB_ERR_BAD_ARG;
        pc[t].read0 = r;
        pc[t].base0 = b;
        r += pc[t].n_reads;
        b += pc[t].n_bases;
        if (cut[t + 1] > cut[t]) done = piece_end[t];
        // an inner piece that stopped early (no trailing newline found etc.) ends the usable text
        if (piece_end[t] != cut[t + 1]) {
            for (int u = t + 1; u < T; u++) pc[u].n_reads = pc[u].n_bases = 0, cut[u] = cut[u + 1] = e;
            T = t + 1;
            break;
        }
    }
    *n_reads = r;
    *n_bases = b;
    *consumed = (uint64_t)(done - text);
    if (!bases || !offsets) return KMB_OK;  // count only
    if (b > bases_capacity || r + 1 > offsets_capacity) return KMB_ERR_NOMEM;
    offsets[0] = 0;
    // ---- pass 2: copy
    auto copy_piece = [&](int t) {
        // [cut[t], piece_end[t]) holds exactly the complete records counted in pass 1
        Piece tmp = pc[t];
        if (fastq) fastq_walk<true>(text, cut[t], piece_end[t], true, tmp, bases, offsets);
        else fasta_walk<true>(cut[t], piece_end[t], true, tmp, bases, offsets);
    };
    run_pieces(T, [&](int t) {
        if (pc[t].n_reads) copy_piece(t);
    });
    return KMB_OK;
}
