// kmb_gunzip.cpp -- member-parallel gzip inflate for the chunk reader (include/kmer_mapper_b200.h, "reader").
//
// A .gz file is a sequence of independent members (bgzip/BGZF output, `cat a.gz b.gz`, many sequencer pipelines);
// the reference reads them through Python's single-threaded gzip inside bionumpy (command_line_interface.py:102-103).
// Where a member starts is not recorded anywhere, so starts are found speculatively: every occurrence of the member
// magic (1f 8b 08, reserved flag bits clear) is a candidate that a worker inflates into a private buffer (with the
// library's own DEFLATE decoder, kmb_inflate.cpp: ~2x zlib, CRC-32 checked); the chain
// "a member starts where the previous one ended" then picks the real ones in order, and a second parallel pass
// copies them to the caller's buffer.  False candidates (magic bytes inside compressed data) fail within a few
// bytes.  A member whose output does not fit the caller's limit ends the call: the caller streams it sequentially.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <vector>

#include "../../include/kmer_mapper_b200.h"
#include "kmb_host.h"

namespace {

struct Bytes {  // growable byte buffer without the zero fill of std::vector::resize
    uint8_t *p = nullptr;
    size_t n = 0, cap = 0;
    Bytes() = default;
    Bytes(const Bytes &) = delete;
    Bytes &operator=(const Bytes &) = delete;
    Bytes(Bytes &&o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr, o.n = o.cap = 0; }
    ~Bytes() { free(p); }
    bool reserve(size_t want) {
        if (want <= cap) return true;
        uint8_t *q = (uint8_t *)realloc(p, want);
        if (!q) return false;
        p = q;
        cap = want;
        return true;
    }
    void clear() {
        free(p);
        p = nullptr;
        n = cap = 0;
    }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    const uint8_t *data() const { return p; }
};

struct Member {
    uint64_t start = 0, end = 0;  // [start, end) in the compressed input
    Bytes out;
    int state = 0;  // 0 not tried, 1 ok, 2 not a member / corrupt / truncated, 3 output larger than the limit
};

// inflate one gzip member starting at gz[start]; at most `limit` bytes of output
void inflate_member(const uint8_t *gz, uint64_t n_gz, uint64_t limit, Member &m) {
    uint8_t *buf = nullptr;
    size_t n = 0;
    uint64_t used = 0;
    m.state = kmb_inflate_member(gz + m.start, n_gz - m.start, limit, &buf, &n, &used);
    if (m.state == 1) {
        m.out.p = buf;
        m.out.n = m.out.cap = n;
        m.end = m.start + used;
    }
}

void run_parallel(int n_threads, int n, std::function<void(int)> fn) {
    struct T {
        static void tramp(void *ctx, int part) { (*static_cast<std::function<void(int)> *>(ctx))(part); }
    };
    kmb_host_parallel(n_threads, n, T::tramp, &fn);
}

}  // namespace

extern "C" int kmb_gunzip_members(const uint8_t *gz, uint64_t n_gz, int n_threads, uint8_t *out, uint64_t out_capacity,
                                  uint64_t max_member_bytes, uint64_t *consumed, uint64_t *produced, int *stopped_at_big_member) {
    if (!consumed || !produced || (!gz && n_gz) || (!out && out_capacity)) return KMB_ERR_BAD_ARG;
    *consumed = *produced = 0;
    if (stopped_at_big_member) *stopped_at_big_member = 0;
    if (n_threads <= 0) n_threads = kmb_host_cpus();
    const uint64_t limit = std::min<uint64_t>(max_member_bytes ? max_member_bytes : out_capacity, out_capacity);
    uint64_t expected = 0;  // where the next member must start
    uint64_t scan = 0;      // candidates below this offset have been collected
    std::vector<Member> batch;
    while (expected < n_gz) {
        // ---- candidates from `expected` on: the expected start itself, then magic matches, ~2 per thread and
        //      at least a few MB of input so that small (BGZF) members still fill the threads
        batch.clear();
        {
            Member m;
            m.start = expected;
            batch.push_back(std::move(m));
        }
        scan = std::max(scan, expected + 1);
        const uint64_t span_end = std::min<uint64_t>(n_gz, expected + std::max<uint64_t>((uint64_t)n_threads << 22, 1u << 24));
        // members that start beyond this are unlikely to fit what is left of the output -- inflating them now would be
        // work thrown away: expansion as measured in this call so far, 6x (FASTQ at gzip -1 is ~5.5x) before that
        const double ratio = *consumed ? std::max(1.0, 1.1 * (double)*produced / (double)*consumed) : 6.0;
        const uint64_t fit_end = expected + std::max<uint64_t>((uint64_t)((double)(out_capacity - *produced) / ratio), 1u << 16);
        uint64_t p = scan;
        while (p + 4 <= n_gz && p < fit_end && (batch.size() < (size_t)(2 * n_threads) || p < span_end) && batch.size() < 65536) {
            const uint8_t *q = (const uint8_t *)memchr(gz + p, 0x1f, (size_t)(n_gz - 3 - p));
            if (!q) {
                p = n_gz;
                break;
            }
            p = (uint64_t)(q - gz);
            if (q[1] == 0x8b && q[2] == 8 && (q[3] & 0xE0) == 0) {
                Member m;
                m.start = p;
                batch.push_back(std::move(m));
            }
            p++;
        }
        scan = std::min(p, fit_end);
        run_parallel(n_threads, (int)batch.size(), [&](int i) { inflate_member(gz, n_gz, limit, batch[(size_t)i]); });
        // ---- walk the chain through this batch
        std::vector<size_t> chain;
        uint64_t pos = expected, total = 0;
        size_t i = 0;
        bool stop = false;
        while (i < batch.size()) {
            if (batch[i].start < pos) {  // a candidate inside a member that was just confirmed
                i++;
                continue;
            }
            if (batch[i].start > pos) break;  // the member at `pos` was not among the candidates: next batch starts there
            Member &m = batch[i];
            if (m.state == 3) {
                if (stopped_at_big_member) *stopped_at_big_member = 1;
                stop = true;
                break;
            }
            if (m.state != 1) {
                // nothing valid at the position the chain demands: corrupt file (or trailing garbage)
                if (chain.empty() && *produced == 0 && expected == 0) return KMB_ERR_BAD_ARG;
                stop = true;
                if (stopped_at_big_member) *stopped_at_big_member = 2;  // cannot continue here
                break;
            }
            if (*produced + total + m.out.size() > out_capacity) {
                stop = true;
                break;
            }
            chain.push_back(i);
            total += m.out.size();
            pos = m.end;
            i++;
        }
        // ---- copy the confirmed members to their places
        if (!chain.empty()) {
            std::vector<uint64_t> at(chain.size());
            uint64_t o = *produced;
            for (size_t c = 0; c < chain.size(); c++) {
                at[c] = o;
                o += batch[chain[c]].out.size();
            }
            run_parallel(n_threads, (int)chain.size(), [&](int c) {
                const Member &m = batch[chain[(size_t)c]];
                if (!m.out.empty()) memcpy(out + at[(size_t)c], m.out.data(), m.out.size());
            });
            *produced = o;
            expected = pos;
            *consumed = expected;
        }
        if (stop) break;
        if (chain.empty()) break;  // no progress possible (output full before the first member of this batch)
    }
    return KMB_OK;
}
