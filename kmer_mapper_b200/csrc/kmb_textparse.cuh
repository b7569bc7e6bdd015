// kmb_textparse.cuh -- FASTA / FASTQ record parsing on the device.
//
// Replaces, for the GPU route, what the reference gets from bionumpy on the host at
// command_line_interface.py:102-103,109-111 (bnp.open(f).read_chunks(...) -> chunk.sequence): strip headers,
// '+' lines, qualities and newlines and hand over the bases of the reads back to back plus their offsets.
// The raw text of a chunk (whole records) crosses PCIe once, and everything else happens in HBM:
//
//   T1 kmb_tp_count_newlines   newlines per 4 KB block of text
//   T2 kmb_tp_scan             exclusive scan of the block counts (one CTA; also used for the line scans)
//   T3 kmb_tp_line_ends        position of every newline, in order -> nl_pos[line]
//   T4 kmb_tp_classify         per line: header / sequence / other, its sequence length (CR stripped), format checks;
//                              per 1024 lines: sequence bytes and headers in them
//   T5 kmb_tp_emit             per line: where its bases go (scan of the lengths) and which read it belongs to (scan of
//                              the headers): offsets[read] at headers, (source, destination, length) at sequence lines
//   T6 kmb_tp_copy             one warp per sequence line: the bases into the flat array
//
// after which the fused mapping kernel runs on (bases, offsets) exactly as for device-resident input.  FASTQ: 4-line
// records (line 4r header '@', 4r+1 bases, 4r+2 '+', 4r+3 qualities), as the reference's reader assumes; FASTA: a
// line starting with '>' begins a record, every other line up to the next header is sequence (multi-line allowed,
// empty lines contribute nothing; a header without sequence is an empty read).  Same rules as the host parser
// (kmb_reader.cpp), against which the tests compare it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define KMB_TP_BLOCK_BYTES 4096
#define KMB_TP_LINES_PER_BLOCK 1024
#define KMB_TP_ERR_MALFORMED 1u   // FASTQ: header without '@' / third line without '+' / incomplete record; FASTA: data before '>'
#define KMB_TP_ERR_TOO_MANY_LINES 2u

struct KmbTextResult {  // written by the device, read back by the host before the mapping kernel is launched
    unsigned long long n_reads;
    unsigned long long n_bases;
    unsigned long long n_lines;
    unsigned int error;
    unsigned int pad;
};

// T1: newlines per block of KMB_TP_BLOCK_BYTES (256 threads x 16 bytes)
__global__ void __launch_bounds__(256) kmb_tp_count_newlines(const uint8_t *__restrict__ text, uint64_t n_text,
                                                              uint32_t *__restrict__ block_count) {
    __shared__ uint32_t s_w[8];
    const uint64_t base = (uint64_t)blockIdx.x * KMB_TP_BLOCK_BYTES + (uint64_t)threadIdx.x * 16;
    uint32_t c = 0;
    if (base + 16 <= n_text && (reinterpret_cast<uintptr_t>(text + base) & 15u) == 0u) {
        const uint4 v = *reinterpret_cast<const uint4 *>(text + base);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t x = w[i] ^ 0x0A0A0A0Au;                        // zero byte <=> newline
            const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
            c += __popc(z);
        }
    } else {
        for (uint64_t p = base; p < n_text && p < base + 16; p++) c += text[p] == '\n';
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; w++) t += s_w[w];
        block_count[blockIdx.x] = t;
    }
}

// T2: in-place exclusive scan of v[0, n) by one CTA of 1024 threads; the total goes to *total.  (n <= a few million.)
__global__ void __launch_bounds__(1024) kmb_tp_scan(uint32_t *v, uint64_t n, unsigned long long *total) {
    __shared__ unsigned long long s_part[1024];
    const int t = threadIdx.x;
    const uint64_t per = (n + 1023) / 1024;
    const uint64_t lo = (uint64_t)t * per, hi = lo + per < n ? lo + per : n;
    unsigned long long sum = 0;
    for (uint64_t i = lo; i < hi; i++) sum += v[i];
    s_part[t] = sum;
    __syncthreads();
    if (t == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; i++) {
            const unsigned long long x = s_part[i];
            s_part[i] = run;
            run += x;
        }
        if (total) *total = run;
    }
    __syncthreads();
    unsigned long long run = s_part[t];
    for (uint64_t i = lo; i < hi; i++) {
        const uint32_t x = v[i];
        v[i] = (uint32_t)run;
        run += x;
    }
}
// 64-bit variant for the base offsets of the lines' blocks
__global__ void __launch_bounds__(1024) kmb_tp_scan64(unsigned long long *v, uint64_t n, unsigned long long *total) {
    __shared__ unsigned long long s_part[1024];
    const int t = threadIdx.x;
    const uint64_t per = (n + 1023) / 1024;
    const uint64_t lo = (uint64_t)t * per, hi = lo + per < n ? lo + per : n;
    unsigned long long sum = 0;
    for (uint64_t i = lo; i < hi; i++) sum += v[i];
    s_part[t] = sum;
    __syncthreads();
    if (t == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; i++) {
            const unsigned long long x = s_part[i];
            s_part[i] = run;
            run += x;
        }
        if (total) *total = run;
    }
    __syncthreads();
    unsigned long long run = s_part[t];
    for (uint64_t i = lo; i < hi; i++) {
        const unsigned long long x = v[i];
        v[i] = run;
        run += x;
    }
}

// T3: nl_pos[block_off[b] + rank within the block] = position of the newline.  One thread per byte, 256 per CTA, 16
// rounds per block: ballot + popcount give the rank inside a warp, a running counter in shared memory the rest.
__global__ void __launch_bounds__(256) kmb_tp_line_ends(const uint8_t *__restrict__ text, uint64_t n_text,
                                                         const uint32_t *__restrict__ block_off, uint32_t *__restrict__ nl_pos,
                                                         uint64_t nl_capacity, KmbTextResult *res) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = block_off[blockIdx.x];
    __syncthreads();
    const uint64_t b0 = (uint64_t)blockIdx.x * KMB_TP_BLOCK_BYTES;
    for (int round = 0; round < KMB_TP_BLOCK_BYTES / 256; round++) {
        const uint64_t p = b0 + (uint64_t)round * 256 + threadIdx.x;
        const bool nl = p < n_text && text[p] == '\n';
        const unsigned m = __ballot_sync(0xFFFFFFFFu, nl);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        uint32_t before = s_run;
        for (int w = 0; w < warp; w++) before += s_warp[w];
        if (nl) {
            const uint64_t at = (uint64_t)before + __popc(m & ((1u << lane) - 1u));
            if (at < nl_capacity) nl_pos[at] = (uint32_t)p;
            else atomicOr(&res->error, KMB_TP_ERR_TOO_MANY_LINES);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < 8; w++) t += s_warp[w];
            s_run += t;
        }
        __syncthreads();
    }
}

// Lines of the text: a last line without newline counts, an empty tail after the last newline does not -- except
// that a FASTQ text ending "...\n+\n" has an empty quality line there (an empty read at the very end of a file).
__device__ __forceinline__ uint64_t kmb_tp_n_lines(const uint32_t *__restrict__ nl_pos, uint64_t n_nl, uint64_t n_text, int format) {
    const bool open_end = n_text && (n_nl == 0 || (uint64_t)nl_pos[n_nl - 1] + 1 < n_text);
    uint64_t n = n_nl + (open_end ? 1 : 0);
    if (format == 1 && !open_end && (n & 3ull) == 3ull) n++;
    return n;
}
// Line i of the text: [start, end) without its newline; n_nl = number of newlines, the text may end without one.
__device__ __forceinline__ void kmb_tp_line(const uint32_t *__restrict__ nl_pos, uint64_t n_nl, uint64_t n_text, uint64_t i,
                                            uint64_t &start, uint64_t &end) {
    start = i ? (i - 1 < n_nl ? (uint64_t)nl_pos[i - 1] + 1 : n_text) : 0;
    end = i < n_nl ? (uint64_t)nl_pos[i] : n_text;
}
// What a line is: 0 = nothing to keep, 1 = header (begins a read), 2 = sequence.  len = sequence bytes ('\r' stripped).
__device__ __forceinline__ int kmb_tp_kind(const uint8_t *__restrict__ text, int format, uint64_t i, uint64_t n_lines_fq,
                                           uint64_t start, uint64_t end, uint32_t &len, unsigned &err) {
    uint64_t l = end - start;
    if (l && text[end - 1] == '\r') l--;
    len = 0;
    if (format == 1) {                       // FASTQ
        if (i >= n_lines_fq) {               // beyond the last whole record: only blank lines may follow
            if (l) err |= KMB_TP_ERR_MALFORMED;
            return 0;
        }
        const unsigned which = (unsigned)(i & 3u);
        if (which == 0) {
            if (end == start || text[start] != '@') err |= KMB_TP_ERR_MALFORMED;
            return 1;
        }
        if (which == 1) {
            len = (uint32_t)l;
            return 2;
        }
        if (which == 2 && (end == start || text[start] != '+')) err |= KMB_TP_ERR_MALFORMED;
        return 0;
    }
    if (end > start && text[start] == '>') return 1;   // FASTA header
    if (i == 0 && l) err |= KMB_TP_ERR_MALFORMED;      // data before the first '>'
    len = (uint32_t)l;
    return 2;
}

// T4: per block of 1024 lines, the sequence bytes and the headers in it.  n_lines = lines of the text (a last line
// without newline counts, an empty tail after the last newline does not).
__global__ void __launch_bounds__(256) kmb_tp_classify(const uint8_t *__restrict__ text, uint64_t n_text, int format,
                                                        const uint32_t *__restrict__ nl_pos, const unsigned long long *n_nl_ptr,
                                                        uint64_t nl_capacity, unsigned long long *__restrict__ blk_bases,
                                                        uint32_t *__restrict__ blk_reads, KmbTextResult *res) {
    __shared__ unsigned long long s_b[8];
    __shared__ uint32_t s_r[8];
    const uint64_t n_nl = *n_nl_ptr < nl_capacity ? *n_nl_ptr : nl_capacity;  // beyond the capacity: flagged, retried by the host
    const uint64_t n_lines = kmb_tp_n_lines(nl_pos, n_nl, n_text, format);
    const uint64_t n_lines_fq = n_lines & ~3ull;   // FASTQ: whole records; what follows must be blank
    unsigned long long bases = 0;
    uint32_t reads = 0;
    unsigned err = 0;
    const uint64_t first = (uint64_t)blockIdx.x * KMB_TP_LINES_PER_BLOCK;
    for (int j = threadIdx.x; j < KMB_TP_LINES_PER_BLOCK; j += 256) {
        const uint64_t i = first + (uint64_t)j;
        if (i >= n_lines) break;
        uint64_t s, e;
        kmb_tp_line(nl_pos, n_nl, n_text, i, s, e);
        uint32_t len;
        const int kind = kmb_tp_kind(text, format, i, n_lines_fq, s, e, len, err);
        bases += len;
        reads += kind == 1;
    }
    for (int o = 16; o > 0; o >>= 1) {
        bases += __shfl_xor_sync(0xFFFFFFFFu, bases, o);
        reads += __shfl_xor_sync(0xFFFFFFFFu, reads, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_b[threadIdx.x >> 5] = bases;
        s_r[threadIdx.x >> 5] = reads;
    }
    if (err) atomicOr(&res->error, err);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tb = 0;
        uint32_t tr = 0;
        for (int w = 0; w < 8; w++) {
            tb += s_b[w];
            tr += s_r[w];
        }
        blk_bases[blockIdx.x] = tb;
        blk_reads[blockIdx.x] = tr;
        if (blockIdx.x == 0) res->n_lines = n_lines;
    }
}

// T5: one CTA per 1024 lines, one thread per 4 consecutive lines: block-local exclusive prefixes of (bases, reads) on top
// of the scanned block sums give every line its destination and its read.
struct KmbTextCopy {  // a sequence line to copy
    uint32_t src;      // offset in the text
    uint32_t len;
    unsigned long long dst;  // offset in bases[]
};
__global__ void __launch_bounds__(256) kmb_tp_emit(const uint8_t *__restrict__ text, uint64_t n_text, int format,
                                                    const uint32_t *__restrict__ nl_pos, const unsigned long long *n_nl_ptr,
                                                    uint64_t nl_capacity,
                                                    const unsigned long long *__restrict__ blk_bases, const uint32_t *__restrict__ blk_reads,
                                                    const unsigned long long *total_bases, const unsigned long long *total_reads,
                                                    int64_t *__restrict__ offsets, uint64_t offsets_capacity,
                                                    KmbTextCopy *__restrict__ copies, KmbTextResult *res) {
    __shared__ unsigned long long s_b[256];
    __shared__ uint32_t s_r[256];
    const uint64_t n_nl = *n_nl_ptr < nl_capacity ? *n_nl_ptr : nl_capacity;  // beyond the capacity: flagged, retried by the host
    const uint64_t n_lines = kmb_tp_n_lines(nl_pos, n_nl, n_text, format);
    const uint64_t n_lines_fq = n_lines & ~3ull;
    const uint64_t first = (uint64_t)blockIdx.x * KMB_TP_LINES_PER_BLOCK + (uint64_t)threadIdx.x * 4;
    uint32_t len[4];
    int kind[4];
    uint64_t start[4];
    unsigned long long my_b = 0;
    uint32_t my_r = 0;
    unsigned err = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint64_t i = first + (uint64_t)j;
        kind[j] = 0;
        len[j] = 0;
        start[j] = 0;
        if (i < n_lines) {
            uint64_t e;
            kmb_tp_line(nl_pos, n_nl, n_text, i, start[j], e);
            kind[j] = kmb_tp_kind(text, format, i, n_lines_fq, start[j], e, len[j], err);
        }
        my_b += len[j];
        my_r += kind[j] == 1;
    }
    s_b[threadIdx.x] = my_b;
    s_r[threadIdx.x] = my_r;
    __syncthreads();
    // exclusive prefix over the 256 threads (serial per thread over shared memory: 256 x 256 reads, cheap next to the text)
    unsigned long long b = blk_bases[blockIdx.x];
    uint32_t r = blk_reads[blockIdx.x];
    for (int t = 0; t < (int)threadIdx.x; t++) {
        b += s_b[t];
        r += s_r[t];
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint64_t i = first + (uint64_t)j;
        if (i >= n_lines) break;
        if (kind[j] == 1) {
            if (r < offsets_capacity) offsets[r] = (int64_t)b;
            r++;
        }
        copies[i].src = (uint32_t)start[j];
        copies[i].len = kind[j] == 2 ? len[j] : 0u;
        copies[i].dst = b;
        b += len[j];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned long long nr = *total_reads, nb = *total_bases;
        if (nr < offsets_capacity) offsets[nr] = (int64_t)nb;
        else atomicOr(&res->error, KMB_TP_ERR_TOO_MANY_LINES);
        res->n_reads = nr;
        res->n_bases = nb;
        if (format == 1 && nr * 4 != n_lines_fq) atomicOr(&res->error, KMB_TP_ERR_MALFORMED);
    }
}

// T6: one warp per line: copy its bases.  Consecutive lanes, consecutive bytes.
__global__ void __launch_bounds__(256) kmb_tp_copy(const uint8_t *__restrict__ text, const KmbTextCopy *__restrict__ copies,
                                                    const KmbTextResult *res, uint8_t *__restrict__ bases) {
    const uint64_t n_lines = res->n_lines;
    const int lane = threadIdx.x & 31;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n_lines; i += warps) {
        const KmbTextCopy c = copies[i];
        for (uint32_t j = (uint32_t)lane; j < c.len; j += 32u) bases[c.dst + j] = text[c.src + j];
    }
}
