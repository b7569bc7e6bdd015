/* CPU oracle (plain C) for kmer_mapper's read -> k-mer -> lookup -> count path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the checker or as the
 * timed CPU baseline ("port").  The product library (kmer_mapper_b200/csrc) never links it.
 *
 * Each function restates a piece of the reference; citations are paths under /root/reference.
 * Pinning: see the header of oracle/oracle.py (lookup/count pinned against the compiled
 * reference mapper.pyx and its golden vector; hash formula pinned by tests/test_hashing.py:13-26;
 * ASCII->code table / invalid-byte policy / short reads are PARITY UNPINNED because bionumpy is
 * an absent third-party dependency).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

/* bionumpy DNAEncoding as used at kmer_mapper/util.py:72-73: A,C,G,T -> 0..3, case-insensitive.
 * 'N' (upper case only) -> 'A' first when n_to_a (command_line_interface.py:40-41). */
static inline int ko_code(uint8_t b, int n_to_a) {
    switch (b) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        case 'N': return n_to_a ? 0 : -1;
        default: return -1;
    }
}

/* util.py:71-75 get_kmer_hashes_from_chunk_sequence: hash[r,p] = sum_j code(r[p+j]) * 4^j
 * (first base in the lowest bits, tests/test_hashing.py:13-26), flat read-major order, no window
 * across reads, reads shorter than k give nothing.  Returns the number of hashes written, or -1
 * and *bad_offset = flat offset of the first invalid byte. `out` may be NULL to count only. */
int64_t ko_kmer_hashes(const uint8_t *bases, const int64_t *offsets, int64_t n_reads, int k,
                       int n_to_a, uint64_t *out, int64_t *bad_offset) {
    const uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1ull);
    int64_t w = 0;
    for (int64_t r = 0; r < n_reads; r++) {
        uint64_t h = 0;
        int64_t s = offsets[r], e = offsets[r + 1];
        for (int64_t i = s; i < e; i++) {
            int c = ko_code(bases[i], n_to_a);
            if (c < 0) { if (bad_offset) *bad_offset = i; return -1; }
            h = (h >> 2) | ((uint64_t)c << (2 * (k - 1)));
            h &= mask;
            if (i - s + 1 >= k) { if (out) out[w] = h; w++; }
        }
    }
    return w;
}

/* mapper.pyx:53-69 -- the loop itself. counts must hold max_node_id+1 zero-initialised (or running)
 * uint32 values; += wraps mod 2^32 like the reference's np.uint32. */
void ko_map_kmers(const int32_t *hashes_to_index, const int32_t *n_kmers, const int32_t *nodes,
                  const uint64_t *index_kmers, const uint16_t *frequencies, uint64_t modulo,
                  const uint64_t *kmers, int64_t n, int max_frequency, uint32_t *counts) {
    for (int64_t i = 0; i < n; i++) {
        uint64_t q = kmers[i];
        uint64_t h = q % modulo;                      /* :54 */
        int nl = n_kmers[h];                          /* :55 */
        int64_t l = hashes_to_index[h];               /* :56 */
        for (int j = 0; j < nl; j++, l++) {           /* :58 no break */
            if (index_kmers[l] != q) continue;        /* :60 */
            if ((int)frequencies[l] > max_frequency) continue; /* :64 */
            counts[nodes[l]] += 1;                    /* :68 */
        }
    }
}

/* mapper.pyx:81-130 in_graph_index: first key match -> 1; frequency ignored (:112-127). */
void ko_in_graph_index(const int32_t *hashes_to_index, const int32_t *n_kmers,
                       const uint64_t *index_kmers, uint64_t modulo,
                       const uint64_t *kmers, int64_t n, uint8_t *out) {
    for (int64_t i = 0; i < n; i++) {
        uint64_t q = kmers[i];
        uint64_t h = q % modulo;
        int nl = n_kmers[h];
        int64_t l = hashes_to_index[h];
        uint8_t hit = 0;
        for (int j = 0; j < nl; j++, l++)
            if (index_kmers[l] == q) { hit = 1; break; }
        out[i] = hit;
    }
}

/* command_line_interface.py:32-56 map_cpu body (N->A :41, hash :42, lookup :51) for one chunk of
 * reads, hashing and probing fused per read; OpenMP over reads with atomic uint32 adds so that the
 * "port" CPU baseline can use every host core (the reference scales by worker processes,
 * command_line_interface.py:124-130; wrap-around sums make any order bit-exact).
 * Returns 0, or -1 with *bad_offset set when an invalid byte was met (counts are then partial). */
int ko_map_reads(const int32_t *hashes_to_index, const int32_t *n_kmers, const int32_t *nodes,
                 const uint64_t *index_kmers, const uint16_t *frequencies, uint64_t modulo,
                 const uint8_t *bases, const int64_t *offsets, int64_t n_reads, int k,
                 int max_frequency, int n_threads, uint32_t *counts, int64_t *n_kmers_mapped,
                 int64_t *bad_offset) {
    const uint64_t mask = (1ull << (2 * k)) - 1ull;
    int64_t bad = -1, total = 0;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(n_threads) reduction(+:total)
    for (int64_t r = 0; r < n_reads; r++) {
        uint64_t q = 0;
        int64_t s = offsets[r], e = offsets[r + 1];
        for (int64_t i = s; i < e; i++) {
            int c = ko_code(bases[i], 1);
            if (c < 0) {
#pragma omp critical
                { if (bad < 0 || i < bad) bad = i; }
                break;
            }
            q = ((q >> 2) | ((uint64_t)c << (2 * (k - 1)))) & mask;
            if (i - s + 1 < k) continue;
            total++;
            uint64_t h = q % modulo;
            int nl = n_kmers[h];
            int64_t l = hashes_to_index[h];
            for (int j = 0; j < nl; j++, l++) {
                if (index_kmers[l] != q) continue;
                if ((int)frequencies[l] > max_frequency) continue;
                __atomic_fetch_add(&counts[nodes[l]], 1u, __ATOMIC_RELAXED);
            }
        }
    }
    if (n_kmers_mapped) *n_kmers_mapped = total;
    if (bad >= 0) { if (bad_offset) *bad_offset = bad; return -1; }
    return 0;
}
