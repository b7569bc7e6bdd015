// Host build of kmb_core.cuh for unit tests (g++, no CUDA): exposes the pure bit-level primitives
// the kernels are made of so that tests/test_core_host.py can check them against Python integers.
// This is a test harness for arithmetic only -- it contains no mapping path.
#include "../../kmer_mapper_b200/csrc/kmb_core.cuh"

extern "C" {
void h_divmod(const uint64_t *n, int64_t cnt, uint64_t d, uint64_t *q, uint64_t *r) {
    KmbMod m = kmb_mod_make(d);
    for (int64_t i = 0; i < cnt; i++) kmb_divmod(n[i], m, q[i], r[i]);
}
// encode 16 ASCII bytes: returns packed 32-bit code word, *invalid = per-base invalid bit mask
uint32_t h_encode16(const uint8_t *b, int n_to_a, uint32_t *invalid) {
    uint32_t w[4];
    for (int a = 0; a < 4; a++) w[a] = (uint32_t)b[4 * a] | ((uint32_t)b[4 * a + 1] << 8) | ((uint32_t)b[4 * a + 2] << 16) | ((uint32_t)b[4 * a + 3] << 24);
    return kmb_encode16(w[0], w[1], w[2], w[3], n_to_a != 0, *invalid);
}
uint64_t h_window(uint64_t lo, uint64_t hi, int i, int k) { return kmb_window(lo, hi, i, kmb_kmer_mask(k)); }
uint64_t h_revcomp(uint64_t x, int k) { return kmb_revcomp(x, k); }
uint64_t h_chain_line(uint64_t main_line, uint32_t ovf_base, uint32_t s) { return kmb_chain_line(main_line, ovf_base, s); }
uint32_t h_chain_slot(uint32_t s) { return kmb_chain_slot(s); }
uint32_t h_chain_extra_lines(uint32_t n) { return kmb_chain_extra_lines(n); }
void h_locate(uint64_t key, uint32_t n_main, uint32_t n_words, uint32_t two, uint32_t *out) {
    KmbAddr a;
    a.n_main = n_main;
    a.n_filter_words = n_words;
    a.n_probes = two ? 2u : 1u;
    KmbLoc l = kmb_locate(key, a);
    out[0] = l.sector;
    out[1] = l.fword;
    out[2] = l.fmask;
}
uint32_t h_sector_header(uint32_t remaining, uint32_t next) { return kmb_sector_header(remaining, next); }
uint32_t h_header_count(uint32_t hdr) { return kmb_header_count(hdr); }
uint32_t h_log_bin(uint32_t node, uint32_t shift) { return kmb_log_bin(node, shift); }

}
