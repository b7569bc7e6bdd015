// kmb_host.h -- host-side helpers shared by kmb_capi.cu and kmb_hostpack.cpp (not part of the C ABI).
#pragma once
#include <stdint.h>

// Worker threads available to this process (CPU affinity mask), at least 1.
int kmb_host_cpus();

// Run fn(ctx, part) for part = 0 .. n_parts-1 on the library's persistent worker pool (the calling thread takes
// parts too) and return when all are done.  n_threads <= 0: every CPU of the affinity mask.  Calls from several
// threads are serialised.
void kmb_host_parallel(int n_threads, int n_parts, void (*fn)(void *ctx, int part), void *ctx);

// 2-bit transport encoding of bases[0, n_bases) (the arithmetic of kmb_encode16, kmb_core.cuh): word j of
// `words` = bases 16j .. 16j+15, base 16j in the lowest bits; positions >= n_bases read as 'A'.  Writes
// kmb_packed_words(n_bases) words (the last KMB_PACK_PAD_WORDS are zero: the kernels read a halo past the end).
// Returns the offset of the first invalid byte or ~0.
#define KMB_PACK_PAD_WORDS 4
static inline uint64_t kmb_packed_words(uint64_t n_bases) { return (n_bases + 15) / 16 + KMB_PACK_PAD_WORDS; }
uint64_t kmb_host_pack(const uint8_t *bases, uint64_t n_bases, bool n_to_a, int n_threads, uint32_t *words);
// Streaming (non-temporal) stores for the packed words from now on (option "host_pack_streaming").
void kmb_host_pack_streaming(bool on);

// out[i] = offsets[i] - base for i in [0, n): the chunk-relative 32-bit read offsets that travel with a packed chunk.
void kmb_host_rel_offsets(const int64_t *offsets, uint64_t n, int64_t base, int n_threads, uint32_t *out);

// One gzip member from gz[0, n_gz) into a malloc'd buffer (kmb_inflate.cpp's DEFLATE decoder, CRC-32 checked):
// 1 = ok, 2 = not a member / corrupt / truncated, 3 = inflates to more than `limit` bytes.  The caller frees *buf.
int kmb_inflate_member(const uint8_t *gz, uint64_t n_gz, uint64_t limit, uint8_t **buf, size_t *n_out, uint64_t *consumed);
