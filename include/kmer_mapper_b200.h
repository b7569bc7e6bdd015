/* kmer_mapper_b200 -- C ABI of the B200-native k-mer mapping path.
 *
 * This is the drop-in boundary for kmer_mapper's read -> k-mer -> index lookup -> per-node count
 * path.  Every entry point below replaces one reference interface (cited as file:line under
 * ivargr/kmer_mapper v0.0.37) and is what a ctypes/cffi binding on the reference side would bind;
 * INTEGRATION.md shows that binding.  Plain pointers and sizes only: no torch, numpy or C++ types.
 *
 * Conventions
 *   - Every function returns an int status: KMB_OK (0) or a negative KMB_ERR_* class.  Nothing
 *     throws across the boundary; kmb_last_error() returns a thread-local message for the last
 *     failure on the calling thread.
 *   - "buf" pointers may be HOST or DEVICE memory of the handle's GPU; the library asks the CUDA
 *     driver which (cudaPointerGetAttributes).  Host buffers are copied with asynchronous
 *     host-to-device copies on a side stream, double buffered against the kernels; device buffers
 *     are used in place (they must be 16-byte aligned).
 *   - One handle per GPU, not re-entrant: one host thread drives a handle at a time.  All device
 *     allocations are owned by the handle that made them.  The caller keeps ownership of every
 *     buffer it passes in.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *     KMB_ERR_CUDA.
 */
#ifndef KMER_MAPPER_B200_H
#define KMER_MAPPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMB_OK 0
#define KMB_ERR_BAD_ARG (-1)       /* null pointer, bad k, misaligned device buffer, size mismatch ... */
#define KMB_ERR_INVALID_BASE (-2)  /* a byte outside ACGTacgt/N in the reads; see kmb_mapper_bad_offset */
#define KMB_ERR_CUDA (-3)          /* CUDA runtime error (message in kmb_last_error) */
#define KMB_ERR_BAD_INDEX (-4)     /* index arrays inconsistent: bucket out of range, negative node ... */
#define KMB_ERR_NOMEM (-5)

#define KMB_FLAG_REVCOMP 1u        /* also look up the reverse complement of every window
                                      (command_line_interface.py:74,180-182, GPU route's -r) */
#define KMB_FLAG_NO_N_TO_A 2u      /* do NOT apply the CPU route's N->A policy
                                      (command_line_interface.py:40-41): 'N' becomes an invalid byte */

typedef struct kmb_index kmb_index;   /* device-resident, re-laid-out k-mer index (one GPU) */
typedef struct kmb_mapper kmb_mapper; /* a count buffer + streams + staging bound to one index */

/* ---- library / device ---------------------------------------------------------------------- */

const char *kmb_last_error(void);
const char *kmb_version(void);
int kmb_device_count(int *n_devices);

/* ---- index: replaces the six attributes mapper.pyx:22-29 reads from a KmerIndex -------------
 * hashes_to_index int32[modulo], n_kmers int32[modulo], nodes int32[n_entries],
 * kmers uint64[n_entries], frequencies uint16[n_entries], modulo scalar.
 * The arrays are copied to `device`, validated once (the reference runs with boundscheck off,
 * mapper.pyx:15-18: every bucket must lie inside [0, n_entries], nodes must be >= 0) and re-laid
 * out on the device into read-only 32-byte sectors (keys, nodes and frequencies) addressed by a hash
 * of the key, plus an L2-resident filter.  Limits: modulo < 2^32, n_entries < 2^31. */
int kmb_index_create(int device,
                     const int32_t *hashes_to_index, const int32_t *n_kmers, uint64_t modulo,
                     const int32_t *nodes, const uint64_t *kmers, const uint16_t *frequencies,
                     uint64_t n_entries, kmb_index **out);
int kmb_index_destroy(kmb_index *index);
/* max_node_id = nodes.max() (KmerIndex.max_node_id(), command_line_interface.py:51,79,117) */
int kmb_index_info(const kmb_index *index, int64_t *max_node_id, uint64_t *n_entries,
                   uint64_t *modulo, uint64_t *device_bytes);
/* Size of the L2-resident bucket filter (probe level 0), 0 when not in use. */
int kmb_index_filter_bytes(const kmb_index *index, uint64_t *bytes);
/* Geometry of the sector table: main sectors, overflow sectors and the number of live entries
 * (entries that lie inside the bucket range of their own key -- the only ones the reference's scan
 * can ever match). */
int kmb_index_layout(const kmb_index *index, uint64_t *n_main_sectors, uint64_t *n_overflow_sectors,
                     uint64_t *n_live_entries);

/* ---- mapper: replaces map_kmers_to_graph_index (mapper.pyx:19-72), the per-chunk worker map_cpu
 * (command_line_interface.py:32-56) and cucounter's count() as driven by GpuCounter
 * (gpu_counter.py:23-24).  Holds node_counts uint32[n_counts], n_counts = max_node_id+1
 * (mapper.pyx:37), cumulative across calls like the reference's additive map-reduce
 * (command_line_interface.py:124-130); sums wrap mod 2^32.
 * counts_device: NULL -> the mapper allocates and zeroes its own buffer; otherwise a caller-owned
 * DEVICE buffer of n_counts uint32 (e.g. a torch tensor that will be NCCL-all-reduced), used as is.
 * max_index_lookup_frequency: entries with frequency > this are skipped (mapper.pyx:19,64). */
int kmb_mapper_create(kmb_index *index, uint64_t n_counts, uint32_t *counts_device,
                      int max_index_lookup_frequency, kmb_mapper **out);
int kmb_mapper_destroy(kmb_mapper *mapper);
/* Run the mapper's kernels on a caller stream (a cudaStream_t, e.g. torch's current stream);
 * NULL restores the mapper's own stream. */
int kmb_mapper_set_stream(kmb_mapper *mapper, void *cuda_stream);

/* mapper.pyx:19 map_kmers_to_graph_index(index, max_node_id, kmers): add the counts of n uint64
 * k-mers (host or device buffer) to the mapper's node counts. */
/* One chunk of raw FASTA (format 0) or FASTQ (format 1) TEXT -- whole records: it begins at a record start and ends
 * at a record end (the last line may lack its newline) -- parsed ON THE DEVICE and mapped: replaces
 * bnp.open(f).read_chunks(...) -> chunk.sequence -> map (command_line_interface.py:102-111 + :32-56) for the GPU
 * route.  The text crosses PCIe once (host text is staged through pinned memory; pageable text is first copied into
 * it by every core); newline scan, line classification, offsets and the compaction of the bases are kernels
 * (csrc/kmb_textparse.cuh); then the fused kernel runs as for device-resident input.  Record rules as
 * kmb_parse_reads.  The previous call's kernels overlap this call's copy and parse.  At most 4 GiB per call. */
int kmb_mapper_map_text(kmb_mapper *mapper, const uint8_t *text, uint64_t n_text, int format, int k, uint32_t flags);
/* The same with the text taken from bytes [offset, offset + n_text) of an open file descriptor: the cores pread their
 * shares straight into the pinned staging buffer. */
int kmb_mapper_map_text_fd(kmb_mapper *mapper, int fd, uint64_t offset, uint64_t n_text, int format, int k, uint32_t flags);
/* A multi-member .gz of FASTA / FASTQ (bgzip / BGZF, concatenated .gz files) -- gz[0, n_gz) begins at a member --
 * inflated, parsed and mapped ON THE DEVICE: the compressed bytes cross PCIe, every member is decoded by one warp
 * (csrc/kmb_gzdev.cuh), checked against the length and CRC-32 of its trailer, and the text goes straight into the
 * device parser.  The members are taken in batches (gz_device_batch_bytes of text); of these the call processes every
 * shard_count-th one, starting with shard_index (the ranks of a multi-GPU job), and a record that straddles two
 * batches is mapped by the batch it begins in (each batch inflates up to 1 MB of the next one's text to complete it).
 * A batch the device decoder gets wrong is inflated again by the host decoder (kmb_gunzip_members); data neither can
 * decode is KMB_ERR_BAD_ARG.  *resume_offset = n_gz when everything was done here; otherwise the offset of the member
 * from which the host decoders (kmb_gunzip_members / kmb_gzstream_*) have to continue because a member further on is
 * too large to give to one warp (a plain single-member .gz: *resume_offset = 0) -- a function of the data alone, so all
 * ranks get the same answer -- skipping, unless that offset is 0, the text up to the first record start after the
 * first newline (kmb_find_record_start), which this call has already mapped.
 * Replaces the .gz side of bnp.open(path) (command_line_interface.py:102-103, Readme.md:11). */
int kmb_mapper_map_gz(kmb_mapper *mapper, const uint8_t *gz, uint64_t n_gz, int format, int k, uint32_t flags,
                      int shard_index, int shard_count, uint64_t *resume_offset);
/* gzip members and text bytes inflated by the device so far in this process; batches the host decoder had to redo. */
int kmb_gz_device_stats(uint64_t *n_members, uint64_t *text_bytes, uint64_t *host_batches);
/* The device parser on its own (the counterpart of kmb_parse_reads below, run by GPU kernels): bases of the reads back
 * to back and offsets[0..n_reads], into host or device buffers.  KMB_ERR_NOMEM (with the sizes in n_reads / n_bases)
 * when a capacity is too small, KMB_ERR_BAD_ARG for malformed records. */
int kmb_parse_text_device(int device, const uint8_t *text, uint64_t n_text, int format, uint8_t *bases,
                          uint64_t bases_capacity, int64_t *offsets, uint64_t offsets_capacity, uint64_t *n_reads,
                          uint64_t *n_bases);
/* Reads and bases parsed by kmb_mapper_map_text in this process so far. */
int kmb_text_parsed(uint64_t *n_reads, uint64_t *n_bases);

int kmb_mapper_map_kmers(kmb_mapper *mapper, const uint64_t *kmers, uint64_t n, uint32_t flags, int k);

/* command_line_interface.py:32-56 map_cpu body, fused: N->A (:41), 2-bit encode + every in-read
 * window hashed as sum_j code[p+j]*4^j (util.py:71-75), lookup + count (mapper.pyx:53-69).
 * bases   uint8[n_bases]    ASCII bases of all reads back to back (no separators)
 * offsets int64[n_reads+1]  read r = bases[offsets[r] : offsets[r+1]]; offsets[0] == 0,
 *                           offsets[n_reads] == n_bases, non-decreasing
 * Both buffers on the host, or both on the device. 0 < k < 32. */
int kmb_mapper_map_reads(kmb_mapper *mapper, const uint8_t *bases, uint64_t n_bases,
                         const int64_t *offsets, uint64_t n_reads, int k, uint32_t flags);

/* Hits (node ids that passed the frequency cut-off, mapper.pyx:64) are first appended to a device-side
 * log in groups tagged by node range; the flush plays the log into the node counts (mapper.pyx:68) one
 * L2-sized window at a time.  kmb_mapper_flush queues that pass on the mapper's stream without waiting
 * (use it before handing the count buffer to an all-reduce on the same stream); sync, read_counts,
 * stats and lookup_counts flush implicitly. */
int kmb_mapper_flush(kmb_mapper *mapper);
/* Wait for all queued work of the mapper (flushing first); returns KMB_ERR_INVALID_BASE if any kernel met an
 * invalid byte since the last reset (counts are then undefined until kmb_mapper_reset). */
int kmb_mapper_sync(kmb_mapper *mapper);
/* Flat offset (within the call that failed) of the first invalid byte, or -1. */
int kmb_mapper_bad_offset(kmb_mapper *mapper, int64_t *offset);
/* Copy the n_counts node counts to a host (or device) buffer; implies kmb_mapper_sync. */
int kmb_mapper_read_counts(kmb_mapper *mapper, uint32_t *out, uint64_t n_counts);
/* Replace the counts by n_counts (= the mapper's) values from a host or device buffer; hits mapped before the
 * call are dropped.  This is `counter._values = node_counts` of the CounterKmerIndex route
 * (command_line_interface.py:136) and the `initial_data` of the additive map-reduce (:116-119). */
int kmb_mapper_write_counts(kmb_mapper *mapper, const uint32_t *values, uint64_t n_counts);
/* Zero the counts, the statistics and the error state. */
int kmb_mapper_reset(kmb_mapper *mapper);
/* Device pointer of the count buffer (for an in-place NCCL all-reduce by the host framework). */
int kmb_mapper_counts_device(kmb_mapper *mapper, uint32_t **counts_device, uint64_t *n_counts);
/* Windows looked up and index entries counted since the last reset (implies kmb_mapper_sync). */
int kmb_mapper_stats(kmb_mapper *mapper, uint64_t *n_kmers_mapped, uint64_t *n_entries_counted);

/* ---- multi-GPU reduction: replaces the additive map-reduce of command_line_interface.py:124-130 ----
 * The reference's only parallelism is data parallelism over read chunks: every worker process returns a
 * uint32[max_node_id+1] array and shared_memory_wrapper's additative_shared_array_map_reduce sums them
 * element-wise.  Here every GPU (one process per GPU) maps its share of the reads into the private count
 * buffer of its mapper and the buffers are summed in place by ONE ncclAllReduce(ncclUint32, ncclSum) over
 * NVLink; uint32 addition wraps, so the result has the reference's bits whatever the reduction order.
 * NCCL is loaded at run time (dlopen of libnccl.so.2 -- the copy a host framework such as PyTorch already
 * loaded, else $KMB_NCCL_LIB, else the loader path); without it these calls return KMB_ERR_NCCL.
 *   rank 0:     kmb_comm_unique_id(id)  and sends the 128 bytes to the other ranks by any means
 *   every rank: kmb_comm_init_rank(device, n_ranks, rank, id, &comm)      (collective, blocking)
 *               ... kmb_mapper_map_reads(...) on its own reads ...
 *               kmb_mapper_allreduce(mapper, comm)   queues flush + all-reduce on the mapper's stream
 *               kmb_mapper_read_counts(mapper, ...)  now yields the job's total on every rank
 * The count buffer is registered with the communicator (ncclCommRegister) on first use, so NCCL can use it
 * in place.  n_counts must be equal on all ranks. */
#define KMB_ERR_NCCL (-6)
#define KMB_COMM_ID_BYTES 128
typedef struct kmb_comm kmb_comm;
int kmb_comm_unique_id(uint8_t id[KMB_COMM_ID_BYTES]);
int kmb_comm_init_rank(int device, int n_ranks, int rank, const uint8_t id[KMB_COMM_ID_BYTES], kmb_comm **comm);
int kmb_comm_destroy(kmb_comm *comm);
int kmb_mapper_allreduce(kmb_mapper *mapper, kmb_comm *comm);

/* ---- membership: replaces in_graph_index / in_graph_index_no_memory_maps (mapper.pyx:81,137) --
 * out[i] = 1 iff some entry of bucket kmers[i] % modulo has key kmers[i]; frequency ignored. */
int kmb_in_graph_index(kmb_index *index, const uint64_t *kmers, uint64_t n, uint8_t *out);

/* ---- hashing: replaces get_kmer_hashes_from_chunk_sequence (util.py:71-75) --------------------
 * Writes the hash of every in-read window, read-major then position order, to out (host or device,
 * capacity out_capacity); *n_out = number written = sum_r max(0, L_r - k + 1).
 * flags: KMB_FLAG_NO_N_TO_A as above (the bare util.py function has no N policy; the CPU route
 * applies it before calling, command_line_interface.py:41-42). */
int kmb_hash_reads(int device, const uint8_t *bases, uint64_t n_bases, const int64_t *offsets,
                   uint64_t n_reads, int k, uint32_t flags, uint64_t *out, uint64_t out_capacity,
                   uint64_t *n_out, int64_t *bad_offset);

/* ---- per-key counter: replaces cucounter.Counter as seen from gpu_counter.py:16,24,33 ----------
 * A counter over unique keys is a mapper whose index maps key i -> "node" i (built by the host
 * shim), so count() is kmb_mapper_map_kmers and Counter.__getitem__ (gpu_counter.py:33) is this
 * lookup: out[i] = counts[node of the first entry whose key equals keys[i]], 0 when absent. */
int kmb_mapper_lookup_counts(kmb_mapper *mapper, const uint64_t *keys, uint64_t n, uint32_t *out);

/* ---- legacy 2-bit codec: replaces encodings.py:25-112 on the device -------------------------- */
int kmb_codec_actg_from_bytes(int device, const uint8_t *seq, uint64_t n, uint8_t *out);   /* :51-59 */
int kmb_codec_simple_from_bytes(int device, const uint8_t *seq, uint64_t n, uint8_t *out); /* :96-102 */
int kmb_codec_to_bytes(int device, const uint8_t *packed, uint64_t n, uint8_t *out);       /* :70-75 */
int kmb_codec_complement(int device, const uint8_t *in, uint64_t n_bytes, uint8_t *out);   /* :44-48 */
int kmb_codec_twobit_swap(int device, const void *in, uint64_t n_words, int word_bytes, void *out); /* :104-112 */

/* ---- chunked read parsing: replaces bnp.open(path).read_chunks(min_chunk_size) -> chunk.sequence
 * (command_line_interface.py:102-111).  Host-side, multi-threaded; no GPU involved.
 * text[0, n_text) is a piece of a FASTA (format 0; multi-line allowed) or FASTQ (format 1; 4-line
 * records) file that starts at a record boundary.  Every COMPLETE record is parsed (all of them when
 * final_chunk): bases of the reads back to back into bases[], offsets[0..n_reads] (offsets[0] = 0);
 * *consumed = bytes of text used, the caller carries the rest over to the next chunk.  With
 * bases == NULL only the counts are returned.  Returns KMB_ERR_BAD_ARG for malformed records
 * (FASTQ record not starting with '@' / third line not '+', FASTA data before the first '>'),
 * KMB_ERR_NOMEM when an output capacity is too small. */
int kmb_parse_reads(const uint8_t *text, uint64_t n_text, int format, int final_chunk, int n_threads,
                    uint8_t *bases, uint64_t bases_capacity, int64_t *offsets, uint64_t offsets_capacity,
                    uint64_t *n_reads, uint64_t *n_bases, uint64_t *consumed);

/* Member-parallel gzip inflate for the same reader (.fa.gz / .fq.gz: Readme.md:11, command_line_interface.py:166).
 * gz[0, n_gz) starts at a member boundary of a .gz file.  Whole members are inflated in order, several at a time
 * (bgzip/BGZF blocks, concatenated .gz files), into out[]; stops before a member that would overflow out_capacity,
 * or at a member that inflates to more than max_member_bytes (0 = out_capacity): *stopped_at_big_member = 1, the
 * caller streams that one sequentially (a plain single-member .gz cannot be inflated in parallel), = 2 when the
 * bytes at *consumed are not a gzip member (corrupt file / trailing garbage).  *consumed / *produced = compressed
 * bytes used / text bytes written.  KMB_ERR_BAD_ARG when gz[0] itself is not a gzip member. */
int kmb_gunzip_members(const uint8_t *gz, uint64_t n_gz, int n_threads, uint8_t *out, uint64_t out_capacity,
                       uint64_t max_member_bytes, uint64_t *consumed, uint64_t *produced, int *stopped_at_big_member);

/* Streaming decoder for what kmb_gunzip_members cannot split: a .gz file that is one long deflate stream (plain
 * `gzip reads.fq`).  A from-scratch DEFLATE decoder tuned for one core (csrc/kmb_inflate.cpp), every member checked
 * against the CRC-32 and length in its trailer.  gz[0, n_gz) must stay mapped while the stream is open and start at
 * a member boundary.  kmb_gzstream_read continues the stream into out[0, out_capacity) (>= 64 KB); the last
 * min(32768, bytes produced so far) bytes of the previous call's output must sit directly in front of `out`
 * (history_bytes says how many the caller put there).  *finished = 1 after the last member.  Returns
 * KMB_ERR_BAD_ARG on corrupt or truncated input; kmb_gzstream_error says why. */
typedef struct kmb_gzstream kmb_gzstream;
int kmb_gzstream_open(const uint8_t *gz, uint64_t n_gz, int n_threads, kmb_gzstream **out);
int kmb_gzstream_read(kmb_gzstream *stream, uint8_t *out, uint64_t out_capacity, uint64_t history_bytes,
                      uint64_t *produced, int *finished);
const char *kmb_gzstream_error(const kmb_gzstream *stream);
int kmb_gzstream_close(kmb_gzstream *stream);

/* First record start AFTER the first newline of text[0, n_text) (n_text when there is none): how a rank of a
 * multi-GPU job finds the beginning of its byte range of a plain FASTA/FASTQ file -- pass the text from one byte
 * before the nominal cut, so that a cut that falls exactly on a record start is found.  FASTQ: a line starting with
 * '@' whose next-but-one line starts with '+' (a quality line may start with '@', but is then followed by a header
 * and a base line); FASTA: a line starting with '>'.  (No counterpart in the reference, whose workers all receive
 * chunks from one reader process: command_line_interface.py:124-130.) */
int kmb_find_record_start(const uint8_t *text, uint64_t n_text, int format, uint64_t *offset);

/* ---- packed transport of host-resident reads -----------------------------------------------------
 * kmb_mapper_map_reads on HOST buffers can encode the bases to 2 bits each on the CPU (all cores, AVX2 when the CPU has
 * it), straight into pinned staging, and send a quarter of the bytes over PCIe.  Option "host_pack": 1 every chunk packed,
 * 0 every chunk as ASCII, 2 hybrid -- a chunk goes as ASCII straight from the caller's pinned buffer whenever the bus is
 * about to run dry (costs no CPU time) and is packed by the cores otherwise, so bases arrive at about the sum of the two
 * rates (config 2, 16 cores: 38.7 GK/s ASCII, 51.2 packed, 66.8 hybrid) --, default -1 = packed for a pageable source with
 * >= 2 threads ("host_threads", default 0 = every CPU of the affinity mask); for a pinned one hybrid when the process has the
 * host to itself ("host_ranks" = 1, set by distributed.init_process_group) and ASCII when several ranks share the host's memory
 * system (the cores' streaming reads slow every rank's DMA down: 8 ranks, 557 ms per step hybrid, 380 ASCII).  Same table as the kernels apply to
 * unpacked input (DNAEncoding as used at util.py:71-75; N -> A of command_line_interface.py:41), same invalid-byte
 * report.  kmb_pack_bases is that encoder on its own: word j of words[] = bases 16j..16j+15, base 16j in the lowest
 * bits, positions past n_bases read as 'A'; words_capacity >= (n_bases + 15) / 16 + 4 (the last 4 are zero padding).
 * Returns KMB_ERR_INVALID_BASE and the offset of the first byte outside ACGTacgt (and N unless
 * KMB_FLAG_NO_N_TO_A) -- the words are still all written.  Host only, no GPU involved. */
int kmb_pack_bases(const uint8_t *bases, uint64_t n_bases, uint32_t flags, int n_threads, uint32_t *words,
                   uint64_t words_capacity, int64_t *first_bad_offset);

/* ---- pinned host memory for the chunk reader (command_line_interface.py:102-111 replacement) -- */
int kmb_host_alloc(void **ptr, size_t bytes);
int kmb_host_free(void *ptr);

/* ---- measurement helpers (bench.py) ---------------------------------------------------------------
 * Host memory read bandwidth in GB/s: n_threads (0 = all) workers sum buf[0, n_bytes) once.  What the packed transport
 * and the DMA engines of all GPUs of a host share: the bound of the end-to-end numbers. */
int kmb_host_read_bandwidth(const void *buf, uint64_t n_bytes, int n_threads, double *gb_per_s);
/* The random-gather micro-roofline of SURVEY.md 8(d).
 * Issues n_loads uniform random loads of load_bytes (8, 16 or 32) from a table of table_bytes,
 * `unroll` independent loads in flight per thread; returns the CUDA-event time in ms. */
int kmb_bench_gather(int device, uint64_t table_bytes, uint64_t n_loads, int load_bytes, int unroll,
                     int threads_per_block, int blocks_per_sm, float *ms);

/* Look-ups that passed the filter and fetched an index sector since the last reset (implies kmb_mapper_sync):
 * the number of random DRAM transactions the probes made. */
int kmb_mapper_candidates(kmb_mapper *mapper, uint64_t *n_candidates);

/* Sum of the device durations (ms) of the mapping kernels launched on this mapper since the last
 * call -- the fused reads kernel / the k-mer kernel only, not the mask or memset launches -- when
 * option "time_kernels" is 1 (CUDA events on the mapper's stream).  Implies a stream synchronize. */
int kmb_mapper_kernel_time(kmb_mapper *mapper, double *ms_total, uint64_t *n_kernels);

/* The same for the apply passes (hit log -> node counts, the `node_counts[...] += 1` of mapper.pyx:68 played one
 * L2-sized window of nodes at a time). */
int kmb_mapper_apply_time(kmb_mapper *mapper, double *ms_total, uint64_t *n_kernels);

/* Tuning knobs (process-wide, read at launch / index-creation time): name in
 * {"map_reads_blocks_per_sm", "map_kmers_blocks_per_sm", "probe_variant", "gathers_in_flight",
 *  "use_filter", "filter_l2_budget_bytes", "sectors_per_100_entries", "l2_persist", "ablate", "policy_filter", "policy_line", "log_max_entries",
 *  "time_kernels",
 *  "l2_fetch_granularity", "bench_grid_blocks", "bench_load_mode", "chunk_bytes", "host_pack", "host_threads", "host_ranks", "filter_probes" (filter bits per key, 0 = by density),
 *  "direct_counts_max_nodes" (count arrays up to this size -- default 4 Mi nodes -- are reduced onto directly: no hit log, no apply pass),
 *  "apply_window_log2" (nodes per apply window; default 0 = auto: 2^23 = 32 MB of counters, 2^24 for count arrays beyond 128 M nodes),
 *  "gz_device_max_member_bytes", "gz_device_batch_bytes", "gz_device_crc", "gz_device_max_mean_member_bytes" (kmb_mapper_map_gz:
 *  files whose members average more text than the last one are left to the host decoders, *resume_offset = 0),
 *  "host_pack" (host chunks of kmb_mapper_map_reads: 1 packed to 2 bits per base by the CPU, 0 ASCII, 2 hybrid -- both pipes side
 *  by side --, -1 default: hybrid for a pinned source, packed for a pageable one), "host_hybrid_backlog_bytes",
 *  "host_pack_streaming" (default 1: the host encoder writes its 2-bit words with non-temporal stores),
 *  "apply_slabs_per_sm" (8: CTAs of the apply pass per SM and node window; 16 and 32 measured slower),
 *  "read_table" (k = 31 reads through the minimizer-bucketed second table, see csrc/kmb_core.cuh: 1 always, 0 never,
 *  default -1 = for every index of at least "read_table_min_entries" live entries when there is room for the second table),
 *  "read_table_min_entries" (8 Mi: auto never builds the table for smaller indexes),
 *  "read_table_buckets_per_100_entries" (150)};
 * read-only: "h2d_bytes" (bytes the mapping calls have copied host -> device so far), "last_reads_kernel" (0 = the
 * key-addressed fused kernel, 1 = the read-path table kernel served the last kmb_mapper_map_reads launch), "bounds_failures" (-1 unless
 * built with -DKMB_BOUNDS_CHECKS). */
int kmb_set_option(const char *name, int64_t value);
int kmb_get_option(const char *name, int64_t *value);
/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
int kmb_launch_count(uint64_t *n_launches);

#ifdef __cplusplus
}
#endif
#endif /* KMER_MAPPER_B200_H */
