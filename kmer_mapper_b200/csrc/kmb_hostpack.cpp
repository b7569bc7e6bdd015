// kmb_hostpack.cpp -- host half of the packed transport of host-resident reads.
//
// A base is one byte in the caller's buffer and two bits in a k-mer.  PCIe (about 50 GB/s on the B200 boxes) is
// the narrowest pipe between a host buffer and the mapping kernel (which consumes > 100 G bases/s), so host
// input is encoded to 2 bits per base on the CPU, by every core, straight into the pinned staging buffer the
// DMA engine reads: 4x fewer bytes cross the bus, and a pageable caller buffer is read exactly once.  The
// arithmetic is kmb_encode16 (kmb_core.cuh), i.e. the same ASCII -> A,C,G,T = 0..3 table the fused kernel
// applies to unpacked input (reference: bionumpy's DNAEncoding as used at util.py:71-75, plus N -> A of
// command_line_interface.py:132), so both transports give the same k-mers and flag the same bytes.
//
// Also here: the persistent worker pool (kmb_host_parallel).
#include <sched.h>
#include <stdint.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "kmb_core.cuh"
#include "kmb_host.h"

// ------------------------------------------------------------------------------------------------
// worker pool
// ------------------------------------------------------------------------------------------------
int kmb_host_cpus() {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        int n = CPU_COUNT(&set);
        if (n > 0) return n;
    }
    unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 1;
}

namespace {

struct Job {
    void (*fn)(void *, int) = nullptr;
    void *ctx = nullptr;
    int n_parts = 0;
    int n_workers = 0;  // pool workers allowed to help (the caller always does)
    std::atomic<int> next{0};
    std::atomic<int> done{0};
};

inline void cpu_relax() {
#if defined(__x86_64__)
    _mm_pause();
#else
    std::this_thread::yield();
#endif
}

void run_parts(Job &job) {
    for (;;) {
        int p = job.next.fetch_add(1, std::memory_order_relaxed);
        if (p >= job.n_parts) return;
        job.fn(job.ctx, p);
        job.done.fetch_add(1, std::memory_order_release);
    }
}

class Pool {
  public:
    explicit Pool(int n_workers) : pid_(getpid()) {
        for (int i = 0; i < n_workers; i++) threads_.emplace_back([this, i] { worker(i); });
        for (auto &t : threads_) t.detach();  // they live as long as the process
    }
    pid_t pid() const { return pid_; }
    int workers() const { return (int)threads_.size(); }

    void run(int n_threads, int n_parts, void (*fn)(void *, int), void *ctx) {
        std::lock_guard<std::mutex> api(api_mu_);
        auto job = std::make_shared<Job>();
        job->fn = fn;
        job->ctx = ctx;
        job->n_parts = n_parts;
        job->n_workers = std::min(workers(), std::max(0, std::min(n_threads, n_parts) - 1));
        if (job->n_workers > 0) {
            {
                std::lock_guard<std::mutex> l(mu_);
                job_ = job;
                generation_.fetch_add(1, std::memory_order_release);
            }
            cv_.notify_all();
        }
        run_parts(*job);
        for (unsigned spins = 0; job->done.load(std::memory_order_acquire) < n_parts; spins++) {
            if (spins < 4096) cpu_relax();
            else std::this_thread::yield();
        }
    }

  private:
    void worker(int index) {
        uint64_t seen = 0;
        for (;;) {
            // a short spin keeps the wake-up latency of back-to-back chunks low; then sleep
            for (int spins = 0; spins < 20000 && generation_.load(std::memory_order_acquire) == seen; spins++) cpu_relax();
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return generation_.load(std::memory_order_acquire) != seen; });
                seen = generation_.load(std::memory_order_acquire);
                job = job_;
            }
            if (job && index < job->n_workers) run_parts(*job);
        }
    }

    pid_t pid_;
    std::vector<std::thread> threads_;
    std::mutex api_mu_, mu_;
    std::condition_variable cv_;
    std::atomic<uint64_t> generation_{0};
    std::shared_ptr<Job> job_;
};

std::mutex g_pool_mu;
Pool *g_pool = nullptr;

Pool *pool() {
    std::lock_guard<std::mutex> l(g_pool_mu);
    // after fork() the child has the object but none of its threads: build a new pool (the old one is leaked)
    if (!g_pool || g_pool->pid() != getpid()) g_pool = new Pool(std::max(0, kmb_host_cpus() - 1));
    return g_pool;
}

}  // namespace

void kmb_host_parallel(int n_threads, int n_parts, void (*fn)(void *, int), void *ctx) {
    if (n_parts <= 0) return;
    if (n_threads <= 0) n_threads = kmb_host_cpus();
    if (n_threads == 1 || n_parts == 1) {
        for (int p = 0; p < n_parts; p++) fn(ctx, p);
        return;
    }
    pool()->run(n_threads, n_parts, fn, ctx);
}

// ------------------------------------------------------------------------------------------------
// packing
// ------------------------------------------------------------------------------------------------
namespace {

const uint64_t NO_BAD = ~0ull;

// words [w_lo, w_hi) from whole 16-base groups; portable SWAR (the device routine, compiled for the host)
std::atomic<bool> g_pack_streaming{true};

uint64_t pack_words_swar(const uint8_t *bases, uint64_t w_lo, uint64_t w_hi, bool n_to_a, uint32_t *words) {
    uint64_t bad = NO_BAD;
    for (uint64_t j = w_lo; j < w_hi; j++) {
        uint32_t w[4];
        memcpy(w, bases + 16 * j, 16);
        uint32_t inv;
        words[j] = kmb_encode16(w[0], w[1], w[2], w[3], n_to_a, inv);
        if (inv && bad == NO_BAD) bad = 16 * j + (uint64_t)__builtin_ctz(inv);
    }
    return bad;
}

#if defined(__x86_64__)
// 64 bases per iteration.  Validity by two nibble look-ups (a byte is fine iff the classes of its low and high
// nibble intersect: x1/x3/x7 with 4x/6x = ACG acg, x4 with 5x/7x = T t, xE with 4x = N), accumulated over a block
// and checked once per block; a block that fails is rescanned by the scalar routine, which finds the byte.
// Codes: ((c >> 1) & 3) ^ (((c >> 1) & 3) >> 1), N -> 0; then two multiply-adds squeeze 4 codes into a byte and a
// shuffle + lane permute collect the bytes.
__attribute__((target("avx2"))) uint64_t pack_words_avx2(const uint8_t *bases, uint64_t w_lo, uint64_t w_hi, bool n_to_a,
                                                         uint32_t *words) {
    uint64_t bad = NO_BAD;
    const __m256i low4 = _mm256_set1_epi8(0x0F);
    const char n_class = n_to_a ? 4 : 0;
    const __m256i lut_lo = _mm256_setr_epi8(0, 1, 0, 1, 2, 0, 0, 1, 0, 0, 0, 0, 0, 0, 4, 0, 0, 1, 0, 1, 2, 0, 0, 1, 0, 0, 0, 0, 0, 0, 4, 0);
    const __m256i lut_hi = _mm256_setr_epi8(0, 0, 0, 0, (char)(1 | n_class), 2, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0,
                                            0, 0, 0, 0, (char)(1 | n_class), 2, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i cN = _mm256_set1_epi8('N');
    const __m256i three = _mm256_set1_epi8(3), one = _mm256_set1_epi8(1);
    const __m256i mul1 = _mm256_set1_epi16(0x0401);      // b0 + 4 b1
    const __m256i mul2 = _mm256_set1_epi32(0x00100001);  // w0 + 16 w1
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            -1, -1, -1, -1, 0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1);
    // Streaming stores (kmb_host_pack_streaming): the packed words are written once and read next by the DMA engine, so
    // they need not pass through the caches -- and a write-allocating store first READS the line it is about to
    // overwrite, a quarter of a byte of host DRAM traffic per base on top of the byte read and the quarter written.
    const bool nt = g_pack_streaming.load(std::memory_order_relaxed) && (reinterpret_cast<uintptr_t>(words) & 15u) == 0;
    uint64_t j = w_lo;
    while (j < w_hi && (j & 3)) {  // leading words up to a multiple of four
        uint64_t b = pack_words_swar(bases, j, j + 1, n_to_a, words);
        if (b < bad) bad = b;
        j++;
    }
    const uint64_t BLOCK = 64;  // words per validity check (1 KB of bases)
    while (j + 4 <= w_hi) {
        const uint64_t j_end = std::min(j + BLOCK, w_hi - ((w_hi - j) & 3));
        const uint64_t j0 = j;
        __m256i all_ok = _mm256_set1_epi8(-1);
        for (; j < j_end; j += 4) {
            __m256i packed[2];
#pragma GCC unroll 2
            for (int h = 0; h < 2; h++) {
                const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(bases + 16 * j + 32 * h));
                const __m256i cls = _mm256_and_si256(_mm256_shuffle_epi8(lut_lo, _mm256_and_si256(v, low4)),
                                                     _mm256_shuffle_epi8(lut_hi, _mm256_and_si256(_mm256_srli_epi16(v, 4), low4)));
                all_ok = _mm256_and_si256(all_ok, _mm256_cmpgt_epi8(cls, _mm256_setzero_si256()));
                __m256i x = _mm256_and_si256(_mm256_srli_epi16(v, 1), three);
                x = _mm256_xor_si256(x, _mm256_and_si256(_mm256_srli_epi16(x, 1), one));
                if (n_to_a) x = _mm256_andnot_si256(_mm256_cmpeq_epi8(v, cN), x);
                const __m256i t = _mm256_madd_epi16(_mm256_maddubs_epi16(x, mul1), mul2);
                packed[h] = _mm256_shuffle_epi8(t, gather);  // dword 0 = bases 0-15, dword 5 = bases 16-31
            }
            // [a0 . . . . a1 . .] and [b0 . . . . b1 . .] -> a0 a1 b0 b1
            const __m256i both = _mm256_or_si256(packed[0], _mm256_slli_si256(packed[1], 8));
            const __m256i q = _mm256_permutevar8x32_epi32(both, _mm256_setr_epi32(0, 5, 2, 7, 0, 0, 0, 0));
            if (nt) _mm_stream_si128(reinterpret_cast<__m128i *>(words + j), _mm256_castsi256_si128(q));
            else _mm_storeu_si128(reinterpret_cast<__m128i *>(words + j), _mm256_castsi256_si128(q));
        }
        if ((uint32_t)_mm256_movemask_epi8(all_ok) != 0xFFFFFFFFu && bad == NO_BAD) {
            uint32_t scratch[BLOCK];
            const uint64_t b = pack_words_swar(bases + 16 * j0, 0, j - j0, n_to_a, scratch);
            if (b != NO_BAD) bad = 16 * j0 + b;
        }
    }
    if (j < w_hi) {
        uint64_t b = pack_words_swar(bases, j, w_hi, n_to_a, words);
        if (b < bad) bad = b;
    }
    if (nt) _mm_sfence();  // streaming stores are weakly ordered: make them visible before this part reports done
    return bad;
}
#endif

typedef uint64_t (*PackFn)(const uint8_t *, uint64_t, uint64_t, bool, uint32_t *);
PackFn pick_pack() {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2")) return pack_words_avx2;
#endif
    return pack_words_swar;
}

struct PackCtx {
    const uint8_t *bases;
    uint64_t n_full_words;  // words made of 16 real bases
    uint64_t per_part;      // words per part (even)
    bool n_to_a;
    uint32_t *words;
    PackFn fn;
    std::atomic<uint64_t> bad;
};

void pack_part(void *p, int part) {
    PackCtx *c = static_cast<PackCtx *>(p);
    const uint64_t lo = (uint64_t)part * c->per_part;
    const uint64_t hi = std::min(lo + c->per_part, c->n_full_words);
    if (lo >= hi) return;
    const uint64_t b = c->fn(c->bases, lo, hi, c->n_to_a, c->words);
    if (b != NO_BAD) {
        uint64_t cur = c->bad.load(std::memory_order_relaxed);
        while (b < cur && !c->bad.compare_exchange_weak(cur, b, std::memory_order_relaxed)) {
        }
    }
}

}  // namespace

void kmb_host_pack_streaming(bool on) { g_pack_streaming.store(on, std::memory_order_relaxed); }

uint64_t kmb_host_pack(const uint8_t *bases, uint64_t n_bases, bool n_to_a, int n_threads, uint32_t *words) {
    const uint64_t n_full = n_bases / 16;
    PackCtx c;
    c.bases = bases;
    c.n_full_words = n_full;
    c.n_to_a = n_to_a;
    c.words = words;
    c.fn = pick_pack();
    c.bad.store(NO_BAD);
    if (n_threads <= 0) n_threads = kmb_host_cpus();
    // parts of >= 256 KB of bases, a few per thread so that a late starter does not hold the others up
    const uint64_t min_words = 1ull << 14;
    int n_parts = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)n_threads * 4, n_full / min_words));
    c.per_part = (((n_full + (uint64_t)n_parts - 1) / (uint64_t)n_parts) + 1) & ~1ull;
    if (n_full) kmb_host_parallel(n_threads, n_parts, pack_part, &c);
    uint64_t bad = c.bad.load();
    // the ragged last word: bytes past the end read as 'A'
    uint64_t w = n_full;
    if (n_bases % 16) {
        uint8_t tail[16];
        memset(tail, 'A', sizeof(tail));
        memcpy(tail, bases + 16 * n_full, (size_t)(n_bases % 16));
        uint32_t t[4], inv;
        memcpy(t, tail, 16);
        words[w++] = kmb_encode16(t[0], t[1], t[2], t[3], n_to_a, inv);
        if (inv) bad = std::min(bad, 16 * n_full + (uint64_t)__builtin_ctz(inv));
    }
    for (int i = 0; i < KMB_PACK_PAD_WORDS; i++) words[w++] = 0;
    return bad;
}

namespace {
struct RelCtx {
    const int64_t *offsets;
    uint64_t n, per_part;
    int64_t base;
    uint32_t *out;
};
void rel_part(void *p, int part) {
    RelCtx *c = static_cast<RelCtx *>(p);
    const uint64_t lo = (uint64_t)part * c->per_part, hi = std::min(lo + c->per_part, c->n);
    for (uint64_t i = lo; i < hi; i++) c->out[i] = (uint32_t)(c->offsets[i] - c->base);
}
}  // namespace

void kmb_host_rel_offsets(const int64_t *offsets, uint64_t n, int64_t base, int n_threads, uint32_t *out) {
    if (n_threads <= 0) n_threads = kmb_host_cpus();
    RelCtx c = {offsets, n, 0, base, out};
    int n_parts = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)n_threads, n >> 15));
    c.per_part = (n + (uint64_t)n_parts - 1) / (uint64_t)n_parts;
    kmb_host_parallel(n_threads, n_parts, rel_part, &c);
}
