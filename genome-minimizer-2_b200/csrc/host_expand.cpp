// host_expand.cpp — part of libgm2.so.  HOST ONLY: the receiving end of the two-bit wire format.
//
// gm2_emit_host / gm2_minimize_host deliver the FASTA image (minimizer_2.py:476-477) into CPU memory.
// The plain way copies the finished image over PCIe (57 GB/s per GPU here).  For an ACGT-only
// reference the GPU can instead ship every kept base as 2 bits (k_emit_packed, k5_emit_packed.cuh)
// and the host expands them into the caller's buffer while the next chunk is in flight: 4x fewer
// PCIe bytes, and the expansion runs at host-memory speed (120 GB/s with 16 threads on the
// measurement box, tools/host_expand_probe.cpp).  The compaction itself — which bases survive, and
// where they go — is done on the GPU; this file only decodes the transport encoding and writes the
// '>' header line and the trailing newline of each record.
#include "host_expand.hpp"

#include <immintrin.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace gm2host {

static const char ACGT[4] = {'A', 'C', 'G', 'T'};

static inline uint8_t base_at(const uint8_t* p, int64_t i) {
    return (uint8_t)ACGT[(p[i >> 2] >> (2 * (i & 3))) & 3];
}

static void expand_scalar(uint8_t* dst, const uint8_t* p, int64_t b0, int64_t n) {
    for (int64_t i = 0; i < n; ++i) dst[i] = base_at(p, b0 + i);
}

// 32 bases (64 bits of the stream, starting at base b0 + 32 k) -> 32 ASCII bytes per step, written with
// non-temporal stores to 32-byte aligned destinations: the image is not read again by this process,
// and regular stores would read every destination line first.
__attribute__((target("avx2")))
static void expand_avx2_aligned(uint8_t* dst, const uint8_t* p, int64_t b0, int64_t ngroups) {
    const __m256i spread = _mm256_setr_epi8(0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3,
                                            4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7);
    const __m256i lut0 = _mm256_setr_epi8('A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T',
                                          'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T');
    const __m256i lut1 = _mm256_setr_epi8('A', 'A', 'A', 'A', 'C', 'C', 'C', 'C', 'G', 'G', 'G', 'G', 'T', 'T', 'T', 'T',
                                          'A', 'A', 'A', 'A', 'C', 'C', 'C', 'C', 'G', 'G', 'G', 'G', 'T', 'T', 'T', 'T');
    const __m256i m0f = _mm256_set1_epi8(0x0f);
    const __m256i k1 = _mm256_set1_epi32(0x0000ff00), k2 = _mm256_set1_epi32(0x00ff0000), k3 = _mm256_set1_epi32((int)0xff000000u);
    const int sh = 2 * (int)(b0 & 3);                      // bit phase of the stream inside its first byte: constant
    const uint8_t* q = p + (b0 >> 2);
    for (int64_t g = 0; g < ngroups; ++g, q += 8, dst += 32) {
        uint64_t x;
        memcpy(&x, q, 8);
        if (sh) x = (x >> sh) | ((uint64_t)q[8] << (64 - sh));
        const __m256i v = _mm256_shuffle_epi8(_mm256_set1_epi64x((long long)x), spread);   // every source byte four times
        const __m256i lo = _mm256_and_si256(v, m0f);
        const __m256i hi = _mm256_and_si256(_mm256_srli_epi16(v, 4), m0f);
        __m256i r = _mm256_blendv_epi8(_mm256_shuffle_epi8(lut0, lo), _mm256_shuffle_epi8(lut1, lo), k1);
        r = _mm256_blendv_epi8(r, _mm256_shuffle_epi8(lut0, hi), k2);
        r = _mm256_blendv_epi8(r, _mm256_shuffle_epi8(lut1, hi), k3);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst), r);
    }
    _mm_sfence();                                          // this thread's non-temporal stores ordered before it signals completion
}

static bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}

void expand_bases(uint8_t* dst, const uint32_t* words, int64_t nbases, bool allow_simd) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(words);
    if (!allow_simd || !have_avx2() || nbases < 96) { expand_scalar(dst, p, 0, nbases); return; }
    int64_t head = (int64_t)((32 - ((uintptr_t)dst & 31)) & 31);
    expand_scalar(dst, p, 0, head);
    const int64_t groups = (nbases - head) / 32;
    expand_avx2_aligned(dst + head, p, head, groups);
    const int64_t done = head + 32 * groups;
    expand_scalar(dst + done, p, done, nbases - done);
}

static int write_header(uint8_t* dst, const char* prefix, int prefix_len, unsigned long long num) {
    memcpy(dst, prefix, (size_t)prefix_len);
    char dig[24]; int nd = 0;
    do { dig[nd++] = (char)('0' + num % 10); num /= 10; } while (num);
    for (int i = 0; i < nd; ++i) dst[prefix_len + i] = (uint8_t)dig[nd - 1 - i];
    dst[prefix_len + nd] = '\n';
    return prefix_len + nd + 1;
}

// One task = a sixteenth of one sample's tiles (finer than a sample so that the tail of a chunk does not
// leave threads idle); the first part also writes the header line, the last one the final newline.
static const int TASKS_PER_SAMPLE = 16;

static void expand_task(const ChunkView& v, int64_t task) {
    const int64_t i = task / TASKS_PER_SAMPLE;
    const int part = (int)(task % TASKS_PER_SAMPLE);
    const int64_t s = v.s0 + i;
    const int64_t roff = v.rec_off[s] - v.rec_off[v.s0];
    uint8_t* rec = v.out + roff;
    const unsigned long long num = (unsigned long long)(v.first_idx + s + 1);
    int nd = 1;
    for (unsigned long long x = num; x >= 10; x /= 10) ++nd;
    const int hl = v.prefix_len + nd + 1;
    if (part == 0) write_header(rec, v.prefix, v.prefix_len, num);
    uint8_t* seq = rec + hl;
    const int64_t len = v.lengths[s];
    const int32_t* toff = v.tile_off + i * v.ntiles;
    const int64_t wbase = (roff >> 4) + i * (int64_t)(v.ntiles + 2);
    const int t0 = (int)((int64_t)v.ntiles * part / TASKS_PER_SAMPLE), t1 = (int)((int64_t)v.ntiles * (part + 1) / TASKS_PER_SAMPLE);
    for (int t = t0; t < t1; ++t) {
        const int64_t a = toff[t], b = t + 1 < v.ntiles ? toff[t + 1] : len;
        if (b > a) expand_bases(seq + a, v.packed + wbase + (a >> 4) + t, b - a, v.simd);
    }
    if (part == TASKS_PER_SAMPLE - 1) seq[len] = '\n';
}

// ---- persistent workers: a chunk is decoded every few milliseconds, thread start-up would show -----
struct Pool::Impl {
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable wake, done;
    const ChunkView* job = nullptr;
    int64_t ntasks = 0;
    std::atomic<int64_t> next{0};
    int active = 0;                 // workers still inside the current job
    uint64_t generation = 0;
    bool stop = false;

    void drain() {
        for (;;) {
            const int64_t t = next.fetch_add(1, std::memory_order_relaxed);
            if (t >= ntasks) break;
            expand_task(*job, t);
        }
    }
    void worker() {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            wake.wait(lk, [&] { return stop || generation != seen; });
            if (stop) return;
            seen = generation;
            lk.unlock();
            drain();
            lk.lock();
            if (--active == 0) done.notify_one();
        }
    }
};

Pool::Pool(int threads) : impl_(new Impl), threads_(threads < 1 ? 1 : threads) {
    for (int t = 1; t < threads_; ++t) impl_->workers.emplace_back([this] { impl_->worker(); });
}

Pool::~Pool() {
    {
        std::lock_guard<std::mutex> lk(impl_->m);
        impl_->stop = true;
    }
    impl_->wake.notify_all();
    for (auto& th : impl_->workers) th.join();
    delete impl_;
}

void Pool::expand_chunk(const ChunkView& v) {
    const int64_t n = (v.s1 - v.s0) * TASKS_PER_SAMPLE;
    if (n <= 0) return;
    Impl& p = *impl_;
    {
        std::lock_guard<std::mutex> lk(p.m);
        p.job = &v; p.ntasks = n; p.next.store(0, std::memory_order_relaxed);
        p.active = (int)p.workers.size();
        ++p.generation;
    }
    p.wake.notify_all();
    p.drain();                                             // the calling thread works too
    std::unique_lock<std::mutex> lk(p.m);
    p.done.wait(lk, [&] { return p.active == 0; });
    p.job = nullptr;
}

void expand_chunk(const ChunkView& v, int threads) {
    Pool pool(threads);
    pool.expand_chunk(v);
}

int local_ranks() {
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) { const int r = atoi(e); if (r > 1) return r; }
    return 1;
}

int default_threads() {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    int t = (int)hw / local_ranks();
    if (t < 1) t = 1;
    if (t > 32) t = 32;
    return t;
}

}  // namespace gm2host
