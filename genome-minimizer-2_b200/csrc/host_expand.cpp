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
#include <sched.h>

#include <atomic>
#include <chrono>
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
}

// AVX-512 VBMI: 64 bases (128 bits of the stream) -> one whole 64-byte cache line per step, one
// non-temporal store each (a full line leaves the core in one piece; two 32-byte halves may not).
// vpermb puts stream bytes 2k, 2k+1, 2k+2 at the bottom of qword k, vpmultishiftqb takes from each
// qword the 8 fields starting at bit (phase + 2 i) — the bit phase of the stream costs nothing —
// and vpshufb maps the two low bits of every byte to its letter.  The partial line in front of the
// first aligned one and the one behind the last are the same decode with a byte-masked load (only the
// stream bytes that exist are touched) and a byte-masked store: a piece costs no scalar work at all.
// No fence here: the caller fences once per task (expand_task), not once per piece — draining the
// write-combining buffers ~130 times per record cost a fifth of the decoder's time.
struct Avx512Tables { alignas(64) uint8_t idx[64], ctl[4][64], lut[64]; };
static const Avx512Tables& avx512_tables() {
    static const Avx512Tables t = [] {
        Avx512Tables x;
        for (int k = 0; k < 8; ++k)
            for (int i = 0; i < 8; ++i) {
                x.idx[8 * k + i] = (uint8_t)(2 * k + (i < 3 ? i : 0));
                for (int ph = 0; ph < 4; ++ph) x.ctl[ph][8 * k + i] = (uint8_t)(2 * ph + 2 * i);
            }
        for (int i = 0; i < 64; ++i) x.lut[i] = (uint8_t)ACGT[i & 3];
        return x;
    }();
    return t;
}

// n bases starting at base b of the stream p (n <= 64) -> d, byte-masked on both sides: touches only the
// stream bytes that hold those bases and only the n destination bytes
__attribute__((target("avx512f,avx512bw,avx512vbmi")))
static inline void avx512_partial(const Avx512Tables& T, uint8_t* d, const uint8_t* p, int64_t b, int n) {
    const int nb = (int)(((b & 3) + n + 3) >> 2);                          // stream bytes holding those bases (<= 17)
    const __m512i src = _mm512_maskz_loadu_epi8(((__mmask64)1 << nb) - 1, p + (b >> 2));
    const __m512i rep = _mm512_permutexvar_epi8(_mm512_load_si512(T.idx), src);
    const __m512i code = _mm512_and_si512(_mm512_multishift_epi64_epi8(_mm512_load_si512(T.ctl[b & 3]), rep), _mm512_set1_epi8(3));
    _mm512_mask_storeu_epi8(d, n >= 64 ? ~(__mmask64)0 : (((__mmask64)1 << n) - 1),
                            _mm512_shuffle_epi8(_mm512_load_si512(T.lut), code));
}

__attribute__((target("avx512f,avx512bw,avx512vbmi")))
static void expand_avx512(uint8_t* dst, const uint8_t* p, int64_t nbases) {
    const Avx512Tables& T = avx512_tables();
    const __m512i idx = _mm512_load_si512(T.idx), lut = _mm512_load_si512(T.lut), m3 = _mm512_set1_epi8(3);
    int64_t head = (int64_t)((64 - ((uintptr_t)dst & 63)) & 63);
    if (head > nbases) head = nbases;
    if (head) avx512_partial(T, dst, p, 0, (int)head);
    const int64_t groups = (nbases - head) / 64;
    const __m512i ctl = _mm512_load_si512(T.ctl[head & 3]);
    const uint8_t* q = p + (head >> 2);
    uint8_t* d = dst + head;
    // every full group reads 32 stream bytes from its first one (17 used); the last group reads exactly what exists
    for (int64_t g = 0; g + 1 < groups; ++g, q += 16, d += 64) {
        // the stream was written by the GPU's DMA: it comes from DRAM, and a piece (a few KB) is too short for the
        // hardware prefetcher to get ahead, so ask for it ~1 KB (64 steps) ahead; past the piece it is the next one's
        _mm_prefetch(reinterpret_cast<const char*>(q) + 1024, _MM_HINT_T0);
        const __m512i src = _mm512_castsi256_si512(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(q)));
        const __m512i rep = _mm512_permutexvar_epi8(idx, src);
        const __m512i code = _mm512_and_si512(_mm512_multishift_epi64_epi8(ctl, rep), m3);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(d), _mm512_shuffle_epi8(lut, code));
    }
    int64_t done = head + 64 * (groups > 0 ? groups - 1 : 0);
    if (groups > 0) {                                                       // last whole line: exact load, still one NT store
        const int nb = (int)(((done & 3) + 64 + 3) >> 2);
        const __m512i src = _mm512_maskz_loadu_epi8(((__mmask64)1 << nb) - 1, p + (done >> 2));
        const __m512i rep = _mm512_permutexvar_epi8(idx, src);
        const __m512i code = _mm512_and_si512(_mm512_multishift_epi64_epi8(ctl, rep), m3);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + done), _mm512_shuffle_epi8(lut, code));
        done += 64;
    }
    if (done < nbases) avx512_partial(T, dst + done, p, done, (int)(nbases - done));
}

static bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}
static bool have_avx512() {
    static const bool v = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                          __builtin_cpu_supports("avx512vbmi");
    return v;
}

int simd_level() { return have_avx512() ? 3 : have_avx2() ? 2 : 0; }

void expand_bases(uint8_t* dst, const uint32_t* words, int64_t nbases, int simd) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(words);
    if (simd == 1) simd = simd_level();
    if (simd == 3 && !have_avx512()) simd = 2;
    if (simd == 2 && !have_avx2()) simd = 0;
    if (simd == 3) { if (nbases > 0) expand_avx512(dst, p, nbases); return; }
    if (simd < 2 || nbases < 96) { expand_scalar(dst, p, 0, nbases); return; }
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

// One task = 1/48 of one sample's tiles (a few tiles, ~55 KB of output): a 16 MB chunk is ~300 tasks for 16-32
// threads, so the tail of a chunk leaves little idle time (measured: 16 tasks per sample 166.5 Gbp/s end to end,
// 48 tasks 172.3 on the same host); the first part also writes the header line, the last one the final newline.
#ifndef GM2_TASKS_PER_SAMPLE
#define GM2_TASKS_PER_SAMPLE 48
#endif
static const int TASKS_PER_SAMPLE = GM2_TASKS_PER_SAMPLE;

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
        if (t + 1 < t1) {                                  // the next piece's stream starts somewhere else: ask for it now
            const char* nx = reinterpret_cast<const char*>(v.packed + wbase + (b >> 4) + t + 1);
            _mm_prefetch(nx, _MM_HINT_T0); _mm_prefetch(nx + 64, _MM_HINT_T0); _mm_prefetch(nx + 128, _MM_HINT_T0);
        }
        if (b > a) expand_bases(seq + a, v.packed + wbase + (a >> 4) + t, b - a, v.simd);
    }
    if (part == TASKS_PER_SAMPLE - 1) seq[len] = '\n';
    _mm_sfence();                                          // this thread's non-temporal stores ordered before it signals completion
}

// ---- persistent workers: a chunk is decoded every ~100 microseconds, so workers first SPIN on the job
// generation for a short while (a futex wake-up costs 10-30 us per thread) and only then sleep ----------
struct Pool::Impl {
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable wake, done;
    ChunkView job;
    int64_t ntasks = 0;
    std::atomic<int64_t> next{0};
    std::atomic<int> active{0};             // workers still inside the current job
    std::atomic<uint64_t> generation{0};
    std::atomic<bool> stop{false};
    int sleepers = 0;                       // workers blocked on `wake` (under m)
    bool running = false;

    static constexpr int SPIN = 20000;      // ~100-200 us of pause instructions

    void drain() {
        for (;;) {
            const int64_t t = next.fetch_add(1, std::memory_order_relaxed);
            if (t >= ntasks) break;
            expand_task(job, t);
        }
    }
    void worker() {
        uint64_t seen = 0;
        for (;;) {
            int spins = 0;
            while (generation.load(std::memory_order_acquire) == seen && !stop.load(std::memory_order_relaxed)) {
                if (++spins < SPIN) { _mm_pause(); continue; }
                std::unique_lock<std::mutex> lk(m);
                ++sleepers;
                wake.wait(lk, [&] { return stop.load() || generation.load(std::memory_order_acquire) != seen; });
                --sleepers;
            }
            if (stop.load(std::memory_order_relaxed)) return;
            seen = generation.load(std::memory_order_acquire);
            drain();
            if (active.fetch_sub(1, std::memory_order_acq_rel) == 1) {
                std::lock_guard<std::mutex> lk(m);
                done.notify_one();
            }
        }
    }
};

// CPUs this process may use, cut into LOCAL_WORLD_SIZE equal slices; this rank's slice (LOCAL_RANK).
// Workers are pinned round-robin inside it (GM2_HOST_PIN=0: leave them to the scheduler): ranks that
// share a host then do not migrate onto each other's cores in the middle of a chunk.
static std::vector<int> my_cpus() {
    std::vector<int> all;
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0)
        for (int c = 0; c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &set)) all.push_back(c);
    const int R = local_ranks();
    int r = 0;
    if (const char* e = getenv("LOCAL_RANK")) r = atoi(e);
    if (R <= 1 || r < 0 || r >= R || (int)all.size() < R) return all;
    const size_t a = all.size() * (size_t)r / (size_t)R, b = all.size() * (size_t)(r + 1) / (size_t)R;
    return std::vector<int>(all.begin() + (long)a, all.begin() + (long)b);
}

static bool pin_enabled() {
    const char* e = getenv("GM2_HOST_PIN");
    return !(e && e[0] == '0');
}

Pool::Pool(int threads) : impl_(new Impl), threads_(threads < 1 ? 1 : threads) {
    const std::vector<int> cpus = pin_enabled() ? my_cpus() : std::vector<int>();
    for (int t = 1; t < threads_; ++t) {
        impl_->workers.emplace_back([this] { impl_->worker(); });
        if (!cpus.empty()) {
            cpu_set_t one;
            CPU_ZERO(&one);
            CPU_SET(cpus[(size_t)t % cpus.size()], &one);
            pthread_setaffinity_np(impl_->workers.back().native_handle(), sizeof(one), &one);   // best effort
        }
    }
}

Pool::~Pool() {
    {
        std::lock_guard<std::mutex> lk(impl_->m);
        impl_->stop.store(true);
    }
    impl_->wake.notify_all();
    for (auto& th : impl_->workers) th.join();
    delete impl_;
}

// Publish a chunk and return at once: the workers start decoding while the caller does something else
// (enqueueing the next chunk's GPU work); finish() makes the caller take tasks too and waits for the rest.
void Pool::start(const ChunkView& v) {
    Impl& p = *impl_;
    p.job = v;
    p.ntasks = (v.s1 - v.s0) * TASKS_PER_SAMPLE;
    p.next.store(0, std::memory_order_relaxed);
    p.active.store((int)p.workers.size(), std::memory_order_relaxed);
    p.running = true;
    p.generation.fetch_add(1, std::memory_order_release);
    bool wake_needed;
    {
        std::lock_guard<std::mutex> lk(p.m);
        wake_needed = p.sleepers > 0;
    }
    if (wake_needed) p.wake.notify_all();
}

void Pool::finish() {
    Impl& p = *impl_;
    if (!p.running) return;
    p.drain();                                             // the calling thread works too
    int spins = 0;
    while (p.active.load(std::memory_order_acquire) != 0) {
        if (++spins < Impl::SPIN) { _mm_pause(); continue; }
        std::unique_lock<std::mutex> lk(p.m);
        p.done.wait(lk, [&] { return p.active.load(std::memory_order_acquire) == 0; });
    }
    p.running = false;
}

void Pool::expand_chunk(const ChunkView& v) {
    if (v.s1 <= v.s0) return;
    start(v);
    finish();
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
    if (const char* e = getenv("GM2_HOST_THREADS")) { const int t = atoi(e); if (t > 0) return t > 256 ? 256 : t; }
    // the CPUs this process is allowed on (cgroup / taskset aware), shared equally among the local ranks
    int n = 0;
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
    if (n <= 0) n = (int)std::thread::hardware_concurrency();
    if (n <= 0) n = 1;
    int t = n / local_ranks();
    if (t < 1) t = 1;
    if (t > 64) t = 64;
    return t;
}

// ---- host write ceiling: the expansion's store pattern without the decode ------------------------
__attribute__((target("avx2")))
static void fill_nt(uint8_t* dst, int64_t n) {
    const __m256i v = _mm256_set1_epi8('A');
    int64_t i = 0;
    for (; i < n && ((uintptr_t)(dst + i) & 31); ++i) dst[i] = 'A';
    for (; i + 32 <= n; i += 32) _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), v);
    for (; i < n; ++i) dst[i] = 'A';
    _mm_sfence();
}
__attribute__((target("avx512f")))
static void fill_nt512(uint8_t* dst, int64_t n) {
    const __m512i v = _mm512_set1_epi8('A');
    int64_t i = 0;
    for (; i < n && ((uintptr_t)(dst + i) & 63); ++i) dst[i] = 'A';
    for (; i + 64 <= n; i += 64) _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + i), v);
    for (; i < n; ++i) dst[i] = 'A';
    _mm_sfence();
}

double fill_probe(uint8_t* host, int64_t bytes, int threads, int reps) {
    if (threads < 1) threads = 1;
    const std::vector<int> cpus = pin_enabled() ? my_cpus() : std::vector<int>();
    const int64_t piece = (int64_t)4 << 20;                      // tasks of 4 MiB from a shared counter, as the expansion takes tasks
    const int64_t ntask = (bytes + piece - 1) / piece;
    double best = 0.0;
    for (int rep = 0; rep < reps; ++rep) {
        std::atomic<int64_t> next{0};
        auto work = [&] {
            for (;;) {
                const int64_t t = next.fetch_add(1, std::memory_order_relaxed);
                if (t >= ntask) break;
                const int64_t a = t * piece, n = std::min(piece, bytes - a);
                if (have_avx512()) fill_nt512(host + a, n);
                else if (have_avx2()) fill_nt(host + a, n);
                else memset(host + a, 'A', (size_t)n);
            }
        };
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 1; t < threads; ++t) {
            th.emplace_back(work);
            if (!cpus.empty()) {
                cpu_set_t one;
                CPU_ZERO(&one);
                CPU_SET(cpus[(size_t)t % cpus.size()], &one);
                pthread_setaffinity_np(th.back().native_handle(), sizeof(one), &one);
            }
        }
        work();
        for (auto& x : th) x.join();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (dt > 0) best = std::max(best, (double)bytes / dt / 1e9);
    }
    return best;
}

}  // namespace gm2host
