// Probe (not part of the product): how fast can the host turn a 2-bit packed base stream into ASCII?
// Decides whether shipping the image over PCIe in 2-bit form and expanding it on the host could beat
// the plain pinned D2H copy (57 GB/s on this pool's boxes).  Build + run:
//   g++ -O3 -mavx2 -pthread tools/host_expand_probe.cpp -o /tmp/expand_probe && /tmp/expand_probe
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static uint32_t LUT[256];

static void expand_scalar(const uint8_t* in, uint8_t* out, size_t nbytes_in) {
    uint32_t* o = reinterpret_cast<uint32_t*>(out);
    for (size_t i = 0; i < nbytes_in; ++i) o[i] = LUT[in[i]];
}

// 8 packed bytes -> 32 ASCII bytes per step; non-temporal stores (the image is not re-read here)
__attribute__((target("avx2"))) static void expand_avx2(const uint8_t* in, uint8_t* out, size_t nbytes_in) {
    const __m256i spread = _mm256_setr_epi8(0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3,
                                            4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7);
    const __m256i lut0 = _mm256_setr_epi8('A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T',
                                          'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T', 'A', 'C', 'G', 'T');
    const __m256i lut1 = _mm256_setr_epi8('A', 'A', 'A', 'A', 'C', 'C', 'C', 'C', 'G', 'G', 'G', 'G', 'T', 'T', 'T', 'T',
                                          'A', 'A', 'A', 'A', 'C', 'C', 'C', 'C', 'G', 'G', 'G', 'G', 'T', 'T', 'T', 'T');
    const __m256i m0f = _mm256_set1_epi8(0x0f);
    const __m256i sel_k1 = _mm256_set1_epi32(0x0000ff00), sel_k2 = _mm256_set1_epi32(0x00ff0000),
                  sel_k3 = _mm256_set1_epi32((int)0xff000000u);
    size_t i = 0;
    for (; i + 8 <= nbytes_in; i += 8) {
        const __m128i q = _mm_loadl_epi64(reinterpret_cast<const __m128i*>(in + i));
        const __m256i x = _mm256_shuffle_epi8(_mm256_broadcastsi128_si256(q), spread);   // each source byte 4x (both lanes hold bytes 0..7)
        const __m256i lo = _mm256_and_si256(x, m0f);
        const __m256i hi = _mm256_and_si256(_mm256_srli_epi16(x, 4), m0f);
        const __m256i a0 = _mm256_shuffle_epi8(lut0, lo), a1 = _mm256_shuffle_epi8(lut1, lo);
        const __m256i a2 = _mm256_shuffle_epi8(lut0, hi), a3 = _mm256_shuffle_epi8(lut1, hi);
        __m256i r = _mm256_blendv_epi8(a0, a1, sel_k1);
        r = _mm256_blendv_epi8(r, a2, sel_k2);
        r = _mm256_blendv_epi8(r, a3, sel_k3);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(out + 4 * i), r);
    }
    for (; i < nbytes_in; ++i) reinterpret_cast<uint32_t*>(out)[i] = LUT[in[i]];
    _mm_sfence();
}

int main(int argc, char** argv) {
    const size_t out_bytes = (argc > 1 ? strtoull(argv[1], nullptr, 10) : 8ull) << 30;
    const size_t in_bytes = out_bytes / 4;
    for (int b = 0; b < 256; ++b) {
        const char* acgt = "ACGT";
        LUT[b] = (uint32_t)acgt[b & 3] | (uint32_t)acgt[(b >> 2) & 3] << 8 | (uint32_t)acgt[(b >> 4) & 3] << 16 | (uint32_t)acgt[(b >> 6) & 3] << 24;
    }
    uint8_t* in = static_cast<uint8_t*>(aligned_alloc(4096, in_bytes));
    uint8_t* out = static_cast<uint8_t*>(aligned_alloc(4096, out_bytes));
    for (size_t i = 0; i < in_bytes; ++i) in[i] = (uint8_t)(i * 2654435761u >> 13);
    memset(out, 0, out_bytes);                                  // fault the pages in before timing
    const unsigned hw = std::thread::hardware_concurrency();
    printf("host threads available: %u, output %zu GiB\n", hw, out_bytes >> 30);
    // correctness of the vector form
    {
        uint8_t* a = static_cast<uint8_t*>(aligned_alloc(64, 4096));
        uint8_t* b = static_cast<uint8_t*>(aligned_alloc(64, 4096));
        expand_scalar(in, a, 1024); expand_avx2(in, b, 1024);
        printf("avx2 == scalar: %s\n", memcmp(a, b, 4096) == 0 ? "yes" : "NO");
        free(a); free(b);
    }
    for (int mode = 0; mode < 3; ++mode) {
        for (unsigned T : {1u, 4u, 8u, 16u, 32u}) {
            if (T > hw && T != 1) continue;
            double best = 0;
            for (int rep = 0; rep < 3; ++rep) {
                auto t0 = std::chrono::steady_clock::now();
                std::vector<std::thread> th;
                const size_t per = (in_bytes / T) & ~(size_t)63;
                for (unsigned t = 0; t < T; ++t) {
                    const size_t a = t * per, n = (t + 1 == T) ? in_bytes - a : per;
                    th.emplace_back([=] {
                        if (mode == 0) expand_scalar(in + a, out + 4 * a, n);
                        else if (mode == 1) expand_avx2(in + a, out + 4 * a, n);
                        else memset(out + 4 * a, 'A', 4 * n);
                    });
                }
                for (auto& x : th) x.join();
                const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                best = std::max(best, (double)out_bytes / dt / 1e9);
            }
            printf("%-12s threads %2u: %6.1f GB/s of ASCII out\n", mode == 0 ? "scalar LUT" : mode == 1 ? "avx2 + NT" : "memset", T, best);
        }
    }
    return 0;
}
