// host_expand.hpp — part of libgm2.so.  HOST ONLY: decoder of the two-bit wire format (see host_expand.cpp).
#pragma once

#include <cstdint>

namespace gm2host {

// One chunk of samples [s0, s1) as it arrives from the GPU (layout defined in k5_emit_packed.cuh):
//   packed    32-bit words; the kept bases of (sample s0+i, tile t) start at word
//             ((rec_off[s0+i] - rec_off[s0]) >> 4) + i * (ntiles + 2) + (tile_off[i][t] >> 4) + t,
//             base j of the piece in bits [2j, 2j+2) of the little-endian stream, 0..3 = A C G T
//   tile_off  [s1-s0][ntiles] int32: bases kept before tile t, per sample (k_plan's output rows)
//   rec_off   int64 record offsets of the whole plan (indexed by absolute sample), lengths likewise
//   out       destination of record s0 (records are laid out contiguously from there)
struct ChunkView {
    const uint32_t* packed;
    const int32_t* tile_off;
    const int64_t* rec_off;
    const int64_t* lengths;
    uint8_t* out;
    int64_t s0, s1;
    int64_t first_idx;
    int ntiles;
    const char* prefix;      // '>' + id prefix (no terminator needed)
    int prefix_len;
    int simd;                // decoder: 0 portable scalar (tests), 1 best the CPU has, 2 AVX2, 3 AVX-512 VBMI
};

// nbases bases starting at bit 0 of words[0] -> ASCII at dst (any alignment).  May read up to 16 bytes
// past the last word that holds a base.
void expand_bases(uint8_t* dst, const uint32_t* words, int64_t nbases, int simd);
// decoder the CPU supports: 0 scalar, 2 AVX2, 3 AVX-512 VBMI
int simd_level();

// Headers, sequences and trailing newlines of every record of the chunk, decoded by `threads` threads
// (the caller's included) that take part-of-a-sample tasks from a shared counter.  A Pool keeps its
// workers between chunks; the free function makes a temporary one.
class Pool {
public:
    explicit Pool(int threads);
    ~Pool();
    Pool(const Pool&) = delete;
    Pool& operator=(const Pool&) = delete;
    void expand_chunk(const ChunkView& v);      // start + finish
    void start(const ChunkView& v);             // workers begin at once; the view is copied
    void finish();                              // the caller helps, then waits for the workers
    int threads() const { return threads_; }
private:
    struct Impl;
    Impl* impl_;
    int threads_;
};
void expand_chunk(const ChunkView& v, int threads);

// processes sharing this host (LOCAL_WORLD_SIZE, as torchrun sets it; 1 when absent)
int local_ranks();
// GM2_HOST_THREADS, else the CPUs this process may run on / local_ranks(), clamped to [1, 64]
int default_threads();
// `threads` threads fill `host` with non-temporal stores (4 MiB tasks from a shared counter): best GB/s of `reps` passes
double fill_probe(uint8_t* host, int64_t bytes, int threads, int reps);

}  // namespace gm2host
