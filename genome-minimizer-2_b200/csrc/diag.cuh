// diag.cuh — part of libgm2.so (included by gm2.cu; one translation unit).
// Measurement-only kernels: fills, store-pattern models, device-side range hashes.
#pragma once

#include "device_util.cuh"
#include "k4_emit.cuh"      // st256

// ------------------------------------------------------------------------------------------
// diagnostics
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_fill(uint4* __restrict__ dst, int64_t nvec, uint32_t pattern)
{
    const uint4 v = make_uint4(pattern, pattern, pattern, pattern);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) dst[i] = v;
}

// Store-only model of k_emit's write pattern: CTA = (tile, batch of samples), each warp streams
// `chunk` contiguous bytes of record s at offset tile*chunk, records `stride` bytes apart.
__global__ void __launch_bounds__(256)
k_fill_streams(uint8_t* __restrict__ dst, int64_t nrec, int64_t stride, int ntile, int64_t chunk, int batch, int nbatch,
               int order, int vec32)
{
    const int tile = order ? (int)(blockIdx.x % ntile) : (int)(blockIdx.x / nbatch);
    const int b = order ? (int)(blockIdx.x / ntile) : (int)(blockIdx.x - tile * nbatch);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int64_t sb = (int64_t)b * batch, se = sb + batch < nrec ? sb + batch : nrec;
    const uint4 v = make_uint4(0x41414141u, 0x43434343u, 0x47474747u, 0x54545454u);
    for (int64_t s = sb + warp; s < se; s += nwarps) {
        // vec32 bits: 1 = 256-bit stores; bits 8.. = misalignment of the chunk start in bytes (multiple of 32);
        // bits 16.. = fragment length in bytes (0 = none): after every fragment 32 bytes are skipped,
        // modelling a run boundary whose sector is written separately
        const int mis = (vec32 >> 8) & 0xff, frag = vec32 >> 16;
        uint8_t* p = dst + s * stride + (int64_t)tile * chunk + mis;
        const int64_t n = chunk - mis;
        if (frag) {
            for (int64_t f0 = 0; f0 + frag <= n; f0 += frag) {
                for (int64_t o = 16 * lane; o + 16 <= frag - 32; o += 512) *reinterpret_cast<uint4*>(p + f0 + o) = v;
                if ((vec32 & 4) && lane == ((f0 / frag) & 31)) st256(p + f0 + frag - 32, v, v);     // in-stream, one lane
                if ((vec32 & 8) && lane < 2) *reinterpret_cast<uint4*>(p + f0 + frag - 32 + 16 * lane) = v;   // in-stream, two lanes
            }
            if ((vec32 & 2)) {                                  // the skipped sectors, one lane each, afterwards
                for (int64_t f0 = (int64_t)frag * (lane + 1) - 32; f0 + 32 <= n; f0 += (int64_t)frag * 32) st256(p + f0, v, v);
            }
        }
        else if (vec32 & 1) { for (int64_t o = 32 * lane; o + 32 <= n; o += 1024) st256(p + o, v, v); }
        else                { for (int64_t o = 16 * lane; o + 16 <= n; o += 512) *reinterpret_cast<uint4*>(p + o) = v; }
    }
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ void range_hash_block(const uint8_t* __restrict__ buf, int64_t buf_bytes, int64_t o0, int64_t n,
                                                 unsigned long long* __restrict__ out_r)
{
    const int64_t nwords = (n + 7) >> 3;
    const int m = (int)(o0 & 7);
    const uint8_t* abase = buf + (o0 - m);                       // 8-byte aligned (buf is)
    const unsigned long long* w64 = reinterpret_cast<const unsigned long long*>(abase);
    const int64_t abytes = buf_bytes - (o0 - m);                // bytes readable from abase
    const int64_t avail = abytes >> 3;                          // whole aligned words readable
    auto load_word = [&](int64_t k) -> unsigned long long {
        if (k < avail) return w64[k];
        unsigned long long w = 0;                               // partial word at the buffer's end
        for (int b = 0; b < 8; ++b) {
            const int64_t p = 8 * k + b;
            if (p < abytes) w |= (unsigned long long)abase[p] << (8 * b);
        }
        return w;
    };
    unsigned long long acc = 0;
    for (int64_t k = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; k < nwords; k += (int64_t)gridDim.y * blockDim.x) {
        unsigned long long w = load_word(k);
        if (m) w = (w >> (8 * m)) | (load_word(k + 1) << (64 - 8 * m));
        const int64_t valid = n - 8 * k;
        if (valid < 8) w &= (1ull << (8 * valid)) - 1ull;
        acc += mix64((unsigned long long)k * 0x9E3779B97F4A7C15ull + w);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, d);
    __shared__ unsigned long long part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; ++i) t += part[i];
        atomicAdd(out_r, t);
    }
}

__global__ void __launch_bounds__(256)
k_range_hashes(const uint8_t* __restrict__ buf, int64_t buf_bytes, const int64_t* __restrict__ off,
               unsigned long long* __restrict__ out)
{
    const int64_t r = blockIdx.x;
    range_hash_block(buf, buf_bytes, off[r], off[r + 1] - off[r], out + r);
}

// same, ranges given as (begin, end) pairs
__global__ void __launch_bounds__(256)
k_range_hashes_pairs(const uint8_t* __restrict__ buf, int64_t buf_bytes, const int64_t* __restrict__ pairs,
                     unsigned long long* __restrict__ out)
{
    const int64_t r = blockIdx.x;
    range_hash_block(buf, buf_bytes, pairs[2 * r], pairs[2 * r + 1] - pairs[2 * r], out + r);
}


