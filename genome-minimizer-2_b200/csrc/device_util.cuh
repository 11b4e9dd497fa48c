// device_util.cuh — part of libgm2.so (included by gm2.cu; one translation unit).
// Warp scans / reductions, decimal digit helpers shared by the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL_MASK, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
    return v;
}
__device__ __forceinline__ long long warp_incl_scan64(long long v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        long long t = __shfl_up_sync(FULL_MASK, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__constant__ unsigned long long c_pow10[20] = {
    1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull,
    1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull,
    100000000000000ull, 1000000000000000ull, 10000000000000000ull, 100000000000000000ull,
    1000000000000000000ull, 10000000000000000000ull};

__device__ __forceinline__ int ndigits_u64(unsigned long long v) {
    int n = 1;
    while (n < 20 && v >= c_pow10[n]) ++n;
    return n;
}

