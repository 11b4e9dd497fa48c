// k3_scan.cuh — part of libgm2.so (included by gm2.cu; one translation unit).
// K3b: across-sample exclusive scan of record sizes, decoupled look-back (k_scan_records).
#pragma once

#include "device_util.cuh"

// ------------------------------------------------------------------------------------------
// K3b  across-sample exclusive scan of record sizes (int64): single pass, chained scan
//   with decoupled look-back.  One 64-bit descriptor per tile = {2-bit status, 62-bit value},
//   tile ids handed out by an atomic ticket so every predecessor is already running.
// ------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS   8
#define SCAN_TILE    (SCAN_THREADS * SCAN_ITEMS)
#define ST_INVALID   0ull
#define ST_AGG       1ull
#define ST_PREFIX    2ull
#define ST_SHIFT     62
#define ST_VALMASK   ((1ull << ST_SHIFT) - 1ull)

__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_records(const int64_t* __restrict__ in, int64_t* __restrict__ out /* n+1 */, int64_t n,
               unsigned long long* desc, unsigned int* ticket)
{
    __shared__ unsigned int s_tile;
    __shared__ long long s_warp[SCAN_THREADS / 32];
    __shared__ long long s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;

    long long v[SCAN_ITEMS];
    long long tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const int64_t k = base + i;
        v[i] = k < n ? in[k] : 0;
        tsum += v[i];
    }
    const long long wincl = warp_incl_scan64(tsum, lane);
    if (lane == 31) s_warp[warp] = wincl;
    __syncthreads();
    long long woff = 0, agg = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        const long long x = s_warp[w];
        if (w < warp) woff += x;
        agg += x;
    }
    // look-back by warp 0
    if (warp == 0) {
        long long excl = 0;
        if (tile == 0) {
            if (lane == 0) st_desc(desc, (ST_PREFIX << ST_SHIFT) | ((unsigned long long)agg & ST_VALMASK));
        } else {
            if (lane == 0) st_desc(desc + tile, (ST_AGG << ST_SHIFT) | ((unsigned long long)agg & ST_VALMASK));
            long long look = (long long)tile - 1;
            while (true) {
                const long long idx = look - lane;
                unsigned long long d = (ST_PREFIX << ST_SHIFT);          // virtual tile -1: prefix 0
                if (idx >= 0) {
                    do { d = ld_desc(desc + idx); } while ((d >> ST_SHIFT) == ST_INVALID);
                }
                const unsigned int is_prefix = __ballot_sync(FULL_MASK, (d >> ST_SHIFT) == ST_PREFIX);
                // lanes 0..first-prefix-lane contribute (lane 0 = nearest predecessor)
                const int stop = is_prefix ? (__ffs(is_prefix) - 1) : 31;
                long long val = lane <= stop ? (long long)(d & ST_VALMASK) : 0;
#pragma unroll
                for (int dd = 16; dd > 0; dd >>= 1) val += __shfl_xor_sync(FULL_MASK, val, dd);
                excl += val;
                if (is_prefix) break;
                look -= 32;
            }
            if (lane == 0) st_desc(desc + tile, (ST_PREFIX << ST_SHIFT) | ((unsigned long long)(excl + agg) & ST_VALMASK));
        }
        if (lane == 0) s_prefix = excl;
    }
    __syncthreads();
    long long run = s_prefix + woff + (wincl - tsum);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const int64_t k = base + i;
        if (k < n) out[k] = run;
        run += v[i];
        if (k == n - 1) out[n] = run;
    }
    if (n == 0 && tile == 0 && threadIdx.x == 0) out[0] = 0;
}

