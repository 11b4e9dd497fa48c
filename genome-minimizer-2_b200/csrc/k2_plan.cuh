// k2_plan.cuh — part of libgm2.so (included by gm2.cu; one translation unit).
// K2 + K3a: per-sample segment kept-flags, per-tile kept lengths, within-sample scan (k_plan).
#pragma once

#include "device_util.cuh"

// ------------------------------------------------------------------------------------------
// K2 + K3a  plan: per sample, segment kept-flags and the exclusive scan of kept lengths
//   (minimizer_2.py:75-80 union-of-ranges, :94-96 running output index)
//   One CTA per PLAN_NS samples (the static tables are read once for all of them).  Segment
//   slots are laid out per genome tile, each tile's slots padded to a multiple of 32 so that
//   one ballot == one stored word and k_emit reads whole words.  Per slot the covering genes
//   are inlined as a pair (x, y): -1 = none; y <= -2 points into an overflow list for the rare
//   slot covered by more than two genes.  A warp takes one tile at a time: branch-free bit tests,
//   one ballot per 32 slots and sample, kept lengths accumulated in registers and reduced once per
//   tile (REDUX); warp k then scans the tile sums of sample k.
// ------------------------------------------------------------------------------------------
#ifndef PLAN_NS
#define PLAN_NS 4
#endif
__device__ __forceinline__ bool keep_bit(const uint32_t* row, int g) {
    return g < 0 ? true : ((row[g >> 5] >> (g & 31)) & 1u) != 0u;
}

__global__ void __launch_bounds__(256)
k_plan(int64_t S, int FW, const uint32_t* __restrict__ keep, int ntiles,
       const int32_t* __restrict__ tile_slot, const int32_t* __restrict__ slot_len,
       const int2* __restrict__ slot_cov, const int32_t* __restrict__ cov_ovf,
       int SW, uint32_t* __restrict__ segkept, int32_t* __restrict__ tile_off,
       int64_t* __restrict__ lengths, int64_t* __restrict__ rec_size, int32_t* __restrict__ hdr_len,
       int64_t first_idx, int prefix_len)
{
    extern __shared__ uint32_t plan_sm[];
    uint32_t* rows = plan_sm;                                   // PLAN_NS x FW
    int32_t* tl = (int32_t*)(plan_sm + (size_t)PLAN_NS * FW);   // PLAN_NS x ntiles kept lengths
    const int64_t sbase = (int64_t)blockIdx.x * PLAN_NS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

    for (int i = threadIdx.x; i < PLAN_NS * FW; i += blockDim.x) {
        const int64_t s = sbase + i / FW;
        rows[i] = s < S ? keep[(size_t)s * FW + (i % FW)] : 0u;
    }
    __syncthreads();

    uint32_t* skw = (uint32_t*)(tl + PLAN_NS * ntiles);           // PLAN_NS x SW kept-bit words, staged

    for (int t = warp; t < ntiles; t += nwarps) {
        const int c0 = __ldg(tile_slot + t) >> 5, c1 = __ldg(tile_slot + t + 1) >> 5;   // 32-slot chunks of the tile
        int acc[PLAN_NS];
#pragma unroll
        for (int k = 0; k < PLAN_NS; ++k) acc[k] = 0;
        for (int c = c0; c < c1; ++c) {
            const int slot = 32 * c + lane;
            const int len = __ldg(slot_len + slot);
            const int2 cv = __ldg(slot_cov + slot);
            // branch-free bit tests: word/shift of both covering genes, computed once for all samples
            const int gx = cv.x < 0 ? 0 : cv.x, gy = cv.y < 0 ? 0 : cv.y;
            const int wx = gx >> 5, wy = gy >> 5;
            const uint32_t sx = gx & 31, sy = gy & 31;
            const uint32_t fx = cv.x < 0 ? 1u : 0u, fy = cv.y < 0 ? 1u : 0u;     // "no gene" counts as kept
            const uint32_t live = len > 0 ? 1u : 0u;                            // padding slots have len 0
            uint32_t kb[PLAN_NS];
#pragma unroll
            for (int k = 0; k < PLAN_NS; ++k) {
                const uint32_t* row = rows + k * FW;
                kb[k] = live & ((row[wx] >> sx) | fx) & ((row[wy] >> sy) | fy) & 1u;
            }
            if (__any_sync(FULL_MASK, cv.y < -1)) {                             // rare: > 2 covering genes
                if (cv.y < -1) {
                    const int32_t* o = cov_ovf + (-cv.y - 2);
                    const int n = __ldg(o);
                    for (int j = 1; j <= n; ++j) {
                        const int g = __ldg(o + j);
#pragma unroll
                        for (int k = 0; k < PLAN_NS; ++k) kb[k] &= (rows[k * FW + (g >> 5)] >> (g & 31)) & 1u;
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int k = 0; k < PLAN_NS; ++k) {
                const uint32_t w = __ballot_sync(FULL_MASK, kb[k] != 0u);
                if (lane == 0) skw[k * SW + c] = w;
                acc[k] += kb[k] ? len : 0;
            }
        }
#pragma unroll
        for (int k = 0; k < PLAN_NS; ++k) {
            const int v = __reduce_add_sync(FULL_MASK, acc[k]);
            if (lane == 0) tl[k * ntiles + t] = v;
        }
    }
    __syncthreads();
    // kept-bit rows out, coalesced
    for (int i = threadIdx.x; i < PLAN_NS * SW; i += blockDim.x) {
        const int64_t s = sbase + i / SW;
        if (s < S) segkept[(size_t)s * SW + (i % SW)] = skw[i];
    }
    __syncthreads();
    if (warp < PLAN_NS && sbase + warp < S) {
        const int64_t s = sbase + warp;
        const int32_t* mytl = tl + warp * ntiles;
        int carry = 0;
        int32_t* to = tile_off + (size_t)s * ntiles;
        for (int base = 0; base < ntiles; base += 32) {
            const int t = base + lane;
            const int v = t < ntiles ? mytl[t] : 0;
            const int incl = warp_incl_scan(v, lane);
            if (t < ntiles) to[t] = carry + incl - v;
            carry += __shfl_sync(FULL_MASK, incl, 31);
        }
        if (lane == 0) {
            lengths[s] = carry;
            const int nd = ndigits_u64((unsigned long long)(first_idx + s + 1));
            rec_size[s] = (int64_t)prefix_len + nd + 1 + carry + 1;   // '>'+prefix, digits, '\n', bases, '\n'
            hdr_len[s] = prefix_len + nd + 1;                         // k_emit reads it instead of counting digits per visit
        }
    }
}

