// k1_keep.cuh — part of libgm2.so (included by gm2.cu; one translation unit).
// K1 keep-mask builders: name-id lists (k_keep_from_ids) and dense decoder output (k_keep_from_probs).
#pragma once

#include "device_util.cuh"

// ------------------------------------------------------------------------------------------
// K1  keep-mask builder: name-id lists -> F-bit keep rows          (minimizer_2.py:59-63)
//   One CTA per sample: every thread takes 128-bit vectors of the sample's id list (all of them
//   in flight at once), walks the static name table — a linked list through the genes,
//   first_gene[id] -> next_same_name[g] -> ..., a few KB that stay in L1 — and sets bits in the
//   sample's row in shared memory with atomicOr; the row is written out coalesced.  (A warp per
//   sample with the table staged in shared memory measured 3x slower: with ~2 samples per
//   resident warp the kernel ran for three whole sample latencies.)
// ------------------------------------------------------------------------------------------
//   Two things keep the per-id work short: a V-bit "names a gene" bitmap staged in shared memory filters the ids that
//   match nothing (half of a real list: names of genes the reference does not carry) before any table access, and
//   first_gene[] carries a flag for names borne by exactly one gene (all but a handful), so that the common id costs
//   one table load and one shared-memory atomic and never touches next_same[].
#define K1_THREADS 256
#define K1_SINGLE 0x40000000            // first_gene[id] | K1_SINGLE: the name has exactly one gene
__device__ __forceinline__ void k1_mark(uint32_t* row, const int32_t* __restrict__ first_gene,
                                        const int32_t* __restrict__ next_same, int64_t id) {
    const int e = __ldg(first_gene + id);
    if (e < 0) return;
    int g = e & (K1_SINGLE - 1);
    atomicOr(&row[g >> 5], 1u << (g & 31));
    if (!(e & K1_SINGLE))
        for (g = __ldg(next_same + g); g >= 0; g = __ldg(next_same + g)) atomicOr(&row[g >> 5], 1u << (g & 31));
}

__global__ void __launch_bounds__(K1_THREADS)
k_keep_from_ids(const int32_t* __restrict__ ids, const int64_t* __restrict__ off, int64_t S, int32_t V,
                const int32_t* __restrict__ first_gene, const int32_t* __restrict__ next_same,
                const uint32_t* __restrict__ has_gene, int FW, uint32_t* __restrict__ keep)
{
    extern __shared__ uint32_t k1_row[];                    // FW words of the row, then ceil(V/32) words of has_gene
    uint32_t* hg = k1_row + FW;
    const int VW = (V + 31) >> 5;
    const int64_t s = blockIdx.x;
    for (int i = threadIdx.x; i < FW; i += blockDim.x) k1_row[i] = 0u;
    for (int i = threadIdx.x; i < VW; i += blockDim.x) hg[i] = __ldg(has_gene + i);
    __syncthreads();
    const int64_t b = off[s], e = off[s + 1];
    auto mark = [&](int32_t id) {
        if ((uint32_t)id < (uint32_t)V && ((hg[id >> 5] >> (id & 31)) & 1u)) k1_mark(k1_row, first_gene, next_same, id);
    };
    // head up to a 16-byte boundary, 128-bit body, scalar tail
    const int64_t b4 = min((b + 3) & ~(int64_t)3, e), e4 = b4 + ((e - b4) & ~(int64_t)3);
    if (b + threadIdx.x < b4) mark(__ldg(ids + b + threadIdx.x));
    const int4* v4 = reinterpret_cast<const int4*>(ids + b4);
    const int64_t nv = (e4 - b4) >> 2;
    for (int64_t i0 = 0; i0 < nv; i0 += 4 * K1_THREADS) {
        int4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + (int64_t)u * K1_THREADS + threadIdx.x;
            x[u] = i < nv ? __ldg(v4 + i) : make_int4(-1, -1, -1, -1);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { mark(x[u].x); mark(x[u].y); mark(x[u].z); mark(x[u].w); }
    }
    if (e4 + threadIdx.x < e) mark(__ldg(ids + e4 + threadIdx.x));
    __syncthreads();
    uint32_t* dst = keep + (size_t)s * FW;
    for (int i = threadIdx.x; i < FW; i += blockDim.x) dst[i] = k1_row[i];
}

// ------------------------------------------------------------------------------------------
// K1'  keep-mask builder from dense probabilities (BASELINE config 5):
//   the reference's  decode -> `> 0.5` (utils/extras.py:200-201) -> masks_to_gene_lists `>= 0.5`
//   on the 0/1 matrix (explore_data/binary_converter.py:55,:64) -> check_essential_genes adds the
//   missing essentials (:91-98) -> `name in needed` (minimizer_2.py:62), collapsed: column c is a
//   name id; it is "present" iff probs[s][c] > threshold; a gene is kept iff its name's column is
//   present or it is forced (essential).  counts[s] = length of the list the reference would
//   have built = present columns + forced ids that are not present (+ a host-side constant for
//   essentials that are no column at all).  One CTA per sample, coalesced 128-bit reads.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_keep_from_probs(const float* __restrict__ probs, int64_t S, int64_t V, int64_t ld, float thr,
                  const int32_t* __restrict__ first_gene, const int32_t* __restrict__ next_same,
                  const uint32_t* __restrict__ has_gene, const uint32_t* __restrict__ forced_ids,
                  const uint32_t* __restrict__ force_keep, int FW, int VW, uint32_t* __restrict__ keep,
                  int64_t* __restrict__ counts)
{
    // shared: keep row (FW words) | has-gene bitmap (VW+1 words) | forced-id bitmap (VW+1 words, zeros if none)
    extern __shared__ uint32_t kp_sm[];
    uint32_t* kp_row = kp_sm;
    uint32_t* hg = kp_sm + FW;
    uint32_t* fo = hg + VW + 1;
    __shared__ int s_count;
    const int64_t s = blockIdx.x;
    for (int i = threadIdx.x; i < FW; i += blockDim.x) kp_row[i] = force_keep ? force_keep[i] : 0u;
    for (int i = threadIdx.x; i <= VW; i += blockDim.x) {
        hg[i] = i < VW ? has_gene[i] : 0u;
        fo[i] = (forced_ids && i < VW) ? forced_ids[i] : 0u;
    }
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const float* p = probs + s * ld;
    int cnt = 0;
    auto mark = [&](int64_t c) { k1_mark(kp_row, first_gene, next_same, c); };
    auto visit1 = [&](int64_t c, float v) {
        const uint32_t bit = 1u << (c & 31);
        if (v > thr) { ++cnt; if (hg[c >> 5] & bit) mark(c); }
        else if (fo[c >> 5] & bit) ++cnt;            // an essential the reference appends to the list
    };
    int64_t head = (int64_t)(((16u - (uint32_t)((uintptr_t)p & 15u)) & 15u) >> 2);
    if (head > V) head = V;
    if (threadIdx.x < head) visit1(threadIdx.x, __ldg(p + threadIdx.x));
    const float4* v4 = reinterpret_cast<const float4*>(p + head);
    const int64_t nvec = (V - head) >> 2;
    for (int64_t i = threadIdx.x; i < nvec; i += blockDim.x) {
        const float4 v = __ldg(v4 + i);
        const int64_t c = head + 4 * i;
        const uint32_t m = (v.x > thr ? 1u : 0u) | (v.y > thr ? 2u : 0u) | (v.z > thr ? 4u : 0u) | (v.w > thr ? 8u : 0u);
        const int w = (int)(c >> 5), sh = (int)(c & 31);
        const uint32_t hg4 = __funnelshift_r(hg[w], hg[w + 1], sh) & 0xfu;      // 4 bitmap bits, may straddle words
        const uint32_t fo4 = __funnelshift_r(fo[w], fo[w + 1], sh) & 0xfu;
        cnt += __popc(m) + __popc(~m & fo4);
        uint32_t todo = m & hg4;                                               // present columns that name a gene: rare
        while (todo) { const int j = __ffs(todo) - 1; todo &= todo - 1; mark(c + j); }
    }
    const int64_t tail0 = head + 4 * nvec;
    if (tail0 + threadIdx.x < V) visit1(tail0 + threadIdx.x, __ldg(p + tail0 + threadIdx.x));
    cnt = __reduce_add_sync(FULL_MASK, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_count, cnt);
    __syncthreads();
    uint32_t* dst = keep + (size_t)s * FW;
    for (int i = threadIdx.x; i < FW; i += blockDim.x) dst[i] = kp_row[i];
    if (threadIdx.x == 0) counts[s] = s_count;
}

