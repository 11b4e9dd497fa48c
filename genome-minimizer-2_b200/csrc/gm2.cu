// gm2.cu — libgm2.so: hand-written sm_100a kernels + the C-ABI of include/gm2.h.
//
// Path replaced (reference = ucl-cssb/genome-minimizer-2, pure Python):
//   GenomeMinimiser.__init__            src/genome_minimizer_2/minimizer/minimizer_2.py:20-48
//     _extract_non_essential_genes      :50-66    -> k_keep_from_ids (K1)
//     _get_positions_to_remove          :68-83    -> k_plan          (K2: segment flags)
//     _create_minimized_sequence        :85-101   -> k_plan / k_scan_records (K3) + k_emit (K4)
//   record write  f">{seq_id}\n{seq}\n" :476-477, :544-545 -> fused into k_emit
//
// Formulation ("segment form").  The breakpoints {0,G} ∪ {gene starts}
// ∪ {gene ends} ∪ {multiples of the tile size} cut the genome into elementary
// segments; each has a STATIC cover set of genes.  Per sample a segment is kept iff
// every gene covering it is kept (coverage by any removed gene deletes — the
// reference's set-union, minimizer_2.py:75-80).  No per-base state ever reaches HBM:
// per sample the plan writes one kept-bit per segment slot (~1.2 KB) and one output
// offset per genome tile (~0.3 KB); k_emit turns those into the FASTA image directly.
//
// There is deliberately no CPU fallback in this file.

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>
#include <new>

#include "gm2.h"

#define GM2_API extern "C" __attribute__((visibility("default")))

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------

static const char kDefaultPrefix[] = "Minimized_E_coli_K12_MG1655_";   // minimizer_2.py:476
#define GM2_MAX_PREFIX 95

struct HeaderPrefix {            // passed by value to kernels; text[0] is '>'
    int  len;
    char text[GM2_MAX_PREFIX + 1];
};

struct gm2_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;        // where work is issued (own or adopted)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_emit[2] = {nullptr, nullptr};
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    std::string err;
    uint64_t launches = 0;

    // configuration
    int tile_bytes = 49152;        // 4 CTAs/SM of 8 warps with the 64-register kernel variant: best measured (profiles/)
    int emit_warps = 8;
    int emit_batch = 0;
    int packing_req = 0;
    int store_policy = 1;
    int rt_cap = 64;
    int debug = 0;
    int order = 1;
    HeaderPrefix prefix;

    // reference
    bool have_ref = false;
    int64_t G = 0;
    int32_t F = 0, FW = 0;
    int ntiles = 0, nseg = 0, nslots = 0, SW = 0, max_tile_slots = 0;
    int packing = 1;
    uint8_t* d_seq = nullptr;
    uint8_t* d_seq2 = nullptr;         // 2 bits per base (only when packing == 2)
    int32_t *d_tile_slot = nullptr, *d_slot_src = nullptr, *d_slot_len = nullptr;
    int2* d_slot_cov = nullptr; int32_t* d_cov_ovf = nullptr;

    // name map
    int32_t V = 0;
    int32_t *d_first_gene = nullptr, *d_next_same = nullptr;
    uint32_t *d_forced_ids = nullptr, *d_force_keep = nullptr;     // optional (gm2_set_forced)
    uint32_t *d_has_gene = nullptr;                                // one bit per name id: names at least one gene
    const float* probs = nullptr; int64_t probs_ld = 0; float probs_thr = 0.5f;   // mode 3 (borrowed)
    int64_t* d_counts = nullptr; int64_t counts_cap = 0;

    // samples
    int64_t S = 0;
    int mode = 0;                      // 0 none, 1 ids, 2 keep rows, 3 dense probabilities
    const int32_t* ids = nullptr;      // device (owned or borrowed)
    const int64_t* ids_off = nullptr;
    const uint32_t* keep_in = nullptr; // device keep rows when mode == 2
    int32_t* own_ids = nullptr;   int64_t own_ids_cap = 0;
    int64_t* own_ids_off = nullptr; int64_t own_ids_off_cap = 0;
    uint32_t* own_keep = nullptr; int64_t own_keep_cap = 0;   // words

    // plan outputs (device)
    uint32_t* d_segkept = nullptr; int64_t segkept_cap = 0;    // words
    int32_t* d_tile_off = nullptr; int64_t tile_off_cap = 0;   // elements
    int64_t *d_len = nullptr, *d_rec_size = nullptr, *d_rec_off = nullptr;
    int64_t len_cap = 0, rec_size_cap = 0, rec_off_cap = 0;
    unsigned long long* d_scan_desc = nullptr; int64_t scan_desc_cap = 0;
    unsigned int* d_scan_ticket = nullptr;
    // plan outputs (pinned host mirror)
    int64_t *h_len = nullptr, *h_rec_off = nullptr; int64_t h_cap = 0;
    bool planned = false, host_plan = false;
    int64_t first_idx = 0;

    // staging for gm2_emit_host
    uint8_t* d_stage[2] = {nullptr, nullptr}; int64_t stage_cap = 0;
    // scratch for diag hashes
};

static thread_local std::string g_create_err;

static int fail(gm2_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_err = msg;
    return code;
}
static int cuda_fail(gm2_ctx* c, cudaError_t e, const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return fail(c, GM2_ERR_CUDA, m);
}
#define CU(c, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail((c), e__, #call); } while (0)

template <typename T>
static int dev_reserve(gm2_ctx* c, T** p, int64_t* cap, int64_t need) {
    if (need <= *cap && *p) return GM2_OK;
    if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
    int64_t n = std::max<int64_t>(need, 1);
    cudaError_t e = cudaMalloc((void**)p, (size_t)n * sizeof(T));
    if (e != cudaSuccess) { *p = nullptr; return cuda_fail(c, e, "cudaMalloc"); }
    *cap = n;
    return GM2_OK;
}
template <typename T>
static int dev_upload(gm2_ctx* c, T** p, const std::vector<T>& v) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    size_t n = std::max<size_t>(v.size(), 1);
    CU(c, cudaMalloc((void**)p, n * sizeof(T)));
    if (!v.empty()) CU(c, cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return GM2_OK;
}

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

// ------------------------------------------------------------------------------------------
// K1  keep-mask builder: name-id lists -> F-bit keep rows          (minimizer_2.py:59-63)
//   One warp per sample, CTAs loop over groups of samples.  The static name table is a
//   linked list through the genes (first_gene[id] -> next_same_name[g] -> ...), staged in
//   shared memory when it fits, so an id costs one coalesced global load plus shared-memory
//   lookups; the row is assembled in shared memory with atomicOr and written out coalesced.
// ------------------------------------------------------------------------------------------
#define K1_WARPS 8
__global__ void __launch_bounds__(K1_WARPS * 32, 4)
k_keep_from_ids(const int32_t* __restrict__ ids, const int64_t* __restrict__ off, int64_t S, int32_t V, int32_t F,
                const int32_t* __restrict__ first_gene, const int32_t* __restrict__ next_same,
                int FW, uint32_t* __restrict__ keep, int map_in_smem)
{
    extern __shared__ uint32_t k1_sm[];
    uint32_t* rows = k1_sm;                                          // K1_WARPS x FW
    const int32_t* fg = first_gene;
    const int32_t* nx = next_same;
    if (map_in_smem) {
        int32_t* s_fg = reinterpret_cast<int32_t*>(k1_sm + (size_t)K1_WARPS * FW);
        int32_t* s_nx = s_fg + V;
        for (int i = threadIdx.x; i < V; i += blockDim.x) s_fg[i] = first_gene[i];
        for (int i = threadIdx.x; i < F; i += blockDim.x) s_nx[i] = next_same[i];
        fg = s_fg; nx = s_nx;
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* row = rows + (size_t)warp * FW;
    for (int64_t s = (int64_t)blockIdx.x * K1_WARPS + warp; s < S; s += (int64_t)gridDim.x * K1_WARPS) {
        for (int i = lane; i < FW; i += 32) row[i] = 0u;
        __syncwarp();
        const int64_t b = off[s], e = off[s + 1];
        auto mark = [&](int32_t id) {
            if ((uint32_t)id < (uint32_t)V)
                for (int g = fg[id]; g >= 0; g = nx[g]) atomicOr(&row[g >> 5], 1u << (g & 31));
        };
        // head up to a 16-byte boundary, 128-bit body (two vectors per lane in flight), scalar tail
        const int64_t b4 = min((b + 3) & ~(int64_t)3, e), e4 = b4 + ((e - b4) & ~(int64_t)3);
        if (b + lane < b4) mark(__ldg(ids + b + lane));
        const int4* v4 = reinterpret_cast<const int4*>(ids + b4);
        const int64_t nv = (e4 - b4) >> 2;
        for (int64_t i0 = 0; i0 < nv; i0 += 128) {
            int4 x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = i0 + 32 * u + lane;
                x[u] = i < nv ? __ldg(v4 + i) : make_int4(-1, -1, -1, -1);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { mark(x[u].x); mark(x[u].y); mark(x[u].z); mark(x[u].w); }
        }
        if (e4 + lane < e) mark(__ldg(ids + e4 + lane));
        __syncwarp();
        uint32_t* dst = keep + (size_t)s * FW;
        for (int i = lane; i < FW; i += 32) dst[i] = row[i];
        __syncwarp();
    }
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
    auto mark = [&](int64_t c) {
        for (int g = __ldg(first_gene + c); g >= 0; g = __ldg(next_same + g)) atomicOr(&kp_row[g >> 5], 1u << (g & 31));
    };
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
#define PLAN_NS 4
__device__ __forceinline__ bool keep_bit(const uint32_t* row, int g) {
    return g < 0 ? true : ((row[g >> 5] >> (g & 31)) & 1u) != 0u;
}

__global__ void __launch_bounds__(256)
k_plan(int64_t S, int FW, const uint32_t* __restrict__ keep, int ntiles,
       const int32_t* __restrict__ tile_slot, const int32_t* __restrict__ slot_len,
       const int2* __restrict__ slot_cov, const int32_t* __restrict__ cov_ovf,
       int SW, uint32_t* __restrict__ segkept, int32_t* __restrict__ tile_off,
       int64_t* __restrict__ lengths, int64_t* __restrict__ rec_size,
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
        }
    }
}

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

// ------------------------------------------------------------------------------------------
// K4  emit: stream-compaction gather + FASTA framing              (minimizer_2.py:94-97, :476-477)
//   CTA = (genome tile, batch of samples).  The tile's bases are staged ONCE in shared
//   memory by a 1-D TMA bulk copy (cp.async.bulk + mbarrier) and reused by every sample
//   of the batch.  Each warp owns one sample at a time: it reads the tile's kept-bit
//   words, scans kept segment lengths with shuffles, and copies every maximal kept run
//   shared->global with destination-aligned 128-bit stores (the source is re-phased with
//   funnel shifts); only a run's <16-byte head and tail use byte stores.  The warp that
//   owns tile 0 writes the '>' header, the one that owns the last tile the final '\n'.
// ------------------------------------------------------------------------------------------
struct EmitParams {
    const uint8_t* seq;
    const int32_t* tile_slot;
    const int32_t* slot_src;
    const int32_t* slot_len;
    const uint32_t* segkept;
    const int32_t* tile_off;
    const int64_t* lengths;
    const int64_t* rec_off;
    uint8_t* out;
    int64_t s0, s1;
    int64_t first_idx;
    int tile_bytes, ntiles, SW, batch, nbatch;
    int tile_smem_bytes;   // bytes of the staged tile: tile_bytes (1 byte/base) or tile_bytes/4 (2 bits/base)
    int rt_cap;            // run-table entries per warp (shared memory)
    int slot_cap;          // slot-table entries staged in shared memory (0: read from global)
    int order;             // CTA -> work mapping: 0 tile-major, 1 sample-major
    int debug;             // timing experiments only (wrong output; needs -DGM2_EMIT_DEBUG): 1 no boundary
                           // sectors, 2 no interior stores, 4 interior stores without shared loads
    HeaderPrefix prefix;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// shared-memory accessors on 32-bit shared-window addresses.  Tile / slot-table reads are plain
// (read-only after the CTA barrier, free to be scheduled); run-table accesses are volatile with a
// memory clobber because the table is rewritten per (sample, tile) around __syncwarp().
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int2 rt_load(uint32_t a) {
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void rt_store(uint32_t a, int x, int y) {
    asm volatile("st.shared.v2.s32 [%0], {%1, %2};" :: "r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void rt_store_x(uint32_t a, int x) {
    asm volatile("st.shared.s32 [%0], %1;" :: "r"(a), "r"(x) : "memory");
}

// POLICY 1 = streaming (evict-first) stores: the image is written once and never re-read here.
template <int POLICY>
__device__ __forceinline__ void st128(uint8_t* p, const uint4& v) {
    if (POLICY == 1) __stcs(reinterpret_cast<uint4*>(p), v);
    else *reinterpret_cast<uint4*>(p) = v;
}
template <int POLICY>
__device__ __forceinline__ void st8(uint8_t* p, uint32_t v) {
    if (POLICY == 1) __stcs(p, (uint8_t)v);
    else *p = (uint8_t)v;
}

// One whole 32-byte sector from one lane (STG.256).
__device__ __forceinline__ void st256(uint8_t* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// ---- 2-bit packing (ACGT-only references): base i of the tile sits in bits [2i, 2i+2) of the
// little-endian bit stream, code 0..3 = A, C, G, T.  Sixteen bases = one 32-bit word.
#define ACGT_LUT 0x54474341u                     // 'A' 'C' 'G' 'T' as bytes 0..3
__device__ __forceinline__ uint32_t spread8(uint32_t t) {      // 8 two-bit codes -> 8 nibbles
    t &= 0xffffu;
    t = (t | (t << 8)) & 0x00ff00ffu;
    t = (t | (t << 4)) & 0x0f0f0f0fu;
    t = (t | (t << 2)) & 0x33333333u;
    return t;
}
__device__ __forceinline__ uint4 expand16(uint32_t x) {       // 16 codes -> 16 ASCII bytes (PRMT as a 4-entry LUT)
    const uint32_t lo = spread8(x), hi = spread8(x >> 16);
    return make_uint4(__byte_perm(ACGT_LUT, 0u, lo), __byte_perm(ACGT_LUT, 0u, lo >> 16),
                      __byte_perm(ACGT_LUT, 0u, hi), __byte_perm(ACGT_LUT, 0u, hi >> 16));
}
__device__ __forceinline__ uint32_t base_at_2bit(uint32_t tile_a, int b) {
    const uint32_t w = lds32(tile_a + (uint32_t)((b >> 4) << 2));
    return (ACGT_LUT >> (8u * ((w >> (2 * (b & 15))) & 3u))) & 0xffu;
}

// Interior of one kept run: nb destination-aligned 16-byte vectors, lanes strided by 32.
// qa = this lane's 16-byte aligned shared address at or below its first source byte,
// K = word phase (0..3), sh = byte phase in bits.
template <int POLICY, int K>
__device__ __forceinline__ void copy_vectors(uint32_t qa, uint8_t* __restrict__ d, int nb, int sh, int lane)
{
#pragma unroll 1
    for (int v = lane; v < nb; v += 32, qa += 512, d += 512) {
        const uint4 lo = lds128(qa);
        const uint4 hi = lds128(qa + 16);
        const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        uint4 o;
        o.x = __funnelshift_r(w[K], w[K + 1], sh);
        o.y = __funnelshift_r(w[K + 1], w[K + 2], sh);
        o.z = __funnelshift_r(w[K + 2], w[K + 3], sh);
        o.w = __funnelshift_r(w[K + 3], w[K + 4], sh);
        st128<POLICY>(d, o);
    }
}

// 16 bytes from an arbitrarily aligned shared address (per-lane alignment).
__device__ __forceinline__ uint4 fetch16(uint32_t a) {
    const uint32_t a4 = a & ~3u;
    const int sh = (int)(a & 3u) * 8;
    const uint32_t w0 = lds32(a4), w1 = lds32(a4 + 4), w2 = lds32(a4 + 8), w3 = lds32(a4 + 12), w4 = lds32(a4 + 16);
    return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh),
                      __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}
__device__ __forceinline__ uint32_t low_bytes_mask(int n) {          // n bytes from the low end, n clamped to 0..4
    return n >= 4 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << (8 * n)) - 1u));
}

// One batch of kept runs of a (sample, tile): table A entry r = {Q_r, S_r}, entry nr = {end, -}.
//   Q = destination offset in "Q space" (bytes from base32, a 32-byte aligned global pointer),
//   S = source byte offset inside the shared-memory tile.  Output is contiguous: run r covers
//   [Q_r, Q_{r+1}).
// The warp writes ONE ASCENDING STREAM in units of 32-byte sectors, each sector exactly once and
// in address order (measured with store-only models, profiles/r01_emit_experiments.md: a sector
// written out of stream, microseconds after its neighbours, costs 11-19 % of the bandwidth because
// its line has already left L2; written in stream it is free):
//   for each run r, in order
//     - if the run starts inside a sector, that sector (tail of run r-1 and earlier, head of run r
//       and later) is gathered cooperatively, lane j <-> byte j, and leaves as one coalesced store;
//     - then every whole sector inside the run, 128-bit stores, source re-phased by funnel shifts;
//   finally the partial last sector (pseudo-run nr).  Bytes outside [Q_0, Q_nr) belong to the
//   neighbouring tile / batch (another warp) and are never touched.
// Everything that is uniform per run is computed ONCE, lane r <-> run r, into table B
//   {x: dst offset of the first whole sector, y: #16-byte vectors | flags, z: src offset of that
//    sector, w: Q_r}, so the run loop costs one broadcast 128-bit shared load plus the copy.
#define RUN_HAS_BOUNDARY 0x40000000
#define RUN_SIMPLE       0x20000000
#define RUN_COUNT_MASK   0x00ffffff

__device__ __forceinline__ int4 rtb_load(uint32_t a) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void rtb_store(uint32_t a, int x, int y, int z, int w) {
    asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" :: "r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

template <int POLICY, int PACK>
__device__ __forceinline__ void emit_runs(uint32_t tile_a, uint32_t rt_a, uint32_t rtb_a, int nr,
                                          uint8_t* __restrict__ base32, int lane, int debug)
{
    __syncwarp();
    const int q_first = rt_load(rt_a).x, q_last = rt_load(rt_a + 8 * nr).x;
    // ---- table B, lane-parallel
    for (int r = lane; r <= nr; r += 32) {
        const int2 er = rt_load(rt_a + 8 * r);
        const int qn = r < nr ? rt_load(rt_a + 8 * (r + 1)).x : er.x;
        const int qp = r > 0 ? rt_load(rt_a + 8 * (r - 1)).x : q_first;
        const int W = er.x >> 5;
        int flags = 0;
        if ((er.x & 31) && !(r > 0 && (qp >> 5) == W && (qp & 31))) {        // first boundary inside sector W owns it
            flags = RUN_HAS_BOUNDARY;
            if ((r == 0 || qp <= (W << 5)) && (r == nr || qn >= (W << 5) + 32)) flags |= RUN_SIMPLE;
        }
        const int sa = (er.x + 31) >> 5, sb = qn >> 5;
        const int nb = sb > sa ? (sb - sa) << 1 : 0;
        rtb_store(rtb_a + 16 * r, sa << 5, nb | flags, er.y + ((sa << 5) - er.x), er.x);
    }
    __syncwarp();
    // ---- the stream
    int dA = 0;
    for (int r = 0; r <= nr; ++r) {
        const int4 t = rtb_load(rtb_a + 16 * r);                               // warp-uniform (broadcast)
        const int dB = t.z - t.x;                                            // S_r - Q_r
#ifdef GM2_EMIT_DEBUG
        if ((t.y & RUN_HAS_BOUNDARY) && !(debug & 1)) {
#else
        if (t.y & RUN_HAS_BOUNDARY) {
#endif
            const int pos = (t.w & ~31) + lane;
            if (pos >= q_first && pos < q_last) {
                int src;
                if (t.y & RUN_SIMPLE) {
                    src = pos + (pos < t.w ? dA : dB);
                } else {                                   // three or more runs meet in this sector
                    int rr = r; int2 ec = rt_load(rt_a + 8 * rr);
                    if (pos < ec.x) { do { --rr; ec = rt_load(rt_a + 8 * rr); } while (pos < ec.x); }
                    else { int qn = rt_load(rt_a + 8 * (rr + 1)).x;
                           while (pos >= qn) { ++rr; ec = rt_load(rt_a + 8 * rr); qn = rt_load(rt_a + 8 * (rr + 1)).x; } }
                    src = ec.y + (pos - ec.x);
                }
                st8<POLICY>(base32 + pos, PACK == 2 ? base_at_2bit(tile_a, src) : lds8(tile_a + (uint32_t)src));
            }
        }
        const int nb = t.y & RUN_COUNT_MASK;
        if (nb > 0 && PACK == 2) {
            // two-bit source: one (unaligned) 32-bit window per 16 output bases, expanded in registers
            int bidx = t.z + 16 * lane;                                      // source base index of this lane's vector
            uint8_t* d = base32 + t.x + 16 * lane;
            const int sh = 2 * (t.z & 15);                                   // warp-uniform, loop-invariant
#pragma unroll 1
            for (int v = lane; v < nb; v += 32, bidx += 512, d += 512) {
                const uint32_t wa = tile_a + (uint32_t)((bidx >> 4) << 2);
                st128<POLICY>(d, expand16(__funnelshift_r(lds32(wa), lds32(wa + 4), sh)));
            }
        } else if (nb > 0) {
            const int mis = t.z & 15;
            const uint32_t qa = tile_a + (uint32_t)(t.z - mis) + 16u * lane;
            uint8_t* d = base32 + t.x + 16 * lane;
            const int sh = (mis & 3) * 8;
#ifdef GM2_EMIT_DEBUG
            if (debug & 6) {
                if (debug & 4) { for (int v = lane; v < nb; v += 32, d += 512) st128<POLICY>(d, make_uint4(sh, mis, nb, r)); }
                else { uint32_t q = qa, acc = 0; for (int v = lane; v < nb; v += 32, q += 512) { const uint4 tt = lds128(q); acc ^= tt.x ^ tt.w; }
                       if (acc == 0x12345u) st128<POLICY>(d, make_uint4(acc, 0, 0, 0)); }
            } else
#endif
            if (mis == 0) {
                uint32_t q = qa;
#pragma unroll 1
                for (int v = lane; v < nb; v += 32, q += 512, d += 512) st128<POLICY>(d, lds128(q));
            } else {
                switch (mis >> 2) {
                case 0:  copy_vectors<POLICY, 0>(qa, d, nb, sh, lane); break;
                case 1:  copy_vectors<POLICY, 1>(qa, d, nb, sh, lane); break;
                case 2:  copy_vectors<POLICY, 2>(qa, d, nb, sh, lane); break;
                default: copy_vectors<POLICY, 3>(qa, d, nb, sh, lane); break;
                }
            }
        }
        dA = dB;
    }
    __syncwarp();
}

#define EMIT_FRONT_PAD 32
#define EMIT_BACK_PAD  64

template <int POLICY, int MIN_CTAS, int PACK>
__global__ void __launch_bounds__(256, MIN_CTAS)
k_emit(const EmitParams p)
{
    // dynamic shared memory: 16 B front pad | tile bytes | 32 B over-read pad | slot tables | per-warp run tables
    extern __shared__ __align__(128) uint8_t dsm[];
    __shared__ __align__(8) unsigned long long bar;

    const int ntl = p.ntiles > 0 ? p.ntiles : 1;
    const int tile = p.order ? (int)(blockIdx.x % ntl) : (int)(blockIdx.x / p.nbatch);
    const int b = p.order ? (int)(blockIdx.x / ntl) : (int)(blockIdx.x - tile * p.nbatch);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const bool have_tile = p.ntiles > 0;
    const int sl0 = have_tile ? __ldg(p.tile_slot + tile) : 0;
    const int nslots = have_tile ? __ldg(p.tile_slot + tile + 1) - sl0 : 0;
    const int nwords = nslots >> 5;
    const int tile_base = tile * p.tile_bytes;

    const uint32_t dsm_a = smem_u32(dsm);
    const uint32_t tile_a = dsm_a + EMIT_FRONT_PAD;
    const uint32_t len_a = tile_a + (uint32_t)p.tile_smem_bytes + EMIT_BACK_PAD;
    const uint32_t src_a = len_a + 4u * (uint32_t)p.slot_cap;
    const uint32_t rt_a = src_a + 4u * (uint32_t)p.slot_cap + (uint32_t)warp * (uint32_t)(p.rt_cap + 2) * 24u;   // table A (8 B) + table B (16 B) per entry
    const uint32_t rtb_a = rt_a + (uint32_t)(p.rt_cap + 2) * 8u;
    int32_t* sm_len = reinterpret_cast<int32_t*>(dsm + EMIT_FRONT_PAD + p.tile_smem_bytes + EMIT_BACK_PAD);
    int32_t* sm_src = sm_len + p.slot_cap;
    const bool slots_staged = nslots <= p.slot_cap;

    if (have_tile) {
        const uint32_t bar_a = smem_u32(&bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const uint32_t bytes = (uint32_t)p.tile_smem_bytes;
            const uint8_t* src = p.seq + (size_t)tile * p.tile_smem_bytes;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(tile_a), "l"(src), "r"(bytes), "r"(bar_a) : "memory");
        }
        if (slots_staged) {
            for (int i = threadIdx.x; i < nslots; i += blockDim.x) {
                sm_len[i] = __ldg(p.slot_len + sl0 + i);
                sm_src[i] = __ldg(p.slot_src + sl0 + i) - tile_base;
            }
        }
        __syncthreads();                      // barrier init + slot tables visible to every thread
        uint32_t done = 0;                    // wait for phase 0 of the mbarrier (the TMA's complete_tx)
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar_a), "r"(0u) : "memory");
        }
    }

    const int64_t sb = p.s0 + (int64_t)b * p.batch;
    const int64_t se = sb + p.batch < p.s1 ? sb + p.batch : p.s1;
    const int last_tile = p.ntiles > 0 ? p.ntiles - 1 : 0;
    const int64_t img0 = __ldg(p.rec_off + p.s0);
    const uint32_t lt_mask = (1u << lane) - 1u;

    // per-sample metadata is fetched one sample ahead: record offset, this tile's output offset,
    // the sample's length (last tile only) and ALL kept-bit words of the tile in one coalesced load
    int64_t m_roff = 0; int m_toff = 0; uint32_t m_words = 0u;
    auto load_meta = [&](int64_t s) {
        m_roff = __ldg(p.rec_off + s);
        if (have_tile) {
            m_toff = __ldg(p.tile_off + (size_t)s * p.ntiles + tile);
            m_words = lane < nwords ? __ldg(p.segkept + (size_t)s * p.SW + (sl0 >> 5) + lane) : 0u;
        }
    };
    int64_t s = sb + warp;
    if (s < se) load_meta(s);
    while (s < se) {
        const int64_t roff = m_roff; const int toff = m_toff; const uint32_t words = m_words;
        const int64_t sn = s + nwarps;
        if (sn < se) load_meta(sn);

        uint8_t* rec = p.out + (roff - img0);
        const unsigned long long num = (unsigned long long)(p.first_idx + s + 1);
        const int nd = ndigits_u64(num);
        const int hl = p.prefix.len + nd + 1;
        if (tile == 0) {
            for (int i = lane; i < hl; i += 32) {
                char ch;
                if (i < p.prefix.len) ch = p.prefix.text[i];
                else if (i == hl - 1) ch = '\n';
                else ch = (char)('0' + (int)((num / c_pow10[nd - 1 - (i - p.prefix.len)]) % 10ull));
                rec[i] = (uint8_t)ch;
            }
        }
        uint8_t* seqout = rec + hl;
        if (have_tile) {
            const int A = (int)((uintptr_t)seqout & 31u);
            uint8_t* base32 = seqout - A;
            asm volatile("" : "+l"(base32));                   // keep the 64-bit base in registers (no re-derivation per run)
            int q = toff + A;
            int nr = 0;
            uint32_t carry = 0u;
            for (int c = 0; ; ++c) {
                const bool done = c >= nwords;
                if (done || nr + 17 > p.rt_cap) {               // tile finished, or table full: emit what we have
                    if (nr > 0) {
                        if (lane == 0) rt_store_x(rt_a + 8u * nr, q);
                        emit_runs<POLICY, PACK>(tile_a, rt_a, rtb_a, nr, base32, lane, p.debug);
                        nr = 0; carry = 0u;
                    }
                    if (done) break;
                }
                const uint32_t w = c < 32 ? __shfl_sync(FULL_MASK, words, c)
                                          : __ldg(p.segkept + (size_t)s * p.SW + (sl0 >> 5) + c);   // warp-uniform
                int len, src;
                if (slots_staged) { len = (int)lds32(len_a + 4u * (32 * c + lane)); src = (int)lds32(src_a + 4u * (32 * c + lane)); }
                else { len = __ldg(p.slot_len + sl0 + 32 * c + lane); src = __ldg(p.slot_src + sl0 + 32 * c + lane) - tile_base; }
                const int x = ((w >> lane) & 1u) ? len : 0;
                const int incl = warp_incl_scan(x, lane);
                const uint32_t starts = w & ~((w << 1) | carry);
                carry = w >> 31;
                if ((starts >> lane) & 1u) rt_store(rt_a + 8u * (nr + __popc(starts & lt_mask)), q + incl - x, src);
                nr += __popc(starts);
                q += __shfl_sync(FULL_MASK, incl, 31);
            }
        }
        if (tile == last_tile && lane == 0) seqout[__ldg(p.lengths + s)] = (uint8_t)'\n';
        s = sn;
    }
}

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


// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------

GM2_API int gm2_abi_version(void) { return GM2_ABI_VERSION; }

GM2_API int gm2_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cuda_fail(nullptr, e, "cudaGetDeviceCount"); return GM2_ERR_CUDA; }
    return n;
}

GM2_API const char* gm2_last_error(const gm2_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

static void set_prefix(gm2_ctx* c, const char* text) {
    size_t n = strlen(text);
    if (n > GM2_MAX_PREFIX - 1) n = GM2_MAX_PREFIX - 1;
    memset(&c->prefix, 0, sizeof(c->prefix));
    c->prefix.text[0] = '>';
    memcpy(c->prefix.text + 1, text, n);
    c->prefix.len = (int)n + 1;
}

GM2_API int gm2_create(int device, gm2_ctx** out) {
    if (!out) return fail(nullptr, GM2_ERR_INVALID, "gm2_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "gm2_create: no usable CUDA device (there is no CPU fallback); cudaGetDeviceCount");
    if (device < 0 || device >= n) return fail(nullptr, GM2_ERR_INVALID, "gm2_create: device ordinal out of range");
    gm2_ctx* c = new (std::nothrow) gm2_ctx();
    if (!c) return fail(nullptr, GM2_ERR_NOMEM, "gm2_create: out of host memory");
    c->device = device;
    set_prefix(c, kDefaultPrefix);
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        cuda_fail(nullptr, e, "gm2_create: cudaSetDevice/cudaGetDeviceProperties"); delete c; return GM2_ERR_CUDA;
    }
    if (prop.major < 9) {
        delete c;
        return fail(nullptr, GM2_ERR_CUDA, "gm2_create: device is not Hopper/Blackwell class; this library is built for sm_100a only");
    }
    c->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        cuda_fail(nullptr, e, "gm2_create: cudaStreamCreate"); delete c; return GM2_ERR_CUDA;
    }
    c->stream = c->own_stream;
    for (int i = 0; i < 2; ++i) {
        cudaEventCreateWithFlags(&c->ev_emit[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming);
    }
    if ((e = cudaMalloc((void**)&c->d_scan_ticket, sizeof(unsigned int))) != cudaSuccess) {
        cuda_fail(nullptr, e, "gm2_create: cudaMalloc"); delete c; return GM2_ERR_CUDA;
    }
    *out = c;
    return GM2_OK;
}

GM2_API int gm2_destroy(gm2_ctx* c) {
    if (!c) return GM2_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    void* frees[] = {c->d_seq, c->d_seq2, c->d_tile_slot, c->d_slot_src, c->d_slot_len, c->d_slot_cov, c->d_cov_ovf,
                     c->d_first_gene, c->d_next_same, c->d_forced_ids, c->d_force_keep, c->d_has_gene, c->d_counts, c->own_ids, c->own_ids_off, c->own_keep, c->d_segkept,
                     c->d_tile_off, c->d_len, c->d_rec_size, c->d_rec_off, c->d_scan_desc, c->d_scan_ticket,
                     c->d_stage[0], c->d_stage[1]};
    for (void* p : frees) if (p) cudaFree(p);
    if (c->h_len) cudaFreeHost(c->h_len);
    if (c->h_rec_off) cudaFreeHost(c->h_rec_off);
    for (int i = 0; i < 2; ++i) {
        if (c->ev_emit[i]) cudaEventDestroy(c->ev_emit[i]);
        if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]);
    }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
    return GM2_OK;
}

GM2_API int gm2_configure(gm2_ctx* c, int key, int64_t value) {
    if (!c) return GM2_ERR_INVALID;
    switch (key) {
    case GM2_CFG_TILE_BYTES:
        if (c->have_ref) return fail(c, GM2_ERR_STATE, "GM2_CFG_TILE_BYTES must be set before gm2_set_reference");
        if (value < 4096 || value > 196608 || (value % 4096) != 0)
            return fail(c, GM2_ERR_INVALID, "tile bytes must be a multiple of 4096 in [4096, 196608]");
        c->tile_bytes = (int)value; return GM2_OK;
    case GM2_CFG_EMIT_WARPS:
        if (value < 1 || value > 8) return fail(c, GM2_ERR_INVALID, "emit warps must be in 1..8");
        c->emit_warps = (int)value; return GM2_OK;
    case GM2_CFG_EMIT_BATCH:
        if (value < 0 || value > (1 << 20)) return fail(c, GM2_ERR_INVALID, "emit batch out of range");
        c->emit_batch = (int)value; return GM2_OK;
    case GM2_CFG_PACKING:
        if (c->have_ref) return fail(c, GM2_ERR_STATE, "GM2_CFG_PACKING must be set before gm2_set_reference");
        if (value < 0 || value > 2) return fail(c, GM2_ERR_INVALID, "packing must be 0 (auto: byte), 1 (byte) or 2 (two-bit, ACGT only)");
        c->packing_req = (int)value; return GM2_OK;
    case GM2_CFG_STORE_POLICY:
        if (value != 0 && value != 1) return fail(c, GM2_ERR_INVALID, "store policy must be 0 or 1");
        c->store_policy = (int)value; return GM2_OK;
    case GM2_CFG_ORDER:
        if (value != 0 && value != 1) return fail(c, GM2_ERR_INVALID, "order must be 0 (tile-major) or 1 (sample-major)");
        c->order = (int)value; return GM2_OK;
    case GM2_CFG_DEBUG:
        c->debug = (int)value; return GM2_OK;
    case GM2_CFG_RUN_TABLE:
        if (value < 32 || value > 1024 || (value & 1)) return fail(c, GM2_ERR_INVALID, "run table entries must be even and in 32..1024");
        c->rt_cap = (int)value; return GM2_OK;
    default:
        return fail(c, GM2_ERR_INVALID, "gm2_configure: unknown key");
    }
}

GM2_API int gm2_query(const gm2_ctx* c, int key, int64_t* out) {
    if (!c || !out) return GM2_ERR_INVALID;
    switch (key) {
    case GM2_Q_SM_COUNT:     *out = c->sm_count; return GM2_OK;
    case GM2_Q_LAUNCHES:     *out = (int64_t)c->launches; return GM2_OK;
    case GM2_Q_NUM_SEGMENTS: *out = c->nseg; return GM2_OK;
    case GM2_Q_NUM_TILES:    *out = c->ntiles; return GM2_OK;
    case GM2_Q_PACKING:      *out = c->packing; return GM2_OK;
    case GM2_Q_NUM_SLOTS:    *out = c->nslots; return GM2_OK;
    case GM2_Q_KEEP_WORDS:   *out = c->FW; return GM2_OK;
    default: return GM2_ERR_INVALID;
    }
}

GM2_API int gm2_set_stream(gm2_ctx* c, void* s) {
    if (!c) return GM2_ERR_INVALID;
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return GM2_OK;
}

GM2_API int gm2_sync(gm2_ctx* c) {
    if (!c) return GM2_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaStreamSynchronize(c->copy_stream));
    return GM2_OK;
}

GM2_API int gm2_set_header_prefix(gm2_ctx* c, const char* prefix) {
    if (!c || !prefix) return GM2_ERR_INVALID;
    if (strlen(prefix) > GM2_MAX_PREFIX - 1) return fail(c, GM2_ERR_INVALID, "header prefix too long");
    set_prefix(c, prefix);
    c->planned = false; c->host_plan = false;
    return GM2_OK;
}

GM2_API int gm2_set_reference(gm2_ctx* c, const uint8_t* seq, int64_t G,
                              const int64_t* gs, const int64_t* ge, int32_t F)
{
    if (!c) return GM2_ERR_INVALID;
    if (G < 0 || F < 0 || (G > 0 && !seq) || (F > 0 && (!gs || !ge)))
        return fail(c, GM2_ERR_INVALID, "gm2_set_reference: bad arguments");
    if (G > (int64_t)0x7fff0000) return fail(c, GM2_ERR_INVALID, "gm2_set_reference: G must be below 2^31 - 65536");
    CU(c, cudaSetDevice(c->device));
    const int64_t T = c->tile_bytes;
    const int ntiles = (int)((G + T - 1) / T);

    // breakpoints
    std::vector<int64_t> bp;
    bp.reserve((size_t)2 * F + ntiles + 2);
    bp.push_back(0); bp.push_back(G);
    std::vector<int64_t> ga((size_t)F), gb((size_t)F);
    for (int32_t g = 0; g < F; ++g) {
        int64_t a = std::min(std::max<int64_t>(gs[g], 0), G), b = std::min(std::max<int64_t>(ge[g], 0), G);
        if (a >= b) { a = b = 0; }                      // empty range: contributes nothing
        ga[g] = a; gb[g] = b;
        if (a < b) { bp.push_back(a); bp.push_back(b); }
    }
    for (int t = 1; t < ntiles; ++t) bp.push_back((int64_t)t * T);
    std::sort(bp.begin(), bp.end());
    bp.erase(std::unique(bp.begin(), bp.end()), bp.end());
    const int nseg = (int)bp.size() - 1;               // 0 when G == 0

    // cover lists (CSR over segments)
    std::vector<int32_t> cnt((size_t)nseg + 1, 0);
    std::vector<int32_t> ia((size_t)F), ib((size_t)F);
    for (int32_t g = 0; g < F; ++g) {
        if (ga[g] >= gb[g]) { ia[g] = ib[g] = 0; continue; }
        ia[g] = (int32_t)(std::lower_bound(bp.begin(), bp.end(), ga[g]) - bp.begin());
        ib[g] = (int32_t)(std::lower_bound(bp.begin(), bp.end(), gb[g]) - bp.begin());
        for (int32_t j = ia[g]; j < ib[g]; ++j) cnt[j]++;
    }
    std::vector<int64_t> seg_cov_off((size_t)nseg + 1, 0);
    for (int j = 0; j < nseg; ++j) seg_cov_off[j + 1] = seg_cov_off[j] + cnt[j];
    if (seg_cov_off[nseg] > (int64_t)0x7fffffff) return fail(c, GM2_ERR_INVALID, "gm2_set_reference: cover table too large");
    std::vector<int32_t> seg_cov((size_t)seg_cov_off[nseg]);
    {
        std::vector<int64_t> fill(seg_cov_off.begin(), seg_cov_off.end() - 1);
        for (int32_t g = 0; g < F; ++g)
            for (int32_t j = ia[g]; j < ib[g]; ++j) seg_cov[(size_t)fill[j]++] = g;
    }

    // slot layout: each tile's segments padded to a multiple of 32 slots
    std::vector<int32_t> tile_slot((size_t)ntiles + 1, 0);
    std::vector<int32_t> slot_src, slot_len, cov_ovf;
    std::vector<int2> slot_cov;
    slot_src.reserve((size_t)nseg + 32 * (size_t)ntiles);
    slot_len.reserve(slot_src.capacity());
    slot_cov.reserve(slot_src.capacity());
    int j = 0;
    for (int t = 0; t < ntiles; ++t) {
        tile_slot[t] = (int32_t)slot_src.size();
        const int64_t tend = std::min<int64_t>((int64_t)(t + 1) * T, G);
        while (j < nseg && bp[j] < tend) {
            slot_src.push_back((int32_t)bp[j]);
            slot_len.push_back((int32_t)(bp[j + 1] - bp[j]));
            {
                const int64_t k0 = seg_cov_off[j], nc = seg_cov_off[j + 1] - k0;
                int2 cv = make_int2(-1, -1);
                if (nc >= 1) cv.x = seg_cov[(size_t)k0];
                if (nc == 2) cv.y = seg_cov[(size_t)k0 + 1];
                if (nc > 2) {                               // rare: spill the rest to the overflow list
                    if (cov_ovf.size() + (size_t)nc > (size_t)0x7ffffff0) return fail(c, GM2_ERR_INVALID, "gm2_set_reference: cover table too large");
                    cv.y = -2 - (int32_t)cov_ovf.size();
                    cov_ovf.push_back((int32_t)(nc - 1));
                    for (int64_t k = 1; k < nc; ++k) cov_ovf.push_back(seg_cov[(size_t)(k0 + k)]);
                }
                slot_cov.push_back(cv);
            }
            ++j;
        }
        while (slot_src.size() % 32) {
            slot_src.push_back((int32_t)tend); slot_len.push_back(0);
            slot_cov.push_back(make_int2(-1, -1));
        }
    }
    tile_slot[ntiles] = (int32_t)slot_src.size();
    int max_tile_slots = 0;
    for (int t = 0; t < ntiles; ++t) max_tile_slots = std::max(max_tile_slots, tile_slot[t + 1] - tile_slot[t]);

    // upload
    if (c->d_seq) { cudaFree(c->d_seq); c->d_seq = nullptr; }
    const size_t seq_alloc = (size_t)ntiles * (size_t)T + 256;
    CU(c, cudaMalloc((void**)&c->d_seq, seq_alloc));
    CU(c, cudaMemset(c->d_seq, 0, seq_alloc));
    if (G > 0) CU(c, cudaMemcpy(c->d_seq, seq, (size_t)G, cudaMemcpyHostToDevice));
    int rc;
    if ((rc = dev_upload(c, &c->d_tile_slot, tile_slot))) return rc;
    if ((rc = dev_upload(c, &c->d_slot_src, slot_src))) return rc;
    if ((rc = dev_upload(c, &c->d_slot_len, slot_len))) return rc;
    if ((rc = dev_upload(c, &c->d_slot_cov, slot_cov))) return rc;
    if ((rc = dev_upload(c, &c->d_cov_ovf, cov_ovf))) return rc;

    c->G = G; c->F = F; c->FW = (F + 31) / 32;
    c->ntiles = ntiles; c->nseg = nseg; c->nslots = (int)slot_src.size(); c->SW = c->nslots / 32;
    c->packing = 1; c->max_tile_slots = max_tile_slots;
    if (c->d_seq2) { cudaFree(c->d_seq2); c->d_seq2 = nullptr; }
    if (c->packing_req == 2) {
        // measured slower than bytes (profiles/r01_emit_experiments.md) — kept as a selectable form
        std::vector<uint8_t> packed((size_t)ntiles * (size_t)(T / 4) + 256, 0);
        for (int64_t i = 0; i < G; ++i) {
            uint8_t code;
            switch (seq[i]) {
            case 'A': code = 0; break; case 'C': code = 1; break; case 'G': code = 2; break; case 'T': code = 3; break;
            default: return fail(c, GM2_ERR_INVALID, "gm2_set_reference: two-bit packing needs an upper-case ACGT-only sequence");
            }
            packed[(size_t)(i >> 2)] |= (uint8_t)(code << (2 * (i & 3)));
        }
        if ((rc = dev_upload(c, &c->d_seq2, packed))) return rc;
        c->packing = 2;
    }
    c->have_ref = true; c->planned = false; c->host_plan = false; c->mode = 0; c->S = 0;
    return GM2_OK;
}

GM2_API int gm2_set_name_map(gm2_ctx* c, const int32_t* off, const int32_t* idx, int32_t V) {
    if (!c) return GM2_ERR_INVALID;
    if (!c->have_ref) return fail(c, GM2_ERR_STATE, "gm2_set_name_map: call gm2_set_reference first");
    if (V < 0 || !off) return fail(c, GM2_ERR_INVALID, "gm2_set_name_map: bad arguments");
    if (off[0] != 0) return fail(c, GM2_ERR_INVALID, "gm2_set_name_map: off[0] must be 0");
    for (int32_t i = 0; i < V; ++i) if (off[i + 1] < off[i]) return fail(c, GM2_ERR_INVALID, "gm2_set_name_map: offsets must be non-decreasing");
    const int32_t n = off[V];
    if (n > 0 && !idx) return fail(c, GM2_ERR_INVALID, "gm2_set_name_map: idx is NULL");
    for (int32_t i = 0; i < n; ++i) if (idx[i] < 0 || idx[i] >= c->F) return fail(c, GM2_ERR_INVALID, "gm2_set_name_map: gene index out of range");
    CU(c, cudaSetDevice(c->device));
    // device form: first_gene[id] -> next_same_name[g] -> ... -> -1 (list in CSR order)
    std::vector<int32_t> first((size_t)V, -1), next((size_t)c->F, -1);
    std::vector<char> seen((size_t)c->F, 0);
    for (int32_t id = 0; id < V; ++id) {
        int32_t prev = -1;
        for (int32_t k = off[id]; k < off[id + 1]; ++k) {
            const int32_t g = idx[k];
            if (seen[g]) return fail(c, GM2_ERR_INVALID, "gm2_set_name_map: a gene is listed under two names");
            seen[g] = 1;
            if (prev < 0) first[id] = g; else next[prev] = g;
            prev = g;
        }
    }
    int rc;
    if ((rc = dev_upload(c, &c->d_first_gene, first))) return rc;
    if ((rc = dev_upload(c, &c->d_next_same, next))) return rc;
    {
        std::vector<uint32_t> hg(((size_t)V + 31) / 32 + 1, 0u);
        for (int32_t id = 0; id < V; ++id) if (first[id] >= 0) hg[(size_t)id >> 5] |= 1u << (id & 31);
        if ((rc = dev_upload(c, &c->d_has_gene, hg))) return rc;
    }
    c->V = V;
    if (c->d_forced_ids) { cudaFree(c->d_forced_ids); c->d_forced_ids = nullptr; }
    if (c->d_force_keep) { cudaFree(c->d_force_keep); c->d_force_keep = nullptr; }
    return GM2_OK;
}

static int begin_samples(gm2_ctx* c, int64_t S, const char* who) {
    if (!c) return GM2_ERR_INVALID;
    if (!c->have_ref) return fail(c, GM2_ERR_STATE, std::string(who) + ": call gm2_set_reference first");
    if (S < 0) return fail(c, GM2_ERR_INVALID, std::string(who) + ": S < 0");
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaSetDevice");
    c->planned = false; c->host_plan = false;
    return GM2_OK;
}

GM2_API int gm2_load_ids_host(gm2_ctx* c, const int32_t* ids, const int64_t* off, int64_t S) {
    int rc = begin_samples(c, S, "gm2_load_ids_host"); if (rc) return rc;
    if (!c->d_first_gene) return fail(c, GM2_ERR_STATE, "gm2_load_ids_host: call gm2_set_name_map first");
    if (!off) return fail(c, GM2_ERR_INVALID, "gm2_load_ids_host: off is NULL");
    if (off[0] != 0) return fail(c, GM2_ERR_INVALID, "gm2_load_ids_host: off[0] must be 0");
    for (int64_t s = 0; s < S; ++s) if (off[s + 1] < off[s]) return fail(c, GM2_ERR_INVALID, "gm2_load_ids_host: offsets must be non-decreasing");
    const int64_t n = off[S];
    if (n > 0 && !ids) return fail(c, GM2_ERR_INVALID, "gm2_load_ids_host: ids is NULL");
    if ((rc = dev_reserve(c, &c->own_ids, &c->own_ids_cap, n))) return rc;
    if ((rc = dev_reserve(c, &c->own_ids_off, &c->own_ids_off_cap, S + 1))) return rc;
    if (n > 0) CU(c, cudaMemcpyAsync(c->own_ids, ids, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->own_ids_off, off, (size_t)(S + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));       // host buffers are the caller's: do not outlive the call
    c->ids = c->own_ids; c->ids_off = c->own_ids_off; c->S = S; c->mode = 1;
    return GM2_OK;
}

GM2_API int gm2_load_ids_dev(gm2_ctx* c, const int32_t* ids, const int64_t* off, int64_t S, int64_t n_ids) {
    int rc = begin_samples(c, S, "gm2_load_ids_dev"); if (rc) return rc;
    if (!c->d_first_gene) return fail(c, GM2_ERR_STATE, "gm2_load_ids_dev: call gm2_set_name_map first");
    if (!off || (n_ids > 0 && !ids) || n_ids < 0) return fail(c, GM2_ERR_INVALID, "gm2_load_ids_dev: bad arguments");
    if (((uintptr_t)ids & 15) || ((uintptr_t)off & 7)) return fail(c, GM2_ERR_INVALID, "gm2_load_ids_dev: ids must be 16-byte aligned, off 8-byte aligned");
    c->ids = ids; c->ids_off = off; c->S = S; c->mode = 1;
    return GM2_OK;
}

GM2_API int gm2_set_forced(gm2_ctx* c, const uint32_t* force_keep, const uint32_t* forced_ids) {
    if (!c) return GM2_ERR_INVALID;
    if (!c->d_first_gene) return fail(c, GM2_ERR_STATE, "gm2_set_forced: call gm2_set_name_map first");
    CU(c, cudaSetDevice(c->device));
    if (c->d_forced_ids) { cudaFree(c->d_forced_ids); c->d_forced_ids = nullptr; }
    if (c->d_force_keep) { cudaFree(c->d_force_keep); c->d_force_keep = nullptr; }
    int rc;
    if (force_keep) {
        std::vector<uint32_t> v(force_keep, force_keep + c->FW);
        if (c->F & 31) { if (c->FW > 0) v[c->FW - 1] &= (1u << (c->F & 31)) - 1u; }     // bits beyond F stay clear
        if ((rc = dev_upload(c, &c->d_force_keep, v))) return rc;
    }
    if (forced_ids) {
        std::vector<uint32_t> v(forced_ids, forced_ids + (c->V + 31) / 32);
        if ((rc = dev_upload(c, &c->d_forced_ids, v))) return rc;
    }
    c->planned = false; c->host_plan = false;
    return GM2_OK;
}

GM2_API int gm2_load_probs_dev(gm2_ctx* c, const float* probs, int64_t S, int64_t ld, float threshold) {
    int rc = begin_samples(c, S, "gm2_load_probs_dev"); if (rc) return rc;
    if (!c->d_first_gene) return fail(c, GM2_ERR_STATE, "gm2_load_probs_dev: call gm2_set_name_map first (columns are name ids)");
    if ((S > 0 && !probs) || ld < c->V || ((uintptr_t)probs & 3)) return fail(c, GM2_ERR_INVALID, "gm2_load_probs_dev: bad arguments (ld >= V, 4-byte aligned pointer)");
    c->probs = probs; c->probs_ld = ld; c->probs_thr = threshold; c->S = S; c->mode = 3;
    return GM2_OK;
}

GM2_API int gm2_get_counts(gm2_ctx* c, int64_t* out) {
    if (!c) return GM2_ERR_INVALID;
    if (!c->planned || c->mode != 3) return fail(c, GM2_ERR_STATE, "gm2_get_counts: needs a plan over gm2_load_probs_dev samples");
    CU(c, cudaSetDevice(c->device));
    if (c->S > 0) {
        if (!out) return fail(c, GM2_ERR_INVALID, "gm2_get_counts: out is NULL");
        CU(c, cudaMemcpyAsync(out, c->d_counts, (size_t)c->S * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    return GM2_OK;
}

GM2_API int gm2_load_keep_host(gm2_ctx* c, const uint32_t* rows, int64_t S) {
    int rc = begin_samples(c, S, "gm2_load_keep_host"); if (rc) return rc;
    const int64_t n = S * c->FW;
    if (n > 0 && !rows) return fail(c, GM2_ERR_INVALID, "gm2_load_keep_host: rows is NULL");
    if ((rc = dev_reserve(c, &c->own_keep, &c->own_keep_cap, n))) return rc;
    if (n > 0) {
        CU(c, cudaMemcpyAsync(c->own_keep, rows, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    c->keep_in = c->own_keep; c->S = S; c->mode = 2;
    return GM2_OK;
}

GM2_API int gm2_load_keep_dev(gm2_ctx* c, const uint32_t* rows, int64_t S) {
    int rc = begin_samples(c, S, "gm2_load_keep_dev"); if (rc) return rc;
    if (S * c->FW > 0 && !rows) return fail(c, GM2_ERR_INVALID, "gm2_load_keep_dev: rows is NULL");
    c->keep_in = rows; c->S = S; c->mode = 2;
    return GM2_OK;
}

#define LAUNCH_CHECK(c, name) do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return cuda_fail((c), e__, name); (c)->launches++; } while (0)

GM2_API int gm2_plan_async(gm2_ctx* c, int64_t first_idx) {
    if (!c) return GM2_ERR_INVALID;
    if (!c->have_ref || c->mode == 0) return fail(c, GM2_ERR_STATE, "gm2_plan: load a reference and samples first");
    if (first_idx < 0) return fail(c, GM2_ERR_INVALID, "gm2_plan: first_idx < 0");
    CU(c, cudaSetDevice(c->device));
    const int64_t S = c->S;
    c->first_idx = first_idx; c->planned = false; c->host_plan = false;
    int rc;
    if ((rc = dev_reserve(c, &c->d_segkept, &c->segkept_cap, S * c->SW))) return rc;
    if ((rc = dev_reserve(c, &c->d_tile_off, &c->tile_off_cap, S * (int64_t)c->ntiles))) return rc;
    if ((rc = dev_reserve(c, &c->d_len, &c->len_cap, S + 1))) return rc;
    if ((rc = dev_reserve(c, &c->d_rec_size, &c->rec_size_cap, S + 1))) return rc;
    if ((rc = dev_reserve(c, &c->d_rec_off, &c->rec_off_cap, S + 1))) return rc;
    const uint32_t* keep = c->keep_in;
    if (c->mode == 3) {
        if ((rc = dev_reserve(c, &c->own_keep, &c->own_keep_cap, S * c->FW))) return rc;
        if ((rc = dev_reserve(c, &c->d_counts, &c->counts_cap, S))) return rc;
        keep = c->own_keep;
        if (S > 0) {
            const int VW = (c->V + 31) / 32;
            const size_t sm = ((size_t)c->FW + 2 * ((size_t)VW + 1)) * 4;
            if (sm > 200 * 1024) return fail(c, GM2_ERR_INVALID, "gm2_plan: too many name ids for the dense keep builder's bitmaps");
            if (sm > 48 * 1024) CU(c, cudaFuncSetAttribute(k_keep_from_probs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            k_keep_from_probs<<<(unsigned)S, 256, sm, c->stream>>>(c->probs, S, c->V, c->probs_ld, c->probs_thr, c->d_first_gene,
                                                                c->d_next_same, c->d_has_gene, c->d_forced_ids, c->d_force_keep,
                                                                c->FW, VW, c->own_keep, c->d_counts);
            LAUNCH_CHECK(c, "k_keep_from_probs");
        }
    }
    if (c->mode == 1) {
        if ((rc = dev_reserve(c, &c->own_keep, &c->own_keep_cap, S * c->FW))) return rc;
        keep = c->own_keep;
        if (S > 0 && c->FW > 0) {
            const size_t rows_b = (size_t)K1_WARPS * c->FW * 4;
            const size_t map_b = ((size_t)c->V + (size_t)c->F) * 4;
            const int map_in_smem = rows_b + map_b <= 96 * 1024;
            const size_t sm = rows_b + (map_in_smem ? map_b : 0);
            if (sm > 200 * 1024) return fail(c, GM2_ERR_INVALID, "gm2_plan: too many genes for the keep-row staging buffer");
            CU(c, cudaFuncSetAttribute(k_keep_from_ids, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            const int64_t groups = (S + K1_WARPS - 1) / K1_WARPS;
            const int64_t blocks = std::min<int64_t>(groups, (int64_t)c->sm_count * (map_in_smem ? 4 : 8));
            k_keep_from_ids<<<(unsigned)blocks, K1_WARPS * 32, sm, c->stream>>>(c->ids, c->ids_off, S, c->V, c->F,
                                                                              c->d_first_gene, c->d_next_same, c->FW,
                                                                              c->own_keep, map_in_smem);
            LAUNCH_CHECK(c, "k_keep_from_ids");
        }
    }
    if (S > 0) {
        const size_t sm = ((size_t)c->FW + (size_t)c->ntiles + (size_t)c->SW) * 4 * PLAN_NS;
        if (sm > 200 * 1024) return fail(c, GM2_ERR_INVALID, "gm2_plan: genome has too many genes/tiles for one CTA's shared memory");
        CU(c, cudaFuncSetAttribute(k_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        const int64_t blocks = (S + PLAN_NS - 1) / PLAN_NS;
        k_plan<<<(unsigned)blocks, 256, sm, c->stream>>>(S, c->FW, keep, c->ntiles, c->d_tile_slot, c->d_slot_len,
                                                         c->d_slot_cov, c->d_cov_ovf, c->SW, c->d_segkept, c->d_tile_off,
                                                         c->d_len, c->d_rec_size, first_idx, c->prefix.len);
        LAUNCH_CHECK(c, "k_plan");
    }
    {
        const int64_t ntile = std::max<int64_t>((S + SCAN_TILE - 1) / SCAN_TILE, 1);
        if ((rc = dev_reserve(c, &c->d_scan_desc, &c->scan_desc_cap, ntile))) return rc;
        CU(c, cudaMemsetAsync(c->d_scan_desc, 0, (size_t)ntile * 8, c->stream));
        CU(c, cudaMemsetAsync(c->d_scan_ticket, 0, sizeof(unsigned int), c->stream));
        k_scan_records<<<(unsigned)ntile, SCAN_THREADS, 0, c->stream>>>(c->d_rec_size, c->d_rec_off, S, c->d_scan_desc, c->d_scan_ticket);
        LAUNCH_CHECK(c, "k_scan_records");
    }
    c->planned = true;
    return GM2_OK;
}

static int pull_plan(gm2_ctx* c) {
    if (!c->planned) return fail(c, GM2_ERR_STATE, "no plan: call gm2_plan first");
    if (c->host_plan) return GM2_OK;
    const int64_t S = c->S;
    if (S + 1 > c->h_cap) {
        if (c->h_len) cudaFreeHost(c->h_len);
        if (c->h_rec_off) cudaFreeHost(c->h_rec_off);
        c->h_len = c->h_rec_off = nullptr; c->h_cap = 0;
        CU(c, cudaMallocHost((void**)&c->h_len, (size_t)(S + 1) * 8));
        CU(c, cudaMallocHost((void**)&c->h_rec_off, (size_t)(S + 1) * 8));
        c->h_cap = S + 1;
    }
    if (S > 0) CU(c, cudaMemcpyAsync(c->h_len, c->d_len, (size_t)S * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(c->h_rec_off, c->d_rec_off, (size_t)(S + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->host_plan = true;
    return GM2_OK;
}

GM2_API int gm2_plan(gm2_ctx* c, int64_t first_idx) {
    int rc = gm2_plan_async(c, first_idx); if (rc) return rc;
    return pull_plan(c);
}

GM2_API int gm2_get_lengths(gm2_ctx* c, int64_t* out) {
    if (!c) return GM2_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    int rc = pull_plan(c); if (rc) return rc;
    if (c->S > 0) { if (!out) return fail(c, GM2_ERR_INVALID, "gm2_get_lengths: out is NULL"); memcpy(out, c->h_len, (size_t)c->S * 8); }
    return GM2_OK;
}
GM2_API int gm2_get_record_offsets(gm2_ctx* c, int64_t* out) {
    if (!c || !out) return GM2_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    int rc = pull_plan(c); if (rc) return rc;
    memcpy(out, c->h_rec_off, (size_t)(c->S + 1) * 8);
    return GM2_OK;
}
GM2_API int gm2_get_keep_rows(gm2_ctx* c, uint32_t* out) {
    if (!c) return GM2_ERR_INVALID;
    if (!c->planned) return fail(c, GM2_ERR_STATE, "gm2_get_keep_rows: call gm2_plan first");
    CU(c, cudaSetDevice(c->device));
    const uint32_t* keep = c->mode == 2 ? c->keep_in : c->own_keep;
    const int64_t n = c->S * c->FW;
    if (n > 0) {
        if (!out) return fail(c, GM2_ERR_INVALID, "gm2_get_keep_rows: out is NULL");
        CU(c, cudaMemcpyAsync(out, keep, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    return GM2_OK;
}
GM2_API int gm2_image_bytes(gm2_ctx* c, int64_t s0, int64_t s1, int64_t* out) {
    if (!c || !out) return GM2_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    int rc = pull_plan(c); if (rc) return rc;
    if (s0 < 0 || s1 < s0 || s1 > c->S) return fail(c, GM2_ERR_INVALID, "gm2_image_bytes: bad sample range");
    *out = c->h_rec_off[s1] - c->h_rec_off[s0];
    return GM2_OK;
}

static int launch_emit(gm2_ctx* c, int64_t s0, int64_t s1, uint8_t* dev_out) {
    const int64_t n = s1 - s0;
    if (n <= 0) return GM2_OK;
    const int tiles = std::max(c->ntiles, 1);
    const int warps = c->emit_warps;
    int64_t batch = c->emit_batch;
    if (batch <= 0) {
        // aim for >= ~8 CTAs per SM over the whole grid, at least one sample per warp
        const int64_t want_ctas = (int64_t)c->sm_count * 8;
        batch = (n * tiles + want_ctas - 1) / want_ctas;
        batch = std::max<int64_t>(batch, warps);
        batch = std::min<int64_t>(batch, 64);
        batch = ((batch + warps - 1) / warps) * warps;
    }
    const int64_t nbatch = (n + batch - 1) / batch;
    const int64_t blocks = nbatch * tiles;
    if (blocks > 0x7fffffffLL) return fail(c, GM2_ERR_INVALID, "gm2_emit: grid too large; emit a smaller sample range");
    EmitParams p;
    const bool two_bit = c->packing == 2;
    p.seq = two_bit ? c->d_seq2 : c->d_seq; p.tile_smem_bytes = two_bit ? c->tile_bytes / 4 : c->tile_bytes;
    p.tile_slot = c->d_tile_slot; p.slot_src = c->d_slot_src; p.slot_len = c->d_slot_len;
    p.segkept = c->d_segkept; p.tile_off = c->d_tile_off; p.lengths = c->d_len; p.rec_off = c->d_rec_off;
    p.out = dev_out; p.s0 = s0; p.s1 = s1; p.first_idx = c->first_idx;
    p.tile_bytes = c->tile_bytes; p.ntiles = c->ntiles; p.SW = c->SW; p.batch = (int)batch; p.nbatch = (int)nbatch;
    p.rt_cap = c->rt_cap;
    p.slot_cap = c->max_tile_slots <= 4096 ? c->max_tile_slots : 0;     // else: slot tables read from global
    p.prefix = c->prefix; p.debug = c->debug; p.order = c->order;
    const size_t sm = 32 + (size_t)p.tile_smem_bytes + 64 + (size_t)p.slot_cap * 8 + (size_t)warps * (p.rt_cap + 2) * 24;
    if (sm > 227 * 1024) return fail(c, GM2_ERR_INVALID, "gm2_emit: shared memory budget exceeded; lower tile bytes / emit warps");
    // register budget follows the shared-memory footprint: small tiles -> 4+ CTAs/SM (64 regs),
    // large tiles -> 3 CTAs/SM (up to 85 regs)
    const bool dense = sm <= 56 * 1024;
    void (*kern)(const EmitParams);
    if (two_bit)
        kern = c->store_policy == 1 ? (dense ? k_emit<1, 4, 2> : k_emit<1, 3, 2>) : (dense ? k_emit<0, 4, 2> : k_emit<0, 3, 2>);
    else
        kern = c->store_policy == 1 ? (dense ? k_emit<1, 4, 1> : k_emit<1, 3, 1>) : (dense ? k_emit<0, 4, 1> : k_emit<0, 3, 1>);
    CU(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kern<<<(unsigned)blocks, warps * 32, sm, c->stream>>>(p);
    LAUNCH_CHECK(c, "k_emit");
    return GM2_OK;
}

GM2_API int gm2_emit_dev(gm2_ctx* c, int64_t s0, int64_t s1, uint8_t* dev_out, int64_t cap) {
    if (!c) return GM2_ERR_INVALID;
    if (!c->planned) return fail(c, GM2_ERR_STATE, "gm2_emit_dev: call gm2_plan first");
    if (s0 < 0 || s1 < s0 || s1 > c->S) return fail(c, GM2_ERR_INVALID, "gm2_emit_dev: bad sample range");
    if (s1 > s0 && !dev_out) return fail(c, GM2_ERR_INVALID, "gm2_emit_dev: dev_out is NULL");
    CU(c, cudaSetDevice(c->device));
    if (c->host_plan) {
        const int64_t need = c->h_rec_off[s1] - c->h_rec_off[s0];
        if (need > cap) return fail(c, GM2_ERR_CAPACITY, "gm2_emit_dev: output buffer too small");
    }
    return launch_emit(c, s0, s1, dev_out);
}

GM2_API int gm2_emit_host(gm2_ctx* c, int64_t s0, int64_t s1, uint8_t* host_out, int64_t cap, int64_t chunk_bytes) {
    if (!c) return GM2_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    int rc = pull_plan(c); if (rc) return rc;
    if (s0 < 0 || s1 < s0 || s1 > c->S) return fail(c, GM2_ERR_INVALID, "gm2_emit_host: bad sample range");
    const int64_t total = c->h_rec_off[s1] - c->h_rec_off[s0];
    if (total > cap) return fail(c, GM2_ERR_CAPACITY, "gm2_emit_host: output buffer too small");
    if (total > 0 && !host_out) return fail(c, GM2_ERR_INVALID, "gm2_emit_host: host_out is NULL");
    if (chunk_bytes <= 0) chunk_bytes = (int64_t)256 << 20;
    // staging must hold the largest single record of the range
    int64_t need = std::min(chunk_bytes, total);
    for (int64_t s = s0; s < s1; ++s) need = std::max(need, c->h_rec_off[s + 1] - c->h_rec_off[s]);
    if (need > c->stage_cap) {
        for (int i = 0; i < 2; ++i) { if (c->d_stage[i]) cudaFree(c->d_stage[i]); c->d_stage[i] = nullptr; }
        c->stage_cap = 0;
        for (int i = 0; i < 2; ++i) CU(c, cudaMalloc((void**)&c->d_stage[i], (size_t)need));
        c->stage_cap = need;
    }
    int64_t a = s0; int i = 0;
    while (a < s1) {
        int64_t b = a + 1;
        while (b < s1 && c->h_rec_off[b + 1] - c->h_rec_off[a] <= c->stage_cap &&
               c->h_rec_off[b + 1] - c->h_rec_off[a] <= chunk_bytes) ++b;
        const int buf = i & 1;
        if (i >= 2) CU(c, cudaStreamWaitEvent(c->stream, c->ev_copy[buf], 0));
        if ((rc = launch_emit(c, a, b, c->d_stage[buf]))) return rc;
        CU(c, cudaEventRecord(c->ev_emit[buf], c->stream));
        CU(c, cudaStreamWaitEvent(c->copy_stream, c->ev_emit[buf], 0));
        const int64_t bytes = c->h_rec_off[b] - c->h_rec_off[a];
        CU(c, cudaMemcpyAsync(host_out + (c->h_rec_off[a] - c->h_rec_off[s0]), c->d_stage[buf], (size_t)bytes,
                              cudaMemcpyDeviceToHost, c->copy_stream));
        CU(c, cudaEventRecord(c->ev_copy[buf], c->copy_stream));
        a = b; ++i;
    }
    CU(c, cudaStreamSynchronize(c->copy_stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return GM2_OK;
}

GM2_API int gm2_minimize_host(gm2_ctx* c, const int32_t* ids, const int64_t* off, const uint32_t* keep_rows,
                              int64_t S, int64_t first_idx, int64_t* lengths, int64_t* rec_off,
                              uint8_t* host_out, int64_t cap, int64_t chunk_bytes)
{
    if (!c) return GM2_ERR_INVALID;
    int rc;
    if (keep_rows) rc = gm2_load_keep_host(c, keep_rows, S);
    else           rc = gm2_load_ids_host(c, ids, off, S);
    if (rc) return rc;
    if ((rc = gm2_plan(c, first_idx))) return rc;
    if (lengths && S > 0) memcpy(lengths, c->h_len, (size_t)S * 8);
    if (rec_off) memcpy(rec_off, c->h_rec_off, (size_t)(S + 1) * 8);
    return gm2_emit_host(c, 0, S, host_out, cap, chunk_bytes);
}

GM2_API int gm2_host_alloc(void** out, int64_t bytes) {
    if (!out || bytes < 0) return GM2_ERR_INVALID;
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, (size_t)std::max<int64_t>(bytes, 1));
    if (e != cudaSuccess) { cuda_fail(nullptr, e, "cudaMallocHost"); return GM2_ERR_NOMEM; }
    return GM2_OK;
}
GM2_API int gm2_host_free(void* p) {
    if (p) cudaFreeHost(p);
    return GM2_OK;
}

GM2_API int gm2_diag_fill(gm2_ctx* c, uint8_t* dev, int64_t bytes, uint32_t pattern) {
    if (!c || !dev || bytes < 0 || ((uintptr_t)dev & 15)) return fail(c, GM2_ERR_INVALID, "gm2_diag_fill: bad arguments (16-byte aligned pointer required)");
    CU(c, cudaSetDevice(c->device));
    const int64_t nvec = bytes / 16;
    if (nvec == 0) return GM2_OK;
    const int blocks = c->sm_count * 16;
    k_fill<<<blocks, 256, 0, c->stream>>>(reinterpret_cast<uint4*>(dev), nvec, pattern);
    LAUNCH_CHECK(c, "k_fill");
    return GM2_OK;
}

// Hashes of the minimized SEQUENCES (bases only, no header / newline) of records [s0,s1), computed
// on the device from staged emits: backs the reference's duplicate report
// (check_sequence_duplicates, minimizer_2.py:273-303) without moving any base to the host.
GM2_API int gm2_sequence_hashes(gm2_ctx* c, int64_t s0, int64_t s1, uint64_t* out) {
    if (!c) return GM2_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    int rc = pull_plan(c); if (rc) return rc;
    if (s0 < 0 || s1 < s0 || s1 > c->S) return fail(c, GM2_ERR_INVALID, "gm2_sequence_hashes: bad sample range");
    if (s1 > s0 && !out) return fail(c, GM2_ERR_INVALID, "gm2_sequence_hashes: out is NULL");
    const int64_t chunk = (int64_t)1 << 30;
    int64_t need = 0;
    for (int64_t s = s0; s < s1; ++s) need = std::max(need, c->h_rec_off[s + 1] - c->h_rec_off[s]);
    need = std::max(need, std::min(chunk, c->h_rec_off[s1] - c->h_rec_off[s0]));
    need = (need + 7) & ~(int64_t)7;
    if (need > c->stage_cap) {
        for (int i = 0; i < 2; ++i) { if (c->d_stage[i]) cudaFree(c->d_stage[i]); c->d_stage[i] = nullptr; }
        c->stage_cap = 0;
        for (int i = 0; i < 2; ++i) CU(c, cudaMalloc((void**)&c->d_stage[i], (size_t)need));
        c->stage_cap = need;
    }
    int64_t a = s0;
    while (a < s1) {
        int64_t b = a + 1;
        while (b < s1 && c->h_rec_off[b + 1] - c->h_rec_off[a] <= c->stage_cap) ++b;
        if ((rc = launch_emit(c, a, b, c->d_stage[0]))) return rc;
        // sequence of record s = [rec start + header, rec end - 1); header = record size - L - 1
        const int64_t n = b - a;
        int64_t* d_off = nullptr; unsigned long long* d_out = nullptr;
        std::vector<int64_t> ranges((size_t)n * 2);
        for (int64_t i = 0; i < n; ++i) {
            const int64_t r1 = c->h_rec_off[a + i + 1] - c->h_rec_off[a];
            ranges[(size_t)(2 * i)] = r1 - 1 - c->h_len[a + i];
            ranges[(size_t)(2 * i + 1)] = r1 - 1;
        }
        CU(c, cudaMalloc((void**)&d_off, (size_t)n * 16));
        cudaError_t e = cudaMalloc((void**)&d_out, (size_t)n * 8);
        if (e != cudaSuccess) { cudaFree(d_off); return cuda_fail(c, e, "cudaMalloc"); }
        cudaMemcpyAsync(d_off, ranges.data(), (size_t)n * 16, cudaMemcpyHostToDevice, c->stream);
        cudaMemsetAsync(d_out, 0, (size_t)n * 8, c->stream);
        dim3 grid((unsigned)n, 32, 1);
        k_range_hashes_pairs<<<grid, 256, 0, c->stream>>>(c->d_stage[0], c->stage_cap, d_off, d_out);
        e = cudaGetLastError();
        if (e == cudaSuccess) { c->launches++; e = cudaMemcpyAsync(out + (a - s0), d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream); }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        cudaFree(d_off); cudaFree(d_out);
        if (e != cudaSuccess) return cuda_fail(c, e, "gm2_sequence_hashes");
        a = b;
    }
    return GM2_OK;
}

GM2_API int gm2_diag_fill_streams(gm2_ctx* c, uint8_t* dev, int64_t nrec, int64_t stride, int32_t ntile, int64_t chunk,
                                  int32_t batch, int32_t warps, int32_t order, int32_t vec32)
{
    if (!c || !dev || nrec <= 0 || ntile <= 0 || batch <= 0 || warps < 1 || warps > 8 || ((uintptr_t)dev & 31) || (stride & 31) || (chunk & 31))
        return fail(c, GM2_ERR_INVALID, "gm2_diag_fill_streams: bad arguments");
    CU(c, cudaSetDevice(c->device));
    const int64_t nbatch = (nrec + batch - 1) / batch;
    k_fill_streams<<<(unsigned)(nbatch * ntile), warps * 32, 0, c->stream>>>(dev, nrec, stride, ntile, chunk, batch, (int)nbatch, order, vec32);
    LAUNCH_CHECK(c, "k_fill_streams");
    return GM2_OK;
}

GM2_API int gm2_diag_range_hashes(gm2_ctx* c, const uint8_t* dev, int64_t dev_bytes, const int64_t* off,
                                  int64_t n, uint64_t* out)
{
    if (!c || n < 0 || (n > 0 && (!dev || !off || !out)) || ((uintptr_t)dev & 7))
        return fail(c, GM2_ERR_INVALID, "gm2_diag_range_hashes: bad arguments (8-byte aligned pointer required)");
    if (n == 0) return GM2_OK;
    for (int64_t i = 0; i < n; ++i)
        if (off[i] < 0 || off[i + 1] < off[i] || off[i + 1] > dev_bytes)
            return fail(c, GM2_ERR_INVALID, "gm2_diag_range_hashes: offsets out of range");
    CU(c, cudaSetDevice(c->device));
    int64_t* d_off = nullptr; unsigned long long* d_out = nullptr;
    CU(c, cudaMalloc((void**)&d_off, (size_t)(n + 1) * 8));
    cudaError_t e = cudaMalloc((void**)&d_out, (size_t)n * 8);
    if (e != cudaSuccess) { cudaFree(d_off); return cuda_fail(c, e, "cudaMalloc"); }
    cudaMemcpyAsync(d_off, off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->stream);
    cudaMemsetAsync(d_out, 0, (size_t)n * 8, c->stream);
    int rc = GM2_OK;
    const int64_t maxx = 1 << 30;
    for (int64_t r0 = 0; r0 < n && rc == GM2_OK; r0 += maxx) {
        const int64_t cnt = std::min(maxx, n - r0);
        dim3 grid((unsigned)cnt, 32, 1);
        k_range_hashes<<<grid, 256, 0, c->stream>>>(dev, dev_bytes, d_off + r0, d_out + r0);
        e = cudaGetLastError();
        if (e != cudaSuccess) rc = cuda_fail(c, e, "k_range_hashes"); else c->launches++;
    }
    if (rc == GM2_OK) {
        e = cudaMemcpyAsync(out, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(c, e, "gm2_diag_range_hashes: copy back");
    }
    cudaFree(d_off); cudaFree(d_out);
    return rc;
}
