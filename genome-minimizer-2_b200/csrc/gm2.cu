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
    int tile_bytes = 65536;
    int emit_warps = 8;
    int emit_batch = 0;
    int packing_req = 0;
    int store_policy = 0;
    HeaderPrefix prefix;

    // reference
    bool have_ref = false;
    int64_t G = 0;
    int32_t F = 0, FW = 0;
    int ntiles = 0, nseg = 0, nslots = 0, SW = 0;
    int packing = 1;
    uint8_t* d_seq = nullptr;
    int32_t *d_tile_slot = nullptr, *d_slot_src = nullptr, *d_slot_len = nullptr;
    int32_t *d_cov_off = nullptr, *d_cov_idx = nullptr;

    // name map
    int32_t V = 0;
    int32_t *d_map_off = nullptr, *d_map_idx = nullptr;

    // samples
    int64_t S = 0;
    int mode = 0;                      // 0 none, 1 ids, 2 keep rows
    const int32_t* ids = nullptr;      // device (owned or borrowed)
    const int64_t* ids_off = nullptr;
    const uint32_t* keep_in = nullptr; // device keep rows when mode == 2
    int32_t* own_ids = nullptr;   int64_t own_ids_cap = 0;
    int64_t* own_ids_off = nullptr; int64_t own_ids_off_cap = 0;
    uint32_t* own_keep = nullptr; int64_t own_keep_cap = 0;   // words

    // plan outputs (device)
    uint32_t* d_segkept = nullptr; int64_t segkept_cap = 0;    // words
    int32_t* d_tile_off = nullptr; int64_t tile_off_cap = 0;   // elements
    int64_t *d_len = nullptr, *d_rec_size = nullptr, *d_rec_off = nullptr; int64_t rec_cap = 0;
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
//   one warp per sample; the row is assembled in shared memory with atomicOr (an id
//   maps to 0..n genes through the static CSR), then written out coalesced.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_keep_from_ids(const int32_t* __restrict__ ids, const int64_t* __restrict__ off, int64_t S, int32_t V,
                const int32_t* __restrict__ map_off, const int32_t* __restrict__ map_idx,
                int FW, uint32_t* __restrict__ keep)
{
    extern __shared__ uint32_t k1_rows[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int64_t s = (int64_t)blockIdx.x * wpb + warp;
    if (s >= S) return;                                   // warp-uniform, no block barriers below
    uint32_t* row = k1_rows + (size_t)warp * FW;
    for (int i = lane; i < FW; i += 32) row[i] = 0u;
    __syncwarp();
    const int64_t b = off[s], e = off[s + 1];
    for (int64_t i = b + lane; i < e; i += 32) {
        const int32_t id = __ldg(ids + i);
        if ((uint32_t)id < (uint32_t)V) {
            const int k1 = __ldg(map_off + id + 1);
            for (int k = __ldg(map_off + id); k < k1; ++k) {
                const int g = __ldg(map_idx + k);
                atomicOr(&row[g >> 5], 1u << (g & 31));
            }
        }
    }
    __syncwarp();
    uint32_t* dst = keep + (size_t)s * FW;
    for (int i = lane; i < FW; i += 32) dst[i] = row[i];
}

// ------------------------------------------------------------------------------------------
// K2 + K3a  plan: per sample, segment kept-flags and the exclusive scan of kept lengths
//   (minimizer_2.py:75-80 union-of-ranges, :94-96 running output index)
//   One CTA per sample.  Segment slots are laid out per genome tile, each tile's slots
//   padded to a multiple of 32 so that one ballot == one stored word and k_emit reads
//   whole words.  A warp reduces one tile at a time; warp 0 then scans the tile sums.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_plan(int64_t S, int FW, const uint32_t* __restrict__ keep, int ntiles,
       const int32_t* __restrict__ tile_slot, const int32_t* __restrict__ slot_len,
       const int32_t* __restrict__ cov_off, const int32_t* __restrict__ cov_idx,
       int SW, uint32_t* __restrict__ segkept, int32_t* __restrict__ tile_off,
       int64_t* __restrict__ lengths, int64_t* __restrict__ rec_size,
       int64_t first_idx, int prefix_len)
{
    extern __shared__ uint32_t plan_sm[];
    uint32_t* row = plan_sm;                       // FW words
    int32_t* tl = (int32_t*)(plan_sm + FW);        // ntiles
    const int64_t s = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

    const uint32_t* krow = keep + (size_t)s * FW;
    for (int i = threadIdx.x; i < FW; i += blockDim.x) row[i] = krow[i];
    __syncthreads();

    uint32_t* sk = segkept + (size_t)s * SW;
    for (int t = warp; t < ntiles; t += nwarps) {
        const int sb = __ldg(tile_slot + t), se = __ldg(tile_slot + t + 1);   // multiples of 32
        int sum = 0;
        for (int slot = sb + lane; slot < se; slot += 32) {
            const int len = __ldg(slot_len + slot);
            bool kept = len > 0;                      // padding slots have len 0
            if (kept) {
                const int k1 = __ldg(cov_off + slot + 1);
                for (int k = __ldg(cov_off + slot); k < k1; ++k) {
                    const int g = __ldg(cov_idx + k);
                    if (!((row[g >> 5] >> (g & 31)) & 1u)) { kept = false; break; }
                }
            }
            const uint32_t w = __ballot_sync(FULL_MASK, kept);
            if (lane == 0) sk[slot >> 5] = w;
            sum += kept ? len : 0;
        }
        sum = warp_sum(sum);
        if (lane == 0) tl[t] = sum;
    }
    __syncthreads();
    if (warp == 0) {
        int carry = 0;
        int32_t* to = tile_off + (size_t)s * ntiles;
        for (int base = 0; base < ntiles; base += 32) {
            const int t = base + lane;
            const int v = t < ntiles ? tl[t] : 0;
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
    int tile_bytes, ntiles, SW, batch, nbatch, store_policy;
    HeaderPrefix prefix;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

template <int POLICY>
__device__ __forceinline__ void st128(uint8_t* p, const uint4& v) {
    if (POLICY == 1) __stcs(reinterpret_cast<uint4*>(p), v);
    else *reinterpret_cast<uint4*>(p) = v;
}

// Warp-cooperative copy of n bytes from shared memory (byte offset a in `tile`) to global d.
template <int POLICY>
__device__ __forceinline__ void copy_run(const uint8_t* __restrict__ tile, int a, uint8_t* __restrict__ d,
                                         int n, int lane)
{
    int h = (int)((16u - (uint32_t)((uintptr_t)d & 15u)) & 15u);
    if (h > n) h = n;
    if (lane < h) d[lane] = tile[a + lane];
    a += h; d += h; n -= h;
    const int nb = n >> 4;
    const int mis = a & 15;
    const uint8_t* q = tile + (a - mis);
    const int k = mis >> 2;
    const int sh = (mis & 3) * 8;
    if (mis == 0) {
        for (int v = lane; v < nb; v += 32)
            st128<POLICY>(d + 16 * v, *reinterpret_cast<const uint4*>(q + 16 * v));
    } else {
        for (int v = lane; v < nb; v += 32) {
            const uint4 lo = *reinterpret_cast<const uint4*>(q + 16 * v);
            const uint4 hi = *reinterpret_cast<const uint4*>(q + 16 * v + 16);
            uint4 o;
            switch (k) {
            case 0:
                o.x = __funnelshift_r(lo.x, lo.y, sh); o.y = __funnelshift_r(lo.y, lo.z, sh);
                o.z = __funnelshift_r(lo.z, lo.w, sh); o.w = __funnelshift_r(lo.w, hi.x, sh); break;
            case 1:
                o.x = __funnelshift_r(lo.y, lo.z, sh); o.y = __funnelshift_r(lo.z, lo.w, sh);
                o.z = __funnelshift_r(lo.w, hi.x, sh); o.w = __funnelshift_r(hi.x, hi.y, sh); break;
            case 2:
                o.x = __funnelshift_r(lo.z, lo.w, sh); o.y = __funnelshift_r(lo.w, hi.x, sh);
                o.z = __funnelshift_r(hi.x, hi.y, sh); o.w = __funnelshift_r(hi.y, hi.z, sh); break;
            default:
                o.x = __funnelshift_r(lo.w, hi.x, sh); o.y = __funnelshift_r(hi.x, hi.y, sh);
                o.z = __funnelshift_r(hi.y, hi.z, sh); o.w = __funnelshift_r(hi.z, hi.w, sh); break;
            }
            st128<POLICY>(d + 16 * v, o);
        }
    }
    const int t = n & 15;
    if (lane < t) d[16 * nb + lane] = tile[a + 16 * nb + lane];
}

template <int POLICY>
__global__ void __launch_bounds__(1024)
k_emit(const EmitParams p)
{
    extern __shared__ __align__(128) uint8_t tile_sm[];   // tile_bytes + 32 (over-read pad)
    __shared__ __align__(8) unsigned long long bar;

    const int tile = blockIdx.x / p.nbatch;
    const int b = blockIdx.x - tile * p.nbatch;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const bool have_tile = p.ntiles > 0;

    if (have_tile) {
        const uint32_t bar_a = smem_u32(&bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)p.tile_bytes;
            const uint8_t* src = p.seq + (size_t)tile * p.tile_bytes;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(tile_sm)), "l"(src), "r"(bytes), "r"(bar_a) : "memory");
        }
        // every thread waits for phase 0 of the barrier (the TMA's complete_tx)
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar_a), "r"(0u) : "memory");
        }
    }

    const int64_t sb = p.s0 + (int64_t)b * p.batch;
    const int64_t se = sb + p.batch < p.s1 ? sb + p.batch : p.s1;
    const int sl0 = have_tile ? __ldg(p.tile_slot + tile) : 0;
    const int sl1 = have_tile ? __ldg(p.tile_slot + tile + 1) : 0;
    const int tile_base = tile * p.tile_bytes;
    const int last_tile = p.ntiles > 0 ? p.ntiles - 1 : 0;
    const int64_t img0 = __ldg(p.rec_off + p.s0);

    for (int64_t s = sb + warp; s < se; s += nwarps) {
        uint8_t* rec = p.out + (__ldg(p.rec_off + s) - img0);
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
            uint8_t* d = seqout + __ldg(p.tile_off + (size_t)s * p.ntiles + tile);
            const uint32_t* sk = p.segkept + (size_t)s * p.SW;
            for (int slot0 = sl0; slot0 < sl1; slot0 += 32) {
                const uint32_t w = __ldg(sk + (slot0 >> 5));          // warp-uniform
                const int slot = slot0 + lane;
                const int len = __ldg(p.slot_len + slot);
                const int src = __ldg(p.slot_src + slot) - tile_base;
                const int x = ((w >> lane) & 1u) ? len : 0;
                const int incl = warp_incl_scan(x, lane);
                const int excl = incl - x;
                uint32_t m = w;
                while (m) {
                    const int a = __ffs(m) - 1;
                    const uint32_t t = ~(m >> a);
                    const int cnt = t ? (__ffs(t) - 1) : 32;
                    const int bl = a + cnt - 1;
                    const int rsrc = __shfl_sync(FULL_MASK, src, a);
                    const int rdst = __shfl_sync(FULL_MASK, excl, a);
                    const int rend = __shfl_sync(FULL_MASK, incl, bl);
                    copy_run<POLICY>(tile_sm, rsrc, d + rdst, rend - rdst, lane);
                    m = cnt >= 32 ? 0u : (m & ~(((1u << cnt) - 1u) << a));
                }
                d += __shfl_sync(FULL_MASK, incl, 31);
            }
        }
        if (tile == last_tile && lane == 0) seqout[__ldg(p.lengths + s)] = (uint8_t)'\n';
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

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
k_range_hashes(const uint8_t* __restrict__ buf, int64_t buf_bytes, const int64_t* __restrict__ off,
               unsigned long long* __restrict__ out)
{
    const int64_t r = blockIdx.x;
    const int64_t o0 = off[r], n = off[r + 1] - o0;
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
        atomicAdd(out + r, t);
    }
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
    void* frees[] = {c->d_seq, c->d_tile_slot, c->d_slot_src, c->d_slot_len, c->d_cov_off, c->d_cov_idx,
                     c->d_map_off, c->d_map_idx, c->own_ids, c->own_ids_off, c->own_keep, c->d_segkept,
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
        if (value < 1 || value > 32) return fail(c, GM2_ERR_INVALID, "emit warps must be in 1..32");
        c->emit_warps = (int)value; return GM2_OK;
    case GM2_CFG_EMIT_BATCH:
        if (value < 0 || value > (1 << 20)) return fail(c, GM2_ERR_INVALID, "emit batch out of range");
        c->emit_batch = (int)value; return GM2_OK;
    case GM2_CFG_PACKING:
        if (value != 0 && value != 1) return fail(c, GM2_ERR_INVALID, "packing: only 0 (auto) and 1 (byte) are implemented");
        c->packing_req = (int)value; return GM2_OK;
    case GM2_CFG_STORE_POLICY:
        if (value != 0 && value != 1) return fail(c, GM2_ERR_INVALID, "store policy must be 0 or 1");
        c->store_policy = (int)value; return GM2_OK;
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
    std::vector<int32_t> slot_src, slot_len, cov_off, cov_idx;
    slot_src.reserve((size_t)nseg + 32 * (size_t)ntiles);
    slot_len.reserve(slot_src.capacity());
    cov_off.reserve(slot_src.capacity() + 1);
    cov_idx.reserve(seg_cov.size());
    cov_off.push_back(0);
    int j = 0;
    for (int t = 0; t < ntiles; ++t) {
        tile_slot[t] = (int32_t)slot_src.size();
        const int64_t tend = std::min<int64_t>((int64_t)(t + 1) * T, G);
        while (j < nseg && bp[j] < tend) {
            slot_src.push_back((int32_t)bp[j]);
            slot_len.push_back((int32_t)(bp[j + 1] - bp[j]));
            for (int64_t k = seg_cov_off[j]; k < seg_cov_off[j + 1]; ++k) cov_idx.push_back(seg_cov[(size_t)k]);
            cov_off.push_back((int32_t)cov_idx.size());
            ++j;
        }
        while (slot_src.size() % 32) {
            slot_src.push_back((int32_t)tend); slot_len.push_back(0);
            cov_off.push_back((int32_t)cov_idx.size());
        }
    }
    tile_slot[ntiles] = (int32_t)slot_src.size();

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
    if ((rc = dev_upload(c, &c->d_cov_off, cov_off))) return rc;
    if ((rc = dev_upload(c, &c->d_cov_idx, cov_idx))) return rc;

    c->G = G; c->F = F; c->FW = (F + 31) / 32;
    c->ntiles = ntiles; c->nseg = nseg; c->nslots = (int)slot_src.size(); c->SW = c->nslots / 32;
    c->packing = 1;
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
    std::vector<int32_t> o(off, off + V + 1), x(idx, idx + n);
    int rc;
    if ((rc = dev_upload(c, &c->d_map_off, o))) return rc;
    if ((rc = dev_upload(c, &c->d_map_idx, x))) return rc;
    c->V = V;
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
    if (!c->d_map_off) return fail(c, GM2_ERR_STATE, "gm2_load_ids_host: call gm2_set_name_map first");
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
    if (!c->d_map_off) return fail(c, GM2_ERR_STATE, "gm2_load_ids_dev: call gm2_set_name_map first");
    if (!off || (n_ids > 0 && !ids) || n_ids < 0) return fail(c, GM2_ERR_INVALID, "gm2_load_ids_dev: bad arguments");
    c->ids = ids; c->ids_off = off; c->S = S; c->mode = 1;
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
    if (S + 1 > c->rec_cap || !c->d_len) {
        int64_t dummy;
        dummy = 0; if ((rc = dev_reserve(c, &c->d_len, &dummy, S + 1))) return rc;
        dummy = 0; if ((rc = dev_reserve(c, &c->d_rec_size, &dummy, S + 1))) return rc;
        dummy = 0; if ((rc = dev_reserve(c, &c->d_rec_off, &dummy, S + 1))) return rc;
        c->rec_cap = S + 1;
    }
    const uint32_t* keep = c->keep_in;
    if (c->mode == 1) {
        if ((rc = dev_reserve(c, &c->own_keep, &c->own_keep_cap, S * c->FW))) return rc;
        keep = c->own_keep;
        if (S > 0 && c->FW > 0) {
            const int wpb = 8;
            const size_t sm = (size_t)wpb * c->FW * 4;
            if (sm > 48 * 1024) CU(c, cudaFuncSetAttribute(k_keep_from_ids, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            const int64_t blocks = (S + wpb - 1) / wpb;
            k_keep_from_ids<<<(unsigned)blocks, wpb * 32, sm, c->stream>>>(c->ids, c->ids_off, S, c->V, c->d_map_off,
                                                                         c->d_map_idx, c->FW, c->own_keep);
            LAUNCH_CHECK(c, "k_keep_from_ids");
        }
    }
    if (S > 0) {
        const size_t sm = ((size_t)c->FW + (size_t)c->ntiles) * 4;
        if (sm > 48 * 1024) CU(c, cudaFuncSetAttribute(k_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_plan<<<(unsigned)S, 256, sm, c->stream>>>(S, c->FW, keep, c->ntiles, c->d_tile_slot, c->d_slot_len,
                                                    c->d_cov_off, c->d_cov_idx, c->SW, c->d_segkept, c->d_tile_off,
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
    const uint32_t* keep = c->mode == 1 ? c->own_keep : c->keep_in;
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
    p.seq = c->d_seq; p.tile_slot = c->d_tile_slot; p.slot_src = c->d_slot_src; p.slot_len = c->d_slot_len;
    p.segkept = c->d_segkept; p.tile_off = c->d_tile_off; p.lengths = c->d_len; p.rec_off = c->d_rec_off;
    p.out = dev_out; p.s0 = s0; p.s1 = s1; p.first_idx = c->first_idx;
    p.tile_bytes = c->tile_bytes; p.ntiles = c->ntiles; p.SW = c->SW; p.batch = (int)batch; p.nbatch = (int)nbatch;
    p.store_policy = c->store_policy; p.prefix = c->prefix;
    const size_t sm = (size_t)c->tile_bytes + 32;
    if (c->store_policy == 1) {
        CU(c, cudaFuncSetAttribute(k_emit<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_emit<1><<<(unsigned)blocks, warps * 32, sm, c->stream>>>(p);
    } else {
        CU(c, cudaFuncSetAttribute(k_emit<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_emit<0><<<(unsigned)blocks, warps * 32, sm, c->stream>>>(p);
    }
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
