// k5_emit_packed.cuh — part of libgm2.so (included by gm2.cu; one translation unit).
// K5: the same stream compaction as k_emit, delivered in the two-bit WIRE FORMAT for the host path.
#pragma once

#include "k4_emit.cuh"

// ------------------------------------------------------------------------------------------
// K5  emit, two bits per base                                    (minimizer_2.py:94-97)
//   Used only by gm2_emit_host / gm2_minimize_host when the reference is ACGT-only: the image's
//   destination is CPU memory, PCIe is the bottleneck (4.1 ms to produce 26 GB on the device, 460 ms
//   to copy it), and the kept bases are all the information there is.  The kernel writes, per
//   (sample, tile), the tile's kept bases as a 2-bit stream ("piece"); the host decodes pieces into
//   the FASTA image and adds headers and newlines (host_expand.cpp).
//
//   Wire layout of a chunk of samples [s0, s1): 32-bit words.  The piece of (sample s, tile t) starts
//   at word   W(s,t) = ((rec_off[s] - rec_off[s0]) >> 4) + (s - s0) * (ntiles + 2) + (tile_off[s][t] >> 4) + t
//   and holds base j of the piece in bits [2j, 2j+2) of the little-endian stream, codes 0..3 = ACGT.
//   W is closed-form in numbers the plan already produced (no extra scan), pieces never overlap
//   (each piece may waste less than one word, which the "+ t" and "+ (ntiles + 2)" terms pay for),
//   and every piece is word-aligned, so no word is shared between warps.
//
//   CTA = (genome tile, batch of samples) as in k_emit; the tile is staged by TMA in its two-bit form
//   (a quarter of the bytes).  Each warp builds the same kept-run table as k_emit and then produces
//   the piece word by word, lane <-> word: a private cursor walks the runs, each run contributes
//   `funnelshift(two packed words) & mask`.  The kernel is far from any roofline that matters: it has
//   to beat the PCIe copy of its own (4x smaller) output, not HBM.
// ------------------------------------------------------------------------------------------
struct PackedParams {
    const uint8_t* seq2;
    const int32_t* tile_slot;
    const int32_t* slot_src;
    const int32_t* slot_len;
    const uint32_t* segkept;
    const int32_t* tile_off;
    const int64_t* rec_off;
    uint32_t* out;
    int64_t s0, s1;
    int tile_bytes, ntiles, SW, batch, nbatch;
    int rt_cap, slot_cap, order;
};

// One batch of kept runs of a piece (table A as in emit_runs: entry r = {Q_r, S_r} in bases, entry nr =
// {end, -}).  Words [Q_0 >> 4, ceil(Q_nr / 16)) of the piece are written; bits outside [Q_0, Q_nr) are
// zero, so when the run table had to be flushed in the middle of a piece (`cont`) the next batch ORs
// its first word onto what the previous batch stored.
__device__ __forceinline__ void emit_runs_packed(uint32_t tile_a, uint32_t rt_a, int nr, uint32_t* __restrict__ pw,
                                                 int lane, bool cont)
{
    __syncwarp();
    const int q_first = rt_load(rt_a).x, q_last = rt_load(rt_a + 8 * nr).x;
    const int w_lo = q_first >> 4, w_hi = (q_last + 15) >> 4;
    uint32_t ra = rt_a;
    int2 e = rt_load(ra);
    int qn = rt_load(ra + 8).x;
#pragma unroll 1
    for (int k = w_lo + lane; k < w_hi; k += 32) {
        const int pos0 = k << 4;
        const int lo0 = max(pos0, q_first), hi0 = min(pos0 + 16, q_last);       // this word's bases inside the batch
        while (qn <= lo0) { ra += 8; e = rt_load(ra); qn = rt_load(ra + 8).x; }  // run holding base lo0
        uint32_t acc = 0u;
        for (;;) {
            const int lo = max(e.x, lo0) - pos0, hi = min(qn, hi0) - pos0;       // 0 <= lo <= hi <= 16
            const int srcb = e.y + (pos0 - e.x);                                 // tile base that lands on bit 0 (>= -15)
            const uint32_t wa = tile_a + (uint32_t)((srcb >> 4) * 4);
            const uint32_t bits = __funnelshift_r(lds32(wa), lds32(wa + 4), 2 * (srcb & 15));
            const uint32_t below_hi = hi >= 16 ? 0xffffffffu : ((1u << (2 * hi)) - 1u);
            const uint32_t below_lo = lo >= 16 ? 0xffffffffu : ((1u << (2 * lo)) - 1u);
            acc |= bits & below_hi & ~below_lo;
            if (qn >= hi0) break;
            ra += 8; e = rt_load(ra); qn = rt_load(ra + 8).x;
        }
        if (cont && k == w_lo && (q_first & 15)) acc |= __ldcv(pw + k);
        pw[k] = acc;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(256)
k_emit_packed(const PackedParams p)
{
    // dynamic shared memory: 32 B front pad | staged two-bit tile | 64 B over-read pad | slot tables | per-warp run table
    extern __shared__ __align__(128) uint8_t dsm[];
    __shared__ __align__(8) unsigned long long bar;

    const int ntl = p.ntiles;                                    // > 0: the host never launches this for an empty genome
    const int tile = p.order ? (int)(blockIdx.x % ntl) : (int)(blockIdx.x / p.nbatch);
    const int b = p.order ? (int)(blockIdx.x / ntl) : (int)(blockIdx.x - tile * p.nbatch);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int sl0 = __ldg(p.tile_slot + tile);
    const int nslots = __ldg(p.tile_slot + tile + 1) - sl0;
    const int nwords = nslots >> 5;
    const int tile_base = tile * p.tile_bytes;
    const int tile_smem_bytes = p.tile_bytes >> 2;

    const uint32_t dsm_a = smem_u32(dsm);
    const uint32_t tile_a = dsm_a + EMIT_FRONT_PAD;
    const uint32_t len_a = tile_a + (uint32_t)tile_smem_bytes + EMIT_BACK_PAD;
    const uint32_t src_a = len_a + 4u * (uint32_t)p.slot_cap;
    const uint32_t rt_a = src_a + 4u * (uint32_t)p.slot_cap + (uint32_t)warp * (uint32_t)(p.rt_cap + 2) * 8u;
    int32_t* sm_len = reinterpret_cast<int32_t*>(dsm + EMIT_FRONT_PAD + tile_smem_bytes + EMIT_BACK_PAD);
    int32_t* sm_src = sm_len + p.slot_cap;
    const bool slots_staged = nslots <= p.slot_cap;

    {
        const uint32_t bar_a = smem_u32(&bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const uint32_t bytes = (uint32_t)tile_smem_bytes;
            const uint8_t* src = p.seq2 + (size_t)tile * tile_smem_bytes;
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
        __syncthreads();
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
    const int64_t img0 = __ldg(p.rec_off + p.s0);
    const uint32_t lt_mask = (1u << lane) - 1u;

    for (int64_t s = sb + warp; s < se; s += nwarps) {
        const int64_t roff = __ldg(p.rec_off + s);
        const int toff = __ldg(p.tile_off + (size_t)s * p.ntiles + tile);
        const uint32_t words = lane < nwords ? __ldg(p.segkept + (size_t)s * p.SW + (sl0 >> 5) + lane) : 0u;
        uint32_t* pw = p.out + (((roff - img0) >> 4) + (s - p.s0) * (int64_t)(p.ntiles + 2) + (toff >> 4) + tile);
        int q = 0, nr = 0;
        uint32_t carry = 0u;
        bool cont = false;
        for (int c = 0; ; ++c) {
            const bool done = c >= nwords;
            if (done || nr + 17 > p.rt_cap) {
                if (nr > 0) {
                    if (lane == 0) rt_store_x(rt_a + 8u * nr, q);
                    emit_runs_packed(tile_a, rt_a, nr, pw, lane, cont);
                    cont = true; nr = 0; carry = 0u;
                }
                if (done) break;
            }
            const uint32_t w = c < 32 ? __shfl_sync(FULL_MASK, words, c)
                                      : __ldg(p.segkept + (size_t)s * p.SW + (sl0 >> 5) + c);
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
}
