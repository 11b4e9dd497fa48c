// k4_emit.cuh — part of libgm2.so (included by gm2.cu; one translation unit).
// K4: TMA-staged stream-compaction gather with fused FASTA framing (k_emit).
#pragma once

#include "device_util.cuh"

#define GM2_MAX_PREFIX 95

struct HeaderPrefix {            // passed by value to kernels; text[0] is '>'
    int  len;
    char text[GM2_MAX_PREFIX + 1];
};


// ------------------------------------------------------------------------------------------
// K4  emit: stream-compaction gather + FASTA framing              (minimizer_2.py:94-97, :476-477)
//   CTA = (genome tile, batch of samples).  The tile's bases are staged ONCE in shared memory by a
//   1-D TMA bulk copy (cp.async.bulk + mbarrier) and reused by every sample of the batch; the tile's
//   static slot tables are staged beside it.  Each warp owns one sample at a time (the next sample's
//   metadata is prefetched): it turns the tile's kept-bit words into a table of kept runs with
//   shuffle scans, derives every per-run constant lane-parallel into a second table, and then writes
//   the runs as one ascending stream of 32-byte sectors — see emit_runs.  The warp that owns tile 0
//   writes the '>' header, the one that owns the last tile the final '\n'.
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
    const int32_t* hdr_len;   // bytes of each record's header line ('>' prefix digits '\n'), from k_plan
    uint8_t* out;
    int64_t s0, s1;
    int64_t first_idx;
    int tile_bytes, ntiles, SW, batch, nbatch;
    int tile_smem_bytes;   // bytes of the staged tile: tile_bytes (1 byte/base) or tile_bytes/4 (2 bits/base)
    int rt_cap;            // run-table entries per warp (shared memory)
    int slot_cap;          // slot-table entries staged in shared memory (0: read from global)
    int order;             // CTA -> work mapping: 0 tile-major, 1 sample-major
    int flat_run_bytes;    // (sample, tile) visits whose mean run length is below this take the flat form (0: never)
    int flat_cap;          // FLAT == 2: runs the packed per-warp table of visit_flat holds
    int flat_bm_words;     // FLAT == 2: words of its vector bitmap
    int debug;             // timing experiments only (wrong output; needs -DGM2_EMIT_DEBUG): 1 no boundary
                           // sectors, 2 no interior stores, 4 interior stores without shared loads
    HeaderPrefix prefix;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// shared-memory accessors on 32-bit shared-window addresses.  Tile / slot-table reads are `volatile`
// asm WITHOUT a memory clobber: that keeps them behind the __syncthreads() / mbarrier wait that
// publishes the staged data (a plain asm has no memory dependency the compiler would respect) while
// leaving ptxas free to schedule them; run-table accesses add the memory clobber because the table
// is rewritten per (sample, tile) around __syncwarp().
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
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

// One destination-aligned vector per lane, warp-uniform source phase (K, sh): qa as in copy_vectors.
template <int POLICY, int K>
__device__ __forceinline__ void copy_one(uint32_t qa, uint8_t* __restrict__ d, int sh)
{
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


// Flat form of emit_runs for batches of SHORT runs (mean run length below EmitParams::flat_run_bytes).
// The run-by-run stream above pays ~50 warp instructions per run besides the copy, which bounds the
// kernel once runs are a few hundred bytes (low retention: every intergenic gap between two deleted
// genes is a run of its own).  Here the unit of work is a destination-aligned 16-byte vector and the
// lanes do not share a run:
//   phase B  lane <-> run boundary: the vector a boundary falls into is assembled byte by byte in
//            registers (private cursor into table A) and leaves as one 128-bit store;
//   phase I  lane <-> vector, stride 32: each lane walks table A with its own cursor; a vector that
//            lies inside one run is two aligned 128-bit shared loads, a per-lane word select + byte
//            funnel shift, one 128-bit store.  Lanes 2j, 2j+1 still fill one sector per instruction;
//   phase E  the partial first / last vector of the batch (shared with the neighbouring tile's warp),
//            lane <-> byte as in the stream form.
// Both halves of a sector are written within the same (sample, tile) visit, i.e. well inside the time
// a line stays in L2, so no partial sector reaches DRAM.  Table B is not built.
template <int POLICY>
__device__ __forceinline__ void emit_runs_flat(uint32_t tile_a, uint32_t rt_a, int nr,
                                               uint8_t* __restrict__ base32, int lane)
{
    const int q_first = rt_load(rt_a).x, q_last = rt_load(rt_a + 8 * nr).x;
    // ---- phase B: vectors holding a run boundary (whole vectors only; the batch's edges are phase E)
    for (int r0 = 1; r0 < nr; r0 += 32) {
        const int r = r0 + lane;
        if (r < nr) {
            const int qr = rt_load(rt_a + 8 * r).x;
            const int p0 = qr & ~15;
            bool own = (qr & 15) && p0 >= q_first && p0 + 16 <= q_last;
            uint32_t ra = rt_a + 8 * (r - 1);
            int2 e = rt_load(ra);
            if ((e.x & ~15) == p0 && (e.x & 15)) own = false;                // an earlier boundary owns this vector
            if (own) {
                int qn = qr;
                uint32_t ab = tile_a + (uint32_t)(e.y - e.x + p0);              // shared address of byte p0 if it were in run rr
                uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    while (p0 + j >= qn) {
                        ra += 8; e = rt_load(ra); qn = rt_load(ra + 8).x;
                        ab = tile_a + (uint32_t)(e.y - e.x + p0);
                    }
                    w[j >> 2] |= lds8(ab + j) << (8 * (j & 3));
                }
                st128<POLICY>(base32 + p0, make_uint4(w[0], w[1], w[2], w[3]));
            }
        }
    }
    // ---- phase I: vectors inside one run
    {
        const int v_hi = q_last >> 4;
        uint32_t ra = rt_a;
        int2 e = rt_load(ra);
        int qn = rt_load(ra + 8).x;
#pragma unroll 1
        for (int v = ((q_first + 15) >> 4) + lane; v < v_hi; v += 32) {
            const int pos = v << 4;
            while (pos >= qn) { ra += 8; e = rt_load(ra); qn = rt_load(ra + 8).x; }
            if (pos + 16 <= qn) {
                const uint32_t a = tile_a + (uint32_t)(e.y + (pos - e.x));
                const uint32_t qa = a & ~15u;
                const int sh = (int)(a & 3u) * 8;
                const uint4 lo = lds128(qa), hi = lds128(qa + 16);
                uint32_t w0 = lo.x, w1 = lo.y, w2 = lo.z, w3 = lo.w, w4 = hi.x, w5 = hi.y, w6 = hi.z;
                if (a & 8u) { w0 = w2; w1 = w3; w2 = w4; w3 = w5; w4 = w6; w5 = hi.w; }
                if (a & 4u) { w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; }
                uint4 o;
                o.x = __funnelshift_r(w0, w1, sh);
                o.y = __funnelshift_r(w1, w2, sh);
                o.z = __funnelshift_r(w2, w3, sh);
                o.w = __funnelshift_r(w3, w4, sh);
                st128<POLICY>(base32 + pos, o);
            }
        }
    }
    // ---- phase E: partial vectors at the two ends of the batch
    {
        const int vf = q_first >> 4, vl = q_last >> 4;
        int pos = -1;
        if (lane < 16) { if (q_first & 15) pos = (vf << 4) + lane; }
        else if ((q_last & 15) && !((q_first & 15) && vl == vf)) pos = (vl << 4) + (lane - 16);
        if (pos >= q_first && pos < q_last) {
            int2 e;
            if (lane < 16) {
                uint32_t ra = rt_a; e = rt_load(ra); int qn = rt_load(ra + 8).x;
                while (pos >= qn) { ra += 8; e = rt_load(ra); qn = rt_load(ra + 8).x; }
            } else {
                uint32_t ra = rt_a + 8 * (nr - 1); e = rt_load(ra);
                while (pos < e.x) { ra -= 8; e = rt_load(ra); }
            }
            st8<POLICY>(base32 + pos, lds8(tile_a + (uint32_t)(e.y + (pos - e.x))));
        }
    }
    __syncwarp();
}


// ------------------------------------------------------------------------------------------
// visit_flat (FLAT == 2): the whole (sample, tile) visit for SHORT runs, chosen up front from the
// visit's byte count and its number of runs (both known before any table is built).
//
// The work unit is again a destination-aligned 16-byte vector, lane <-> vector, 32 consecutive
// vectors (one 512-byte row of the output) per warp iteration — the store pattern the memory system
// wants (profiles/r01_emit_experiments.md).  What the first flat form paid for was FINDING each
// vector's run (a private linear cursor per lane: ~500 of 2,150 warp instructions per visit at 10 %
// gene retention) and the byte-wise assembly of vectors that hold a run boundary (~400).  Here:
//
//   R space    byte offsets from the 16-byte aligned address at or below the visit's first output
//              byte; everything fits 16 bits (tile <= 60 KB), so a run is ONE packed word
//              A[r] = R_r | S_r << 16  (S = source offset in the staged tile), A[nr] = R_end.
//   events     an "event" is a vector in which at least one run starts.  One pass, lane <-> run,
//              writes EV[e] = tile address + S - R of the LAST run starting in that vector (ballot + popc
//              ranks), i.e. "source address of output byte 0 if it belonged to that run", and sets two
//              bits per vector in a per-warp bitmap pair (shared-memory atomic OR): `event` (a run starts
//              in it) and `skip` (a run starts inside it, not on its first byte: phase B's vector).  The
//              skip word also masks everything from the visit's last partial vector on, so phase I has
//              no range checks at all.
//   phase I    row i (512 output bytes), lane L: {event, skip} = BM[i] (one broadcast 64-bit load); the
//              vector's source is EV[base + popc(event & le) - 1] + pos — no search, no compare; unless
//              its skip bit is set the lane copies it: two aligned LDS.128, word select, funnel shift,
//              one STG.128.  ~40 warp instructions per row whatever the number of runs in it.
//   phase B    same pass as the events, lane <-> run: the first run starting in a vector owns it; up to
//              three sources (tail of the previous run, one or two starts) are fetched as unaligned
//              16-byte windows and merged under byte masks; more than three
//              (runs of a few bytes in a row) fall back to a byte loop.
//   phase E    the partial first / last vector of the visit (shared with the neighbouring tile's warp),
//              lane <-> byte.
// Each vector is written exactly once and whole, boundary vectors a few hundred cycles before the
// rows around them, i.e. well inside the time the line stays in L2.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tb_load(uint32_t a) {                 // per-warp tables (rewritten per visit)
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void tb_store(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint2 tb_load2(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 tb_load4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void tb_store2(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tb_or(uint32_t a, uint32_t v) {
    asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}

// 16 bytes from an arbitrary shared-memory address: two aligned 128-bit loads, word select, byte funnel shift
__device__ __forceinline__ uint4 lds_unaligned16(uint32_t a) {
    const uint32_t qa = a & ~15u;
    const int sh = (int)(a & 3u) * 8;
    const uint4 lo = lds128(qa), hi = lds128(qa + 16);
    uint32_t w0 = lo.x, w1 = lo.y, w2 = lo.z, w3 = lo.w, w4 = hi.x, w5 = hi.y, w6 = hi.z;
    if (a & 8u) { w0 = w2; w1 = w3; w2 = w4; w3 = w5; w4 = w6; w5 = hi.w; }
    if (a & 4u) { w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; }
    uint4 o;
    o.x = __funnelshift_r(w0, w1, sh);
    o.y = __funnelshift_r(w1, w2, sh);
    o.z = __funnelshift_r(w2, w3, sh);
    o.w = __funnelshift_r(w3, w4, sh);
    return o;
}
// lds_unaligned16 when only bytes [b0, b1) of the window will be used (0 <= b0 < b1 <= 16): an aligned half that
// holds none of them is not loaded.  Lanes of the boundary-vector pass read unrelated addresses, so every lane
// that drops out of a load saves a wavefront on the L1 data pipe (the limiter at low retention).
__device__ __forceinline__ uint4 lds_unaligned16_part(uint32_t a, int b0, int b1) {
    const uint32_t qa = a & ~15u;
    const int m = (int)(a & 15u), sh = (int)(a & 3u) * 8;
    uint4 lo = make_uint4(0u, 0u, 0u, 0u), hi = lo;
    if (m + b0 < 16) lo = lds128(qa);
    if (m + b1 > 16) hi = lds128(qa + 16);
    uint32_t w0 = lo.x, w1 = lo.y, w2 = lo.z, w3 = lo.w, w4 = hi.x, w5 = hi.y, w6 = hi.z;
    if (a & 8u) { w0 = w2; w1 = w3; w2 = w4; w3 = w5; w4 = w6; w5 = hi.w; }
    if (a & 4u) { w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; }
    uint4 o;
    o.x = __funnelshift_r(w0, w1, sh);
    o.y = __funnelshift_r(w1, w2, sh);
    o.z = __funnelshift_r(w2, w3, sh);
    o.w = __funnelshift_r(w3, w4, sh);
    return o;
}
// bytes [0, k) of x, bytes [k, 16) of y  (0 <= k <= 16).  Word w takes its low clamp(8k - 32w, 0, 32) bits from x:
// the clamped funnel shift turns that bit count into the mask without a table.
__device__ __forceinline__ uint4 merge16(const uint4& x, const uint4& y, int k) {
    const int b = 8 * k;
    const uint32_t m0 = __funnelshift_lc(0xffffffffu, 0u, b);
    const uint32_t m1 = __funnelshift_lc(0xffffffffu, 0u, max(b - 32, 0));
    const uint32_t m2 = __funnelshift_lc(0xffffffffu, 0u, max(b - 64, 0));
    const uint32_t m3 = __funnelshift_lc(0xffffffffu, 0u, max(b - 96, 0));
    return make_uint4((x.x & m0) | (y.x & ~m0), (x.y & m1) | (y.y & ~m1),
                      (x.z & m2) | (y.z & ~m2), (x.w & m3) | (y.w & ~m3));
}

#define FLAT_MAX_TILE 61440          // R and S must fit 16 bits

// The 16 output bytes whose first one is base `off` of the staged tile (off may be negative by < 16, or run past the
// tile: both land in the pads or in this CTA's other shared memory and are never stored).  PACK 1: bytes, an unaligned
// 16-byte window (two LDS.128 + word select + funnel shift).  PACK 2: two bits per base — the 32 bits holding the 16
// codes are a funnel shift of two CONSECUTIVE words, and the lanes of a run read consecutive words: one wavefront
// per load instead of four, which is what the L1 data pipe is short of at low retention; the price is the
// expansion (spread + PRMT), ~5 more instructions per vector.
template <int PACK>
__device__ __forceinline__ uint4 tile_window16(uint32_t tile_a, int off) {
    if (PACK == 2) {
        const uint32_t wa = tile_a + (uint32_t)((off >> 4) << 2);
        return expand16(__funnelshift_r(lds32(wa), lds32(wa + 4u), 2 * (off & 15)));
    }
    return lds_unaligned16(tile_a + (uint32_t)off);
}
template <int PACK>
__device__ __forceinline__ uint4 tile_window16_part(uint32_t tile_a, int off, int b0, int b1) {
    if (PACK == 2) return tile_window16<2>(tile_a, off);
    return lds_unaligned16_part(tile_a + (uint32_t)off, b0, b1);
}
template <int PACK>
__device__ __forceinline__ uint32_t tile_base_at(uint32_t tile_a, int off) {
    return PACK == 2 ? base_at_2bit(tile_a, off) : lds8(tile_a + (uint32_t)off);
}

template <int POLICY, int PACK>
__device__ __forceinline__ void visit_flat(uint32_t tile_a, uint32_t a_a, uint32_t ev_a, uint32_t bm_a,
                                           uint32_t len_a, uint32_t src_a, uint32_t words, int nwords,
                                           uint8_t* __restrict__ out0 /* first output byte of the visit */,
                                           int bytes, int lane)
{
    uint32_t lt_mask = (1u << lane) - 1u;
    const int o = (int)((uintptr_t)out0 & 15u);
    uint8_t* base16 = out0 - o;
    asm volatile("" : "+l"(base16), "+r"(tile_a), "+r"(lt_mask));       // keep these in registers (no re-derivation per row)
    const int r_end = o + bytes;
    const int nbm = ((r_end - 1) >> 9) + 1;                          // rows (= bitmap entries) the visit touches
    const int vl = r_end >> 4;                                       // first vector that is not whole inside the visit
    // bitmap entry i = {event bits, skip bits} of row i; the skip words mask everything from vector vl on,
    // including the whole entry after the last row (phase I works on pairs of rows)
    for (int i = lane; i <= nbm; i += 32)
        tb_store2(bm_a + 8u * i, 0u, i > (vl >> 5) ? 0xffffffffu : i == (vl >> 5) ? 0xffffffffu << (vl & 31) : 0u);

    // ---- table A: one packed word per kept run
    int nr = 0;
    if (nwords <= 4) {
        // lane <-> four consecutive slots (a tile of the usual shapes has 64-128 slots): the kept lengths and the
        // run starts of the whole tile go through ONE warp scan, packed {bytes : 20 | starts : 12} — a lane holds
        // at most two run starts (starts are never adjacent).  Lanes past the tile's slots see kept bits 0 and
        // read table entries that are never used.
        const uint32_t wsrc = __shfl_sync(FULL_MASK, words, lane >> 3);           // word c sits in lane c
        const uint32_t nib = (wsrc >> (4 * (lane & 7))) & 0xfu;
        const uint4 l4 = lds128(len_a + 16u * (uint32_t)lane);
        const uint4 s4 = lds128(src_a + 16u * (uint32_t)lane);
        uint32_t up = __shfl_up_sync(FULL_MASK, nib, 1);
        if (lane == 0) up = 0u;
        const uint32_t st = nib & ~((nib << 1) | (up >> 3));
        const int p1 = (nib & 1u) ? (int)l4.x : 0;                                // kept bytes before slot 1, 2, 3 of this lane
        const int p2 = p1 + ((nib & 2u) ? (int)l4.y : 0);
        const int p3 = p2 + ((nib & 4u) ? (int)l4.z : 0);
        const int tot = p3 + ((nib & 8u) ? (int)l4.w : 0);
        const int mine = tot | (__popc(st) << 20);
        const int incl = warp_incl_scan(mine, lane);
        const int excl = incl - mine;
        nr = (int)((uint32_t)__shfl_sync(FULL_MASK, incl, 31) >> 20);
        if (st) {
            const int q = o + (excl & 0xfffff);
            const uint32_t ra = a_a + 4u * ((uint32_t)excl >> 20);
            const int j0 = __ffs((int)st) - 1;                                    // first start: slot 0..3
            const int pj = j0 == 0 ? 0 : j0 == 1 ? p1 : j0 == 2 ? p2 : p3;
            const uint32_t sj = j0 == 0 ? s4.x : j0 == 1 ? s4.y : j0 == 2 ? s4.z : s4.w;
            tb_store(ra, (uint32_t)(q + pj) | (sj << 16));
            if (st & (st - 1u)) {                                                 // a second start: slot 2 or 3
                const bool at3 = (st & 8u) != 0u;
                tb_store(ra + 4u, (uint32_t)(q + (at3 ? p3 : p2)) | ((at3 ? s4.w : s4.z) << 16));
            }
        }
        if (lane == 0) tb_store(a_a + 4u * nr, (uint32_t)r_end);
    } else {
        int q = o;
        uint32_t carry = 0u;
        for (int c = 0; c < nwords; ++c) {
            const uint32_t w = __shfl_sync(FULL_MASK, words, c);                  // warp-uniform
            const int len = (int)lds32(len_a + 4u * (32 * c + lane));
            const int src = (int)lds32(src_a + 4u * (32 * c + lane));
            const int x = ((w >> lane) & 1u) ? len : 0;
            const int incl = warp_incl_scan(x, lane);
            const uint32_t starts = w & ~((w << 1) | carry);
            carry = w >> 31;
            if ((starts >> lane) & 1u)
                tb_store(a_a + 4u * (nr + __popc(starts & lt_mask)), (uint32_t)(q + incl - x) | ((uint32_t)src << 16));
            nr += __popc(starts);
            q += __shfl_sync(FULL_MASK, incl, 31);
        }
        if (lane == 0) tb_store(a_a + 4u * nr, (uint32_t)r_end);
    }
    __syncwarp();

    // ---- events + phase B, lane <-> run
    int ne = 0;
    for (int r0 = 0; r0 < nr; r0 += 32) {
        const int r = r0 + lane;
        bool last = false;
        uint32_t x = 0u;
        if (r < nr) {
            x = tb_load(a_a + 4u * r);
            const int R = (int)(x & 0xffffu), v = R >> 4, p0 = v << 4;
            const uint32_t x1 = tb_load(a_a + 4u * (r + 1));                      // run r+1, or the end sentinel
            const bool in1 = r + 1 < nr && (int)((x1 & 0xffffu) >> 4) == v;
            last = !in1;
            const uint32_t xm = r > 0 ? tb_load(a_a + 4u * (r - 1)) : 0u;
            const bool first = r == 0 || (int)((xm & 0xffffu) >> 4) != v;
            if (first && p0 >= o && p0 + 16 <= r_end && (R > p0 || in1)) {        // this lane owns a boundary vector
                const uint32_t x2 = r + 2 <= nr ? tb_load(a_a + 4u * (r + 2)) : 0u;
                const bool in2 = in1 && r + 2 < nr && (int)((x2 & 0xffffu) >> 4) == v;
                const bool has_prev = R > p0;                                     // then r >= 1: R_0 = o and p0 >= o
                const int nsrc = (has_prev ? 1 : 0) + 1 + (in1 ? 1 : 0) + (in2 ? 1 : 0);
                bool slow = nsrc > 3;
                if (!slow && in2) {                                               // three starts: a fourth would need the loop
                    const uint32_t x3 = r + 3 <= nr ? tb_load(a_a + 4u * (r + 3)) : 0u;
                    slow = r + 3 < nr && (int)((x3 & 0xffffu) >> 4) == v;
                }
                uint4 ov;
                if (!slow) {
                    const uint32_t sa = has_prev ? xm : x, sb = has_prev ? x : x1, sc = has_prev ? x1 : x2;
                    const int kb = (int)(sb & 0xffffu) - p0;                      // bytes [0, kb) from source a, [kb, kc) from b
                    const int kc = nsrc == 3 ? (int)(sc & 0xffffu) - p0 : 16;     // ... and [kc, 16) from c
                    const uint4 X = tile_window16_part<PACK>(tile_a, (int)(sa >> 16) + p0 - (int)(sa & 0xffffu), 0, kb);
                    const uint4 Y = tile_window16_part<PACK>(tile_a, (int)(sb >> 16) + p0 - (int)(sb & 0xffffu), kb, kc);
                    ov = merge16(X, Y, kb);
                    if (nsrc == 3) {
                        const uint4 Z = tile_window16_part<PACK>(tile_a, (int)(sc >> 16) + p0 - (int)(sc & 0xffffu), kc, 16);
                        ov = merge16(ov, Z, kc);
                    }
                } else {                                                          // many tiny runs in one vector
                    int rc = has_prev ? r - 1 : r;
                    uint32_t xc = tb_load(a_a + 4u * rc);
                    int rn = (int)(tb_load(a_a + 4u * (rc + 1)) & 0xffffu);
                    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        while (p0 + j >= rn) { ++rc; xc = tb_load(a_a + 4u * rc); rn = (int)(tb_load(a_a + 4u * (rc + 1)) & 0xffffu); }
                        w[j >> 2] |= tile_base_at<PACK>(tile_a, (int)(xc >> 16) + p0 + j - (int)(xc & 0xffffu)) << (8 * (j & 3));
                    }
                    ov = make_uint4(w[0], w[1], w[2], w[3]);
                }
                st128<POLICY>(base16 + p0, ov);
            }
            if (last) tb_or(bm_a + 8u * (uint32_t)(v >> 5), 1u << (v & 31));
            if (R & 15) tb_or(bm_a + 8u * (uint32_t)(v >> 5) + 4u, 1u << (v & 31));
            x = (PACK == 1 ? tile_a : 0u) + (x >> 16) - (uint32_t)R;              // source of output byte 0 in run r's frame:
                                                                                  // shared address (bytes) / base index (two-bit)
        }
        const uint32_t bal = __ballot_sync(FULL_MASK, last);
        if (last) tb_store(ev_a + 4u * (ne + __popc(bal & lt_mask)), x);
        ne += __popc(bal);
    }
    __syncwarp();

    // ---- phase I: interior vectors, two rows (1 KB of output) per step.  Nothing here branches on data: the
    // two rows' bitmap entries arrive in one 128-bit load, fetched one step ahead (the table reads are ordered
    // asm statements: nothing else overlaps their latency); both source windows are loaded unconditionally
    // (a skipped lane reads valid shared memory near its run and drops it) and only the store is predicated,
    // so the two rows' instruction streams interleave.  Entries past the last row are skip-all / never used.
    // (Measured and dropped: issuing the next step's event look-ups before this step's select/shift/store,
    // -2.5 % — at this point the L1 data pipe, shared loads + global stores, is ~80 % busy, not latency.)
    {
        uint32_t le_mask = lt_mask | (1u << lane), lane_bit = 1u << lane;
        asm volatile("" : "+r"(le_mask), "+r"(lane_bit));
        uint32_t evp = ev_a - 4u;                                                 // &EV[events of earlier rows - 1]
        uint32_t pa = (uint32_t)(lane << 4);                                      // this lane's output offset
        uint8_t* d = base16 + (lane << 4);
        uint32_t bmp = bm_a;
        auto rows2 = [&](const uint4& mk) {
            const int n0 = __popc(mk.x);
            const uint32_t ea0 = tb_load(evp + 4u * (uint32_t)__popc(mk.x & le_mask));
            const uint32_t ea1 = tb_load(evp + 4u * (uint32_t)(n0 + __popc(mk.z & le_mask)));
            evp += 4u * (uint32_t)(n0 + __popc(mk.z));
            const uint4 v0 = PACK == 1 ? lds_unaligned16(ea0 + pa) : tile_window16<2>(tile_a, (int)(ea0 + pa));
            const uint4 v1 = PACK == 1 ? lds_unaligned16(ea1 + pa + 512u) : tile_window16<2>(tile_a, (int)(ea1 + pa + 512u));
            if ((mk.y & lane_bit) == 0u) st128<POLICY>(d, v0);
            if ((mk.w & lane_bit) == 0u) st128<POLICY>(d + 512, v1);
            pa += 1024u; d += 1024;
        };
        uint4 mk = tb_load4(bmp);                                                 // row 0 always holds event 0 (bit 0)
#pragma unroll 1
        for (int i = 0; ; i += 4) {
            const uint4 nx = tb_load4(bmp + 16u);
            rows2(mk);
            if (i + 2 >= nbm) break;
            mk = tb_load4(bmp + 32u);
            bmp += 32u;
            rows2(nx);
            if (i + 4 >= nbm) break;
        }
    }

    // ---- phase E: the partial first / last vector of the visit, lane <-> byte
    {
        const int vl = r_end >> 4;
        int pos = -1;
        if (lane < 16) { if (o) pos = lane; }
        else if ((r_end & 15) && !(o && vl == 0)) pos = (vl << 4) + (lane - 16);
        if (pos >= o && pos < r_end) {
            uint32_t xc;
            if (lane < 16) {
                int rc = 0; xc = tb_load(a_a);
                int rn = (int)(tb_load(a_a + 4u) & 0xffffu);
                while (pos >= rn) { ++rc; xc = tb_load(a_a + 4u * rc); rn = (int)(tb_load(a_a + 4u * (rc + 1)) & 0xffffu); }
            } else {
                int rc = nr - 1; xc = tb_load(a_a + 4u * rc);
                while (pos < (int)(xc & 0xffffu)) { --rc; xc = tb_load(a_a + 4u * rc); }
            }
            st8<POLICY>(base16 + pos, tile_base_at<PACK>(tile_a, (int)(xc >> 16) + pos - (int)(xc & 0xffffu)));
        }
    }
    __syncwarp();
}

#define EMIT_FRONT_PAD 32
#define EMIT_BACK_PAD  64

template <int POLICY, int MIN_CTAS, int PACK, int FLAT>
__global__ void __launch_bounds__(256, MIN_CTAS)
k_emit(const EmitParams p)
{
    // dynamic shared memory: 32 B front pad | staged tile | 64 B over-read pad | slot tables | per-warp run tables A + B
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
    int64_t m_roff = 0; int m_toff = 0, m_tend = 0, m_hl = 0; uint32_t m_words = 0u;
    // the flat form needs the tile's slot tables in shared memory and all its kept-bit words in one register per lane
    const bool flat_ok = FLAT == 2 && p.flat_cap > 0 && p.flat_run_bytes > 0 && slots_staged && nwords <= 32;
    auto load_meta = [&](int64_t s) {
        m_roff = __ldg(p.rec_off + s);
        m_hl = __ldg(p.hdr_len + s);
        if (have_tile) {
            m_toff = __ldg(p.tile_off + (size_t)s * p.ntiles + tile);
            if (FLAT == 2)                       // where this tile's output ends: the next tile's offset / the sample's length
                m_tend = tile == p.ntiles - 1 ? (int)__ldg(p.lengths + s) : __ldg(p.tile_off + (size_t)s * p.ntiles + tile + 1);
            m_words = lane < nwords ? __ldg(p.segkept + (size_t)s * p.SW + (sl0 >> 5) + lane) : 0u;
        }
    };
    int64_t s = sb + warp;
    if (s < se) load_meta(s);
    while (s < se) {
        const int64_t roff = m_roff; const int toff = m_toff, tend = m_tend, hl = m_hl; const uint32_t words = m_words;
        const int64_t sn = s + nwarps;
        if (sn < se) load_meta(sn);

        uint8_t* rec = p.out + (roff - img0);
        if (tile == 0) {
            const unsigned long long num = (unsigned long long)(p.first_idx + s + 1);
            const int nd = hl - p.prefix.len - 1;
            for (int i = lane; i < hl; i += 32) {
                char ch;
                if (i < p.prefix.len) ch = p.prefix.text[i];
                else if (i == hl - 1) ch = '\n';
                else ch = (char)('0' + (int)((num / c_pow10[nd - 1 - (i - p.prefix.len)]) % 10ull));
                rec[i] = (uint8_t)ch;
            }
        }
        uint8_t* seqout = rec + hl;
        bool flat = false;
        if (FLAT == 2 && flat_ok) {
            // runs of the visit = run starts in its kept-bit words (lane c holds word c); bytes = tend - toff
            uint32_t pm = __shfl_up_sync(FULL_MASK, words >> 31, 1);
            if (lane == 0) pm = 0u;
            const int nrt = __reduce_add_sync(FULL_MASK, __popc(words & ~((words << 1) | pm)));
            const int bytes = tend - toff;
            flat = nrt > 0 && nrt <= p.flat_cap && bytes < nrt * p.flat_run_bytes && ((bytes + 14) >> 9) + 1 <= p.flat_bm_words;
            if (flat)
                visit_flat<POLICY, PACK>(tile_a, rt_a, rt_a + 4u * (uint32_t)(p.flat_cap + 2), rt_a + 8u * (uint32_t)(p.flat_cap + 2),
                                   len_a, src_a, words, nwords, seqout + toff, bytes, lane);
        }
        if (have_tile && !flat) {
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
                        __syncwarp();
                        if (FLAT == 1 && PACK == 1 && q - rt_load(rt_a).x < nr * p.flat_run_bytes)
                            emit_runs_flat<POLICY>(tile_a, rt_a, nr, base32, lane);
                        else
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

