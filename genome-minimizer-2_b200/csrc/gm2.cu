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
#include "host_tokenize.hpp"
#include "host_genbank.hpp"
#include "host_expand.hpp"

#define GM2_API extern "C" __attribute__((visibility("default")))

// ------------------------------------------------------------------------------------------
// kernels (one translation unit; each header documents the reference lines it replaces)
// ------------------------------------------------------------------------------------------
#include "device_util.cuh"
#include "k1_keep.cuh"
#include "k2_plan.cuh"
#include "k3_scan.cuh"
#include "k4_emit.cuh"
#include "k5_emit_packed.cuh"
#include "diag.cuh"

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------

static const char kDefaultPrefix[] = "Minimized_E_coli_K12_MG1655_";   // minimizer_2.py:476

struct gm2_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;        // where work is issued (own or adopted)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_emit[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_copy[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_order = nullptr;       // gm2_order_after
    std::string err;
    uint64_t launches = 0;

    // configuration
    int tile_bytes = 49152;        // bases staged per CTA: what gm2_set_reference last used (configured or chosen)
    int tile_bytes_cfg = 0;        // GM2_CFG_TILE_BYTES; 0 = chosen from the reference's gene density (auto_tile_bytes)
    int emit_warps = 8;
    int emit_batch = 0;
    int packing_req = 0;
    int store_policy = 1;
    int rt_cap = 64;
    int debug = 0;
    int order = 1;
    int wire = 0;                  // gm2_emit_host transport: 0 auto, 1 bytes, 2 two-bit + host expansion
    int host_threads = 0;          // host expansion threads (0: hardware threads / LOCAL_WORLD_SIZE)
    int flat_run_bytes = 640;      // visits with a shorter mean kept run take the flat form; crossover measured in profiles/
    int flat_mode = 0;             // short-run form: 0 auto (from the kept fraction, see launch_emit), 1 emit_runs_flat
                                   // (cursor per lane, chosen per flushed batch), 2 visit_flat (bitmap-indexed, chosen per visit)
    int emit_occupancy = 0;        // k_emit CTAs per SM: 0 auto (from the kept fraction, see launch_emit), 3, 4
    HeaderPrefix prefix;

    // reference
    bool have_ref = false;
    int64_t G = 0;
    int32_t F = 0, FW = 0;
    int ntiles = 0, nseg = 0, nslots = 0, SW = 0, max_tile_slots = 0;
    int packing = 1;
    bool acgt_only = false;            // the reference holds nothing but upper-case A, C, G, T (two-bit forms usable)
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
    int32_t* d_hdr_len = nullptr; int64_t hdr_len_cap = 0;      // bytes of every record's header line (k_plan)
    unsigned long long* d_scan_desc = nullptr; int64_t scan_desc_cap = 0;
    unsigned int* d_scan_ticket = nullptr;
    // plan outputs (pinned host mirror)
    int64_t *h_len = nullptr, *h_rec_off = nullptr; int64_t h_cap = 0;
    bool planned = false, host_plan = false;
    int64_t first_idx = 0;
    // kept fraction (kept bases / (S * G)) of the most recent plan whose total has reached the host; -1 unknown.
    // Feeds the occupancy choice of k_emit only, never the output.  A plan's image size travels to h_total
    // asynchronously (ev_total), so gm2_plan_async stays asynchronous.
    double kept_frac = -1.0;
    int64_t* h_total = nullptr; cudaEvent_t ev_total = nullptr;
    bool total_pending = false; int64_t total_S = 0;
    int last_emit_ctas = 0;            // CTAs per SM the last k_emit launch was configured for
    int last_flat_mode = 1;            // short-run form it was built with (1 or 2)

    // staging for gm2_emit_host
    uint8_t* d_stage[2] = {nullptr, nullptr}; int64_t stage_cap = 0;
    // two-bit wire format (k_emit_packed -> pinned host staging -> host_expand): device / pinned words, tile_off rows
    uint32_t* d_pstage[3] = {nullptr, nullptr, nullptr}; uint32_t* h_pstage[3] = {nullptr, nullptr, nullptr}; int64_t pstage_words = 0;
    int32_t* h_toff[3] = {nullptr, nullptr, nullptr}; int64_t h_toff_cap = 0;
    bool packed_attr_set = false; size_t packed_attr_smem = 0;
    gm2host::Pool* pool = nullptr;     // expansion workers, created on first use
    int64_t last_d2h_bytes = 0;        // device->host bytes moved by the last gm2_emit_host
    int last_wire = 1;                 // wire format it used
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

// A C++ exception must not cross the C ABI: entry points that allocate host memory are function-try-blocks.
#define GM2_CATCH(c, who) \
    catch (const std::bad_alloc&) { return fail((c), GM2_ERR_NOMEM, who ": out of host memory"); } \
    catch (const std::exception& e__) { return fail((c), GM2_ERR_INVALID, std::string(who ": ") + e__.what()); }


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
// Uploads go through the stream the kernels run on (it is created non-blocking, so nothing orders it
// behind the legacy default stream) and are complete when the call returns.
template <typename T>
static int dev_upload(gm2_ctx* c, T** p, const std::vector<T>& v) {
    if (*p) { CU(c, cudaStreamSynchronize(c->stream)); cudaFree(*p); *p = nullptr; }
    size_t n = std::max<size_t>(v.size(), 1);
    CU(c, cudaMalloc((void**)p, n * sizeof(T)));
    if (!v.empty()) {
        CU(c, cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    return GM2_OK;
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
    for (int i = 0; i < 3; ++i) {
        cudaEventCreateWithFlags(&c->ev_emit[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming);
    }
    if ((e = cudaMalloc((void**)&c->d_scan_ticket, sizeof(unsigned int))) != cudaSuccess ||
        (e = cudaMallocHost((void**)&c->h_total, 8)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->ev_total, cudaEventDisableTiming)) != cudaSuccess) {
        cuda_fail(nullptr, e, "gm2_create: cudaMalloc"); gm2_destroy(c); return GM2_ERR_CUDA;
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
                     c->d_tile_off, c->d_len, c->d_rec_size, c->d_rec_off, c->d_hdr_len, c->d_scan_desc, c->d_scan_ticket,
                     c->d_stage[0], c->d_stage[1]};
    for (void* p : frees) if (p) cudaFree(p);
    if (c->h_len) cudaFreeHost(c->h_len);
    if (c->h_rec_off) cudaFreeHost(c->h_rec_off);
    if (c->h_total) cudaFreeHost(c->h_total);
    if (c->ev_total) cudaEventDestroy(c->ev_total);
    if (c->ev_order) cudaEventDestroy(c->ev_order);
    delete c->pool;
    for (int i = 0; i < 3; ++i) {
        if (c->d_pstage[i]) cudaFree(c->d_pstage[i]);
        if (c->h_pstage[i]) cudaFreeHost(c->h_pstage[i]);
        if (c->h_toff[i]) cudaFreeHost(c->h_toff[i]);
    }
    for (int i = 0; i < 3; ++i) {
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
        if (value != 0 && (value < 4096 || value > 196608 || (value % 4096) != 0))
            return fail(c, GM2_ERR_INVALID, "tile bytes must be 0 (auto) or a multiple of 4096 in [4096, 196608]");
        c->tile_bytes_cfg = (int)value; if (value) c->tile_bytes = (int)value; return GM2_OK;
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
    case GM2_CFG_FLAT_RUN_BYTES:
        if (value < 0 || value > (1 << 20)) return fail(c, GM2_ERR_INVALID, "flat run bytes must be in 0..1048576");
        c->flat_run_bytes = (int)value; return GM2_OK;
    case GM2_CFG_FLAT_MODE:
        if (value < 0 || value > 2) return fail(c, GM2_ERR_INVALID, "flat mode must be 0 (auto), 1 (cursor per lane) or 2 (bitmap-indexed)");
        c->flat_mode = (int)value; return GM2_OK;
    case GM2_CFG_WIRE:
        if (value < 0 || value > 2) return fail(c, GM2_ERR_INVALID, "wire must be 0 (auto), 1 (bytes) or 2 (two-bit)");
        c->wire = (int)value; return GM2_OK;
    case GM2_CFG_HOST_THREADS:
        if (value < 0 || value > 256) return fail(c, GM2_ERR_INVALID, "host threads must be in 0..256");
        c->host_threads = (int)value; return GM2_OK;
    case GM2_CFG_EMIT_OCCUPANCY:
        if (value != 0 && value != 3 && value != 4) return fail(c, GM2_ERR_INVALID, "emit occupancy must be 0 (auto), 3 or 4");
        c->emit_occupancy = (int)value; return GM2_OK;
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
    case GM2_Q_LAST_WIRE:    *out = c->last_wire; return GM2_OK;
    case GM2_Q_LAST_D2H_BYTES: *out = c->last_d2h_bytes; return GM2_OK;
    case GM2_Q_LAST_EMIT_CTAS: *out = c->last_emit_ctas; return GM2_OK;
    case GM2_Q_LAST_FLAT_MODE: *out = c->last_flat_mode; return GM2_OK;
    case GM2_Q_TILE_BYTES: *out = c->tile_bytes; return GM2_OK;
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

GM2_API int gm2_order_after(gm2_ctx* c, gm2_ctx* other) {
    if (!c || !other) return GM2_ERR_INVALID;
    if (c == other) return GM2_OK;                                   // a stream is already ordered with itself
    if (c->device != other->device) return fail(c, GM2_ERR_INVALID, "gm2_order_after: the contexts are on different devices");
    CU(c, cudaSetDevice(c->device));
    if (!c->ev_order) CU(c, cudaEventCreateWithFlags(&c->ev_order, cudaEventDisableTiming));
    // the event is re-recorded per call: a wait captures the record that precedes it, later records do not move it
    CU(c, cudaEventRecord(c->ev_order, other->stream));
    CU(c, cudaStreamWaitEvent(c->stream, c->ev_order, 0));
    return GM2_OK;
}

GM2_API int gm2_set_header_prefix(gm2_ctx* c, const char* prefix) {
    if (!c || !prefix) return GM2_ERR_INVALID;
    if (strlen(prefix) > GM2_MAX_PREFIX - 1) return fail(c, GM2_ERR_INVALID, "header prefix too long");
    set_prefix(c, prefix);
    c->planned = false; c->host_plan = false;
    return GM2_OK;
}

// Tile size when none is configured.  k_emit's short-run form handles the kept runs of a (sample, tile) visit 32 at
// a time, lane <-> run, and a visit meets about one run per gene of the tile when few genes are kept; so the
// tile is sized for ~34 genes (the 33rd run costs a second, nearly empty pass over the lanes), rounded to 4 KB
// and clamped to [24 KB, 48 KB].  Measured on both benchmark shapes (profiles/r02_emit_low_retention.md): 36 KB
// for 4,400 genes on 4.64 Mbp, 40 KB for 10,000 genes on 12 Mbp, +6-8 % at 10-20 % gene retention against a fixed
// 48 KB; from 50 % retention up every size in that range writes at the same rate.
static int auto_tile_bytes(int64_t G, int64_t genes) {
    if (genes <= 0 || G <= 0) return 49152;
    const int64_t want = 34 * G / genes;
    const int64_t t = ((want + 2048) / 4096) * 4096;
    return (int)std::min<int64_t>(std::max<int64_t>(t, 24576), 49152);
}

GM2_API int gm2_set_reference(gm2_ctx* c, const uint8_t* seq, int64_t G,
                              const int64_t* gs, const int64_t* ge, int32_t F)
try {
    if (!c) return GM2_ERR_INVALID;
    if (G < 0 || F < 0 || (G > 0 && !seq) || (F > 0 && (!gs || !ge)))
        return fail(c, GM2_ERR_INVALID, "gm2_set_reference: bad arguments");
    if (G > (int64_t)0x7fff0000) return fail(c, GM2_ERR_INVALID, "gm2_set_reference: G must be below 2^31 - 65536");
    CU(c, cudaSetDevice(c->device));
    if (c->tile_bytes_cfg == 0) {
        int64_t genes = 0;
        for (int32_t g = 0; g < F; ++g) genes += std::min(std::max<int64_t>(ge[g], 0), G) > std::min(std::max<int64_t>(gs[g], 0), G);
        c->tile_bytes = auto_tile_bytes(G, genes);
    }
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
    CU(c, cudaMemsetAsync(c->d_seq, 0, seq_alloc, c->stream));
    if (G > 0) CU(c, cudaMemcpyAsync(c->d_seq, seq, (size_t)G, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
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
    {
        // two bits per base, for k_emit<.,.,2> (GM2_CFG_PACKING 2: measured slower than bytes,
        // profiles/r01_emit_experiments.md, kept selectable) and for the two-bit wire format of gm2_emit_host
        bool acgt = true;
        std::vector<uint8_t> packed((size_t)ntiles * (size_t)(T / 4) + 256, 0);
        for (int64_t i = 0; i < G && acgt; ++i) {
            uint8_t code = 0;
            switch (seq[i]) {
            case 'A': code = 0; break; case 'C': code = 1; break; case 'G': code = 2; break; case 'T': code = 3; break;
            default: acgt = false;
            }
            packed[(size_t)(i >> 2)] |= (uint8_t)(code << (2 * (i & 3)));
        }
        c->acgt_only = acgt && G > 0;
        if (c->packing_req == 2 && !acgt)
            return fail(c, GM2_ERR_INVALID, "gm2_set_reference: two-bit packing needs an upper-case ACGT-only sequence");
        if (acgt) { if ((rc = dev_upload(c, &c->d_seq2, packed))) return rc; }
        if (c->packing_req == 2) c->packing = 2;
    }
    c->have_ref = true; c->planned = false; c->host_plan = false; c->mode = 0; c->S = 0;
    c->kept_frac = -1.0; c->total_pending = false;       // a new genome: no history for the occupancy choice
    return GM2_OK;
} GM2_CATCH(c, "gm2_set_reference")

GM2_API int gm2_set_name_map(gm2_ctx* c, const int32_t* off, const int32_t* idx, int32_t V) try {
    if (!c) return GM2_ERR_INVALID;
    if (!c->have_ref) return fail(c, GM2_ERR_STATE, "gm2_set_name_map: call gm2_set_reference first");
    if (V < 0 || !off) return fail(c, GM2_ERR_INVALID, "gm2_set_name_map: bad arguments");
    if (off[0] != 0) return fail(c, GM2_ERR_INVALID, "gm2_set_name_map: off[0] must be 0");
    if (c->F >= K1_SINGLE) return fail(c, GM2_ERR_INVALID, "gm2_set_name_map: more than 2^30 genes");
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
    {
        std::vector<uint32_t> hg(((size_t)V + 31) / 32 + 1, 0u);
        for (int32_t id = 0; id < V; ++id) if (first[id] >= 0) hg[(size_t)id >> 5] |= 1u << (id & 31);
        if ((rc = dev_upload(c, &c->d_has_gene, hg))) return rc;
    }
    // names borne by exactly one gene (all but a handful) are flagged: the keep builders then never read next_same[]
    for (int32_t id = 0; id < V; ++id) if (first[id] >= 0 && next[first[id]] < 0) first[id] |= K1_SINGLE;
    if ((rc = dev_upload(c, &c->d_first_gene, first))) return rc;
    if ((rc = dev_upload(c, &c->d_next_same, next))) return rc;
    c->V = V;
    if (c->d_forced_ids) { cudaFree(c->d_forced_ids); c->d_forced_ids = nullptr; }
    if (c->d_force_keep) { cudaFree(c->d_force_keep); c->d_force_keep = nullptr; }
    return GM2_OK;
} GM2_CATCH(c, "gm2_set_name_map")

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

GM2_API int gm2_set_forced(gm2_ctx* c, const uint32_t* force_keep, const uint32_t* forced_ids) try {
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
} GM2_CATCH(c, "gm2_set_forced")

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
    if ((rc = dev_reserve(c, &c->d_hdr_len, &c->hdr_len_cap, S + 1))) return rc;
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
            const size_t sm = ((size_t)c->FW + (size_t)(c->V + 31) / 32) * 4;
            if (sm > 200 * 1024) return fail(c, GM2_ERR_INVALID, "gm2_plan: too many genes / name ids for the keep builder's shared memory");
            if (sm > 48 * 1024) CU(c, cudaFuncSetAttribute(k_keep_from_ids, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            k_keep_from_ids<<<(unsigned)S, K1_THREADS, sm, c->stream>>>(c->ids, c->ids_off, S, c->V, c->d_first_gene,
                                                                        c->d_next_same, c->d_has_gene, c->FW, c->own_keep);
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
                                                         c->d_len, c->d_rec_size, c->d_hdr_len, first_idx, c->prefix.len);
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
    // the image size follows the plan to the host without a synchronisation (see set_kept_frac)
    CU(c, cudaMemcpyAsync(c->h_total, c->d_rec_off + S, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaEventRecord(c->ev_total, c->stream));
    c->total_pending = true; c->total_S = S;
    c->planned = true;
    return GM2_OK;
}

// kept bases / (S * G) from an image size: records are '>' prefix digits '\n' bases '\n'; the digit
// count is taken as 6 — the fraction only steers an occupancy choice.
static void set_kept_frac(gm2_ctx* c, int64_t image_bytes, int64_t S) {
    if (S <= 0 || c->G <= 0) return;
    const double framing = (double)S * (double)(c->prefix.len + 8);
    const double kept = std::max(0.0, (double)image_bytes - framing);
    c->kept_frac = std::min(1.0, kept / ((double)S * (double)c->G));
}
static void poll_kept_frac(gm2_ctx* c) {
    if (c->total_pending && cudaEventQuery(c->ev_total) == cudaSuccess) {
        c->total_pending = false;
        set_kept_frac(c, *c->h_total, c->total_S);
    }
}

static int pull_plan(gm2_ctx* c) try {
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
    c->total_pending = false;
    set_kept_frac(c, c->h_rec_off[S] - c->h_rec_off[0], S);
    return GM2_OK;
} GM2_CATCH(c, "pull_plan")

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
GM2_API int gm2_get_lengths_dev(gm2_ctx* c, int64_t* dev_out) {
    if (!c) return GM2_ERR_INVALID;
    if (!c->planned) return fail(c, GM2_ERR_STATE, "gm2_get_lengths_dev: call gm2_plan / gm2_plan_async first");
    if (c->S > 0 && !dev_out) return fail(c, GM2_ERR_INVALID, "gm2_get_lengths_dev: dev_lengths is NULL");
    CU(c, cudaSetDevice(c->device));
    if (c->S > 0) CU(c, cudaMemcpyAsync(dev_out, c->d_len, (size_t)c->S * 8, cudaMemcpyDeviceToDevice, c->stream));
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
    poll_kept_frac(c);
    // Short-run form: the bitmap-indexed whole-visit form (2) is at least as fast as the per-batch cursor form (1)
    // at every retention measured (0.1 ... 0.9, both genome shapes) and 15-25 % faster below 30 %, so auto means 2
    // wherever it applies (tile <= 60 KB); 1 stays selectable for A/B runs.
    const bool want2 = c->flat_mode == 2 || c->flat_mode == 0;
    const bool flat2 = want2 && c->tile_bytes <= FLAT_MAX_TILE;
    // Packing of the staged tile: bytes unless GM2_CFG_PACKING asked for two bits.  (Measured and dropped: staging the
    // two-bit copy automatically for low-retention launches — from a two-bit tile a stored vector is two 1-wavefront
    // LDS.32 instead of two 4-wavefront LDS.128, which relieves the L1 data pipe, but the expansion in registers costs
    // more than that returns: 0.745 vs 0.811 of peak at 10 % gene retention, profiles/r02_sweep_two_bit_tile_rejected.log.)
    const bool two_bit = c->packing == 2;
    p.seq = two_bit ? c->d_seq2 : c->d_seq; p.tile_smem_bytes = two_bit ? c->tile_bytes / 4 : c->tile_bytes;
    p.tile_slot = c->d_tile_slot; p.slot_src = c->d_slot_src; p.slot_len = c->d_slot_len;
    p.segkept = c->d_segkept; p.tile_off = c->d_tile_off; p.lengths = c->d_len; p.rec_off = c->d_rec_off;
    p.hdr_len = c->d_hdr_len;
    p.out = dev_out; p.s0 = s0; p.s1 = s1; p.first_idx = c->first_idx;
    p.tile_bytes = c->tile_bytes; p.ntiles = c->ntiles; p.SW = c->SW; p.batch = (int)batch; p.nbatch = (int)nbatch;
    p.slot_cap = c->max_tile_slots <= 4096 ? c->max_tile_slots : 0;     // else: slot tables read from global
    p.prefix = c->prefix; p.debug = c->debug; p.order = c->order; p.flat_run_bytes = c->flat_run_bytes;
    // Two builds of k_emit: 64 registers (4 CTAs of 256 threads per SM, needs <= 56 KB of shared memory per CTA:
    // with the default 48 KB tile that means a 32-entry run table) and 72 registers (3 CTAs).  Same plan, same
    // bytes either way, so the choice is made here, per launch, from the kept fraction of the latest plan known
    // to the host.
    auto smem_for = [&](int rt) {
        return 32 + (size_t)p.tile_smem_bytes + 64 + (size_t)p.slot_cap * 8 + (size_t)warps * (rt + 2) * 24;
    };
    const size_t dense_limit = 56 * 1024;
    int rt_cap = c->rt_cap;
    // Launch form, re-measured with the bitmap-indexed short-run form and the gene-density tile
    // (profiles/r02_emit_low_retention.md): with tiles of 28-48 KB the 72-register build is the faster one at
    // every retention (fewer instructions; at low retention the L1 data pipe — shared loads + global stores —
    // is the limit, not latency, and above ~45 % kept the kernel is write-bound and 3 CTAs are 2 % faster).
    // Only small tiles (<= 24 KB: gene-dense references) below ~43 % kept gain from the fourth CTA (+5-7 %).
    const bool want4 = c->emit_occupancy == 4 ||
                       (c->emit_occupancy == 0 && c->tile_bytes <= 24576 && c->kept_frac >= 0.0 && c->kept_frac < 0.43);
    if (want4 && smem_for(rt_cap) > dense_limit && smem_for(32) <= dense_limit) rt_cap = 32;
    p.rt_cap = rt_cap;
    // visit_flat keeps packed 4-byte entries in the same per-warp region ((rt_cap + 2) * 24 bytes): table A and the
    // event table with 2 * rt_cap + 2 entries each, the rest (2 * rt_cap + 8 words) holds rt_cap + 4 bitmap entries
    // {event word, skip word}, one per 512-byte output row; phase I reads pairs of entries one pair ahead, so
    // the last five stay spare.
    p.flat_cap = flat2 ? 2 * rt_cap : 0;
    p.flat_bm_words = flat2 ? rt_cap - 1 : 0;
    const size_t sm = smem_for(rt_cap);
    if (sm > 227 * 1024) return fail(c, GM2_ERR_INVALID, "gm2_emit: shared memory budget exceeded; lower tile bytes / emit warps");
    const bool dense = want4 && sm <= dense_limit;         // which build: 64 registers (4 CTAs of 256 threads) or 72 (3)
    c->last_emit_ctas = dense ? 4 : 3;
    c->last_flat_mode = flat2 ? 2 : 1;
    void (*kern)(const EmitParams);
    if (two_bit && flat2)
        kern = c->store_policy == 1 ? (dense ? k_emit<1, 4, 2, 2> : k_emit<1, 3, 2, 2>) : (dense ? k_emit<0, 4, 2, 2> : k_emit<0, 3, 2, 2>);
    else if (two_bit)
        kern = c->store_policy == 1 ? (dense ? k_emit<1, 4, 2, 1> : k_emit<1, 3, 2, 1>) : (dense ? k_emit<0, 4, 2, 1> : k_emit<0, 3, 2, 1>);
    else if (flat2)
        kern = c->store_policy == 1 ? (dense ? k_emit<1, 4, 1, 2> : k_emit<1, 3, 1, 2>) : (dense ? k_emit<0, 4, 1, 2> : k_emit<0, 3, 1, 2>);
    else
        kern = c->store_policy == 1 ? (dense ? k_emit<1, 4, 1, 1> : k_emit<1, 3, 1, 1>) : (dense ? k_emit<0, 4, 1, 1> : k_emit<0, 3, 1, 1>);
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

// ---- two-bit wire format for the host path (k5_emit_packed.cuh, host_expand.cpp) ------------------
static int64_t env_i64(const char* name, int64_t dflt) {
    const char* e = getenv(name);
    if (!e || !*e) return dflt;
    const long long v = atoll(e);
    return v > 0 ? (int64_t)v : dflt;
}

static inline int64_t packed_words(const gm2_ctx* c, int64_t a, int64_t b) {
    return ((c->h_rec_off[b] - c->h_rec_off[a]) >> 4) + (b - a) * (int64_t)(c->ntiles + 2) + 8;
}

static int launch_emit_packed(gm2_ctx* c, int64_t s0, int64_t s1, uint32_t* dev_out) {
    const int64_t n = s1 - s0;
    if (n <= 0) return GM2_OK;
    const int warps = c->emit_warps;
    int64_t batch = c->emit_batch;
    if (batch <= 0) {
        const int64_t want_ctas = (int64_t)c->sm_count * 8;
        batch = (n * c->ntiles + want_ctas - 1) / want_ctas;
        batch = std::max<int64_t>(batch, warps);
        batch = std::min<int64_t>(batch, 64);
        batch = ((batch + warps - 1) / warps) * warps;
    }
    const int64_t nbatch = (n + batch - 1) / batch;
    const int64_t blocks = nbatch * c->ntiles;
    if (blocks > 0x7fffffffLL) return fail(c, GM2_ERR_INVALID, "gm2_emit_host: grid too large; use a smaller chunk");
    PackedParams p;
    p.seq2 = c->d_seq2; p.tile_slot = c->d_tile_slot; p.slot_src = c->d_slot_src; p.slot_len = c->d_slot_len;
    p.segkept = c->d_segkept; p.tile_off = c->d_tile_off; p.rec_off = c->d_rec_off; p.out = dev_out;
    p.s0 = s0; p.s1 = s1; p.tile_bytes = c->tile_bytes; p.ntiles = c->ntiles; p.SW = c->SW;
    p.batch = (int)batch; p.nbatch = (int)nbatch; p.rt_cap = c->rt_cap;
    p.slot_cap = c->max_tile_slots <= 4096 ? c->max_tile_slots : 0;
    p.order = c->order;
    const size_t sm = 32 + (size_t)(c->tile_bytes / 4) + 64 + (size_t)p.slot_cap * 8 + (size_t)warps * (p.rt_cap + 2) * 8;
    if (sm > 227 * 1024) return fail(c, GM2_ERR_INVALID, "gm2_emit_host: shared memory budget exceeded; lower tile bytes / emit warps");
    if (!c->packed_attr_set || c->packed_attr_smem != sm) {       // once per configuration, not per chunk
        CU(c, cudaFuncSetAttribute(k_emit_packed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        c->packed_attr_set = true; c->packed_attr_smem = sm;
    }
    k_emit_packed<<<(unsigned)blocks, warps * 32, sm, c->stream>>>(p);
    LAUNCH_CHECK(c, "k_emit_packed");
    return GM2_OK;
}

static int emit_host_packed(gm2_ctx* c, int64_t s0, int64_t s1, uint8_t* host_out, int64_t chunk_bytes) try {
    const int nt = c->ntiles;
    const int threads = c->host_threads > 0 ? c->host_threads : gm2host::default_threads();
    if (!c->pool || c->pool->threads() != threads) { delete c->pool; c->pool = nullptr; c->pool = new gm2host::Pool(threads); }
    // Chunks: at most chunk_bytes of image AND at most chunk_bytes/16 words of per-piece slack, so that the
    // staging need is bounded by the chunk size alone (chunk_bytes/8 words, chunk_bytes/16 tile_off entries)
    // and buffers sized once serve every later call; only a single record larger than that grows them.
    std::vector<std::pair<int64_t, int64_t>> chunks;
    const int64_t slack_cap = std::max<int64_t>(chunk_bytes >> 4, 1);
    int64_t max_words = (chunk_bytes >> 3) + 16, max_rows = slack_cap / (nt + 2) + 1;
    for (int64_t a = s0; a < s1;) {
        int64_t b = a + 1;
        while (b < s1 && c->h_rec_off[b + 1] - c->h_rec_off[a] <= chunk_bytes && (b + 1 - a) * (int64_t)(nt + 2) <= slack_cap) ++b;
        chunks.emplace_back(a, b);
        max_words = std::max(max_words, packed_words(c, a, b));
        max_rows = std::max(max_rows, b - a);
        a = b;
    }
    if (max_words > c->pstage_words) {
        CU(c, cudaStreamSynchronize(c->stream)); CU(c, cudaStreamSynchronize(c->copy_stream));
        for (int i = 0; i < 3; ++i) {
            if (c->d_pstage[i]) cudaFree(c->d_pstage[i]);
            if (c->h_pstage[i]) cudaFreeHost(c->h_pstage[i]);
            c->d_pstage[i] = nullptr; c->h_pstage[i] = nullptr;
        }
        c->pstage_words = 0;
        for (int i = 0; i < 3; ++i) {
            CU(c, cudaMalloc((void**)&c->d_pstage[i], (size_t)max_words * 4));
            CU(c, cudaMallocHost((void**)&c->h_pstage[i], (size_t)max_words * 4 + 64));    // + the decoder's over-read
        }
        c->pstage_words = max_words;
    }
    if (max_rows * nt > c->h_toff_cap) {
        CU(c, cudaStreamSynchronize(c->copy_stream));
        for (int i = 0; i < 3; ++i) { if (c->h_toff[i]) cudaFreeHost(c->h_toff[i]); c->h_toff[i] = nullptr; }
        c->h_toff_cap = 0;
        for (int i = 0; i < 3; ++i) CU(c, cudaMallocHost((void**)&c->h_toff[i], (size_t)(max_rows * nt) * 4));
        c->h_toff_cap = max_rows * nt;
    }
    // Three staging buffers: the GPU side (k_emit_packed + the two copies of a chunk) runs up to two chunks ahead
    // of the expansion.  Per chunk the caller's thread waits for the chunk's copy, starts the expansion workers,
    // enqueues the GPU work of the chunk two ahead WHILE they decode, and then decodes with them.
    auto enqueue = [&](size_t i) -> int {
        const int buf = (int)(i % 3);
        const int64_t a = chunks[i].first, b = chunks[i].second;
        // d_pstage[buf] / h_pstage[buf] were last used by chunk i-3: its copy is ordered by the event, its
        // expansion finished on this thread before this call
        if (i >= 3) CU(c, cudaStreamWaitEvent(c->stream, c->ev_copy[buf], 0));
        int rc2 = launch_emit_packed(c, a, b, c->d_pstage[buf]);
        if (rc2) return rc2;
        CU(c, cudaEventRecord(c->ev_emit[buf], c->stream));
        CU(c, cudaStreamWaitEvent(c->copy_stream, c->ev_emit[buf], 0));
        const int64_t words = packed_words(c, a, b), rows = (b - a) * nt;
        CU(c, cudaMemcpyAsync(c->h_pstage[buf], c->d_pstage[buf], (size_t)words * 4, cudaMemcpyDeviceToHost, c->copy_stream));
        CU(c, cudaMemcpyAsync(c->h_toff[buf], c->d_tile_off + (size_t)a * nt, (size_t)rows * 4, cudaMemcpyDeviceToHost, c->copy_stream));
        CU(c, cudaEventRecord(c->ev_copy[buf], c->copy_stream));
        c->last_d2h_bytes += words * 4 + rows * 4;
        return GM2_OK;
    };
    int rc;
    for (size_t i = 0; i < chunks.size() && i < 2; ++i) if ((rc = enqueue(i))) return rc;
    for (size_t i = 0; i < chunks.size(); ++i) {
        const int buf = (int)(i % 3);
        CU(c, cudaEventSynchronize(c->ev_copy[buf]));
        gm2host::ChunkView v;
        v.packed = c->h_pstage[buf]; v.tile_off = c->h_toff[buf]; v.rec_off = c->h_rec_off; v.lengths = c->h_len;
        v.out = host_out + (c->h_rec_off[chunks[i].first] - c->h_rec_off[s0]);
        v.s0 = chunks[i].first; v.s1 = chunks[i].second; v.first_idx = c->first_idx; v.ntiles = nt;
        v.prefix = c->prefix.text; v.prefix_len = c->prefix.len; v.simd = 1;
        c->pool->start(v);
        rc = i + 2 < chunks.size() ? enqueue(i + 2) : GM2_OK;
        c->pool->finish();
        if (rc) return rc;
    }
    CU(c, cudaStreamSynchronize(c->copy_stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return GM2_OK;
} GM2_CATCH(c, "gm2_emit_host")

GM2_API int gm2_emit_host(gm2_ctx* c, int64_t s0, int64_t s1, uint8_t* host_out, int64_t cap, int64_t chunk_bytes) {
    if (!c) return GM2_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    int rc = pull_plan(c); if (rc) return rc;
    if (s0 < 0 || s1 < s0 || s1 > c->S) return fail(c, GM2_ERR_INVALID, "gm2_emit_host: bad sample range");
    const int64_t total = c->h_rec_off[s1] - c->h_rec_off[s0];
    if (total > cap) return fail(c, GM2_ERR_CAPACITY, "gm2_emit_host: output buffer too small");
    if (total > 0 && !host_out) return fail(c, GM2_ERR_INVALID, "gm2_emit_host: host_out is NULL");
    const bool default_chunk = chunk_bytes <= 0;
    if (default_chunk) chunk_bytes = (int64_t)256 << 20;
    c->last_d2h_bytes = 0; c->last_wire = 1;
    if (c->wire == 2 && !c->acgt_only)
        return fail(c, GM2_ERR_STATE, "gm2_emit_host: the two-bit wire format needs an ACGT-only reference (GM2_CFG_WIRE)");
    // auto: a process alone on the host needs ~6 expansion threads to beat its plain copy (measured: 50 / 68 / 78 /
    // 98 Gbp/s with 4 / 6 / 8 / 16 threads against 55 for the copy).  Several ranks on one host share its memory
    // system, which caps their plain copies well below PCIe (17.6 GB/s per GPU at 4 ranks), and 4 threads each are
    // enough (4 ranks: 105 Gbp/s two-bit against 69); with fewer threads per rank the plain copy is kept.
    const int host_threads = c->host_threads > 0 ? c->host_threads : gm2host::default_threads();
    if (total > 0 && (c->wire == 2 || (c->wire == 0 && c->acgt_only && host_threads >= (gm2host::local_ranks() > 1 ? 4 : 6)))) {
        c->last_wire = 2;
        // Chunk size of the pipeline: 16 MB of image per chunk (GM2_WIRE_CHUNK_BYTES).  The copy of chunk i+1 hides
        // under the expansion of chunk i; a caller that asks for one 256 MB range at a time (engine.drain) pays a short
        // pipeline fill per call.  Bigger chunks mean fewer hand-overs but a staging ring (3 x chunk / 4) that no longer
        // stays in the last-level cache: measured, 64 MB chunks are 10 % faster than 16 MB ones on this pool's 24-core
        // hosts and 7 % slower (file outputs 30 % slower) on its 16-core hosts (profiles/r02_host_path.md).
        const int64_t wire_chunk = default_chunk ? env_i64("GM2_WIRE_CHUNK_BYTES", (int64_t)16 << 20) : chunk_bytes;
        return emit_host_packed(c, s0, s1, host_out, wire_chunk);
    }
    c->last_d2h_bytes = total;
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

GM2_API int gm2_device_alloc(gm2_ctx* c, void** out, int64_t bytes) {
    if (!c || !out || bytes < 0) return GM2_ERR_INVALID;
    *out = nullptr;
    CU(c, cudaSetDevice(c->device));
    cudaError_t e = cudaMalloc(out, (size_t)std::max<int64_t>(bytes, 1));
    if (e != cudaSuccess) { cuda_fail(c, e, "cudaMalloc"); return GM2_ERR_NOMEM; }
    return GM2_OK;
}
GM2_API int gm2_device_free(gm2_ctx* c, void* p) {
    if (!c) return GM2_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    if (p) { CU(c, cudaStreamSynchronize(c->stream)); CU(c, cudaFree(p)); }
    return GM2_OK;
}
GM2_API int gm2_upload(gm2_ctx* c, void* dev, const void* host, int64_t bytes) {
    if (!c || bytes < 0 || (bytes > 0 && (!dev || !host))) return GM2_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    if (bytes > 0) {
        CU(c, cudaMemcpyAsync(dev, host, (size_t)bytes, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    return GM2_OK;
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
GM2_API int gm2_sequence_hashes(gm2_ctx* c, int64_t s0, int64_t s1, uint64_t* out) try {
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
} GM2_CATCH(c, "gm2_sequence_hashes")

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
try {
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
} GM2_CATCH(c, "gm2_diag_range_hashes")


// ------------------------------------------------------------------------------------------
// host-only: gene-name lists container -> id CSR (see host_tokenize.hpp)
// ------------------------------------------------------------------------------------------
GM2_API int gm2_tokenize_pickle(const uint8_t* body, int64_t nbytes, int64_t S, int64_t L,
                                const char* names, const int64_t* name_off, int32_t V,
                                int32_t* ids_out, int64_t ids_cap, int64_t* off_out, int64_t* count_out) {
    if (!body || nbytes <= 0 || S < 0 || L < 0 || V < 0 || (V > 0 && (!names || !name_off)) || !off_out || !count_out ||
        (ids_cap > 0 && !ids_out))
        return fail(nullptr, GM2_ERR_INVALID, "gm2_tokenize_pickle: bad argument");
    try {
        gm2tok::Machine m;
        m.p = body; m.end = body + nbytes;
        m.vocab.reserve((size_t)V * 2);
        for (int32_t v = 0; v < V; ++v)
            m.vocab.emplace(std::string_view(names + name_off[v], (size_t)(name_off[v + 1] - name_off[v])), v);
        const int rc = m.run();
        if (rc == 1) return fail(nullptr, GM2_ERR_UNSUPPORTED, "gm2_tokenize_pickle: pickle opcode outside the supported subset");
        if (rc != 0) return fail(nullptr, GM2_ERR_INVALID, "gm2_tokenize_pickle: corrupt pickle stream");
        // the outermost object is built last: its state is (version, shape, dtype, fortran, [elements])
        if (m.last_state < 0 || m.seqs[m.last_state].empty() || !gm2tok::is_seq(m.seqs[m.last_state].back()))
            return fail(nullptr, GM2_ERR_UNSUPPORTED, "gm2_tokenize_pickle: not a pickled object array");
        const std::vector<gm2tok::Item>& flat = m.seqs[gm2tok::seq_index(m.seqs[m.last_state].back())];
        const int64_t want = L > 0 ? S * L : S;
        if ((int64_t)flat.size() != want) return fail(nullptr, GM2_ERR_INVALID, "gm2_tokenize_pickle: element count does not match the shape");
        int64_t n = 0;
        off_out[0] = 0;
        for (int64_t i = 0; i < S; ++i) {
            const int32_t* it; int64_t cnt;
            if (L > 0) { it = flat.data() + i * L; cnt = L; }
            else {
                if (!gm2tok::is_seq(flat[i])) return fail(nullptr, GM2_ERR_UNSUPPORTED, "gm2_tokenize_pickle: an element is not a list");
                const std::vector<gm2tok::Item>& lst = m.seqs[gm2tok::seq_index(flat[i])];
                it = lst.data(); cnt = (int64_t)lst.size();
            }
            for (int64_t k = 0; k < cnt; ++k) {
                const int32_t t = it[k];
                if (t < gm2tok::IT_UNKNOWN_STR) return fail(nullptr, GM2_ERR_UNSUPPORTED, "gm2_tokenize_pickle: a list item is not a str");
                if (t >= 0) {
                    if (n >= ids_cap) return fail(nullptr, GM2_ERR_CAPACITY, "gm2_tokenize_pickle: ids_out too small");
                    ids_out[n++] = t;
                }
            }
            count_out[i] = cnt;
            off_out[i + 1] = n;
        }
        return GM2_OK;
    } catch (const std::bad_alloc&) {
        return fail(nullptr, GM2_ERR_NOMEM, "gm2_tokenize_pickle: out of host memory");
    }
}

// ------------------------------------------------------------------------------------------
// host-only: GenBank flat file -> sequence + gene table (see host_genbank.hpp)
// ------------------------------------------------------------------------------------------
struct gm2_genbank { gm2gb::Genes g; };

GM2_API int gm2_genbank_parse(const uint8_t* text, int64_t nbytes, gm2_genbank** out) {
    if (!out) return fail(nullptr, GM2_ERR_INVALID, "gm2_genbank_parse: out is NULL");
    *out = nullptr;
    if (nbytes < 0 || (nbytes > 0 && !text)) return fail(nullptr, GM2_ERR_INVALID, "gm2_genbank_parse: bad argument");
    try {
        gm2_genbank* h = new gm2_genbank();
        if (gm2gb::scan(std::string_view(reinterpret_cast<const char*>(text), (size_t)nbytes), h->g) != gm2gb::Status::ok) {
            const std::string why = "gm2_genbank_parse: " + h->g.why;
            delete h;
            return fail(nullptr, GM2_ERR_UNSUPPORTED, why);
        }
        *out = h;
        return GM2_OK;
    } catch (const std::bad_alloc&) {
        return fail(nullptr, GM2_ERR_NOMEM, "gm2_genbank_parse: out of host memory");
    }
}
GM2_API int gm2_genbank_sizes(const gm2_genbank* h, int64_t* G, int32_t* F, int64_t* name_bytes, int64_t* n_features) {
    if (!h) return fail(nullptr, GM2_ERR_INVALID, "gm2_genbank_sizes: handle is NULL");
    if (h->g.start.size() > (size_t)0x7fffffff) return fail(nullptr, GM2_ERR_INVALID, "gm2_genbank_sizes: too many gene features");
    if (G) *G = (int64_t)h->g.seq.size();
    if (F) *F = (int32_t)h->g.start.size();
    if (name_bytes) *name_bytes = (int64_t)h->g.names.size();
    if (n_features) *n_features = h->g.n_features;
    return GM2_OK;
}
GM2_API int gm2_genbank_copy(const gm2_genbank* h, uint8_t* seq, int64_t* gene_start, int64_t* gene_end,
                             int64_t* name_off, uint8_t* names) {
    if (!h) return fail(nullptr, GM2_ERR_INVALID, "gm2_genbank_copy: handle is NULL");
    const gm2gb::Genes& g = h->g;
    if (seq && !g.seq.empty()) memcpy(seq, g.seq.data(), g.seq.size());
    if (gene_start && !g.start.empty()) memcpy(gene_start, g.start.data(), g.start.size() * 8);
    if (gene_end && !g.end.empty()) memcpy(gene_end, g.end.data(), g.end.size() * 8);
    if (name_off) memcpy(name_off, g.name_off.data(), g.name_off.size() * 8);
    if (names && !g.names.empty()) memcpy(names, g.names.data(), g.names.size());
    return GM2_OK;
}
GM2_API int gm2_genbank_free(gm2_genbank* h) {
    delete h;
    return GM2_OK;
}

// ------------------------------------------------------------------------------------------
// host-only: write ceiling of the expansion's store pattern (bench.py)
// ------------------------------------------------------------------------------------------
GM2_API int gm2_diag_host_fill(uint8_t* host, int64_t bytes, int32_t threads, int32_t reps, double* gbs) {
    if (!host || bytes <= 0 || !gbs) return fail(nullptr, GM2_ERR_INVALID, "gm2_diag_host_fill: bad argument");
    try {
        *gbs = gm2host::fill_probe(host, bytes, threads > 0 ? threads : gm2host::default_threads(), reps > 0 ? reps : 1);
        return GM2_OK;
    } catch (const std::exception& e) {
        return fail(nullptr, GM2_ERR_NOMEM, std::string("gm2_diag_host_fill: ") + e.what());
    }
}

// ------------------------------------------------------------------------------------------
// host-only: the decoder of the two-bit wire format on caller-supplied data (tests, probes)
// ------------------------------------------------------------------------------------------
GM2_API int gm2_diag_expand(const uint32_t* packed, const int32_t* tile_off, const int64_t* rec_off,
                            const int64_t* lengths, int64_t S, int32_t ntiles, int64_t first_idx,
                            const char* prefix, uint8_t* out, int32_t threads, int32_t simd) {
    if (S < 0 || ntiles <= 0 || !prefix || (S > 0 && (!packed || !tile_off || !rec_off || !lengths || !out)))
        return fail(nullptr, GM2_ERR_INVALID, "gm2_diag_expand: bad argument");
    const size_t n = strlen(prefix);
    if (n > GM2_MAX_PREFIX - 1) return fail(nullptr, GM2_ERR_INVALID, "gm2_diag_expand: prefix too long");
    std::string full = ">" + std::string(prefix);
    gm2host::ChunkView v;
    v.packed = packed; v.tile_off = tile_off; v.rec_off = rec_off; v.lengths = lengths; v.out = out;
    v.s0 = 0; v.s1 = S; v.first_idx = first_idx; v.ntiles = ntiles;
    v.prefix = full.c_str(); v.prefix_len = (int)full.size(); v.simd = simd;
    gm2host::expand_chunk(v, threads > 0 ? threads : gm2host::default_threads());
    return GM2_OK;
}
