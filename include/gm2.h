/* gm2.h — C-ABI of libgm2.so, the B200 (sm_100a) build of genome-minimizer-2's
 * `--mode minimizer` hot path.
 *
 * The reference (ucl-cssb/genome-minimizer-2) is pure Python and has NO native
 * interface of its own: the path is the constructor of
 * `GenomeMinimiser` (src/genome_minimizer_2/minimizer/minimizer_2.py:20-101) and the
 * two batch entry functions around it (:447-495, :499-560).  This header is the
 * boundary a maintainer would bind from those functions with ctypes; each entry
 * point cites the reference lines whose work it takes over.  INTEGRATION.md shows
 * the binding.
 *
 * Conventions
 *   - plain C, no exceptions, no torch types: pointers + sizes only.
 *   - every function returns GM2_OK (0) or a negative GM2_ERR_* code; the text of
 *     the last failure on a context is `gm2_last_error(ctx)`.
 *   - "host" pointers are ordinary (pageable or pinned) CPU memory owned by the
 *     caller and only read/written during the call.  "dev" pointers are CUDA device
 *     pointers on the context's device (e.g. `torch.Tensor.data_ptr()`).
 *   - one context per GPU; a context is not thread-safe, distinct contexts are
 *     independent.  All device work of a context is issued on its stream
 *     (`gm2_set_stream`), and calls that return host data synchronise that stream.
 *   - there is no CPU fallback: without a usable CUDA device `gm2_create` fails.
 *
 * Data model
 *   reference  : G upper-case ASCII bases + F `gene` intervals [start,end) in file order
 *   sample s   : a keep vector over the F genes (bit g set  <=>  name_g in the sample's list)
 *   output     : for s ascending   ">" prefix (first_idx+s+1) "\n"  kept bases  "\n"
 *                a base p is deleted iff some NOT-kept gene has start <= p < end.
 */
#ifndef GM2_H_
#define GM2_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GM2_ABI_VERSION 1

#define GM2_OK            0
#define GM2_ERR_INVALID  -1   /* bad argument                                  */
#define GM2_ERR_CUDA     -2   /* a CUDA runtime call or kernel launch failed   */
#define GM2_ERR_STATE    -3   /* call made in the wrong order                  */
#define GM2_ERR_CAPACITY -4   /* caller's output buffer is too small           */
#define GM2_ERR_NOMEM    -5   /* host or device allocation failed              */
#define GM2_ERR_UNSUPPORTED -6 /* input outside the subset this entry point handles (caller has a slower general path) */

/* gm2_configure keys */
#define GM2_CFG_TILE_BYTES    1  /* reference bases staged per CTA (multiple of 4096, or 0 = chosen from the gene
                                   density at gm2_set_reference, the default; before set_reference) */
#define GM2_CFG_EMIT_WARPS    2  /* warps per emit CTA (1..8)                                */
#define GM2_CFG_EMIT_BATCH    3  /* samples per emit CTA; 0 = choose from S and the SM count */
#define GM2_CFG_PACKING       4  /* 0 auto (= byte, the measured-faster form), 1 byte/base, 2 two-bit (ACGT-only; before set_reference) */
#define GM2_CFG_STORE_POLICY  5  /* 0 plain stores, 1 streaming st.global.cs (default)       */
#define GM2_CFG_RUN_TABLE     6  /* kept-run table entries per warp in shared memory (32..1024) */
#define GM2_CFG_ORDER         8  /* emit CTA order: 0 tile-major, 1 sample-major (default)   */
#define GM2_CFG_FLAT_RUN_BYTES 9 /* (sample, tile) batches whose mean kept-run length is below this many bytes use the
                                    vector-per-lane emit path instead of the run-by-run stream; 0 = never, 1048576 = always
                                    (default 640; byte packing only) */
#define GM2_CFG_WIRE          10 /* transport of gm2_emit_host / gm2_minimize_host: 0 auto (two-bit when the reference is
                                    ACGT-only and enough host threads are available: 6, or 4 per process when
                                    LOCAL_WORLD_SIZE says several processes share the host), 1 image bytes over PCIe,
                                    2 two bits per base over PCIe + expansion by host threads (error if not ACGT-only) */
#define GM2_CFG_HOST_THREADS  11 /* host threads for that expansion; 0 = hardware threads / LOCAL_WORLD_SIZE */
#define GM2_CFG_EMIT_OCCUPANCY 12 /* k_emit build / CTAs per SM: 0 auto (the 72-register build, 3 CTAs; the 64-register
                                   * build, 4 CTAs, only for tiles <= 24 KB when the latest plan kept less than ~43 % of
                                   * the bases), 3, or 4 (when the shared memory fits) */
#define GM2_CFG_FLAT_MODE     13 /* form of that short-run path: 2 = whole visit, vector -> run through a per-warp pair of
                                  * bitmaps (run starts / boundary vectors: no search, no compare), boundary vectors
                                  * merged from <= 3 windows; 1 = per flushed batch, private run cursor per lane;
                                  * 0 (default) = 2 wherever it applies (one byte per base, tile <= 60 KB) */
#define GM2_CFG_DEBUG         7  /* timing knock-outs (WRONG output); only effective in -DGM2_EMIT_DEBUG builds */

/* gm2_query keys */
#define GM2_Q_SM_COUNT        1
#define GM2_Q_LAUNCHES        2  /* kernels launched by this context so far */
#define GM2_Q_NUM_SEGMENTS    3  /* elementary segments between breakpoints */
#define GM2_Q_NUM_TILES       4
#define GM2_Q_PACKING         5  /* packing actually in use (1 or 2)        */
#define GM2_Q_NUM_SLOTS       6
#define GM2_Q_KEEP_WORDS      7  /* 32-bit words per keep row = ceil(F/32)  */
#define GM2_Q_LAST_WIRE       8  /* wire format the last gm2_emit_host used (1 or 2) */
#define GM2_Q_LAST_D2H_BYTES  9  /* device->host bytes the last gm2_emit_host moved  */
#define GM2_Q_LAST_EMIT_CTAS  10 /* CTAs per SM the last k_emit launch was configured for (3 or 4) */
#define GM2_Q_LAST_FLAT_MODE  11 /* short-run form of the last k_emit launch (1 or 2, see GM2_CFG_FLAT_MODE) */
#define GM2_Q_TILE_BYTES      12 /* tile size gm2_set_reference used (configured or chosen)               */

typedef struct gm2_ctx gm2_ctx;

int         gm2_abi_version(void);
/* Number of CUDA devices visible, or a negative error. */
int         gm2_device_count(void);

/* Lifetime.  `device` is a CUDA ordinal.  Fails (no fallback) if CUDA is unusable. */
int         gm2_create(int device, gm2_ctx** out);
int         gm2_destroy(gm2_ctx* ctx);
const char* gm2_last_error(const gm2_ctx* ctx);      /* ctx may be NULL: create-time error */
int         gm2_configure(gm2_ctx* ctx, int key, int64_t value);
int         gm2_query(const gm2_ctx* ctx, int key, int64_t* out);
/* Issue this context's work on the caller's CUDA stream (cudaStream_t passed as
 * void*).  NULL restores the context's own non-blocking stream; to target the legacy
 * default stream pass cudaStreamLegacy ((void*)0x1) explicitly. */
int         gm2_set_stream(gm2_ctx* ctx, void* cuda_stream);
int         gm2_sync(gm2_ctx* ctx);
/* Pipelines over several contexts of one GPU (a context is one plan slot: its plan state is overwritten
 * by its next gm2_plan).  Device work issued on `ctx` AFTER this call starts only when everything issued
 * on `other` BEFORE this call has finished; nothing blocks on the host.  Two contexts on the same
 * reference, chunk i on context i & 1:   load(i); plan_async(i); order_after(ctx[i&1], ctx[~i&1]);
 * emit_dev(i)   keeps the emits in file order on the device while chunk i+1 is planned under emit i
 * (bench.py, engine.ContextPair). */
int         gm2_order_after(gm2_ctx* ctx, gm2_ctx* other);

/* The record:  replaces `record.seq` (minimizer_2.py:35, :94) and, for every feature
 * with type == "gene" in file order, `int(feature.location.start)` /
 * `int(feature.location.end)` (minimizer_2.py:59-60, :78-79).  `seq` is G upper-case
 * ASCII bytes (any letter).  Intervals are clamped to [0,G]; start >= end deletes
 * nothing (Python `range` semantics).  Builds the static segment tables on the host
 * and uploads everything; host pointers are not retained. */
int gm2_set_reference(gm2_ctx* ctx, const uint8_t* seq, int64_t G,
                      const int64_t* gene_start, const int64_t* gene_end, int32_t F);

/* Name table:  replaces `feature.qualifiers.get("gene", [""])[0]` + `name not in
 * needed_genes` (minimizer_2.py:61-62) after the host has interned names to ids
 * 0..V-1.  CSR id -> gene indices (1:many: several genes may share a name). */
int gm2_set_name_map(gm2_ctx* ctx, const int32_t* id2gene_off /* V+1 */,
                     const int32_t* id2gene_idx, int32_t V);

/* Record header: '>' + prefix + decimal(first_idx+s+1) + '\n'.  Default prefix is the
 * reference's literal "Minimized_E_coli_K12_MG1655_" (minimizer_2.py:476, :537). */
int gm2_set_header_prefix(gm2_ctx* ctx, const char* prefix);

/* Samples in: either interned name-id lists (CSR; ids outside 0..V-1 and duplicates
 * are legal and ignored/idempotent) or ready keep rows of ceil(F/32) little-endian
 * 32-bit words each.  *_host copy from CPU memory into context-owned device buffers;
 * *_dev borrow device memory in place (zero copy): the caller keeps ownership and
 * must keep it alive and unchanged until the next gm2_load_* or gm2_destroy. */
int gm2_load_ids_host(gm2_ctx* ctx, const int32_t* ids, const int64_t* off /* S+1 */, int64_t S);
int gm2_load_ids_dev (gm2_ctx* ctx, const int32_t* ids, const int64_t* off /* S+1 */, int64_t S, int64_t n_ids);
int gm2_load_keep_host(gm2_ctx* ctx, const uint32_t* keep_rows, int64_t S);
int gm2_load_keep_dev (gm2_ctx* ctx, const uint32_t* keep_rows, int64_t S);

/* Dense form (BASELINE config 5): the samples are rows of a device-resident
 * float32 matrix probs[S][ld] (the VAE decoder's output, utils/extras.py:192-203); column c is
 * name id c of gm2_set_name_map; a column is present iff probs[s][c] > threshold (strict, as
 * utils/extras.py:200-201; identical to binary_converter.py:55 on the resulting 0/1 matrix).
 * gm2_set_forced registers genes that are always kept (the essentials check_essential_genes
 * appends, binary_converter.py:91-98; one bit per gene, ceil(F/32) words) and, for the list
 * length the reference prints, which ids are forced (one bit per id, ceil(V/32) words); either
 * may be NULL.  The matrix is borrowed like the other *_dev inputs.  gm2_get_counts returns per
 * sample (#ids above threshold) + (#forced ids not above threshold). */
int gm2_set_forced(gm2_ctx* ctx, const uint32_t* force_keep_genes, const uint32_t* forced_ids);
int gm2_load_probs_dev(gm2_ctx* ctx, const float* probs, int64_t S, int64_t ld, float threshold);
int gm2_get_counts(gm2_ctx* ctx, int64_t* counts /* S */);

/* K1 keep-mask builder (ids / dense modes), K2 segment flags, K3 scans: replaces
 * `_extract_non_essential_genes` (minimizer_2.py:50-66), `_get_positions_to_remove`
 * (:68-83) and the running output index of `_create_minimized_sequence` (:94-96).
 * `first_idx` is the global 0-based index of sample 0 (a rank's shard offset).
 * Synchronises; afterwards lengths / record offsets are readable. */
int gm2_plan(gm2_ctx* ctx, int64_t first_idx);
/* As gm2_plan but leaves everything on the device and does not synchronise
 * (gm2_emit_dev can follow on the same stream; gm2_get_* would sync). */
int gm2_plan_async(gm2_ctx* ctx, int64_t first_idx);

/* Results of the plan (host arrays owned by the caller).  lengths[s] = kept bases
 * L_s; rec_off[s] = byte offset of record s in the concatenated image, rec_off[S] =
 * image size.  keep_rows = the S x ceil(F/32) keep matrix (what K1 produced). */
int gm2_get_lengths(gm2_ctx* ctx, int64_t* lengths /* S */);
int gm2_get_record_offsets(gm2_ctx* ctx, int64_t* rec_off /* S+1 */);
int gm2_get_keep_rows(gm2_ctx* ctx, uint32_t* keep_rows /* S*ceil(F/32) */);
/* The same lengths left on the device: an asynchronous device-to-device copy of the S values into
 * `dev_lengths` on the context's stream, no synchronisation — what a multi-GPU caller all-gathers
 * (ncclAllGather / torch.distributed) to derive every rank's file offset. */
int gm2_get_lengths_dev(gm2_ctx* ctx, int64_t* dev_lengths /* S, device */);
/* Total image bytes for records [s0,s1) (needs a synchronised plan). */
int gm2_image_bytes(gm2_ctx* ctx, int64_t s0, int64_t s1, int64_t* out);

/* K4 compaction gather + FASTA framing: replaces `_create_minimized_sequence`
 * (minimizer_2.py:85-101) and the record write `f">{seq_id}\n{seq}\n"` (:476-477,
 * :544-545) for records [s0,s1), record s0 starting at dev_out[0].  Asynchronous on
 * the context's stream.  `cap` is checked against the image size when the plan has
 * been synchronised, otherwise it is trusted. */
int gm2_emit_dev(gm2_ctx* ctx, int64_t s0, int64_t s1, uint8_t* dev_out, int64_t cap);

/* Same, delivered to CPU memory: records [s0,s1) are produced in device staging
 * buffers in pieces of about `chunk_bytes` (0 = default) and copied to `host_out`
 * while the next piece is being produced.  Synchronous.  `host_out` should be pinned
 * (gm2_host_alloc) for full PCIe speed.  Transport (GM2_CFG_WIRE): either the finished image bytes
 * are copied, or — ACGT-only reference — the kept bases cross PCIe as 2 bits each and host threads
 * (GM2_CFG_HOST_THREADS) expand them into `host_out`, headers and newlines included; the bytes in
 * `host_out` are the same.  chunk_bytes 0 = default (256 MiB for the copy, 16 MiB for two-bit). */
int gm2_emit_host(gm2_ctx* ctx, int64_t s0, int64_t s1, uint8_t* host_out, int64_t cap,
                  int64_t chunk_bytes);

/* 64-bit hashes (definition: gm2_diag_range_hashes) of the minimized SEQUENCES of records [s0,s1)
 * — bases only, header and newline excluded — computed on the device from staged emits.  Backs the
 * reference's duplicate report (check_sequence_duplicates, minimizer_2.py:273-303): equal sequences
 * have equal (length, hash).  `out` is a host array of s1-s0 values. */
int gm2_sequence_hashes(gm2_ctx* ctx, int64_t s0, int64_t s1, uint64_t* out);

/* One-call convenience used by the batch entry functions: load keep rows or ids from
 * host memory, plan, and deliver the whole image to host memory. */
int gm2_minimize_host(gm2_ctx* ctx, const int32_t* ids, const int64_t* off,
                      const uint32_t* keep_rows, int64_t S, int64_t first_idx,
                      int64_t* lengths /* S */, int64_t* rec_off /* S+1 */,
                      uint8_t* host_out, int64_t cap, int64_t chunk_bytes);

/* Device memory owned by the caller, for the *_dev entry points when the caller has no CUDA
 * allocator of its own (a torch tensor's data_ptr() works just as well).  gm2_upload is a
 * synchronous host-to-device copy on the context's stream. */
int gm2_device_alloc(gm2_ctx* ctx, void** out, int64_t bytes);
int gm2_device_free(gm2_ctx* ctx, void* p);
int gm2_upload(gm2_ctx* ctx, void* dev, const void* host, int64_t bytes);

/* Pinned host memory for output buffers. */
int gm2_host_alloc(void** out, int64_t bytes);
int gm2_host_free(void* p);

/* Diagnostics used by bench.py only: a write-only fill of `bytes` bytes with 128-bit
 * stores (the write roofline of this device), on the context's stream. */
int gm2_diag_fill(gm2_ctx* ctx, uint8_t* dev, int64_t bytes, uint32_t pattern);
/* Diagnostics: a store-only model of the emit kernel's write pattern (nrec records `stride` bytes
 * apart, each written as ntile chunks of `chunk` bytes by one warp per (record, chunk), CTAs =
 * (chunk index, batch of records)).  Used to separate "what the memory system gives this pattern"
 * from "what the kernel's instructions cost". */
int gm2_diag_fill_streams(gm2_ctx* ctx, uint8_t* dev, int64_t nrec, int64_t stride, int32_t ntile,
                          int64_t chunk, int32_t batch, int32_t warps, int32_t order, int32_t vec32);
/* Position-dependent 64-bit hashes of n byte ranges [off[i], off[i+1]) of a device
 * buffer, reduced on the device,
 * mod 2^64 over the range read as little-endian 8-byte words w_k (zero padded at the
 * end): hash = sum_k mix64(k * 0x9E3779B97F4A7C15 + w_k) with mix64 = the splitmix64
 * finaliser.  `dev` must be 8-byte aligned.  Lets tests compare full-size images record by record with the oracle
 * without copying 26 GB to the host.  `off` and `out` are host arrays. */
int gm2_diag_range_hashes(gm2_ctx* ctx, const uint8_t* dev, int64_t dev_bytes,
                          const int64_t* off /* n+1 */, int64_t n, uint64_t* out /* n */);

/* Host only.  The host-side ceiling of gm2_emit_host's two-bit transport: `threads` threads (0 = what
 * the expansion would use) fill `host` (e.g. the pinned output buffer) with non-temporal stores, the
 * way the expansion writes the image; *gbs = bytes / seconds / 1e9 of the best of `reps` passes. */
int gm2_diag_host_fill(uint8_t* host, int64_t bytes, int32_t threads, int32_t reps, double* gbs);

/* Host only (no context, no GPU).  Gene-name lists container -> id CSR, replacing
 * `np.load(genes_path, allow_pickle=True).tolist()` (minimizer_2.py:456, :518) plus the per-name
 * `name in needed_genes` test of :62 for the file `np.save` wrote at binary_converter.py:71 / :117.
 * `body` is the pickle that follows the .npy header (dtype object); the array has S elements
 * (1-D: each a list/tuple of str, L = 0) or S x L elements (2-D: each a str, L > 0).
 * `names` + `name_off[V+1]` is the vocabulary (UTF-8, id = position; gm2_set_name_map's ids).
 * Out: off_out[S+1], ids_out[off_out[S]] = vocabulary ids of the names that are in the
 * vocabulary, in list order with duplicates; count_out[S] = len() of each list (all items).
 * Returns GM2_ERR_UNSUPPORTED for anything but lists of plain str (the caller then uses NumPy's
 * loader), GM2_ERR_INVALID for a corrupt stream or a shape mismatch, GM2_ERR_CAPACITY when
 * ids_cap is too small (nbytes / 2 always suffices).  Message: gm2_last_error(NULL). */
int gm2_tokenize_pickle(const uint8_t* body, int64_t nbytes, int64_t S, int64_t L,
                        const char* names, const int64_t* name_off, int32_t V,
                        int32_t* ids_out, int64_t ids_cap, int64_t* off_out, int64_t* count_out);

/* Host only (no context, no GPU).  GenBank flat file -> what gm2_set_reference and the name map are
 * fed with, replacing `SeqIO.read(genome_path, "genbank")` (minimizer_2.py:455, :515; Biopython 1.85,
 * a third-party dependency of the reference) plus the feature loop of minimizer_2.py:59-61, :78-79:
 *   G bytes of sequence  = record.seq (ORIGIN lines from column 11, blanks removed, upper-cased);
 *   F gene rows, file order, one per feature whose key is exactly "gene":
 *     gene_start / gene_end = int(location.start) / int(location.end): 0-based half-open span over
 *       all parts of join()/order(), complement() and fuzzy '<' '>' ignored, "N^M" -> [N, N);
 *     name = first /gene qualifier value with its quotes removed ("" when there is none),
 *       names[name_off[g] .. name_off[g+1]).
 * n_features counts every feature-table entry of any key.
 * gm2_genbank_parse returns GM2_ERR_UNSUPPORTED (reason: gm2_last_error(NULL)) for anything outside
 * the plain subset — zero or several LOCUS records, carriage returns, bytes outside printable
 * ASCII / tab / newline, remote, within-position ("N.M") or malformed locations — and the caller
 * then uses its general reader (genome_minimizer_2_b200/genbank.py), which owns the error messages.
 * The handle is immutable after parse; gm2_genbank_copy fills caller arrays sized from
 * gm2_genbank_sizes (seq G, gene_start F, gene_end F, name_off F+1, names name_bytes; any may be NULL). */
typedef struct gm2_genbank gm2_genbank;
int gm2_genbank_parse(const uint8_t* text, int64_t nbytes, gm2_genbank** out);
int gm2_genbank_sizes(const gm2_genbank* h, int64_t* G, int32_t* F, int64_t* name_bytes, int64_t* n_features);
int gm2_genbank_copy(const gm2_genbank* h, uint8_t* seq, int64_t* gene_start, int64_t* gene_end,
                     int64_t* name_off, uint8_t* names);
int gm2_genbank_free(gm2_genbank* h);

/* Host only.  Decoder of the two-bit wire format gm2_emit_host uses internally (GM2_CFG_WIRE), on
 * caller-supplied data: S records whose kept bases arrive as per-(sample, tile) 2-bit pieces; piece
 * (i, t) starts at 32-bit word (rec_off[i] >> 4) + i * (ntiles + 2) + (tile_off[i][t] >> 4) + t of
 * `packed` (rec_off[0] == 0), base j in bits [2j, 2j+2), codes 0..3 = A C G T.  Writes the records
 * ('>' + prefix + (first_idx + i + 1) + '\n' + bases + '\n') at out + rec_off[i].  `packed` must be
 * readable 16 bytes past its last used word.  threads 0 = default; simd: 0 = portable scalar
 * decoder, 1 = the best the CPU has, 2 = AVX2, 3 = AVX-512 VBMI (a level the CPU lacks falls back one down). */
int gm2_diag_expand(const uint32_t* packed, const int32_t* tile_off, const int64_t* rec_off,
                    const int64_t* lengths, int64_t S, int32_t ntiles, int64_t first_idx,
                    const char* prefix, uint8_t* out, int32_t threads, int32_t simd);

#ifdef __cplusplus
}
#endif
#endif /* GM2_H_ */
