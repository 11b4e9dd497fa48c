/* minimizer_c.c — plain-C restatement of the minimizer hot path.  TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py): used by tests/ and bench.py's cpu_baseline leg as the checker,
 * never by the product.
 *
 * Follows the reference's algorithm (ucl-cssb/genome-minimizer-2,
 * src/genome_minimizer_2/minimizer/minimizer_2.py):
 *   :50-66   a gene is removed iff its name is not in the sample's list  -> keep bit clear
 *   :68-83   positions_to_remove = union of range(start, end) over removed genes
 *   :85-101  output = bases at positions not in that union, ascending
 *   :476-477 record = ">" prefix (idx+1) "\n" bases "\n"
 * The union is held as a coverage-count difference array instead of a Python set; the
 * result is identical (a position is deleted iff its coverage count is > 0).
 *
 * Pinned against the reference's own outputs through tests/golden/ (tests/test_oracle.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* Scratch: diff must hold G+1 int32.  Returns L (kept bases written to out, capacity G). */
ORACLE_API int64_t oracle_minimize(const uint8_t* seq, int64_t G,
                                   const int64_t* start, const int64_t* end, int32_t F,
                                   const uint32_t* keep_row, int32_t* diff, uint8_t* out)
{
    memset(diff, 0, (size_t)(G + 1) * sizeof(int32_t));
    for (int32_t g = 0; g < F; ++g) {
        if ((keep_row[g >> 5] >> (g & 31)) & 1u) continue;          /* kept gene: deletes nothing */
        int64_t a = start[g], b = end[g];
        if (a < 0) a = 0;
        if (b < 0) b = 0;
        if (a > G) a = G;
        if (b > G) b = G;
        if (a >= b) continue;                                       /* range(a, b) empty */
        diff[a] += 1; diff[b] -= 1;
    }
    int64_t L = 0; int32_t cover = 0;
    for (int64_t p = 0; p < G; ++p) {
        cover += diff[p];
        if (cover == 0) out[L++] = seq[p];
    }
    return L;
}

static int ndigits(uint64_t v) { int n = 1; while (v >= 10) { v /= 10; ++n; } return n; }

/* Writes record idx (0-based global index) at out; returns bytes written. */
ORACLE_API int64_t oracle_record(const char* prefix, int64_t idx, const uint8_t* bases, int64_t L, uint8_t* out)
{
    int64_t o = 0;
    out[o++] = '>';
    size_t pl = strlen(prefix);
    memcpy(out + o, prefix, pl); o += (int64_t)pl;
    uint64_t num = (uint64_t)(idx + 1);
    int nd = ndigits(num);
    for (int i = nd - 1; i >= 0; --i) { out[o + i] = (uint8_t)('0' + (num % 10)); num /= 10; }
    o += nd;
    out[o++] = '\n';
    memcpy(out + o, bases, (size_t)L); o += L;
    out[o++] = '\n';
    return o;
}

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* Same definition as gm2_diag_range_hashes (include/gm2.h). */
ORACLE_API uint64_t oracle_range_hash(const uint8_t* data, int64_t n)
{
    uint64_t acc = 0; int64_t k = 0, nw = n >> 3;
    for (; k < nw; ++k) { uint64_t w; memcpy(&w, data + 8 * k, 8); acc += mix64((uint64_t)k * 0x9E3779B97F4A7C15ull + w); }
    if (n & 7) { uint64_t w = 0; memcpy(&w, data + 8 * k, (size_t)(n & 7)); acc += mix64((uint64_t)k * 0x9E3779B97F4A7C15ull + w); }
    return acc;
}

/* Batch: for samples s = 0..S-1 (keep rows of FW words) compute L_s and the hash of the
 * full FASTA record (first_idx + s); optionally append the records to `image` (may be NULL).
 * Returns total image bytes, or -1 on allocation failure. */
ORACLE_API int64_t oracle_batch(const uint8_t* seq, int64_t G, const int64_t* start, const int64_t* end,
                                int32_t F, const uint32_t* keep_rows, int64_t S, int64_t first_idx,
                                const char* prefix, int64_t* lengths, uint64_t* hashes, uint8_t* image)
{
    int32_t FW = (F + 31) / 32;
    int32_t* diff = (int32_t*)malloc((size_t)(G + 1) * sizeof(int32_t));
    uint8_t* bases = (uint8_t*)malloc((size_t)(G > 0 ? G : 1));
    uint8_t* rec = (uint8_t*)malloc((size_t)G + strlen(prefix) + 64);
    if (!diff || !bases || !rec) { free(diff); free(bases); free(rec); return -1; }
    int64_t total = 0;
    for (int64_t s = 0; s < S; ++s) {
        int64_t L = oracle_minimize(seq, G, start, end, F, keep_rows + (size_t)s * FW, diff, bases);
        int64_t n = oracle_record(prefix, first_idx + s, bases, L, rec);
        if (lengths) lengths[s] = L;
        if (hashes) hashes[s] = oracle_range_hash(rec, n);
        if (image) memcpy(image + total, rec, (size_t)n);
        total += n;
    }
    free(diff); free(bases); free(rec);
    return total;
}
