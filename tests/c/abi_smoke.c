/* A plain C99 caller of libgm2.so: proves that include/gm2.h is a C header (not C++ in disguise), that
 * every entry point used here links, and that the host-only entry points work without a GPU.
 * Built and run by tests/test_abi_c.py with `gcc -std=c99 -Wall -Wextra -Werror -pedantic`.
 * Exit status 0 = every check passed; otherwise the number of the failed check. */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "gm2.h"

static const char GB[] =
    "LOCUS       C99 20 bp DNA linear\n"
    "FEATURES             Location/Qualifiers\n"
    "     gene            3..6\n"
    "                     /gene=\"aaa\"\n"
    "     CDS             3..6\n"
    "                     /gene=\"aaa\"\n"
    "     gene            complement(join(5..7,\n"
    "                     9..10))\n"
    "                     /locus_tag=\"no name\"\n"
    "ORIGIN\n"
    "        1 acgttgcaag cttaggccat\n"
    "//\n";

int main(void) {
    gm2_genbank* h = NULL;
    int64_t G = -1, name_bytes = -1, n_features = -1;
    int32_t F = -1;
    uint8_t seq[32];
    int64_t start[4], end[4], name_off[5];
    uint8_t names[16];
    gm2_ctx* ctx = NULL;
    int rc;

    if (gm2_abi_version() != GM2_ABI_VERSION) return 1;

    /* host-only: GenBank bytes -> sequence + gene table */
    if (gm2_genbank_parse((const uint8_t*)GB, (int64_t)strlen(GB), &h) != GM2_OK || !h) return 3;
    if (gm2_genbank_sizes(h, &G, &F, &name_bytes, &n_features) != GM2_OK) return 4;
    if (G != 20 || F != 2 || name_bytes != 3 || n_features != 3) return 5;
    if (gm2_genbank_copy(h, seq, start, end, name_off, names) != GM2_OK) return 6;
    if (memcmp(seq, "ACGTTGCAAGCTTAGGCCAT", 20) != 0) return 7;
    if (start[0] != 2 || end[0] != 6 || start[1] != 4 || end[1] != 10) return 8;
    if (name_off[0] != 0 || name_off[1] != 3 || name_off[2] != 3 || memcmp(names, "aaa", 3) != 0) return 9;
    if (gm2_genbank_free(h) != GM2_OK) return 10;

    /* anything outside the scanner's subset is declined, with a reason */
    h = NULL;
    rc = gm2_genbank_parse((const uint8_t*)"no record here\n", 15, &h);
    if (rc != GM2_ERR_UNSUPPORTED || h != NULL || !gm2_last_error(NULL) || !*gm2_last_error(NULL)) return 11;

    /* a context needs a CUDA device: with none, creation fails loudly (there is no CPU fallback) */
    rc = gm2_create(0, &ctx);
    if (gm2_device_count() <= 0) {                  /* 0, or a negative error when there is no driver at all */
        if (rc == GM2_OK || ctx != NULL) return 12;
        if (!gm2_last_error(NULL) || !*gm2_last_error(NULL)) return 13;
        printf("no CUDA device: gm2_create -> %d (%s)\n", rc, gm2_last_error(NULL));
    } else {
        if (rc != GM2_OK || !ctx) return 14;
        if (gm2_set_reference(ctx, seq, G, start, end, F) != GM2_OK) return 15;
        if (gm2_destroy(ctx) != GM2_OK) return 16;
        printf("CUDA device present: context created, reference uploaded, destroyed\n");
    }
    return 0;
}
