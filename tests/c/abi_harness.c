/* abi_harness.c -- the C ABI of libbdx driven from plain C ("the boundary is exercised
 * by a C test harness and a Python ctypes harness").  Builds with `gcc -std=c99 -I include` against libbdx.so.
 *
 *   abi_harness            config validation, error codes, FASTQ scanner, barcode loader: no GPU needed;
 *                          then, if a CUDA device is visible, one small classification whose results are
 *                          printed for the calling test to compare with the oracle
 *   exit code 0 = all checks passed; every failed check prints its line.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bdx.h"

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                                    \
        }                                                                  \
    } while (0)

static bdx_range full_range(void)
{
    bdx_range r;
    r.start_offset = 1; r.start_from_end = 0; r.end_offset = 0; r.end_from_end = 1;   /* "1:end" */
    return r;
}

int main(void)
{
    /* ---- configuration ---- */
    static const char bc_bytes[] = "ACGTACGTAAGGCCTTAACCTTGGAACCGGTTAGGATTCCAATTGGCCGCGCATATGCATGCATTAGCTAGCTAGGATCCGGATCCTT";
    int32_t offsets[12], lens[11];
    const int n_bc = 11, m = 8;
    for (int b = 0; b <= n_bc; b++) offsets[b] = b * m;
    for (int b = 0; b < n_bc; b++) lens[b] = m;

    bdx_params p;
    memset(&p, 0, sizeof(p));
    p.struct_size = sizeof(p);
    p.abi_version = BDX_ABI_VERSION;
    p.max_error_rate = 0.2;
    p.mismatch = 1;
    p.indel = 1;
    p.algorithm = BDX_SEMIGLOBAL;
    p.set1.n_barcodes = n_bc;
    p.set1.bytes = (const uint8_t *)bc_bytes;
    p.set1.offsets = offsets;
    p.set1.lengths_no_n = lens;
    p.set1.trim_side = 5;
    p.set1.ref_search_range = p.set1.barcode_start_range = p.set1.barcode_end_range = full_range();

    bdx_config *cfg = NULL;
    CHECK(bdx_abi_version() == BDX_ABI_VERSION);
    CHECK(bdx_config_create(&p, &cfg) == BDX_OK && cfg != NULL);

    bdx_params bad = p;
    bad.set1.trim_side = 4;                                   /* core.jl:308-313 */
    bdx_config *none = NULL;
    CHECK(bdx_config_create(&bad, &none) == BDX_ERR_INVALID && none == NULL);
    CHECK(strstr(bdx_last_error(), "trim_side must be 3 or 5") != NULL);
    bad = p;
    bad.indel = 0;                                            /* DivideError in the reference */
    CHECK(bdx_config_create(&bad, &none) == BDX_ERR_INVALID);
    bad = p;
    bad.struct_size = 7;
    CHECK(bdx_config_create(&bad, &none) == BDX_ERR_INVALID);
    CHECK(bdx_config_create(NULL, &none) == BDX_ERR_INVALID);

    bdx_stats_layout lay;
    CHECK(bdx_stats_layout_get(cfg, &lay) == BDX_OK && lay.b1 == n_bc && lay.b2 == 0 && lay.total_len > 4);

    /* ---- host-side FASTQ scanner / packer ---- */
    static const char fq[] = "@r1\nGGACGTACGTCC\n+\nIIIIIIIIIIII\n@r2\r\nTTAAGGCCTTAA\r\n+\r\nJJJJJJJJJJJJ\r\n@r3\nACG";
    bdx_fastq_record recs[4];
    int32_t n_rec = 0;
    int64_t consumed = 0;
    CHECK(bdx_fastq_scan((const uint8_t *)fq, (int64_t)strlen(fq), 0, 4, recs, &n_rec, &consumed) == BDX_OK);
    CHECK(n_rec == 2 && consumed == (int64_t)(strstr(fq, "@r3") - fq));
    CHECK(recs[1].seq_len == 12 && memcmp(fq + recs[1].seq_off, "TTAAGGCCTTAA", 12) == 0);
    CHECK(bdx_fastq_scan((const uint8_t *)fq, (int64_t)strlen(fq), 1, 4, recs, &n_rec, &consumed) == BDX_OK);
    CHECK(n_rec == 3 && consumed == (int64_t)strlen(fq) && recs[2].seq_len == 3 && recs[2].qual_len == 0);
    uint8_t packed[64];
    int32_t poff[4];
    CHECK(bdx_fastq_pack((const uint8_t *)fq, recs, 3, packed, sizeof(packed), poff) == BDX_OK);
    CHECK(poff[0] == 0 && poff[1] == 12 && poff[2] == 24 && poff[3] == 27 && memcmp(packed + 24, "ACG", 3) == 0);

    /* ---- barcode loader errors ---- */
    bdx_barcode_table *tab = NULL;
    CHECK(bdx_barcode_table_load("/nonexistent/barcodes.tsv", 0, 0, &tab) == BDX_ERR_INVALID && tab == NULL);
    CHECK(strstr(bdx_barcode_table_error(), "cannot open") != NULL);

    /* ---- streams: no CPU fallback ---- */
    bdx_stream *st = NULL;
    const int n_dev = bdx_device_count();
    printf("devices %d\n", n_dev);
    if (n_dev == 0) {
        CHECK(bdx_stream_create(cfg, 0, 16, 4096, &st) == BDX_ERR_CUDA && st == NULL);
        CHECK(strstr(bdx_last_error(), "no CPU fallback") != NULL);
    } else {
        CHECK(bdx_stream_create(cfg, 0, 16, 4096, &st) == BDX_OK && st != NULL);
        bdx_stream *st2 = NULL;
        CHECK(bdx_stream_create(cfg, n_dev + 3, 16, 4096, &st2) == BDX_ERR_INVALID && st2 == NULL);
        bdx_result res[3];
        uint64_t tag = 0;
        int32_t n_out = 0;
        CHECK(bdx_fetch(st, &tag, &n_out, res, NULL) == BDX_ERR_STATE);                 /* nothing in flight */
        CHECK(bdx_submit(st, packed, poff, 3, 77) == BDX_OK);
        CHECK(bdx_fetch(st, &tag, &n_out, res, NULL) == BDX_OK && tag == 77 && n_out == 3);
        for (int i = 0; i < 3; i++)
            printf("result %d %d %d %d %d\n", res[i].status, res[i].bc1, res[i].bc2, res[i].keep_start, res[i].keep_end);
        /* read 1 = GG + barcode 1 + CC with trim_side 5: keep what follows the barcode */
        CHECK(res[0].status == BDX_MATCH && res[0].bc1 == 1 && res[0].keep_start == 11 && res[0].keep_end == 12);
        CHECK(res[2].status == BDX_UNKNOWN && res[2].keep_start == -1);
        int32_t big_off[2] = {0, 5000};
        CHECK(bdx_submit(st, packed, big_off, 1, 0) == BDX_ERR_TOO_LARGE);
        bdx_stream_destroy(st);
    }
    bdx_config_destroy(cfg);
    printf(failures ? "FAILED %d\n" : "OK\n", failures);
    return failures ? 1 : 0;
}
