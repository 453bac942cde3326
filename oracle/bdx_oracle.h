/*
 * bdx_oracle.h -- CPU oracle for the BioDemuX barcode-assignment hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * Julia algorithm (src/classification.jl, BioDemuX.jl v1.6.0).  It is used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs as the checker and the timed CPU baseline.  Nothing in the
 * product path (biodemux.jl_b200/, include/) may include, link or call it.
 *
 * Parity status: PINNED.  The restatement reproduces the reference's own
 * golden outputs (test/results, 124 files / 2 880 reads) and every
 * known-answer assert in test/unit/{alignment,trimming,hamming,exact}.jl;
 * see tests/test_oracle_golden.py and tests/test_oracle_kat.py.  Julia is not
 * installed in this image, so the reference itself cannot be executed here.
 *
 * All positions are 1-based inclusive, as in the reference.
 */
#ifndef BDX_ORACLE_H
#define BDX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_ALGO_SEMIGLOBAL 0
#define ORC_ALGO_HAMMING 1
#define ORC_ALGO_EXACT 2

#define ORC_STATUS_MATCH 0
#define ORC_STATUS_UNKNOWN 1
#define ORC_STATUS_AMBIGUOUS 2

/* classification.jl:9-14 */
typedef struct orc_range {
    int64_t start_offset;
    int32_t start_from_end;
    int64_t end_offset;
    int32_t end_from_end;
} orc_range;

/* one barcode set + its pass parameters (classification.jl:778-792) */
typedef struct orc_set {
    int32_t n_bc;
    const uint8_t *bc_bytes;   /* concatenated, already preprocessed */
    const int32_t *bc_off;     /* n_bc + 1 */
    const int64_t *bc_len_no_n; /* n_bc */
    orc_range ref_search_range;
    orc_range barcode_start_range;
    orc_range barcode_end_range;
    int32_t trim_side;         /* 0 = nothing, 3, 5 */
} orc_set;

/* hot-path fields of DemuxConfig (classification.jl:16-58) */
typedef struct orc_config {
    double max_error_rate;
    double min_delta;
    int64_t match, mismatch, indel, nindel;
    int32_t has_nindel;
    int32_t algorithm;
    int32_t is_dual;
    int32_t want_stats; /* stats != nothing: forces traceback (classification.jl:812) */
    orc_set set1, set2;
} orc_config;

typedef struct orc_pass {
    int32_t status; /* ORC_STATUS_*; -1 = pass not run */
    int32_t bc;     /* 1-based, 0 none */
    int64_t start, end;
    double score;
} orc_pass;

typedef struct orc_result {
    int32_t status;
    int32_t bc1, bc2;
    int64_t keep_start, keep_end; /* -1,-1 for unknown/ambiguous; 1,0 empty keep */
    orc_pass pass[2];
} orc_result;

/* classification.jl:61-94; returns 0 ok, -1 on malformed input */
int orc_parse_dynamic_range(const char *s, orc_range *out);
/* classification.jl:96-100 (UnitRange normalisation: last = first-1 when empty) */
void orc_resolve(const orc_range *dr, int64_t len, int64_t *first, int64_t *last);

/* classification.jl:238-445 via :447-477.  has_n selects NScoring. Returns score (INFINITY if none). */
double orc_semiglobal(const uint8_t *q, int64_t m, const uint8_t *r, int64_t n,
                      double max_error, int64_t match, int64_t mismatch, int64_t indel,
                      int32_t has_n, int64_t nindel,
                      int64_t range_first, int64_t range_last,
                      int64_t max_start_pos, int64_t min_end_pos,
                      int64_t norm_len, int32_t traceback, int32_t trim_side,
                      int64_t *start, int64_t *end);
/* classification.jl:557-625 */
double orc_hamming(const uint8_t *q, int64_t m, const uint8_t *r, int64_t n,
                   double max_error_rate, int64_t range_first, int64_t range_last,
                   int64_t max_start_pos, int64_t min_end_pos, int32_t trim_side,
                   int64_t *start, int64_t *end);
/* classification.jl:485-548 */
double orc_exact(const uint8_t *q, int64_t m, const uint8_t *r, int64_t n,
                 int64_t range_first, int64_t range_last,
                 int64_t max_start_pos, int64_t min_end_pos, int32_t trim_side,
                 int64_t *start, int64_t *end);

/* classification.jl:632-728.  Returns barcode index (1-based, 0 none). */
int32_t orc_find_best(const orc_config *cfg, const orc_set *set,
                      const uint8_t *r, int64_t n,
                      int64_t range_first, int64_t range_last,
                      int64_t max_start_pos, int64_t min_end_pos,
                      int32_t need_traceback,
                      double *min_score, double *delta, int64_t *start, int64_t *end);

/* classification.jl:776-868 (without the Dict updates; the caller rebuilds
 * them from orc_pass) */
void orc_match_pass(const orc_config *cfg, int is_pass2, const uint8_t *r, int64_t n, orc_pass *out);

/* classification.jl:871-1005 */
void orc_determine(const orc_config *cfg, const uint8_t *r, int64_t n, orc_result *out);

/* core.jl:243-267 loop over a chunk */
void orc_classify(const orc_config *cfg, const uint8_t *seqs, const int64_t *offsets,
                  int64_t n_reads, orc_result *out);

/* Base.round(x, digits=2) for non-negative finite x, RoundNearest (ties to even
 * on the scaled value), as used at classification.jl:835,853 */
double orc_round2(double x);

#ifdef __cplusplus
}
#endif
#endif
