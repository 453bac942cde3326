/*
 * bdx_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see bdx_oracle.h).
 *
 * Restates /root/reference/src/classification.jl (BioDemuX.jl v1.6.0) function
 * by function.  Control flow is kept literal on purpose -- pruning, stale DP
 * cells, tie-break order and early exits all change results --
 * so do not "simplify" it.  Integer widths follow the reference: Int64 cells,
 * IEEE double thresholds and scores.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: the reference's
 * floor(max_error * norm) must not be contracted into an FMA).
 */
#include "bdx_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

/* classification.jl:7  INF_INT = typemax(Int) / 4 */
#define INF_INT (INT64_MAX / 4)

static inline int64_t i64max(int64_t a, int64_t b) { return a > b ? a : b; }
static inline int64_t i64min(int64_t a, int64_t b) { return a < b ? a : b; }

/* ------------------------------------------------------------------ */
/* parse_part / parse_dynamic_range  (classification.jl:61-94)          */
/* ------------------------------------------------------------------ */

/* parse(Int, strip(s)): optional sign, decimal digits, surrounding blanks */
static int parse_int(const char *b, const char *e, int64_t *out)
{
    while (b < e && isspace((unsigned char)*b)) b++;
    while (e > b && isspace((unsigned char)e[-1])) e--;
    if (b >= e) return -1;
    int neg = 0;
    if (*b == '+' || *b == '-') { neg = (*b == '-'); b++; }
    if (b >= e) return -1;
    int64_t v = 0;
    for (; b < e; b++) {
        if (!isdigit((unsigned char)*b)) return -1;
        v = v * 10 + (*b - '0');
    }
    *out = neg ? -v : v;
    return 0;
}

/* classification.jl:61-81 */
static int parse_part(const char *b, const char *e, int64_t *val, int32_t *from_end)
{
    char buf[128];
    size_t len;
    while (b < e && isspace((unsigned char)*b)) b++;
    while (e > b && isspace((unsigned char)e[-1])) e--;
    len = (size_t)(e - b);
    if (len >= sizeof(buf) - 1) return -1;
    /* replace(s, "end" => "0") on every occurrence (:63-66) */
    size_t o = 0;
    *from_end = 0;
    for (size_t i = 0; i < len;) {
        if (i + 3 <= len && memcmp(b + i, "end", 3) == 0) {
            buf[o++] = '0';
            i += 3;
            *from_end = 1;
        } else {
            buf[o++] = b[i++];
        }
    }
    buf[o] = 0;
    const char *s = buf, *se = buf + o;
    const char *minus = memchr(s, '-', o);
    const char *plus = memchr(s, '+', o);
    int64_t a, c;
    if (minus) {
        /* split(s,'-'): p[1] - p[2]; further pieces are ignored (:70-72) */
        const char *second_end = memchr(minus + 1, '-', (size_t)(se - (minus + 1)));
        if (!second_end) second_end = se;
        if (parse_int(s, minus, &a) || parse_int(minus + 1, second_end, &c)) return -1;
        *val = a - c;
    } else if (plus) {
        const char *second_end = memchr(plus + 1, '+', (size_t)(se - (plus + 1)));
        if (!second_end) second_end = se;
        if (parse_int(s, plus, &a) || parse_int(plus + 1, second_end, &c)) return -1;
        *val = a + c;
    } else {
        if (parse_int(s, se, &a)) return -1;
        *val = a;
    }
    return 0;
}

int orc_parse_dynamic_range(const char *s, orc_range *out)
{
    /* split(range_str, ':') must give exactly two parts (:84-87) */
    const char *colon = strchr(s, ':');
    if (!colon) return -1;
    if (strchr(colon + 1, ':')) return -1;
    if (parse_part(s, colon, &out->start_offset, &out->start_from_end)) return -1;
    if (parse_part(colon + 1, s + strlen(s), &out->end_offset, &out->end_from_end)) return -1;
    return 0;
}

/* classification.jl:96-100.  Julia's UnitRange(a,b) stores last = a-1 when b < a,
 * and callers read last(range) of possibly-empty ranges (:801). */
void orc_resolve(const orc_range *dr, int64_t len, int64_t *first, int64_t *last)
{
    int64_t s = dr->start_from_end ? len + dr->start_offset : dr->start_offset;
    int64_t e = dr->end_from_end ? len + dr->end_offset : dr->end_offset;
    int64_t f = i64max(1, s);
    int64_t l = i64min(len, e);
    if (l < f) l = f - 1;
    *first = f;
    *last = l;
}

double orc_round2(double x)
{
    /* Base._round_invstep: round(x * 100) / 100, RoundNearest */
    double y = rint(x * 100.0) / 100.0;
    return isfinite(y) ? y : x;
}

/* ------------------------------------------------------------------ */
/* semiglobal_alignment_core  (classification.jl:238-445)              */
/* ------------------------------------------------------------------ */

typedef struct {
    int64_t match, mismatch, indel, nindel;
    int has_n;
} scoring_t;

/* step_scores (:208-236): used for the first computed row and for row m */
static inline void step_edge(const scoring_t *sc, const uint8_t *q, const uint8_t *r,
                             int64_t i, int64_t j, int64_t prev, const int64_t *DP, int64_t m,
                             int64_t *ins, int64_t *del, int64_t *sub)
{
    if (!sc->has_n) {
        *ins = (i == m) ? INF_INT : DP[i] + sc->indel;                 /* :213 */
        *del = prev + sc->indel;                                       /* :214 */
        *sub = (i == 1 ? 0 : DP[i - 1]) + (q[i] == r[j] ? sc->match : sc->mismatch); /* :215 */
    } else {
        int is_n = q[i] == (uint8_t)'N';                               /* :226 */
        int64_t cost = is_n ? sc->nindel : sc->indel;
        *ins = DP[i] + (i == m ? INF_INT : cost);                      /* :229 */
        *del = prev + cost;
        int is_match = (q[i] == r[j]) || is_n;
        *sub = (i == 1 ? 0 : DP[i - 1]) + (is_match ? sc->match : sc->mismatch);
    }
}

/* step_scores_main (:178-206): interior rows */
static inline void step_main(const scoring_t *sc, const uint8_t *q, const uint8_t *r,
                             int64_t i, int64_t j, int64_t prev, const int64_t *DP,
                             int64_t *ins, int64_t *del, int64_t *sub)
{
    if (!sc->has_n) {
        *ins = DP[i] + sc->indel;
        *del = prev + sc->indel;
        *sub = DP[i - 1] + (q[i] == r[j] ? sc->match : sc->mismatch);
    } else {
        int is_n = q[i] == (uint8_t)'N';
        int64_t cost = is_n ? sc->nindel : sc->indel;
        *ins = DP[i] + cost;
        *del = prev + cost;
        int is_match = (q[i] == r[j]) || is_n;
        *sub = DP[i - 1] + (is_match ? sc->match : sc->mismatch);
    }
}

typedef struct {
    int64_t score, start, end;
} tb_result;

/* q and r are passed 1-based (pointer to element before the first byte). */
static double semiglobal_core(int64_t *DP, int64_t *origin,
                              const uint8_t *q1, const uint8_t *r1, int64_t m, int64_t n,
                              double max_error, const scoring_t *sc,
                              int is_traceback, int trim_side,
                              int64_t range_first, int64_t range_last,
                              int64_t max_start_pos, int64_t min_end_pos,
                              int64_t norm_len, int64_t *out_start, int64_t *out_end)
{
    tb_result res = {INF_INT, -1, -1};      /* init_result (:130-136) */
    *out_start = -1;
    *out_end = -1;
    if (m == 0 || n == 0) return INFINITY;  /* :250-252 */

    int64_t allowed_error = (int64_t)floor(max_error * (double)norm_len);   /* :254 */
    int64_t min_cost = sc->has_n ? i64min(sc->indel, sc->nindel) : sc->indel;
    if (min_cost == 0) return INFINITY; /* Julia: DivideError; the product rejects zero gap costs */
    int64_t max_indel_steps = allowed_error / min_cost;                     /* :170-176 (div truncates) */
    int64_t min_valid_start = min_end_pos - (m + max_indel_steps) + 1;      /* :259 */
    if (min_valid_start > max_start_pos) return INFINITY;                   /* :261-263 */
    if (min_valid_start > range_first) range_first = i64max(range_first, min_valid_start); /* :266-268 */

    int64_t band_offset = i64max(m - n - max_indel_steps, -max_start_pos - max_indel_steps); /* :270 */

    for (int64_t i = 1; i <= m; i++) {      /* :278-283 */
        DP[i] = sc->indel * i;
        if (is_traceback) origin[i] = 1 - i;
    }

    int64_t lact = i64min(allowed_error + 1, m);   /* :286 */
    for (int64_t j = range_first; j <= range_last; j++) {
        int64_t prev_origin = j, prev, fact;
        int64_t cur_origin = 0;
        if (j + band_offset >= 1) {         /* :289-295 */
            fact = j + band_offset;
            prev = allowed_error;
        } else {
            fact = 1;
            prev = 0;
        }
        if (fact > lact) goto finalize;     /* :297-299 */

        int64_t ins, del, sub;
        /* 1. first computed row (:303-335) */
        step_edge(sc, q1, r1, fact, j, prev, DP, m, &ins, &del, &sub);
        if (is_traceback) {
            int64_t best = del, bo = prev_origin;
            int64_t sub_o = (fact == 1) ? j : origin[fact - 1];
            if (sub < best) { best = sub; bo = sub_o; }
            if (ins < best) { best = ins; bo = origin[fact]; }
            cur_origin = bo;
        }
        if (fact != 1) {
            DP[fact - 1] = prev;
            if (is_traceback) origin[fact - 1] = prev_origin;
        }
        prev = i64min(ins, i64min(del, sub));
        if (is_traceback) prev_origin = cur_origin;

        /* 2. interior rows (:338-373) */
        int64_t limit = (lact == m) ? m - 1 : lact;
        for (int64_t i = fact + 1; i <= limit; i++) {
            step_main(sc, q1, r1, i, j, prev, DP, &ins, &del, &sub);
            if (is_traceback) {
                int64_t best = del, bo = prev_origin;
                if (sub < best) { best = sub; bo = origin[i - 1]; }
                if (ins < best) { best = ins; bo = origin[i]; }
                cur_origin = bo;
            }
            DP[i - 1] = prev;
            if (is_traceback) origin[i - 1] = prev_origin;
            prev = i64min(ins, i64min(del, sub));
            if (is_traceback) prev_origin = cur_origin;
        }

        /* 3. last row, no insertion (:376-409) */
        if (lact == m && lact > fact) {
            step_edge(sc, q1, r1, m, j, prev, DP, m, &ins, &del, &sub);
            if (is_traceback) {
                int64_t best = del, bo = prev_origin;
                if (sub < best) { best = sub; bo = origin[m - 1]; }
                if (ins < best) { best = ins; bo = origin[m]; }
                cur_origin = bo;
            }
            DP[m - 1] = prev;
            if (is_traceback) origin[m - 1] = prev_origin;
            prev = i64min(ins, i64min(del, sub));
            if (is_traceback) prev_origin = cur_origin;
        }

        DP[lact] = prev;                    /* :412-415 */
        if (is_traceback) origin[lact] = prev_origin;

        if (lact == m && prev <= allowed_error) {   /* :417-438 */
            lact -= 1;
            if (j >= min_end_pos) {
                if (prev == 0) {
                    int early = !is_traceback || trim_side == 5;   /* :421 */
                    if (early) {
                        if (is_traceback) { *out_start = prev_origin; *out_end = j; }
                        return 0.0 / (double)norm_len;             /* :425-427 */
                    }
                }
                if (is_traceback) {         /* update_result (:142-153) */
                    if (prev < res.score) {
                        res.score = prev; res.start = prev_origin; res.end = j;
                    } else if (prev == res.score) {
                        if (trim_side == 3 && prev_origin > res.start) {
                            res.start = prev_origin; res.end = j;
                        }
                    }
                } else {
                    res.score = i64min(res.score, prev);           /* :138-140 */
                }
            }
        }
        while (lact > 0 && DP[lact] > allowed_error) lact -= 1;    /* :439-442 */
        lact += 1;
    }
finalize:
    if (is_traceback) { *out_start = res.start; *out_end = res.end; }
    if (res.score >= INF_INT) return INFINITY;     /* :155-168 */
    return (double)res.score / (double)norm_len;
}

double orc_semiglobal(const uint8_t *q, int64_t m, const uint8_t *r, int64_t n,
                      double max_error, int64_t match, int64_t mismatch, int64_t indel,
                      int32_t has_n, int64_t nindel,
                      int64_t range_first, int64_t range_last,
                      int64_t max_start_pos, int64_t min_end_pos,
                      int64_t norm_len, int32_t traceback, int32_t trim_side,
                      int64_t *start, int64_t *end)
{
    scoring_t sc = {match, mismatch, indel, nindel, has_n};
    /* the reference sizes the workspace to the longest barcode (core.jl:229-233);
     * DP[lact] with lact <= m is the largest index touched */
    int64_t *DP = (int64_t *)malloc(sizeof(int64_t) * (size_t)(2 * (m + 2)));
    int64_t *origin = DP + (m + 2);
    int64_t s = -1, e = -1;
    /* traceback is on iff trim_side != nothing || need_traceback (:454-458) */
    int tb = traceback || trim_side != 0;
    double score = semiglobal_core(DP, origin, q - 1, r - 1, m, n, max_error, &sc, tb, trim_side,
                                   range_first, range_last, max_start_pos, min_end_pos, norm_len, &s, &e);
    free(DP);
    if (start) *start = s;
    if (end) *end = e;
    return score;
}

/* ------------------------------------------------------------------ */
/* hamming_align  (classification.jl:557-625)                          */
/* ------------------------------------------------------------------ */
double orc_hamming(const uint8_t *q, int64_t m, const uint8_t *r, int64_t n,
                   double max_error_rate, int64_t range_first, int64_t range_last,
                   int64_t max_start_pos, int64_t min_end_pos, int32_t trim_side,
                   int64_t *start, int64_t *end)
{
    double best_score = INFINITY;
    int64_t best_start = -1, best_end = -1;
    *start = -1;
    *end = -1;
    if (m == 0) return INFINITY; /* 0/0 = NaN never updates best (:607-613) */
    int64_t allowed = (int64_t)floor(max_error_rate * (double)m);            /* :567 */
    int64_t first = i64max(range_first, 1);                                   /* :570 */
    int64_t last = i64min(range_last, i64min(max_start_pos, n - m + 1));      /* :571 */
    if (last < first) return INFINITY;
    for (int64_t j = first; j <= last; j++) {
        int64_t end_pos = j + m - 1;
        if (end_pos < min_end_pos) continue;                                  /* :583-586 */
        int64_t mm = 0;
        int failed = 0;
        for (int64_t k = 0; k < m; k++) {
            uint8_t qc = q[k], rc = r[j - 1 + k];
            if (qc != rc && qc != 0x4E) {                                     /* :597 */
                mm++;
                if (mm > allowed) { failed = 1; break; }
            }
        }
        if (!failed) {
            double score = (double)mm / (double)m;                            /* :607 */
            if (score < best_score) {
                best_score = score; best_start = j; best_end = end_pos;
            } else if (score == best_score) {
                if (trim_side == 3 && j > best_start) { best_start = j; best_end = end_pos; }
            }
        }
    }
    *start = best_start;
    *end = best_end;
    return best_score;
}

/* ------------------------------------------------------------------ */
/* exact_align  (classification.jl:485-548)                            */
/* Base.findnext(q, r, i): first occurrence starting at >= i.           */
/* Base.findprev(q, r, k): last occurrence lying inside r[1:k].         */
/* ------------------------------------------------------------------ */
static int64_t find_next(const uint8_t *q, int64_t m, const uint8_t *r, int64_t n, int64_t from)
{
    for (int64_t s = i64max(from, 1); s + m - 1 <= n; s++)
        if (memcmp(r + s - 1, q, (size_t)m) == 0) return s;
    return 0;
}
static int64_t find_prev(const uint8_t *q, int64_t m, const uint8_t *r, int64_t n, int64_t k)
{
    for (int64_t s = i64min(k, n) - m + 1; s >= 1; s--)
        if (memcmp(r + s - 1, q, (size_t)m) == 0) return s;
    return 0;
}

double orc_exact(const uint8_t *q, int64_t m, const uint8_t *r, int64_t n,
                 int64_t range_first, int64_t range_last,
                 int64_t max_start_pos, int64_t min_end_pos, int32_t trim_side,
                 int64_t *start, int64_t *end)
{
    *start = -1;
    *end = -1;
    /* Empty barcodes: findnext("", ...) semantics are not restated; the product
     * rejects empty barcodes at config creation. */
    if (m == 0) return INFINITY;
    int64_t first = i64max(range_first, 1);                                   /* :490 */
    int64_t last = i64min(range_last, i64min(max_start_pos, n - m + 1));      /* :491 */
    if (last < first) return INFINITY;
    if (trim_side == 3) {                                                     /* :499-515 */
        int64_t s = find_prev(q, m, r, n, last + m - 1);
        if (s && s >= first && s + m - 1 >= min_end_pos) {
            *start = s; *end = s + m - 1;
            return 0.0;
        }
        return INFINITY;
    }
    int64_t s = find_next(q, m, r, n, first);                                 /* :519-546 */
    if (s && s <= last) {
        if (s + m - 1 >= min_end_pos) { *start = s; *end = s + m - 1; return 0.0; }
        int64_t next = s + 1;
        while (next <= last) {
            s = find_next(q, m, r, n, next);
            if (!s) break;
            if (s > last) break;
            if (s + m - 1 >= min_end_pos) { *start = s; *end = s + m - 1; return 0.0; }
            next = s + 1;
        }
    }
    return INFINITY;
}

/* ------------------------------------------------------------------ */
/* find_best_matching_bc  (classification.jl:632-728)                  */
/* ------------------------------------------------------------------ */
int32_t orc_find_best(const orc_config *cfg, const orc_set *set,
                      const uint8_t *r, int64_t n,
                      int64_t range_first, int64_t range_last,
                      int64_t max_start_pos, int64_t min_end_pos,
                      int32_t need_traceback,
                      double *out_min, double *out_delta, int64_t *out_start, int64_t *out_end)
{
    double thr = cfg->max_error_rate;
    double min_score = INFINITY, sub_min = INFINITY;
    int32_t min_bc = 0;
    int64_t best_start = -1, best_end = -1;
    int with_delta = cfg->min_delta != 0.0;                /* :723 */
    int trim_side = set->trim_side;

    int64_t max_m = 0;
    for (int32_t b = 0; b < set->n_bc; b++)
        max_m = i64max(max_m, set->bc_off[b + 1] - set->bc_off[b]);
    int64_t *DP = (int64_t *)malloc(sizeof(int64_t) * (size_t)(2 * (max_m + 2)));
    int64_t *origin = DP + (max_m + 2);
    scoring_t sc = {cfg->match, cfg->mismatch, cfg->indel, cfg->nindel, cfg->has_nindel};

    for (int32_t b = 0; b < set->n_bc; b++) {
        const uint8_t *q = set->bc_bytes + set->bc_off[b];
        int64_t m = set->bc_off[b + 1] - set->bc_off[b];
        double score;
        int64_t s = -1, e = -1;
        if (cfg->algorithm == ORC_ALGO_HAMMING) {
            score = orc_hamming(q, m, r, n, thr, range_first, range_last, max_start_pos, min_end_pos, trim_side, &s, &e);
        } else if (cfg->algorithm == ORC_ALGO_EXACT) {
            score = orc_exact(q, m, r, n, range_first, range_last, max_start_pos, min_end_pos, trim_side, &s, &e);
        } else {
            int tb = (trim_side != 0) || need_traceback;   /* :454-458 */
            int64_t norm = cfg->has_nindel ? set->bc_len_no_n[b] : m;   /* :460, :476 */
            score = semiglobal_core(DP, origin, q - 1, r - 1, m, n, thr, &sc, tb, trim_side,
                                    range_first, range_last, max_start_pos, min_end_pos, norm, &s, &e);
            if (!tb) { s = -1; e = -1; }                   /* :651-656, :688-693 */
        }
        if (!with_delta) {
            if (score <= thr && score < min_score) {       /* :658-664 */
                min_score = score;
                min_bc = b + 1;
                thr = fmin(thr, min_score);
                best_start = s;
                best_end = e;
            }
        } else {
            if (score <= thr) {                            /* :696-709 */
                if (score < min_score) {
                    sub_min = min_score;
                    min_score = score;
                    min_bc = b + 1;
                    thr = fmin(thr, sub_min);
                    best_start = s;
                    best_end = e;
                } else if (score < sub_min) {
                    sub_min = score;
                    thr = fmin(thr, sub_min);
                }
            }
        }
    }
    free(DP);
    *out_min = min_score;
    *out_delta = with_delta ? (sub_min - min_score) : INFINITY;   /* :666, :711 */
    *out_start = best_start;
    *out_end = best_end;
    return min_bc;
}

/* ------------------------------------------------------------------ */
/* match_barcode_pass  (classification.jl:776-868)                     */
/* ------------------------------------------------------------------ */
void orc_match_pass(const orc_config *cfg, int is_pass2, const uint8_t *r, int64_t n, orc_pass *out)
{
    const orc_set *set = is_pass2 ? &cfg->set2 : &cfg->set1;
    int64_t rs_f, rs_l, bs_f, bs_l, be_f, be_l;
    orc_resolve(&set->ref_search_range, n, &rs_f, &rs_l);      /* :795-797 */
    orc_resolve(&set->barcode_start_range, n, &bs_f, &bs_l);
    orc_resolve(&set->barcode_end_range, n, &be_f, &be_l);

    int64_t start_j = i64max(rs_f, i64max(bs_f, 1));           /* :799 */
    int64_t end_j = i64min(rs_l, i64min(be_l, n));             /* :800 */
    int64_t max_start_pos = bs_l;                              /* :801 */
    int64_t min_end_pos = be_f;                                /* :802 */

    out->status = ORC_STATUS_UNKNOWN;
    out->bc = 0;
    out->start = -1;
    out->end = -1;
    out->score = INFINITY;
    if (start_j > end_j || start_j > max_start_pos || end_j < min_end_pos) return;   /* :805-807 */

    int need_tb = (set->trim_side != 0) || cfg->want_stats;    /* :812 */
    double score, delta;
    int64_t s, e;
    int32_t bc = orc_find_best(cfg, set, r, n, start_j, end_j, max_start_pos, min_end_pos, need_tb,
                               &score, &delta, &s, &e);
    if (bc == 0) return;                                       /* :820-821 */
    if (delta < cfg->min_delta) {                              /* :822-823 */
        out->status = ORC_STATUS_AMBIGUOUS;
        return;
    }
    out->status = ORC_STATUS_MATCH;
    out->bc = bc;
    out->start = s;
    out->end = e;
    out->score = score;
}

/* ------------------------------------------------------------------ */
/* determine_filename[_and_stats]  (classification.jl:871-1005)        */
/* ------------------------------------------------------------------ */
void orc_determine(const orc_config *cfg, const uint8_t *r, int64_t n, orc_result *out)
{
    memset(out, 0, sizeof(*out));
    out->keep_start = -1;
    out->keep_end = -1;
    out->pass[1].status = -1;
    out->pass[1].start = -1;
    out->pass[1].end = -1;
    out->pass[1].score = INFINITY;

    orc_match_pass(cfg, 0, r, n, &out->pass[0]);
    if (out->pass[0].status != ORC_STATUS_MATCH) {             /* :879-883 */
        out->status = out->pass[0].status;
        return;
    }
    int64_t start1 = out->pass[0].start, end1 = out->pass[0].end;
    int64_t start2 = -1, end2 = -1;
    if (cfg->is_dual) {                                        /* :887-897 */
        orc_match_pass(cfg, 1, r, n, &out->pass[1]);
        if (out->pass[1].status != ORC_STATUS_MATCH) {
            out->status = out->pass[1].status;
            return;
        }
        out->bc2 = out->pass[1].bc;
        start2 = out->pass[1].start;
        end2 = out->pass[1].end;
    }
    out->status = ORC_STATUS_MATCH;
    out->bc1 = out->pass[0].bc;

    int64_t keep_start = 1, keep_end = n;                      /* :907-908 */
    if (cfg->set1.trim_side == 3) keep_end = i64max(1, start1) - 1;        /* :914 */
    else if (cfg->set1.trim_side == 5) keep_start = end1 + 1;              /* :917 */
    if (cfg->is_dual && cfg->set2.trim_side != 0) {            /* :921-929 */
        if (cfg->set2.trim_side == 3) keep_end = i64min(keep_end, i64max(1, start2) - 1);
        else if (cfg->set2.trim_side == 5) keep_start = i64max(keep_start, end2 + 1);
    }
    if (keep_start > keep_end) { keep_start = 1; keep_end = 0; }           /* :932-935 */
    out->keep_start = keep_start;
    out->keep_end = keep_end;
}

void orc_classify(const orc_config *cfg, const uint8_t *seqs, const int64_t *offsets,
                  int64_t n_reads, orc_result *out)
{
    for (int64_t i = 0; i < n_reads; i++)
        orc_determine(cfg, seqs + offsets[i], offsets[i + 1] - offsets[i], &out[i]);
}
