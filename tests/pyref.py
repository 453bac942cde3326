"""Second, independent restatement of the reference's semi-global path -- TEST INFRASTRUCTURE.

Written from the Julia text of /root/reference/src/classification.jl (lines cited per function), NOT from
oracle/bdx_oracle.c: it exists so that the C oracle can be differential-fuzzed in the regimes no reference
test pins (SURVEY.md section 8c "gaps": traceback through indels, barcode_start_range with errors allowed,
barcode_end_range, variable-length sets, nindel != indel, match != 0, m > n).  Plain Python ints / floats
(Julia Int64 / Float64 semantics for these magnitudes); pure-Python loops, so only small cases.

  semiglobal_core      classification.jl:238-445  (+ step_scores* :178-236, policies :130-176)
  find_best            classification.jl:632-713
  match_pass           classification.jl:776-824  (range prologue + status decision)
"""
from __future__ import annotations

import math

INF_INT = (2 ** 63 - 1) // 4          # classification.jl:7  typemax(Int) >> 2
N_BYTE = ord("N")


def _jdiv(a: int, b: int) -> int:
    """Julia's div: truncation toward zero."""
    qt = abs(a) // abs(b)
    return qt if (a < 0) == (b < 0) else -qt


def _fdiv(a: int, b: int) -> float:
    """Float64(a) / Float64(b) with IEEE results for b == 0."""
    if b == 0:
        return math.nan if a == 0 else math.copysign(math.inf, a)
    return a / b


def semiglobal_core(q: bytes, r: bytes, max_error: float, match: int, mismatch: int, indel: int, nindel,
                    rng, max_start_pos: int, min_end_pos: int, norm: int, traceback: bool, trim_side):
    """semiglobal_alignment_core.  rng = (first, last) 1-based inclusive.  Returns a float score (ScoreOnly) or
    (score, start, end) (TracebackOutput(trim_side)); float('inf') when nothing is within the threshold."""
    m, n = len(q), len(r)
    nscoring = nindel is not None

    def finalize(res):                                               # :155-168
        if traceback:
            sc, s, e = res
            return (math.inf if sc >= INF_INT else _fdiv(sc, norm), s, e)
        return math.inf if res >= INF_INT else _fdiv(res, norm)

    result = (INF_INT, -1, -1) if traceback else INF_INT            # :130-136
    if m == 0 or n == 0:                                             # :250-252
        return finalize(result)
    allowed_error = math.floor(max_error * norm)                     # :254
    gap_floor = min(indel, nindel) if nscoring else indel            # :170-176
    max_indel_steps = _jdiv(allowed_error, gap_floor)
    min_valid_start = min_end_pos - (m + max_indel_steps) + 1        # :259
    if min_valid_start > max_start_pos:                              # :261-263
        return finalize(result)
    lo, hi = rng
    if min_valid_start > lo:                                         # :266-268
        lo = max(lo, min_valid_start)
    band_offset = max(m - n - max_indel_steps, -max_start_pos - max_indel_steps)   # :270

    DP = [0] * (m + 2)                                               # 1-based like the workspace
    origin = [0] * (m + 2)
    for i in range(1, m + 1):                                        # :278-283
        DP[i] = indel * i
        origin[i] = 1 - i
    lact = min(allowed_error + 1, m)                                 # :286

    for j in range(lo, hi + 1):                                      # :287
        prev_origin = j
        if j + band_offset >= 1:                                     # :289-295
            fact = j + band_offset
            prev = allowed_error
        else:
            fact = 1
            prev = 0
        if fact > lact:                                              # :297-299
            return finalize(result)
        rj = r[j - 1]

        def cell(i, first_or_last):
            """one DP cell; returns (score, origin) -- :178-236 and the origin choice :306-324 / :343-362 / :379-398"""
            qi = q[i - 1]
            is_n = nscoring and qi == N_BYTE
            cost = nindel if is_n else indel
            if first_or_last:                                        # step_scores :208-236
                if nscoring:
                    ins = DP[i] + (INF_INT if i == m else cost)      # :229
                else:
                    ins = INF_INT if i == m else DP[i] + indel       # :213
                diag = 0 if i == 1 else DP[i - 1]
            else:                                                    # step_scores_main :178-206
                ins = DP[i] + cost
                diag = DP[i - 1]
            dele = prev + cost
            sub = diag + (match if (qi == rj or is_n) else mismatch)
            best, bo = dele, prev_origin
            sub_o = j if (first_or_last and i == 1) else origin[i - 1]
            if sub < best:
                best, bo = sub, sub_o
            if ins < best:
                best, bo = ins, origin[i]
            return min(ins, dele, sub), bo

        # 1. first computed row (:301-335)
        sc, co = cell(fact, True)
        if fact != 1:
            DP[fact - 1] = prev
            origin[fact - 1] = prev_origin
        prev, prev_origin = sc, co
        # 2. interior rows (:338-373)
        limit = m - 1 if lact == m else lact
        for i in range(fact + 1, limit + 1):
            sc, co = cell(i, False)
            DP[i - 1] = prev
            origin[i - 1] = prev_origin
            prev, prev_origin = sc, co
        # 3. last row (:376-409)
        if lact == m and lact > fact:
            sc, co = cell(m, True)
            DP[m - 1] = prev
            origin[m - 1] = prev_origin
            prev, prev_origin = sc, co
        DP[lact] = prev                                              # :412-415
        origin[lact] = prev_origin

        if lact == m and prev <= allowed_error:                      # :417-438
            lact -= 1
            if j >= min_end_pos:
                if prev == 0 and (not traceback or trim_side == 5):  # :420-429
                    return finalize((0, prev_origin, j)) if traceback else finalize(0)
                if traceback:                                        # update_result :142-153
                    bs, bst, _ = result
                    if prev < bs:
                        result = (prev, prev_origin, j)
                    elif prev == bs and trim_side == 3 and prev_origin > bst:
                        result = (prev, prev_origin, j)
                else:
                    result = min(result, prev)
        while lact > 0 and DP[lact] > allowed_error:                 # :439-442
            lact -= 1
        lact += 1
    return finalize(result)


def find_best(read: bytes, barcodes, norms, max_error_rate: float, min_delta: float, match: int, mismatch: int,
              indel: int, nindel, rng, max_start_pos: int, min_end_pos: int, trim_side, need_traceback: bool):
    """find_best_matching_bc (:722-728) -> (bc index 1-based or 0, min_score, delta, start, end)."""
    tb = trim_side is not None or need_traceback
    thr = max_error_rate
    min_score, sub_min, best_bc, bs, be = math.inf, math.inf, 0, -1, -1
    for i, bc in enumerate(barcodes, start=1):
        norm = norms[i - 1] if nindel is not None else len(bc)       # :460, :476, :647
        res = semiglobal_core(bc, read, thr, match, mismatch, indel, nindel, rng, max_start_pos, min_end_pos, norm,
                              tb, trim_side)
        score, s, e = res if tb else (res, -1, -1)
        if min_delta == 0.0:                                         # :658-664
            if score <= thr and score < min_score:
                min_score, best_bc, bs, be = score, i, s, e
                thr = min(thr, min_score)
        elif score <= thr:                                           # :696-708
            if score < min_score:
                sub_min = min_score
                min_score, best_bc, bs, be = score, i, s, e
                thr = min(thr, sub_min)
            elif score < sub_min:
                sub_min = score
                thr = min(thr, sub_min)
    delta = math.inf if min_delta == 0.0 else sub_min - min_score    # :666, :711
    return best_bc, min_score, delta, bs, be


def resolve(dr, length: int):
    """resolve(::DynamicRange, len) (:96-100); dr = (start_offset, start_from_end, end_offset, end_from_end)."""
    so, sf, eo, ef = dr
    s = length + so if sf else so
    e = length + eo if ef else eo
    first, last = max(1, s), min(length, e)
    if last < first:                                                 # Julia's UnitRange normalisation
        last = first - 1
    return first, last


def match_pass(read: bytes, barcodes, norms, opts, ranges, trim_side, want_stats: bool):
    """match_barcode_pass (:776-824) -> (status, bc, start, end, score).  ranges = (ref_search, bc_start, bc_end)."""
    n = len(read)
    rs, bs_, be_ = (resolve(x, n) for x in ranges)
    start_j = max(rs[0], bs_[0], 1)                                  # :799-802
    end_j = min(rs[1], be_[1], n)
    max_start_pos, min_end_pos = bs_[1], be_[0]
    if start_j > end_j or start_j > max_start_pos or end_j < min_end_pos:   # :805-807
        return "unknown", 0, -1, -1, math.inf
    need_tb = trim_side is not None or want_stats                    # :812
    bc, score, delta, s, e = find_best(read, barcodes, norms, opts["max_error_rate"], opts["min_delta"], opts["match"],
                                       opts["mismatch"], opts["indel"], opts["nindel"], (start_j, end_j), max_start_pos,
                                       min_end_pos, trim_side, need_tb)
    if bc == 0:                                                      # :820-824
        return "unknown", 0, -1, -1, math.inf
    if delta < opts["min_delta"]:
        return "ambiguous", 0, -1, -1, math.inf
    return "match", bc, s, e, score
