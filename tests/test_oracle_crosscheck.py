"""Second, independent restatement used to cross-check the C oracle where the reference's own tests pin
nothing (SURVEY.md 8c "gaps"): a textbook unpruned semi-global DP, a brute-force Hamming scan and a
brute-force substring search, written from the definitions rather than from the reference's control flow.
In the regimes where the reference's pruning is exact (benign costs) the oracle must agree
with them; reported alignment positions are checked as properties (the reported substring really aligns
globally to the barcode at the reported cost), which is independent of tie-breaking."""
import math

import numpy as np
import pytest

import orc

INF = float("inf")


def clean_semiglobal(q, r, a, b, match, mismatch, indel, nindel):
    """min over columns j in [a, b] of D[m][j]; D[0][j] = 0 (free start), D[i][a-1] = indel * i
    (classification.jl:278-279 uses `indel` there even under NScoring), no insertion in the last row."""
    m = len(q)
    big = 10 ** 9
    prev = [indel * i for i in range(m + 1)]
    best = big
    for j in range(a, b + 1):
        cur = [0] * (m + 1)
        for i in range(1, m + 1):
            is_n = nindel is not None and q[i - 1] == ord("N")
            gap = nindel if is_n else indel
            same = is_n or q[i - 1] == r[j - 1]
            sub = prev[i - 1] + (match if same else mismatch)
            dele = cur[i - 1] + gap
            ins = prev[i] + gap if i < m else big
            cur[i] = min(sub, dele, ins)
        best = min(best, cur[m])
        prev = cur
    return best


def global_cost(q, s, match, mismatch, indel, nindel):
    """weighted global alignment cost of barcode q against the read substring s"""
    m, n = len(q), len(s)
    gap = [nindel if (nindel is not None and c == ord("N")) else indel for c in q]
    D = [[0] * (n + 1) for _ in range(m + 1)]
    for i in range(1, m + 1):
        D[i][0] = D[i - 1][0] + gap[i - 1]
    for j in range(1, n + 1):
        D[0][j] = 10 ** 9                      # the barcode must be consumed from its first base: no leading read-only columns
    D[0][0] = 0
    for i in range(1, m + 1):
        for j in range(1, n + 1):
            same = (nindel is not None and q[i - 1] == ord("N")) or q[i - 1] == s[j - 1]
            D[i][j] = min(D[i - 1][j - 1] + (match if same else mismatch), D[i - 1][j] + gap[i - 1],
                          D[i][j - 1] + gap[i - 1] if 0 < i < m else 10 ** 9)
    return D[m][n]


def _rand_case(rng, with_n):
    m = int(rng.integers(3, 13))
    n = int(rng.integers(1, 41))
    alpha = b"ACGT"
    q = bytes(alpha[k] for k in rng.integers(0, 4, m))
    r = bytearray(alpha[k] for k in rng.integers(0, 4, n))
    if n >= m and rng.random() < 0.8:                       # plant a mutated copy
        mut = bytearray(q)
        for _ in range(int(rng.integers(0, 4))):
            op, pos = int(rng.integers(0, 3)), int(rng.integers(0, max(len(mut), 1)))
            if op == 0 and mut:
                mut[pos] = alpha[int(rng.integers(0, 4))]
            elif op == 1:
                mut.insert(pos, alpha[int(rng.integers(0, 4))])
            elif mut:
                del mut[pos]
        st = int(rng.integers(0, max(n - len(mut), 0) + 1))
        r[st:st + len(mut)] = mut
        r = r[:n]
    if with_n:
        q = bytearray(q)
        for _ in range(int(rng.integers(1, 3))):
            q[int(rng.integers(0, m))] = ord("N")
        q = bytes(q)
    if rng.random() < 0.2 and len(r):
        r[int(rng.integers(0, len(r)))] = ord("N")
    return q, bytes(r)


@pytest.mark.parametrize("costs", [(0, 1, 1, None), (0, 1, 2, None), (0, 2, 1, None), (0, 3, 2, None),
                                   (0, 1, 1, 1), (0, 1, 1, 2), (0, 2, 1, 3)])
def test_semiglobal_score_equals_clean_dp(costs):
    match, mismatch, indel, nindel = costs
    rng = np.random.default_rng(sum(x or 0 for x in costs) * 7919 + 13)
    n_hit = 0
    for _ in range(1500):
        q, r = _rand_case(rng, nindel is not None)
        n = len(r)
        a = int(rng.integers(1, n + 1))
        b = int(rng.integers(a, n + 1))
        if rng.random() < 0.5:
            a, b = 1, n
        thr = float(rng.choice([0.0, 0.1, 0.2, 0.34, 0.5, 0.75]))
        norm = sum(1 for c in q if c != ord("N")) if nindel is not None else len(q)
        if norm == 0:
            continue
        allowed = math.floor(thr * norm)
        d = clean_semiglobal(q, r, a, b, match, mismatch, indel, nindel)
        want = d / norm if d <= allowed else INF
        got = orc.semiglobal(q, r, thr, match, mismatch, indel, nindel, rng=(a, b), max_start_pos=n, min_end_pos=1,
                             norm=norm)
        assert got == want, (q, r, a, b, thr, costs, got, want)
        n_hit += want != INF
        for trim in (None, 3, 5):                               # positions: property check
            sc, s, e = orc.semiglobal(q, r, thr, match, mismatch, indel, nindel, rng=(a, b), max_start_pos=n,
                                      min_end_pos=1, norm=norm, traceback=True, trim_side=trim)
            assert sc == want
            if want != INF and s >= 1:
                assert a <= s and e <= b and s <= e + 1
                # An alignment that opens by skipping barcode bases at column j is labelled with start j by the
                # reference (the row-0 origin of the deletion move, classification.jl:305-306), one column
                # before the first read base it consumes: accept either reading of the label.
                costs_se = [global_cost(q, r[s - 1:e], match, mismatch, indel, nindel)]
                if s <= e:
                    costs_se.append(global_cost(q, r[s:e], match, mismatch, indel, nindel))
                assert d in costs_se, (q, r, s, e, d, trim, costs_se)
    assert n_hit > 200


def test_hamming_equals_brute_force():
    rng = np.random.default_rng(77)
    for _ in range(3000):
        q, r = _rand_case(rng, rng.random() < 0.3)
        m, n = len(q), len(r)
        a = int(rng.integers(1, n + 1))
        b = int(rng.integers(a, n + 1))
        max_start = int(rng.integers(1, n + 1)) if rng.random() < 0.5 else n
        min_end = int(rng.integers(1, n + 1)) if rng.random() < 0.5 else 1
        thr = float(rng.choice([0.0, 0.1, 0.2, 0.34, 0.5]))
        allowed = math.floor(thr * m)
        for trim in (None, 3, 5):
            cands = []
            for s in range(max(a, 1), min(b, max_start, n - m + 1) + 1):
                if s + m - 1 < min_end:
                    continue
                mm = sum(1 for i in range(m) if q[i] != ord("N") and q[i] != r[s - 1 + i])
                if mm <= allowed:
                    cands.append((mm, s))
            if cands:
                best = min(c[0] for c in cands)
                starts = [s for mm, s in cands if mm == best]
                s = max(starts) if trim == 3 else min(starts)
                want = (best / m, s, s + m - 1)
            else:
                want = (INF, -1, -1)
            got = orc.hamming(q, r, thr, (a, b), max_start, min_end, trim)
            assert got == want, (q, r, a, b, max_start, min_end, thr, trim, got, want)


def test_exact_equals_brute_force():
    rng = np.random.default_rng(78)
    for _ in range(3000):
        q, r = _rand_case(rng, False)
        q = q[:int(rng.integers(2, 6))]
        m, n = len(q), len(r)
        a = int(rng.integers(1, n + 1))
        b = int(rng.integers(a, n + 1))
        max_start = int(rng.integers(1, n + 1)) if rng.random() < 0.5 else n
        min_end = int(rng.integers(1, n + 1)) if rng.random() < 0.5 else 1
        for trim in (None, 3, 5):
            starts = [s for s in range(max(a, 1), min(b, max_start, n - m + 1) + 1)
                      if r[s - 1:s - 1 + m] == q and s + m - 1 >= min_end]
            if trim == 3:
                # findprev takes the rightmost occurrence in the window and gives up if THAT one is invalid
                # (classification.jl:499-515): brute force over occurrences ignoring min_end first
                occ = [s for s in range(1, n - m + 2) if r[s - 1:s - 1 + m] == q and s <= min(b, max_start, n - m + 1)]
                want = (INF, -1, -1)
                if occ and min(b, max_start, n - m + 1) >= max(a, 1):
                    s = max(occ)
                    if s >= max(a, 1) and s + m - 1 >= min_end:
                        want = (0.0, s, s + m - 1)
            else:
                want = (0.0, min(starts), min(starts) + m - 1) if starts else (INF, -1, -1)
            got = orc.exact(q, r, (a, b), max_start, min_end, trim)
            assert got == want, (q, r, a, b, max_start, min_end, trim, got, want)
