"""Differential fuzz of the C oracle (oracle/bdx_oracle.c) against an INDEPENDENT Python transcription of
classification.jl (tests/pyref.py), in the regimes that no reference test pins (SURVEY.md section 8c "gaps"):
traceback through indels (trim 3 / 5 / stats-only), max_start_pos < n with errors allowed, min_end_pos > 1,
sub-ranges, variable-length sets, nindel != indel (also nindel < indel), match != 0 (also negative), m > n,
thresholds whose IEEE floor differs from the rational one.  >= 200 000 alignments in the default run
(BDX_PYREF_SCALE multiplies the case counts for soak runs)."""
import math
import os
import random

import pytest

import orc
import pyref

SCALE = float(os.environ.get("BDX_PYREF_SCALE", "1"))
THRESHOLDS = [0.0, 0.1, 0.2, 0.25, 0.29, 3 / 11, 0.34, 0.4, 0.5, 0.6, 15 / 22, 1.0]


def _seq(rnd, lo, hi, alphabet=b"ACGT"):
    return bytes(rnd.choice(alphabet) for _ in range(rnd.randint(lo, hi)))


def _planted(rnd, q, n_lo, n_hi, n_frac=0.0):
    """a read with a mutated copy of q somewhere (so that alignments with a few edits exist)"""
    r = bytearray(_seq(rnd, n_lo, n_hi))
    if rnd.random() < 0.8 and len(r):
        mq = bytearray(q.replace(b"N", b"A"))
        for _ in range(rnd.choice([0, 0, 1, 1, 2, 3])):
            k = rnd.randint(0, 2)
            if k == 0 and mq:
                mq[rnd.randrange(len(mq))] = rnd.choice(b"ACGT")
            elif k == 1:
                mq.insert(rnd.randint(0, len(mq)), rnd.choice(b"ACGT"))
            elif len(mq) > 1:
                del mq[rnd.randrange(len(mq))]
        st = rnd.randint(0, max(len(r) - 1, 0))
        r[st:st + len(mq)] = mq
    if n_frac and len(r) and rnd.random() < n_frac:
        r[rnd.randrange(len(r))] = ord("N")
    return bytes(r)


def _same(a, b):
    if isinstance(a, tuple):
        return _same(a[0], b[0]) and a[1:] == b[1:]
    return a == b or (math.isinf(a) and math.isinf(b)) or (math.isnan(a) and math.isnan(b))


REGIMES = {
    # name: (cases, kwargs of the generator)
    "traceback_indels": (30000, dict(tb=True)),
    "start_bound": (35000, dict(start_bound=True)),
    "end_bound": (30000, dict(end_bound=True)),
    "both_bounds_subrange": (25000, dict(start_bound=True, end_bound=True, sub=True)),
    "nindel": (25000, dict(nscoring=True)),
    "match_nonzero": (20000, dict(match=True)),
    "m_gt_n": (15000, dict(long_q=True)),
    "weighted_costs_tb": (20000, dict(costs=True, tb=True, start_bound=True)),
}


@pytest.mark.parametrize("name", sorted(REGIMES))
def test_alignment_core_regimes(name):
    cases, g = REGIMES[name]
    rnd = random.Random(sum(map(ord, name)))
    for case in range(int(cases * SCALE)):
        nscoring = g.get("nscoring", False)
        alpha = b"ACGTN" if nscoring else b"ACGT"
        q = _seq(rnd, 2, 14, alpha)
        if g.get("long_q"):
            q = _seq(rnd, 8, 20)
            r = _planted(rnd, q[:rnd.randint(2, 8)], 1, 10)
        else:
            r = _planted(rnd, q, 4, 28, n_frac=0.2)
        n = len(r)
        indel = rnd.choice([1, 1, 2, 3]) if (g.get("costs") or nscoring or g.get("match")) else 1
        mismatch = rnd.choice([1, 1, 2, 3]) if (g.get("costs") or g.get("match")) else 1
        match = rnd.choice([0, 1, -1, 2]) if g.get("match") else 0
        nindel = rnd.choice([1, 2, 3]) if nscoring else None
        norm = len(q) if not nscoring else max(sum(1 for c in q if c != ord("N")), 1)
        lo, hi = 1, n
        if g.get("sub") or rnd.random() < 0.2:
            lo = rnd.randint(1, max(n // 2, 1))
            hi = rnd.randint(lo, n)
        max_start = rnd.randint(lo, min(lo + 8, n)) if g.get("start_bound") else n
        min_end = rnd.randint(max(hi - 8, 1), hi) if g.get("end_bound") else rnd.choice([1, 1, lo])
        tb = g.get("tb", False) or rnd.random() < 0.5
        trim = rnd.choice([None, 3, 5]) if tb else None
        thr = rnd.choice(THRESHOLDS)
        want = pyref.semiglobal_core(q, r, thr, match, mismatch, indel, nindel, (lo, hi), max_start, min_end, norm,
                                     tb, trim)
        got = orc.semiglobal(q, r, thr, match=match, mismatch=mismatch, indel=indel, nindel=nindel, rng=(lo, hi),
                             max_start_pos=max_start, min_end_pos=min_end, norm=norm, traceback=tb, trim_side=trim)
        assert _same(got, want), (name, case, q, r, thr, match, mismatch, indel, nindel, (lo, hi), max_start, min_end,
                                  norm, tb, trim, got, want)


class _Cfg:
    """the fields orc.Oracle reads from a DemuxConfig"""

    def __init__(self, bcs, norms, **kw):
        import bdx_b200 as bdx
        self.__dict__.update(bdx.DemuxConfig(bc_seqs=[b.decode() for b in bcs], bc_lengths_no_N=norms,
                                             ids=[str(i) for i in range(len(bcs))], **kw).__dict__)


@pytest.mark.parametrize("with_delta", [False, True])
def test_find_best_order_dependence(with_delta):
    """find_best_matching_bc_{no,with}_delta over variable-length sets with a constrained start: the running
    threshold changes what later barcodes may return (vector V9)."""
    import bdx_b200 as bdx
    rnd = random.Random(99 + with_delta)
    for case in range(int(6000 * SCALE)):
        nb = rnd.randint(2, 7)
        nscoring = rnd.random() < 0.2
        bcs = [_seq(rnd, 3, 12, b"ACGTN" if nscoring else b"ACGT") for _ in range(nb)]
        if rnd.random() < 0.3:
            bcs[rnd.randrange(nb)] = bcs[rnd.randrange(nb)]
        norms = [max(sum(1 for c in b if c != ord("N")), 1) for b in bcs]
        read = _planted(rnd, rnd.choice(bcs), 8, 30, n_frac=0.1)
        n = len(read)
        lo = rnd.randint(1, 4)
        hi = rnd.randint(max(lo, n - 6), n)
        max_start = rnd.randint(lo, min(lo + 9, n))
        min_end = rnd.randint(1, hi)
        kw = dict(max_error_rate=rnd.choice(THRESHOLDS), min_delta=rnd.choice([0.05, 0.1, 0.2, 0.34]) if with_delta else 0.0,
                  mismatch=rnd.choice([1, 1, 2]), indel=rnd.choice([1, 1, 2]), nindel=rnd.choice([1, 2]) if nscoring else None)
        trim = rnd.choice([None, None, 3, 5])
        need_tb = rnd.random() < 0.3
        want = pyref.find_best(read, bcs, norms, kw["max_error_rate"], kw["min_delta"], 0, kw["mismatch"], kw["indel"],
                               kw["nindel"], (lo, hi), max_start, min_end, trim, need_tb)
        cfg = bdx.DemuxConfig(bc_seqs=[b.decode() for b in bcs], bc_lengths_no_N=norms, ids=[str(i) for i in range(nb)],
                              trim_side=trim, **kw)
        got = orc.Oracle(cfg).find_best(read, (lo, hi), max_start, min_end, need_traceback=need_tb)
        assert got[0] == want[0] and _same(got[1], want[1]) and _same(got[2], want[2]) and got[3:] == want[3:], (
            case, bcs, read, kw, (lo, hi), max_start, min_end, trim, need_tb, got, want)


def test_match_pass_dual_ranges():
    """match_barcode_pass through orc_classify: dynamic ranges resolved per read length, dual passes on the same
    read, status decisions -- against pyref.match_pass."""
    import numpy as np
    import bdx_b200 as bdx
    R = bdx.parse_dynamic_range
    rnd = random.Random(4242)
    for case in range(int(400 * SCALE)):
        b1 = [_seq(rnd, 4, 10) for _ in range(rnd.randint(2, 5))]
        b2 = [_seq(rnd, 4, 10) for _ in range(rnd.randint(1, 4))]
        a, w = rnd.randint(1, 4), rnd.randint(8, 20)
        rs = [f"{a}:{a + w}", f"1:end", f"end-{w}:end", f"{a}:end-{a}"]
        kw = dict(max_error_rate=rnd.choice(THRESHOLDS), min_delta=rnd.choice([0.0, 0.1, 0.2]),
                  ref_search_range=R(rnd.choice(rs)), barcode_start_range=R(rnd.choice([f"1:{a + 5}", "1:end", f"{a}:end-3"])),
                  barcode_end_range=R(rnd.choice(["1:end", f"{a + 3}:end", "end-12:end"])),
                  ref_search_range2=R(rnd.choice(rs)), barcode_start_range2=R(rnd.choice(["1:end", f"end-{w}:end"])),
                  barcode_end_range2=R(rnd.choice(["1:end", "end-5:end"])), trim_side=rnd.choice([None, 3, 5]),
                  trim_side2=rnd.choice([None, 3, 5]))
        cfg = bdx.DemuxConfig(bc_seqs=[b.decode() for b in b1], bc_lengths_no_N=[len(b) for b in b1],
                              ids=[str(i) for i in range(len(b1))], is_dual=True, bc_seqs2=[b.decode() for b in b2],
                              bc_lengths_no_N2=[len(b) for b in b2], ids2=[str(i) for i in range(len(b2))], **kw)
        reads = []
        for _ in range(12):
            r = bytearray(_planted(rnd, rnd.choice(b1), 20, 44))
            t = _planted(rnd, rnd.choice(b2), 10, 12)
            r[len(r) - len(t):] = t
            reads.append(bytes(r))
        got = orc.Oracle(cfg).classify_reads(reads)
        opts = dict(max_error_rate=kw["max_error_rate"], min_delta=kw["min_delta"], match=0, mismatch=1, indel=1, nindel=None)

        def rng4(d):
            return (d.start_offset, d.start_from_end, d.end_offset, d.end_from_end)

        for i, read in enumerate(reads):
            p1 = pyref.match_pass(read, b1, None, opts, tuple(rng4(kw[k]) for k in ("ref_search_range", "barcode_start_range", "barcode_end_range")),
                                  kw["trim_side"], False)
            code = {"match": 0, "unknown": 1, "ambiguous": 2}
            g1 = got["passes"][i, 0]
            assert g1["status"] == code[p1[0]] and (p1[0] != "match" or (g1["bc"], g1["start"], g1["end"], g1["score"]) == p1[1:]), (case, i, p1, g1)
            if p1[0] != "match":
                assert got["status"][i] == code[p1[0]]
                continue
            p2 = pyref.match_pass(read, b2, None, opts, tuple(rng4(kw[k]) for k in ("ref_search_range2", "barcode_start_range2", "barcode_end_range2")),
                                  kw["trim_side2"], False)
            g2 = got["passes"][i, 1]
            assert g2["status"] == code[p2[0]] and (p2[0] != "match" or (g2["bc"], g2["start"], g2["end"], g2["score"]) == p2[1:]), (case, i, p2, g2)
            assert got["status"][i] == code[p2[0]]
