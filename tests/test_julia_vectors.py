"""Loader for golden vectors produced by the UNMODIFIED reference (tools/make_golden.jl; needs Julia, which this
image does not have).  When tests/golden/julia_vectors.jsonl exists, every vector is replayed through the C
oracle and the Python transcription; until then the loader itself is exercised on a file of the same format
written by the Python transcription."""
import json
import math
import os
import random

import pytest

import bdx_b200 as bdx
import orc
import pyref

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "julia_vectors.jsonl")


def _f(x):
    return {"Inf": math.inf, "NaN": math.nan}.get(x, x) if isinstance(x, str) else float(x)


def _same(a, b):
    return a == b or (math.isinf(a) and math.isinf(b) and (a > 0) == (b > 0)) or (math.isnan(a) and math.isnan(b))


def replay(path, impls=("oracle", "pyref")):
    """Checks every vector of `path`; returns the number of vectors."""
    n = 0
    for line in open(path):
        v = json.loads(line)
        n += 1
        if v["kind"] == "align":
            args = (v["q"].encode(), v["r"].encode(), _f(v["max_error"]), v["match"], v["mismatch"], v["indel"], v["nindel"],
                    (v["lo"], v["hi"]), v["max_start"], v["min_end"], v["norm"], v["traceback"], v["trim"])
            want = (_f(v["score"]), v["start"], v["end"])
            for impl in impls:
                if impl == "pyref":
                    got = pyref.semiglobal_core(*args)
                else:
                    got = orc.semiglobal(args[0], args[1], args[2], match=args[3], mismatch=args[4], indel=args[5],
                                         nindel=args[6], rng=args[7], max_start_pos=args[8], min_end_pos=args[9],
                                         norm=args[10], traceback=args[11], trim_side=args[12])
                if not isinstance(got, tuple):
                    got = (got, -1, -1)
                assert _same(got[0], want[0]) and tuple(got[1:]) == want[1:], (impl, v, got)
        else:
            bcs = [b.encode() for b in v["barcodes"]]
            want = (v["bc"], _f(v["score"]), _f(v["delta"]), v["start"], v["end"])
            for impl in impls:
                if impl == "pyref":
                    got = pyref.find_best(v["read"].encode(), bcs, v["norms"], _f(v["max_error_rate"]), _f(v["min_delta"]), 0,
                                          v["mismatch"], v["indel"], None, (v["lo"], v["hi"]), v["max_start"], v["min_end"],
                                          v["trim"], v["need_tb"])
                else:
                    cfg = bdx.DemuxConfig(bc_seqs=v["barcodes"], bc_lengths_no_N=v["norms"], ids=[str(i) for i in range(len(bcs))],
                                          max_error_rate=_f(v["max_error_rate"]), min_delta=_f(v["min_delta"]),
                                          mismatch=v["mismatch"], indel=v["indel"], trim_side=v["trim"])
                    got = orc.Oracle(cfg).find_best(v["read"].encode(), (v["lo"], v["hi"]), v["max_start"], v["min_end"],
                                                    need_traceback=v["need_tb"])
                assert got[0] == want[0] and _same(got[1], want[1]) and _same(got[2], want[2]) and tuple(got[3:]) == want[3:], (impl, v, got)
    return n


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="no vectors from the Julia reference (tools/make_golden.jl was not run)")
def test_julia_reference_vectors():
    assert replay(GOLDEN) > 0


def test_loader_on_transcription_vectors(tmp_path):
    """The same file format, written by the Python transcription: the loader and both replays work."""
    rnd = random.Random(5)

    def js(x):
        return "Inf" if math.isinf(x) else x

    path = tmp_path / "vectors.jsonl"
    with open(path, "w") as fh:
        for _ in range(300):
            q = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(3, 10)))
            r = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(6, 20)))
            r = r[:3] + q[:-1] + r[3:]
            tb, trim = rnd.random() < 0.5, rnd.choice([None, 3, 5])
            tb = tb or trim is not None
            lo, hi = 1, len(r)
            ms, me = rnd.randint(1, len(r)), rnd.randint(1, len(r))
            res = pyref.semiglobal_core(q.encode(), r.encode(), 0.34, 0, 1, 1, None, (lo, hi), ms, me, len(q), tb, trim)
            sc, s, e = res if isinstance(res, tuple) else (res, -1, -1)
            fh.write(json.dumps(dict(kind="align", q=q, r=r, max_error=0.34, match=0, mismatch=1, indel=1, nindel=None, lo=lo,
                                     hi=hi, max_start=ms, min_end=me, norm=len(q), traceback=tb, trim=trim, score=js(sc),
                                     start=s, end=e)) + "\n")
            bcs = ["".join(rnd.choice("ACGT") for _ in range(rnd.randint(3, 9))) for _ in range(4)]
            bc, sc, dl, s, e = pyref.find_best(r.encode(), [b.encode() for b in bcs], [len(b) for b in bcs], 0.4, 0.1, 0, 1, 1,
                                               None, (lo, hi), ms, me, trim, tb)
            fh.write(json.dumps(dict(kind="find_best", read=r, barcodes=bcs, norms=[len(b) for b in bcs], max_error_rate=0.4,
                                     min_delta=0.1, mismatch=1, indel=1, lo=lo, hi=hi, max_start=ms, min_end=me, trim=trim,
                                     need_tb=tb, bc=bc, score=js(sc), delta=js(dl), start=s, end=e)) + "\n")
    assert replay(str(path)) == 600
