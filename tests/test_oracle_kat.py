"""Known-answer tests of the reference's own unit tests, replayed on the oracle.
Sources: test/unit/alignment.jl, trimming.jl, hamming.jl, exact.jl, plus the
restatement cross-checks of SURVEY.md section 10 (V1-V10, not pinned by a reference test)."""
import math

import pytest

import bdx_b200 as bdx
import orc

INF = math.inf


def test_alignment_n_support():  # test/unit/alignment.jl:5-47
    kw = dict(match=0, mismatch=1, indel=1, nindel=1, max_start_pos=1, norm=2)
    assert orc.semiglobal("ANNC", "ATTC", 0.5, rng=(1, 4), min_end_pos=4, **kw) == 0.0
    assert orc.semiglobal("ANNC", "ATTG", 0.5, rng=(1, 4), min_end_pos=4, **kw) == 0.5
    assert orc.semiglobal("ANNC", "ATT", 0.5, rng=(1, 3), min_end_pos=3, **kw) == 0.5


@pytest.mark.parametrize("impl", ["oracle", "host"])
def test_position_restriction_logic(impl):  # test/unit/alignment.jl:49-79
    if impl == "oracle":
        parse, resolve = orc.parse_dynamic_range, orc.resolve
    else:
        parse, resolve = bdx.parse_dynamic_range, bdx.resolve
    d1, d2, d3 = parse("1:10"), parse("1:end"), parse("end-5:end")
    assert (d1.start_offset, bool(d1.start_from_end), d1.end_offset, bool(d1.end_from_end)) == (1, False, 10, False)
    assert (d2.start_offset, bool(d2.start_from_end), d2.end_offset, bool(d2.end_from_end)) == (1, False, 0, True)
    assert (d3.start_offset, bool(d3.start_from_end), d3.end_offset, bool(d3.end_from_end)) == (-5, True, 0, True)
    assert resolve(d1, 100) == (1, 10)
    assert resolve(d2, 100) == (1, 100)
    assert resolve(d3, 100) == (95, 100)
    with pytest.raises(ValueError):
        parse("1:2:3")
    with pytest.raises(ValueError):
        parse("17")
    # empty range keeps Julia's UnitRange normalisation last = first - 1
    assert resolve(parse("10:5"), 100) == (10, 9)
    assert resolve(parse("end+3:end"), 20) == (23, 22)


def test_range_optimization():  # test/unit/alignment.jl:81-103
    assert orc.semiglobal("AAAA", "TTTTAAAA", 0.0, rng=(1, 8), max_start_pos=1, min_end_pos=8) == INF


def test_trimming_positions():  # test/unit/trimming.jl:9-67
    read, bc = "AAAAATTTTTCCCCC", "TTTTT"
    for side in (3, 5):
        assert orc.semiglobal(bc, read, 0.0, rng=(1, 15), max_start_pos=100, trim_side=side) == (0.0, 6, 10)
    for side, keep in ((3, (1, 5)), (5, (11, 15))):
        cfg = bdx.DemuxConfig(bc_seqs=[bc], bc_lengths_no_N=[5], ids=["id1"], trim_side=side)
        r = orc.Oracle(cfg).classify_reads([read.encode()])[0]
        assert (r["status"], r["bc1"], r["keep_start"], r["keep_end"]) == (0, 1, keep[0], keep[1])
        assert bdx.output_filename(cfg, 0, 1, 0) == "id1.fastq"


def test_tie_breaking_trim3():  # test/unit/trimming.jl:69-94
    assert orc.semiglobal("ACGT", "ACGTACGT", 0.0, rng=(1, 8), max_start_pos=100, trim_side=3) == (0.0, 5, 8)
    assert orc.semiglobal("AA", "AAAA", 0.0, rng=(1, 4), max_start_pos=100, trim_side=3) == (0.0, 3, 4)


def test_dual_trimming():  # test/unit/trimming.jl:96-130
    cfg = bdx.DemuxConfig(bc_seqs=["TTTTT"], bc_lengths_no_N=[5], ids=["id1"], is_dual=True,
                          bc_seqs2=["GGGGG"], bc_lengths_no_N2=[5], ids2=["id2"], trim_side=5, trim_side2=3)
    r = orc.Oracle(cfg).classify_reads([b"AAAAATTTTTCCCCCGGGGGTTTTT"])[0]
    assert (r["status"], r["bc1"], r["bc2"], r["keep_start"], r["keep_end"]) == (0, 1, 1, 11, 15)
    assert bdx.output_filename(cfg, 0, 1, 1) == "id1.id2.fastq"


def test_score_only_is_float():  # test/unit/trimming.jl:132-143
    s = orc.semiglobal("TTTTT", "AAAAATTTTTCCCCC", 0.0, rng=(1, 15), max_start_pos=100)
    assert isinstance(s, float) and s == 0.0


def test_hamming_vectors():  # test/unit/hamming.jl:1-55
    r = "TTAAAAgg"
    assert orc.hamming("AAAA", r, 0.2, (1, 8), 8, 1) == (0.0, 3, 6)
    r = "TTAATAgg"
    assert orc.hamming("AAAA", r, 0.3, (1, 8), 8, 1) == (0.25, 3, 6)
    assert orc.hamming("AAAA", r, 0.2, (1, 8), 8, 1) == (INF, -1, -1)
    assert orc.hamming("ANNA", r, 0.0, (1, 8), 8, 1) == (0.0, 3, 6)
    assert orc.hamming("AAAA", "TTANAAgg", 0.0, (1, 8), 8, 1) == (INF, -1, -1)
    assert orc.hamming("AAAA", "AAAA", 0.0, (1, 4), 4, 1) == (0.0, 1, 4)
    assert orc.hamming("AA", "AATAA", 0.0, (1, 5), 5, 1, 3) == (0.0, 4, 5)
    assert orc.hamming("AA", "AATAA", 0.0, (1, 5), 5, 1, None) == (0.0, 1, 2)


def test_exact_vectors():  # test/unit/exact.jl:1-89
    assert orc.exact("AAAA", "TTAAAAgg", (1, 8), 8, 1) == (0.0, 3, 6)
    assert orc.exact("AAAA", "TTAATAgg", (1, 8), 8, 1) == (INF, -1, -1)
    assert orc.exact("ANNA", "TTAATAgg", (1, 8), 8, 1) == (INF, -1, -1)
    assert orc.exact("AAAA", "TTANAAgg", (1, 8), 8, 1) == (INF, -1, -1)
    assert orc.exact("AAAA", "AAAA", (1, 4), 4, 1) == (0.0, 1, 4)
    assert orc.exact("AA", "AATAA", (1, 5), 5, 1, 3) == (0.0, 4, 5)
    assert orc.exact("AA", "AATAA", (1, 5), 5, 1, 5) == (0.0, 1, 2)
    assert orc.exact("AA", "AATAA", (1, 5), 5, 4, 3) == (0.0, 4, 5)
    assert orc.exact("AA", "AATAA", (1, 5), 5, 6, 3) == (INF, -1, -1)
    assert orc.exact("AA", "AATAA", (1, 5), 5, 3, 5) == (0.0, 4, 5)


# restatement cross-checks (expected values from the survey's
# independent transcription of classification.jl)
def test_survey_vectors():
    sg = orc.semiglobal
    assert sg("CCT", "TGTAACTGTAATGCTCAAACAGCGT", 0.6, 0, 3, 1, rng=(3, 16), max_start_pos=11,
              min_end_pos=10, norm=3) == INF                                                   # V1
    assert sg("GGNGNA", "GGGA", 0.6, 0, 1, 2, nindel=1, rng=(1, 4), max_start_pos=4, norm=4,
              traceback=True, trim_side=3) == (INF, -1, -1)                                   # V2
    assert sg("ACGTACGTAC", "TTACGTTCGTACGG", 0.2, rng=(1, 14), norm=10, traceback=True) == (0.1, 3, 12)   # V3
    assert sg("ACGTACGTAC", "TTACGTCGTACGG", 0.2, rng=(1, 13), norm=10, trim_side=5) == (0.1, 3, 11)       # V4
    assert sg("ACGTACGTAC", "TTACGTAACGTACGGACGTACGTAC", 0.2, rng=(1, 25), norm=10,
              trim_side=3) == (0.0, 16, 25)                                                   # V5
    assert sg("GGACGT", "ACGTTTTT", 0.34, rng=(1, 8), norm=6, trim_side=3) == (1 / 3, -1, 4)   # V6
    assert sg("ACGTACGTACGT", "GGACGTACTACGTCC", 0.25, 0, 1, 2, rng=(1, 15), norm=12,
              traceback=True) == (2 / 12, 3, 13)                                              # V7
    assert sg("AAAA", "AAAT", 0.3, rng=(1, 4), norm=4, traceback=True) == (0.25, 1, 3)         # V8


def test_survey_v9_order_dependence():
    bcs = ["CGCA", "TAAGGTTTTTT", "CAAACCG", "CGCAC", "TAGAT", "TCGA"]
    cfg = bdx.DemuxConfig(bc_seqs=bcs, bc_lengths_no_N=[len(b) for b in bcs], ids=[str(i) for i in range(6)],
                          max_error_rate=0.5, mismatch=1, indel=2)
    bc, score, delta, s, e = orc.Oracle(cfg).find_best("CCGAGTTTCCGACGTGG", (2, 17), 7, 9)
    assert (bc, score, delta, s, e) == (4, 0.4, INF, -1, -1)


def test_v6_keep_range_is_empty():
    # start label <= 0 from the init column => keep_end = max(1,-1)-1 = 0 => (1, 0)
    cfg = bdx.DemuxConfig(bc_seqs=["GGACGT"], bc_lengths_no_N=[6], ids=["x"], max_error_rate=0.34, trim_side=3)
    r = orc.Oracle(cfg).classify_reads([b"ACGTTTTT"])[0]
    assert (r["status"], r["keep_start"], r["keep_end"]) == (0, 1, 0)


def test_round2():
    assert orc.round2(1 / 3) == 0.33
    assert orc.round2(0.125) == 0.12  # ties-to-even on the scaled value
    assert orc.round2(1 / 6) == 0.17


# Integration expectations with tiny inline inputs (test/integration/*.jl)
def _classify(cfg, reads):
    return orc.Oracle(cfg).classify_reads([r.encode() for r in reads])


def test_integration_n_and_ranges():  # single_barcode.jl:47-97
    cfg = bdx.DemuxConfig(bc_seqs=["ANNC", "TTTT"], bc_lengths_no_N=[2, 4], ids=["BC1", "BC2"],
                          max_error_rate=0.6, nindel=1)
    r = _classify(cfg, ["ATTC", "TTTT", "ATTG", "GGGG"])
    assert list(r["status"]) == [0, 0, 0, 1]
    assert list(r["bc1"]) == [1, 2, 1, 0]


@pytest.mark.parametrize("kw,expect", [
    (dict(ref_search_range=bdx.parse_dynamic_range("1:4")), [0, 1, 1]),
    (dict(ref_search_range=bdx.parse_dynamic_range("5:8")), [1, 0, 1]),
    (dict(barcode_start_range=bdx.parse_dynamic_range("1:1")), [0, 1, 1]),
])
def test_integration_range_restrictions(kw, expect):  # single_barcode.jl:99-173
    cfg = bdx.DemuxConfig(bc_seqs=["AAAA"], bc_lengths_no_N=[4], ids=["BC1"], max_error_rate=0.0, nindel=1, **kw)
    assert list(_classify(cfg, ["AAAATTTT", "TTTTAAAA", "TTAAAATT"])["status"]) == expect


def test_integration_dual():  # dual_barcode.jl:5-162
    cfg = bdx.DemuxConfig(bc_seqs=["AAAA", "CCCC"], bc_lengths_no_N=[4, 4], ids=["ID1_A", "ID1_C"], is_dual=True,
                          bc_seqs2=["TTTT", "GGGG"], bc_lengths_no_N2=[4, 4], ids2=["ID2_T", "ID2_G"],
                          ref_search_range=bdx.parse_dynamic_range("1:4"),
                          ref_search_range2=bdx.parse_dynamic_range("9:12"), max_error_rate=0.0)
    r = _classify(cfg, ["AAAATATATTTTACGT", "CCCCTATAGGGGACGT", "AAAATATAGGGGACGT", "AAAATATAAAAAACGT"])
    assert [(int(a), int(b), int(c)) for a, b, c in zip(r["status"], r["bc1"], r["bc2"])] == \
        [(0, 1, 1), (0, 2, 2), (0, 1, 2), (1, 0, 0)]
    cfg.trim_side, cfg.trim_side2 = 5, 3
    cfg.bc_seqs, cfg.bc_lengths_no_N, cfg.ids = ["AAAA"], [4], ["ID1_A"]
    cfg.bc_seqs2, cfg.bc_lengths_no_N2, cfg.ids2 = ["TTTT"], [4], ["ID2_T"]
    r = _classify(cfg, ["AAAATATATTTTACGT"])[0]
    assert (r["keep_start"], r["keep_end"]) == (5, 8)


def test_integration_hamming_and_exact():  # hamming_demux.jl:30-42, exact_demux.jl:28-42
    reads = ["ACGTAC", "CCCCCC", "ACATAC", "ACGTAG", "ACGGTAC"]
    cfg = bdx.DemuxConfig(bc_seqs=["ACGTAC", "CCCCCC"], bc_lengths_no_N=[6, 6], ids=["BC1", "BC2"],
                          matching_algorithm="hamming", max_error_rate=0.2)
    assert list(_classify(cfg, reads)["bc1"]) == [1, 2, 1, 1, 0]
    cfg.matching_algorithm, cfg.max_error_rate = "exact", 0.0
    assert list(_classify(cfg, reads)["bc1"]) == [1, 2, 0, 0, 0]


def test_integration_summary_counts():  # summary_mode.jl:48-56, summary_distributions.jl
    cfg = bdx.DemuxConfig(bc_seqs=["AAAA", "TTTT"], bc_lengths_no_N=[4, 4], ids=["BC1", "BC2"],
                          max_error_rate=0.2, summary=True)
    r = _classify(cfg, ["AAAA", "TTTT", "GGGG", "AAAT"])
    assert int((r["status"] == 0).sum()) == 2
    cfg = bdx.DemuxConfig(bc_seqs=["AAAA"], bc_lengths_no_N=[4], ids=["BC1"], max_error_rate=0.3, summary=True)
    r = _classify(cfg, ["AAAA", "NAAAA", "NNAAAA", "AAAT", "AAA"])
    p1 = r["passes"][:, 0]
    assert list(p1["start"]) == [1, 2, 3, 1, 1]
    assert list(p1["end"] - p1["start"] + 1) == [4, 4, 4, 3, 3]
