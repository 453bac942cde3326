"""CPU-side checks of the C-ABI library and the host mirror (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import bdx_b200 as bdx
from bdx_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    capi.build_library()
    return capi.load_library()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "bdx.h")).read()
    declared = set(re.findall(r"\b(bdx_[a-z_0-9]+)\s*\(", header))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"libbdx.so does not export {name}"
    assert lib.bdx_abi_version() == capi.ABI_VERSION


def test_struct_layouts_match_header():
    # sizes the C compiler gives include/bdx.h (LP64)
    assert C.sizeof(capi.Range) == 24
    assert C.sizeof(capi.BarcodeSet) == 8 + 3 * 8 + 3 * 24
    assert C.sizeof(capi.Params) == 8 + 16 + 32 + 16 + 2 * C.sizeof(capi.BarcodeSet)
    assert bdx.RESULT_DTYPE.itemsize == 20 and bdx.DETAIL_DTYPE.itemsize == 24


def _cfg(**kw):
    base = dict(bc_seqs=["ACGTACGTAC", "TTTTGGGGCC"], bc_lengths_no_N=[10, 10], ids=["a", "b"])
    base.update(kw)
    return bdx.DemuxConfig(**base)


def test_config_create_and_stats_layout(lib):
    c = capi.Config(_cfg(summary=True))
    L = c.layout
    assert (L.b1, L.b2) == (2, 0)
    assert L.dist_bins == 3            # floor(0.2 * 10) + 1
    assert L.pos_bias == 10 and L.len_bins == 1036 and L.pos_bins == 1036
    assert L.sample_off == 4 and L.pos_off[0] == 4 + 3
    assert L.total_len == L.dist_off[1] + L.dist_bins
    c.close()


@pytest.mark.parametrize("kw,msg", [
    (dict(trim_side=4), "trim_side"),
    (dict(bc_seqs=["ACGT", ""], bc_lengths_no_N=[4, 0]), "empty barcode"),
    (dict(indel=0), "zero gap cost"),
    (dict(mismatch=1 << 21), "cost magnitude"),
    (dict(bc_seqs=["A" * 300, "C"], bc_lengths_no_N=[300, 1]), "longer than 256"),
])
def test_config_validation_errors(lib, kw, msg):
    with pytest.raises(capi.BdxError) as ei:
        capi.Config(_cfg(**kw))
    assert ei.value.code == capi.BDX_ERR_INVALID and msg in str(ei.value)


def test_no_cpu_fallback(lib, refdata, tmp_path):
    if lib.bdx_device_count() > 0:
        pytest.skip("a CUDA device is present")
    c = capi.Config(_cfg())
    with pytest.raises(capi.BdxError) as ei:
        capi.Stream(c)
    assert ei.value.code == capi.BDX_ERR_CUDA and "no CPU fallback" in str(ei.value)
    with pytest.raises(capi.BdxError):
        fq = os.path.join(refdata, "FASTQ_files", "demo1_R1", "demo-001_R1.fastq")
        bdx.execute_demultiplexing(fq, os.path.join(refdata, "reference_files", "demo1.tsv"), str(tmp_path / "o"))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "biodemux.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in text and "bdx_oracle" not in text and "import orc" not in text, f


def test_barcode_table_loader(refdata):
    seqs, lens, ids = bdx.preprocess_bc_file(os.path.join(refdata, "reference_files", "demo1.tsv"), False, False)
    assert len(seqs) == 24 and all(len(s) == 24 for s in seqs) and lens == [24] * 24
    assert ids[0] == "barcode_001" and seqs[0] == "TGACAGGTAACTACTGACCGTCCT"
    s2, _, _ = bdx.preprocess_bc_file(os.path.join(refdata, "reference_files", "demo2.csv"), True, True)
    comp = {"A": "T", "T": "A", "G": "C", "C": "G"}
    full, ann = "TGGATGGGAGAGATCGACCTGGTCCCCCC", "XXXBBBBBBBBBBBBBBBBBBBBBBBBXX"   # demo2.csv row 1
    kept = "".join(c for c, a in zip(full, ann) if a == "B")
    assert s2[0] == "".join(comp[c] for c in kept)[::-1]


def test_fasta_loader_and_filenames(tmp_path):
    p = tmp_path / "bc.fasta"
    p.write_text(">BC1 some description\nANNC\n>BC2\nuuTT\n")
    seqs, lens, ids = bdx.preprocess_bc_file(str(p), False, False)
    assert (seqs, lens, ids) == (["ANNC", "TTTT"], [2, 4], ["BC1", "BC2"])
    cfg = bdx.DemuxConfig(bc_seqs=seqs, bc_lengths_no_N=lens, ids=ids, gzip_output=True)
    assert bdx.output_filename(cfg, bdx.MATCH, 2, 0) == "BC2.fastq.gz"
    assert bdx.output_filename(cfg, bdx.UNKNOWN, 0, 0) == "unknown.fastq.gz"
    assert bdx.output_filename(cfg, bdx.AMBIGUOUS, 0, 0) == "ambiguous_classification.fastq.gz"
    bad = tmp_path / "bad.csv"
    bad.write_text("ID,Full_seq,Full_annotation\nx,ACGT,BBB\n")
    with pytest.raises(ValueError, match="Length mismatch"):
        bdx.preprocess_bc_file(str(bad), False, False)


def test_build_config_validation(refdata):
    bc = os.path.join(refdata, "reference_files", "demo1.tsv")
    with pytest.raises(ValueError, match="trim_side must be 3 or 5"):
        import hostref
        hostref.build_cfg(bc, ["x.fastq"], trim_side=4)
    import hostref
    cfg = hostref.build_cfg(bc, ["x.fastq.gz"])
    assert cfg.gzip_output is True          # core.jl:316
    with pytest.raises(ValueError, match="Invalid range format"):
        hostref.build_cfg(bc, ["x.fastq"], ref_search_range="1-10")


def test_stats_roundtrip_helpers():
    from bdx_b200.stats import julia_round2
    assert julia_round2(1 / 3) == 0.33 and julia_round2(0.125) == 0.12 and julia_round2(1 / 6) == 0.17


@pytest.mark.parametrize("dual,nindel", [(False, None), (True, None), (False, 2)])
def test_stats_entries_bridge(dual, nindel):
    """bdx_stats_entries (C report bridge) == the Python conversion of the same counter buffer,
    == DemuxStats built from per-pass records the way determine_filename_and_stats does."""
    from bdx_b200.stats import stats_from_counters, stats_from_entries, stats_from_passes
    rng = np.random.default_rng(11 + dual)
    bcs = ["ACGTNCGTAC", "TTGGCCAATTGG", "ANNNNNNNNT", "GATTACAGATTACAGA"]
    lens = [sum(1 for c in s if c != "N") for s in bcs]
    kw = dict(bc_seqs=bcs, bc_lengths_no_N=lens, ids=list("abcd"), summary=True, nindel=nindel, max_error_rate=0.5)
    if dual:
        kw.update(is_dual=True, bc_seqs2=bcs[:3], bc_lengths_no_N2=lens[:3], ids2=list("xyz"))
    cfg = bdx.DemuxConfig(**kw)
    config = capi.Config(cfg)
    lay = config.layout
    from bdx_b200.stats import pass_norms
    norms = [pass_norms(cfg, False), pass_norms(cfg, True) if dual else []]
    n = 4000
    buf = np.zeros(lay.total_len, dtype=np.int64)
    status, bc1, bc2, passes = [], [], [], []
    for _ in range(n):
        rec = []
        ok = True
        for p in range(2 if dual else 1):
            if not ok or rng.random() < 0.2:
                rec.append((1, 0, -1, -1, float("inf")))
                ok = False
                continue
            nb = len(norms[p])
            b = int(rng.integers(1, nb + 1))
            dist = int(rng.integers(0, int(0.5 * norms[p][b - 1]) + 1))
            s = int(rng.integers(-3, 60))
            e = s + int(rng.integers(0, 20))
            rec.append((0, b, s, e, dist / norms[p][b - 1]))
            bsz = (lay.pos_bins, lay.len_bins, lay.dist_bins)
            for row in (0, b):
                buf[lay.pos_off[p] + row * bsz[0] + s + lay.pos_bias] += 1
                buf[lay.len_off[p] + row * bsz[1] + (e - s + 1)] += 1
                buf[lay.dist_off[p] + row * bsz[2] + dist + lay.dist_bias] += 1
        while len(rec) < 2:
            rec.append((-1, 0, -1, -1, float("inf")))
        passes.append(rec)
        buf[0] += 1
        if ok:
            status.append(0); bc1.append(rec[0][1]); bc2.append(rec[1][1] if dual else 0)
            buf[1] += 1
            buf[lay.sample_off + bc1[-1] * (lay.b2 + 1) + bc2[-1]] += 1
        else:
            status.append(1); bc1.append(0); bc2.append(0)
            buf[2] += 1
    ent = capi.stats_entries(config, buf)
    got = stats_from_entries(buf, lay, ent)
    assert got == stats_from_counters(buf, lay, cfg)
    assert got == stats_from_passes(status, bc1, bc2, passes, cfg)
    config.close()


def test_stats_overflow_merge():
    """Exact overflow records (bdx_stats_overflow_fetch) land in the same Dicts as histogram bins."""
    from bdx_b200.stats import DemuxStats, merge_overflow
    ovf = np.array([(1, 3, 4000, 24), (1, 3, 4000, 25), (2, 1, 70000, 16)], dtype=capi.STATS_OVERFLOW_DTYPE)
    st = merge_overflow(DemuxStats(), ovf)
    assert st.bc1_pos_counts == {4000: 2} and st.bc1_len_counts == {24: 1, 25: 1}
    assert st.bc1_per_bc_pos_counts == {3: {4000: 2}} and st.bc1_per_bc_len_counts == {3: {24: 1, 25: 1}}
    assert st.bc2_pos_counts == {70000: 1} and st.bc2_per_bc_len_counts == {1: {16: 1}}
    assert merge_overflow(DemuxStats(), None) == DemuxStats()


def test_pack4_matches_numpy_reference():
    """bdx_pack_reads4 (AVX2 and scalar paths) against a numpy restatement: code table = distinct barcode bytes of
    both sets in order of appearance, nibble k = code of byte k, low nibble first; no CUDA needed."""
    rng = np.random.default_rng(5)
    for bcs, bcs2 in ((["ACGTACGT", "TTGGCCAA"], None), (["ACGTN", "RYKM"], ["ACGU", "acgt"]), (["AC.T", "A-GT"], None)):
        cfg = bdx.DemuxConfig(bc_seqs=bcs, bc_lengths_no_N=[len(b) for b in bcs], ids=[str(i) for i in range(len(bcs))])
        if bcs2:
            cfg.is_dual, cfg.bc_seqs2, cfg.bc_lengths_no_N2 = True, bcs2, [len(b) for b in bcs2]
            cfg.ids2 = [str(i) for i in range(len(bcs2))]
        config = capi.Config(cfg)
        table = config.code_table()
        order = []
        for b in "".join(bcs + (bcs2 or [])).encode("latin-1"):
            if b not in order:
                order.append(b)
        want_table = np.zeros(256, np.uint8)
        for k, b in enumerate(order):
            want_table[b] = k + 1
        assert (table == want_table).all()
        for n in (0, 1, 2, 63, 64, 65, 1000, 4097):
            seq = rng.choice(np.frombuffer(b"ACGTNacgtRYKMU.-\n\x00\xff", dtype=np.uint8), n).astype(np.uint8)
            got = config.pack4(seq)
            codes = want_table[seq]
            if n % 2:
                codes = np.append(codes, 0)
            want = (codes[0::2] | (codes[1::2] << 4)).astype(np.uint8)
            assert (got == want).all(), (bcs, n)
        config.close()
    many = ["".join(chr(65 + k) for k in range(16))]          # 16 distinct barcode bytes: no packed input
    config = capi.Config(bdx.DemuxConfig(bc_seqs=many, bc_lengths_no_N=[16], ids=["x"]))
    with pytest.raises(capi.BdxError):
        config.code_table()
    config.close()


@pytest.mark.parametrize("seed", range(6))
def test_table_builders_accept_any_set_shape(lib, seed):
    """bdx_config_create builds every kernel family's tables on the host (csrc/tables.cu) -- no GPU needed.  Random
    set shapes (1 .. 3000 barcodes, 4 .. 64 nt, uniform and mixed lengths, 2 .. 6 distinct bytes, wildcards, all
    three algorithms, constrained ranges, odd thresholds) must give a config or a clean BDX_ERR_* -- never a
    crash, and the same answer when asked twice."""
    import synth
    rng = np.random.default_rng(5150 + seed)
    R = bdx.parse_dynamic_range
    for _ in range(40):
        n_bc = int(rng.choice([1, 2, 7, 8, 33, 96, 384, 1536, 3000]))
        m_lo = int(rng.choice([4, 6, 8, 12, 16, 24, 32, 33, 48, 64]))
        m_hi = min(64, m_lo + int(rng.choice([0, 0, 1, 4, 12])))
        alphabet = [b"ACGT", b"AC", b"ACGTN", b"ACGTacgt", b"ACGTRY"][int(rng.integers(0, 5))]
        bcs = synth.random_barcodes(rng, n_bc, m_lo, m_hi, alphabet=alphabet)
        kw = dict(bc_seqs=bcs, bc_lengths_no_N=[max(1, len(b) - b.count("N")) for b in bcs],
                  ids=[f"b{i}" for i in range(n_bc)],
                  max_error_rate=float(rng.choice([0.0, 0.05, 0.2, 0.34, 0.5, 1.0])),
                  min_delta=float(rng.choice([0.0, 0.0, 0.1])),
                  matching_algorithm=str(rng.choice(["semiglobal", "semiglobal", "hamming", "exact"])),
                  trim_side=[None, None, 3, 5][int(rng.integers(0, 4))], summary=bool(rng.random() < 0.2))
        shape = int(rng.integers(0, 4))
        if shape == 1:
            kw.update(ref_search_range=R("1:40"), barcode_start_range=R(f"1:{int(rng.integers(1, 12))}"))
        elif shape == 2:
            kw.update(ref_search_range=R("end-49:end"), barcode_end_range=R(f"end-{int(rng.integers(0, 12))}:end"))
        elif shape == 3:
            kw.update(ref_search_range=R(f"{int(rng.integers(1, 9))}:{int(rng.integers(60, 200))}"))
        if rng.random() < 0.2:
            kw.update(indel=int(rng.choice([1, 2, 3])), mismatch=int(rng.choice([1, 2])), nindel=int(rng.choice([1, 2])))
        outcomes = []
        for _twice in range(2):
            try:
                c = capi.Config(bdx.DemuxConfig(**kw))
                outcomes.append(("ok", c.layout.total_len))
                c.close()
            except capi.BdxError as e:
                outcomes.append(("err", e.code))
        assert outcomes[0] == outcomes[1], (kw, outcomes)


def _described(cfg, pass_=0):
    c = capi.Config(cfg)
    try:
        return {ln.split(":", 1)[0]: dict(kv.split("=") for kv in ln.split(":", 1)[1].split())
                for ln in c.describe(pass_).splitlines()}
    finally:
        c.close()


def test_level_choices_for_the_benchmark_configs(lib):
    """bdx_config_describe: the stage / level choices DESIGN.md sections 4.2b and 4.2c quote for BASELINE.json's configs
    are what the table builders (csrc/tables.cu) really make -- checked on the host, no GPU."""
    import synth
    rng = np.random.default_rng(1)
    R = bdx.parse_dynamic_range

    def cfg(b1, b2=None, **kw):
        c = bdx.DemuxConfig(bc_seqs=b1, bc_lengths_no_N=[len(x) for x in b1], ids=[f"a{i}" for i in range(len(b1))], **kw)
        if b2:
            c.is_dual = True
            c.bc_seqs2, c.bc_lengths_no_N2, c.ids2 = b2, [len(x) for x in b2], [f"b{i}" for i in range(len(b2))]
        return c

    b96 = synth.random_barcodes(rng, 96, 24, 24)
    # config 2: prefilter, ONE k_seed level (K = 2, 8-mers), complete k_seed_var level of 4- and 5-mers, filter inside the scan
    d = _described(cfg(b96))
    assert "prefilter" in d and d["k_seed level 1"]["K"] == "2" and d["k_seed level 1"]["q"] == "8"
    assert "k_seed level 2" not in d
    tail = d["k_seed_var level 2"]
    assert (tail["complete"], tail["q"], tail["q2"], tail["K"], tail["qgram_filter"]) == ("1", "4", "5", "4..4", "1")
    assert int(tail["group_reads"]) % 32 == 0
    assert _described(cfg(b96), 1) == {}
    # ... with trimming the complete level cannot take the tail: k_seed keeps its second level (K = 3, 6-mers)
    d = _described(cfg(b96, trim_side=5))
    assert d["k_seed level 2"]["K"] == "3" and d["k_seed level 2"]["q"] == "6"
    # config 3: no k_seed (lengths differ); k_seed_var level 1 with 5-mers at depths 2..4, then the complete level;
    # constrained geometry => the filter runs over the finished list (mode 2); a constrained START keeps >= m + 2 rows
    b1, b2 = synth.random_barcodes(rng, 384, 16, 28), synth.random_barcodes(rng, 384, 16, 28)
    c3 = cfg(b1, b2, ref_search_range=R("1:40"), barcode_start_range=R("1:6"), ref_search_range2=R("end-39:end"),
             barcode_end_range2=R("end-5:end"), min_delta=0.1)
    for p in (0, 1):
        d = _described(c3, p)
        assert "k_seed level 1" not in d
        assert d["k_seed_var level 1"]["q"] == "5" and d["k_seed_var level 1"]["K"] == "2..4"
        assert d["k_seed_var level 1"]["qgram_filter"] == "2" and d["k_seed_var level 2"]["complete"] == "1"
    assert int(_described(c3, 0)["k_seed_var level 1"]["hit_rows"]) >= 28 + 2
    # config 5: 1 536 barcodes keep both k_seed levels and get no complete level (its chance hits cost more than k_filter)
    d = _described(cfg(synth.random_barcodes(rng, 1536, 24, 24)))
    assert d["k_seed level 1"]["q"] == "12" and d["k_seed level 2"]["q"] == "8"
    assert not any(v.get("complete") == "1" for v in d.values())
