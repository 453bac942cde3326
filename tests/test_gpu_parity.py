"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI,
against the oracle on the same seeded inputs -- bit-exact on every output field."""
import os

import numpy as np
import pytest

import bdx_b200 as bdx
import hostref
import synth
from gpu_common import compare, run_cuda
from bdx_b200 import capi

pytestmark = pytest.mark.gpu
R = bdx.parse_dynamic_range


def _cfg(bcs, **kw):
    lens = [sum(1 for c in b if c != "N") for b in bcs]
    return bdx.DemuxConfig(bc_seqs=bcs, bc_lengths_no_N=lens, ids=[f"bc{i}" for i in range(len(bcs))], **kw)


def _dual(bcs1, bcs2, **kw):
    c = _cfg(bcs1, **kw)
    c.is_dual = True
    c.bc_seqs2 = bcs2
    c.bc_lengths_no_N2 = [sum(1 for ch in b if ch != "N") for b in bcs2]
    c.ids2 = [f"x{i}" for i in range(len(bcs2))]
    return c


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_golden_files_through_cuda(refdata, tmp_path, idx):
    """Config 1: the reference's fixtures -> byte-identical output files via the CUDA path."""
    case = hostref.demo_cases(refdata)[idx]

    engines = []

    def make(cfg):
        for e in engines:
            e.close()
        engines.clear()
        engines.append(capi.Engine(cfg, max_reads=4000))
        return engines[0].classify_reads

    assert hostref.run_case(case, str(tmp_path / "out"), make) == {0: 24, 1: 24, 2: 76}[idx]
    for e in engines:
        e.close()


def test_unit_vectors_through_cuda():
    cfg = _cfg(["TTTTT"], trim_side=3)
    res, _ = compare(cfg, [b"AAAAATTTTTCCCCC"])
    assert (res["keep_start"][0], res["keep_end"][0]) == (1, 5)
    cfg = _cfg(["TTTTT"], trim_side=5)
    res, _ = compare(cfg, [b"AAAAATTTTTCCCCC"])
    assert (res["keep_start"][0], res["keep_end"][0]) == (11, 15)
    cfg = _dual(["TTTTT"], ["GGGGG"], trim_side=5, trim_side2=3)
    res, _ = compare(cfg, [b"AAAAATTTTTCCCCCGGGGGTTTTT"])
    assert tuple(res[0]) == (0, 1, 1, 11, 15)
    cfg = _cfg(["ANNC", "TTTT"], max_error_rate=0.6, nindel=1)
    res, _ = compare(cfg, [b"ATTC", b"TTTT", b"ATTG", b"GGGG"])
    assert list(res["bc1"]) == [1, 2, 1, 0]
    cfg = _cfg(["GGACGT"], max_error_rate=0.34, trim_side=3)   # start label <= 0
    res, _ = compare(cfg, [b"ACGTTTTT"])
    assert tuple(res[0]) == (0, 1, 0, 1, 0)
    bcs = ["CGCA", "TAAGGTTTTTT", "CAAACCG", "CGCAC", "TAGAT", "TCGA"]  # order dependence
    cfg = _cfg(bcs, max_error_rate=0.5, mismatch=1, indel=2, ref_search_range=R("2:17"),
               barcode_start_range=R("1:7"), barcode_end_range=R("9:end"))
    res, _ = compare(cfg, [b"CCGAGTTTCCGACGTGG"])
    assert res["bc1"][0] == 4


def test_edge_cases():
    cfg = _cfg(["ACGTAC", "TTGGCCAA"])
    reads = [b"", b"A", b"ACGTA", b"ACGTAC", b"N" * 20, b"acgtacacgtac", b"TTGGCCAA" * 3, b"ACGTAC" + b"G" * 300]
    for algo in ("semiglobal", "hamming", "exact"):
        cfg.matching_algorithm = algo
        compare(cfg, reads, label=algo)
        compare(cfg, [], label=algo + "-empty")
    cfg.matching_algorithm = "semiglobal"
    # degenerate ranges
    for rs, bs, be in (("10:5", "1:end", "1:end"), ("1:end", "end+3:end", "1:end"), ("1:end", "1:end", "end+1:end"),
                       ("end-3:end", "1:2", "1:end"), ("1:3", "1:end", "end:end")):
        c = _cfg(["ACGTAC", "TTGGCCAA"], ref_search_range=R(rs), barcode_start_range=R(bs), barcode_end_range=R(be),
                 trim_side=3)
        compare(c, reads, label=f"{rs}|{bs}|{be}")


CONFIGS = {
    "default96": dict(n_bc=96, m=(24, 24)),
    "delta": dict(n_bc=96, m=(24, 24), min_delta=0.1),
    "delta_trim_stats": dict(n_bc=96, m=(24, 24), min_delta=0.08, trim_side=3, want_stats=True),
    "delta_close_pairs": dict(n_bc=96, m=(24, 24), min_delta=0.13, close_pairs=True),
    "weighted": dict(n_bc=48, m=(24, 24), max_error_rate=0.25, min_delta=0.15, mismatch=1, indel=2),
    "weighted_nodelta_trim": dict(n_bc=96, m=(24, 24), mismatch=1, indel=2, trim_side=3, want_stats=True),
    "weighted_nodelta": dict(n_bc=60, m=(16, 28), mismatch=2, indel=3, max_error_rate=0.3),
    "mismatch3": dict(n_bc=40, m=(16, 28), max_error_rate=0.3, mismatch=3, indel=1),
    "match1": dict(n_bc=20, m=(12, 20), max_error_rate=0.4, match=1, mismatch=2, indel=2),
    "negmatch": dict(n_bc=20, m=(12, 20), max_error_rate=0.3, match=-1, mismatch=2, indel=2),
    "varlen": dict(n_bc=100, m=(8, 32), max_error_rate=0.22),
    "long": dict(n_bc=33, m=(33, 64), max_error_rate=0.15, min_delta=0.05),
    "verylong": dict(n_bc=5, m=(65, 120), max_error_rate=0.1),
    "nindel": dict(n_bc=64, m=(20, 26), nindel=1, n_frac=0.15, max_error_rate=0.3),
    "nindel2": dict(n_bc=30, m=(20, 26), nindel=2, indel=1, n_frac=0.15, max_error_rate=0.3),
    "nindel_lt": dict(n_bc=30, m=(10, 16), nindel=1, indel=2, n_frac=0.2, max_error_rate=0.5),
    # tiny sets (no filter kernel) and barcodes of 33..64 nt through the shortcut stages
    "tiny_adapter": dict(n_bc=1, m=(33, 33), trim_side=3),
    "tiny_set5": dict(n_bc=5, m=(20, 20), trim_side=5, want_stats=True),
    "tiny_set3_plain": dict(n_bc=3, m=(12, 18)),
    "uniform40_delta": dict(n_bc=50, m=(40, 40), min_delta=0.06),
    "uniform64_trim": dict(n_bc=12, m=(64, 64), trim_side=3, max_error_rate=0.12),
    "trim3": dict(n_bc=96, m=(24, 24), trim_side=3),
    "trim5": dict(n_bc=96, m=(24, 24), trim_side=5, min_delta=0.05),
    "stats": dict(n_bc=50, m=(18, 24), want_stats=True),
    "startrange": dict(n_bc=96, m=(16, 28), ref_search_range=R("1:40"), barcode_start_range=R("1:6"), min_delta=0.1,
                       start_hi=5, max_error_rate=0.3),
    "endrange": dict(n_bc=64, m=(16, 28), ref_search_range=R("end-39:end"), barcode_end_range=R("end-5:end"),
                     trim_side=3),
    "subrange": dict(n_bc=64, m=(20, 24), ref_search_range=R("5:60"), barcode_start_range=R("3:30"),
                     barcode_end_range=R("20:end-10"), trim_side=5, want_stats=True),
    "iupac": dict(n_bc=24, m=(20, 24), alphabet=b"ACGTRYKM", max_error_rate=0.3),
    "big1536": dict(n_bc=1536, m=(24, 24), n_reads=600),
    "hamming": dict(n_bc=96, m=(24, 24), matching_algorithm="hamming", n_frac=0.05),
    "hamming_trim3": dict(n_bc=40, m=(6, 12), matching_algorithm="hamming", trim_side=3, max_error_rate=0.34,
                          min_delta=0.1, barcode_start_range=R("1:40"), barcode_end_range=R("10:end")),
    "hamming_pf": dict(n_bc=96, m=(24, 24), matching_algorithm="hamming", want_stats=True),
    "hamming_pf_trim3": dict(n_bc=64, m=(10, 20), matching_algorithm="hamming", trim_side=3, max_error_rate=0.15,
                             barcode_start_range=R("1:60"), barcode_end_range=R("15:end")),
    "hamming_pf_trim5": dict(n_bc=64, m=(10, 20), matching_algorithm="hamming", trim_side=5,
                             ref_search_range=R("3:90"), barcode_end_range=R("end-80:end")),
    # uniform length, no N in the barcodes: the bit-plane packed scan (hamming.cu)
    "hamming_packed_delta": dict(n_bc=200, m=(24, 24), matching_algorithm="hamming", min_delta=0.1, trim_side=5,
                                 want_stats=True),
    "hamming_packed_m32": dict(n_bc=64, m=(32, 32), matching_algorithm="hamming", max_error_rate=0.22, trim_side=3,
                               barcode_start_range=R("1:100"), barcode_end_range=R("40:end")),
    "hamming_packed_zero": dict(n_bc=96, m=(16, 16), matching_algorithm="hamming", max_error_rate=0.05),
    "hamming_packed_short": dict(n_bc=300, m=(8, 8), matching_algorithm="hamming", max_error_rate=0.4,
                                 ref_search_range=R("2:end-3")),
    "hamming_packed_many": dict(n_bc=1536, m=(24, 24), matching_algorithm="hamming", n_reads=1500),
    "exact": dict(n_bc=96, m=(24, 24), matching_algorithm="exact"),
    "exact_var": dict(n_bc=200, m=(8, 30), matching_algorithm="exact", trim_side=5, barcode_start_range=R("2:70"),
                      barcode_end_range=R("20:end"), want_stats=True),
    "exact_trim3": dict(n_bc=30, m=(4, 8), matching_algorithm="exact", trim_side=3, min_delta=0.5,
                        barcode_start_range=R("1:60"), barcode_end_range=R("8:end")),
}


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_random_parity(name):
    spec = dict(CONFIGS[name])
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    n_bc, (m_lo, m_hi) = spec.pop("n_bc"), spec.pop("m")
    n_reads = spec.pop("n_reads", 3000)
    want_stats = spec.pop("want_stats", False)
    start_hi = spec.pop("start_hi", None)
    bcs = synth.random_barcodes(rng, n_bc, m_lo, m_hi, alphabet=spec.pop("alphabet", b"ACGT"),
                                n_frac=spec.pop("n_frac", 0.0))
    if spec.pop("close_pairs", False):          # barcodes one or two edits apart: runner-ups that make reads ambiguous
        for k in range(0, n_bc - 1, 4):
            bcs[k + 1] = synth.mutate(rng, bcs[k].encode(), int(rng.integers(1, 3))).decode()[:m_lo].ljust(m_lo, "A")
    cfg = _cfg(bcs, **spec)
    reads = synth.random_reads(rng, n_reads, bcs, min_len=60, max_len=160, start_hi=start_hi, lower_prob=0.02)
    reads += [b"", b"ACG"]
    compare(cfg, reads, want_stats=want_stats, label=name)


DUAL = {
    "dual_default": dict(),
    "dual_trim": dict(trim_side=5, trim_side2=3, min_delta=0.1),
    "dual_ranges": dict(ref_search_range=R("1:40"), barcode_start_range=R("1:6"),
                        ref_search_range2=R("end-39:end"), barcode_end_range2=R("end-5:end"), min_delta=0.1,
                        max_error_rate=0.25),
    "dual_stats": dict(want_stats=True, trim_side2=5),
    "dual_adapter_uniform": dict(ref_search_range=R("1:32"), trim_side=5, trim_side2=3, adapter=True, uniform1=True),
    "dual_hamming": dict(matching_algorithm="hamming", trim_side=5),
    "dual_adapter": dict(ref_search_range=R("1:32"), trim_side=5, trim_side2=3, adapter=True),
}


@pytest.mark.parametrize("name", sorted(DUAL))
def test_random_parity_dual(name):
    spec = dict(DUAL[name])
    rng = np.random.default_rng(sum(map(ord, name)))
    want_stats = spec.pop("want_stats", False)
    adapter = spec.pop("adapter", False)
    uniform1 = spec.pop("uniform1", False)
    b1 = synth.random_barcodes(rng, 96 if adapter else 60, 24 if uniform1 else 16, 24 if uniform1 else 28)
    b2 = ["AGATCGGAAGAGCACACGTCTGAACTCCAGTCA"] if adapter else synth.random_barcodes(rng, 70, 16, 28)
    cfg = _dual(b1, b2, **spec)
    reads = synth.random_reads(rng, 3000, b1, barcodes2=b2, min_len=100, max_len=150, start_hi=4,
                               at_end2=not adapter)
    compare(cfg, reads, want_stats=want_stats, label=name)


def test_filter_matches_literal_only():
    """The bit-parallel filter path and the literal-only path must agree read for read."""
    rng = np.random.default_rng(7)
    bcs = synth.random_barcodes(rng, 96, 24)
    reads = synth.random_reads(rng, 4000, bcs, min_len=150)
    for kw in (dict(), dict(min_delta=0.1), dict(trim_side=3), dict(mismatch=2, indel=3, max_error_rate=0.3)):
        cfg = _cfg(bcs, **kw)
        a, _, _, _ = run_cuda(cfg, reads)
        b, _, _, _ = run_cuda(cfg, reads, debug=capi.DEBUG_NO_FILTER)
        assert (a == b).all(), kw


def test_streaming_double_buffer_and_errors():
    rng = np.random.default_rng(11)
    bcs = synth.random_barcodes(rng, 96, 24)
    cfg = _cfg(bcs)
    batches = [synth.random_reads(rng, 500 + 37 * i, bcs, min_len=150) for i in range(5)]
    import orc
    o = orc.Oracle(cfg)
    with capi.Engine(cfg, max_reads=1000, max_bytes=1000 * 160) as eng:
        st = eng.stream
        packed = [bdx.pack_reads(b) for b in batches]
        got = {}
        st.submit(packed[0][0], packed[0][1], tag=100)
        for i in range(1, 5):
            st.submit(packed[i][0], packed[i][1], tag=100 + i)
            tag, res = st.fetch()
            got[tag] = res
        with pytest.raises(capi.BdxError) as ei:    # more than BDX_MAX_IN_FLIGHT batches are refused
            for _ in range(5):
                st.submit(packed[0][0], packed[0][1])
        assert ei.value.code == capi.BDX_ERR_STATE
        while True:
            try:
                tag, res = st.fetch()
                got.setdefault(tag, res)
            except capi.BdxError:
                break
        for i in range(5):
            ref = o.classify_reads(batches[i])
            for f in ("status", "bc1", "bc2", "keep_start", "keep_end"):
                assert (got[100 + i][f] == ref[f]).all()
        big = bdx.pack_reads(synth.random_reads(rng, 1001, bcs, min_len=10))
        with pytest.raises(capi.BdxError) as ei:
            st.submit(big[0], big[1])
        assert ei.value.code == capi.BDX_ERR_TOO_LARGE
        assert st.launch_count > 0


def test_prefilter_edge_cases():
    """Perfect-occurrence prefilter: duplicate barcodes (lowest index wins), barcodes that are
    substrings of longer ones, several lengths, occurrences cut by the search range, and the
    prefilter-off path must agree."""
    rng = np.random.default_rng(21)
    base = synth.random_barcodes(rng, 40, 12, 12)
    bcs = base + [base[3], base[7] + "ACGTAC", "TT" + base[9], base[11][:8]] + synth.random_barcodes(rng, 30, 16, 18)
    reads = synth.random_reads(rng, 3000, bcs, min_len=30, max_len=120, max_edits=2)
    reads += [b"", b"ACGT", bcs[0].encode(), (bcs[43][:-1]).encode(), ("GG" + bcs[41] + "GG").encode()]
    for kw in (dict(), dict(ref_search_range=R("5:60")), dict(ref_search_range=R("end-30:end")),
               dict(max_error_rate=0.0), dict(min_delta=0.1)):
        cfg = _cfg(bcs, **kw)
        res, _ = compare(cfg, reads, label=f"prefilter {kw}")
        off, _, _, _ = run_cuda(cfg, reads, debug=capi.DEBUG_NO_PREFILTER)
        assert (res == off).all(), kw


def test_prefilter_counters():
    rng = np.random.default_rng(22)
    bcs = synth.random_barcodes(rng, 96, 24)
    reads = [(b"ACGTTGCA" * 3 + bcs[i % 96].encode() + b"TTGACCAT" * 4) for i in range(1000)]
    reads += synth.random_reads(rng, 1000, bcs, min_len=100, plant=0.0)
    cfg = _cfg(bcs)
    blob, off = bdx.pack_reads(reads)
    with capi.Engine(cfg, max_reads=len(reads), max_bytes=int(off[-1])) as eng:
        res = eng.classify_packed(blob, off)
        pre, seed, auto = eng.stream.path_counters()
    assert (res["bc1"][:1000] == (np.arange(1000) % 96) + 1).all()
    assert pre >= 1000 and pre + seed + auto == 2000


def test_hash_paths_with_repeats_and_duplicates():
    """:hamming / :exact hash paths on adversarial inputs: duplicated barcodes, barcodes that are
    prefixes of others, reads with the same barcode several times (rightmost / leftmost rules),
    low-complexity reads with many table hits; hash paths off must agree."""
    rng = np.random.default_rng(33)
    base = synth.random_barcodes(rng, 30, 12, 12)
    bcs = base + [base[2], base[5], base[5] + "ACGT", "A" * 12, "AC" * 6] + synth.random_barcodes(rng, 20, 13, 16)
    reads = synth.random_reads(rng, 1500, bcs, min_len=40, max_len=130, max_edits=2)
    for i in range(200):
        b = bcs[int(rng.integers(0, len(bcs)))].encode()
        reads.append(b"GT" * int(rng.integers(0, 5)) + b + b"TGCA" * int(rng.integers(0, 4)) + b + b"CC")
    reads += [b"A" * 100, b"AC" * 60, b"ACGT" * 30, b"", b"ACGTACGTACG"]
    for algo in ("hamming", "exact"):
        for kw in (dict(), dict(trim_side=3), dict(trim_side=5, barcode_start_range=R("1:30")),
                   dict(min_delta=0.2), dict(barcode_end_range=R("30:end"), trim_side=3)):
            cfg = _cfg(bcs, matching_algorithm=algo, max_error_rate=0.1, **kw)
            res, _ = compare(cfg, reads, label=f"{algo} {kw}")
            off, _, _, _ = run_cuda(cfg, reads, debug=capi.DEBUG_NO_PREFILTER)
            assert (res == off).all(), (algo, kw)


@pytest.mark.parametrize("algo", ["semiglobal", "hamming", "exact"])
def test_long_and_ragged_reads(algo):
    """Reads longer than one staging tile (256 columns) and blocks of reads that exceed the
    prefilter's shared-memory stage (48 KB per 256 reads) take the multi-tile / global-memory
    paths; mixed with very short reads."""
    rng = np.random.default_rng(41)
    bcs = synth.random_barcodes(rng, 40, 20, 24)
    reads = synth.random_reads(rng, 700, bcs, min_len=300, max_len=2000, max_edits=3)
    reads += synth.random_reads(rng, 700, bcs, min_len=5, max_len=600, max_edits=3)
    rng.shuffle(reads)
    for kw in (dict(), dict(trim_side=3), dict(ref_search_range=R("end-400:end"), min_delta=0.05),
               dict(barcode_start_range=R("200:end"), trim_side=5)):
        compare(_cfg(bcs, matching_algorithm=algo, **kw), reads, label=f"long {algo} {kw}")


@pytest.mark.parametrize("streams_per_device", [1, 3])
def test_pool_dispatcher_order_and_stats(streams_per_device):
    """bdx_pool: batches dealt round-robin over streams (and GPUs when there are several) come back in
    submission order with the results of a single stream; the summed counters equal one stream's."""
    rng = np.random.default_rng(31)
    bcs = synth.random_barcodes(rng, 48, 20)
    cfg = _cfg(bcs, trim_side=5, summary=True)
    batches = [synth.random_reads(rng, 300 + 41 * i, bcs, min_len=100) for i in range(11)]
    packed = [bdx.pack_reads(b) for b in batches]
    config = capi.Config(cfg)
    n_dev = capi.load_library().bdx_device_count()
    with capi.Engine(cfg, max_reads=1000, max_bytes=1000 * 110) as eng:
        want = [eng.classify_packed(*p) for p in packed]
        want_stats = eng.stream.stats()
    with capi.Pool(config, list(range(n_dev)), streams_per_device, max_reads=1000, max_bytes=1000 * 110) as pool:
        got, nxt = [], 0
        while len(got) < len(packed):
            while nxt < len(packed) and pool.try_submit(packed[nxt][0], packed[nxt][1], tag=1000 + nxt):
                nxt += 1
            tag, res = pool.fetch()
            assert tag == 1000 + len(got)
            got.append(res)
        assert pool.in_flight == 0
        for g, w in zip(got, want):
            for f in ("status", "bc1", "bc2", "keep_start", "keep_end"):
                assert (g[f] == w[f]).all()
        assert (pool.stats() == want_stats).all()
        with pytest.raises(capi.BdxError) as ei:
            pool.fetch()
        assert ei.value.code == capi.BDX_ERR_STATE
    config.close()


@pytest.mark.parametrize("algo", ["semiglobal", "hamming"])
@pytest.mark.parametrize("rng_expr,at_end", [("1:120", False), ("end-119:end", True), ("30:180", False)])
def test_long_reads_short_search_ranges(algo, rng_expr, at_end):
    """Nanopore-style input: reads of thousands of bases, barcode searched in a short range near one end.  The
    shortcut kernels stage only the columns of the search range, so these reads take them too."""
    rng = np.random.default_rng(len(rng_expr) * 31 + len(algo))
    bcs = synth.random_barcodes(rng, 96, 24)
    reads = []
    for r in synth.random_reads(rng, 1200, bcs, min_len=100, max_len=118, start_hi=60):
        pad = bytes(synth.BASES[rng.integers(0, 4, int(rng.integers(500, 4000)))])
        reads.append(pad + r if at_end else r + pad)
    for kw in (dict(), dict(trim_side=3 if at_end else 5, summary=True), dict(min_delta=0.09)):
        cfg = _cfg(bcs, matching_algorithm=algo, ref_search_range=R(rng_expr), **kw)
        compare(cfg, reads, want_stats=bool(kw.get("summary")), label=f"{algo} {rng_expr} {kw}")
    # the reads really went through the shortcut stages
    cfg = _cfg(bcs, matching_algorithm=algo, ref_search_range=R(rng_expr))
    blob, off = bdx.pack_reads(reads)
    with capi.Engine(cfg, max_reads=len(reads), max_bytes=int(off[-1])) as eng:
        eng.classify_packed(blob, off)
        pre, seed, auto = eng.stream.path_counters()
        if algo == "semiglobal":
            assert pre + seed > 0.4 * len(reads), (pre, seed, auto)   # ("30:180" cuts half of the planted barcodes)


@pytest.mark.parametrize("kw", [dict(), dict(trim_side=5, summary=True), dict(min_delta=0.1),
                                dict(matching_algorithm="hamming")])
def test_graph_replay_of_small_batches(kw):
    """Full-size chunks on a stream replay a captured CUDA graph from the slot's second use on: 24 chunks of 512
    reads through 4 slots must equal one big batch (never captured) read for read, stats included."""
    rng = np.random.default_rng(77)
    bcs = synth.random_barcodes(rng, 96, 24)
    cfg = _cfg(bcs, **kw)
    B, nb = 512, 24
    reads = synth.random_reads(rng, B * nb, bcs, min_len=150)
    blob, off = bdx.pack_reads(reads)
    with capi.Engine(cfg, max_reads=B * nb + 1, max_bytes=int(off[-1]) + 16) as big:
        want = big.classify_packed(blob, off)
        want_stats = big.stream.stats() if kw.get("summary") else None
    sub_off = np.ascontiguousarray(off[:B + 1])
    with capi.Engine(cfg, max_reads=B, max_bytes=B * 150) as eng:
        st, got, q = eng.stream, [], 0
        l0 = st.launch_count
        for k in range(nb):
            st.submit(blob[k * B * 150:(k + 1) * B * 150], sub_off, tag=k)
            q += 1
            if q == 4:
                got.append(st.fetch()[1])
                q -= 1
        while q:
            got.append(st.fetch()[1])
            q -= 1
        assert st.launch_count - l0 >= nb * 2
        res = np.concatenate(got)
        for f in ("status", "bc1", "bc2", "keep_start", "keep_end"):
            assert (res[f] == want[f]).all(), f
        if want_stats is not None:
            assert (st.stats() == want_stats).all()


def test_bad_offsets_are_refused_before_anything_runs():
    """A batch whose offsets decrease would be a negative read length on the device: refused with BDX_ERR_INVALID
    on every host entry (bdx_submit, bdx_submit_pinned via the same check, bdx_commit)."""
    rng = np.random.default_rng(3)
    bcs = synth.random_barcodes(rng, 24, 20)
    cfg = _cfg(bcs)
    blob = np.frombuffer(b"ACGT" * 100, dtype=np.uint8).copy()
    with capi.Engine(cfg, max_reads=16, max_bytes=4096) as eng:
        for off in ([0, 100, 50, 400], [0, 100, 200, 150], [5, 100, 200, 400]):
            with pytest.raises(capi.BdxError) as ei:
                eng.stream.submit(blob, np.asarray(off, np.int32))
            assert ei.value.code == capi.BDX_ERR_INVALID, off
        res = eng.classify_packed(blob, np.asarray([0, 100, 100, 400], np.int64))     # an empty read is fine
        assert len(res) == 3


@pytest.mark.parametrize("algo", ["semiglobal", "exact"])
def test_thousands_of_short_barcodes(algo):
    """~ 9 000 short barcodes: the prefilter's hash table + bitmap would need more shared memory than a block may
    have (the launch used to fail); the stage is dropped for such sets and the results stay the oracle's."""
    rng = np.random.default_rng(77)
    bcs = sorted(set(synth.random_barcodes(rng, 9000, 10, 12)))
    cfg = _cfg(bcs, matching_algorithm=algo, max_error_rate=0.1)
    reads = synth.random_reads(rng, 400, bcs, min_len=40, max_len=60, max_edits=1)
    compare(cfg, reads, label=f"9000 barcodes {algo}")
