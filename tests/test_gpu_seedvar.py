"""k_seed_var (csrc/seed_var.cu): seed-and-verify for barcode sets of different lengths and for constrained
barcode start / end geometries (BASELINE.json config 3) -- CUDA path vs oracle, bit-exact, and proof that the
shortcut stage is the one that ran."""
import numpy as np
import pytest

import bdx_b200 as bdx
import orc
import synth
from gpu_common import compare
from bdx_b200 import capi

pytestmark = pytest.mark.gpu
R = bdx.parse_dynamic_range


def _cfg(b1, b2=None, **kw):
    c = bdx.DemuxConfig(bc_seqs=b1, bc_lengths_no_N=[len(x) for x in b1], ids=[f"a{i}" for i in range(len(b1))], **kw)
    if b2:
        c.is_dual = True
        c.bc_seqs2, c.bc_lengths_no_N2, c.ids2 = b2, [len(x) for x in b2], [f"b{i}" for i in range(len(b2))]
    return c


CONFIG3 = dict(ref_search_range=R("1:40"), barcode_start_range=R("1:6"), ref_search_range2=R("end-39:end"),
               barcode_end_range2=R("end-5:end"), min_delta=0.1)


def _classify_counted(cfg, reads):
    blob, off = bdx.pack_reads(reads)
    with capi.Engine(cfg, max_reads=len(reads), max_bytes=int(off[-1])) as eng:
        eng.stream.path_counters(reset=True)
        got = eng.classify_packed(blob, off)
        _, seed_reads, auto_reads = eng.stream.path_counters()
    return got, seed_reads, auto_reads, blob, off


def test_config3_shape_384x384():
    """SURVEY.md section 8d config 3 at its stated set sizes: dual 384 x 384, lengths 16..28, start / end
    constrained search ranges, min_delta 0.1; 30 000 reads against the multi-threaded oracle."""
    rng = np.random.default_rng(33)
    b1, b2 = synth.random_barcodes(rng, 384, 16, 28), synth.random_barcodes(rng, 384, 16, 28)
    cfg = _cfg(b1, b2, **CONFIG3)
    reads = synth.random_reads(rng, 30000, b1, barcodes2=b2, min_len=150, start_hi=4, max_edits=5, n_prob=0.005)
    got, seed_reads, auto_reads, blob, off = _classify_counted(cfg, reads)
    want = orc.Oracle(cfg).classify_mt(blob, off)
    for f in ("status", "bc1", "bc2", "keep_start", "keep_end"):
        bad = np.nonzero(got[f] != want[f])[0]
        assert bad.size == 0, f"{f} differs at read {bad[0]}: {got[bad[0]]} vs {want[bad[0]]} ({reads[bad[0]]!r})"
    # the seed stage has to carry the bulk: most reads of both passes resolved without the full automaton
    assert seed_reads > 0.45 * 2 * len(reads), (seed_reads, auto_reads)
    assert auto_reads < 0.55 * 2 * len(reads), (seed_reads, auto_reads)


VARIANTS = {
    # name: (n1, (m_lo, m_hi), options, read kwargs)
    "start_only": (200, (16, 28), dict(ref_search_range=R("1:40"), barcode_start_range=R("1:6"), min_delta=0.1),
                   dict(start_hi=5)),
    "start_nodelta": (384, (12, 30), dict(ref_search_range=R("1:50"), barcode_start_range=R("1:10")), dict(start_hi=9)),
    "start_loose_edge": (96, (16, 24), dict(ref_search_range=R("1:45"), barcode_start_range=R("1:3"), min_delta=0.05,
                                             max_error_rate=0.3), dict(start_hi=9)),    # plants beyond max_start_pos
    "end_only": (200, (16, 28), dict(ref_search_range=R("end-39:end"), barcode_end_range=R("end-5:end"),
                                     min_delta=0.1), dict(at_end=True)),
    "end_nodelta": (150, (14, 26), dict(ref_search_range=R("end-49:end"), barcode_end_range=R("end-8:end")),
                    dict(at_end=True)),
    "both_bounds": (120, (16, 24), dict(ref_search_range=R("3:60"), barcode_start_range=R("3:12"),
                                        barcode_end_range=R("20:end"), min_delta=0.08), dict(start_hi=12)),
    "varlen_default_geometry": (300, (14, 30), dict(min_delta=0.1), dict()),
    "varlen_default_nodelta": (96, (10, 32), dict(max_error_rate=0.25), dict()),
    "uniform_start": (96, (24, 24), dict(ref_search_range=R("1:40"), barcode_start_range=R("1:5"), min_delta=0.1),
                      dict(start_hi=6)),
    "tight_threshold": (200, (16, 28), dict(ref_search_range=R("1:40"), barcode_start_range=R("1:6"),
                                            max_error_rate=0.07, min_delta=0.02), dict(start_hi=5)),
    "zero_threshold": (200, (16, 28), dict(ref_search_range=R("1:40"), barcode_start_range=R("1:6"),
                                           max_error_rate=0.0), dict(start_hi=5)),
    "long64": (60, (33, 64), dict(ref_search_range=R("1:90"), barcode_start_range=R("1:8"), min_delta=0.04,
                                  max_error_rate=0.15), dict(start_hi=7)),
    "close_pairs": (200, (18, 26), dict(ref_search_range=R("1:40"), barcode_start_range=R("1:6"), min_delta=0.1),
                    dict(start_hi=5, close=True)),
}


def _reads(rng, n, bcs, start_hi=None, at_end=False, close=False, min_len=100, max_len=150):
    if at_end:        # barcode planted 0..5 bases before the read end
        out = []
        for r in synth.random_reads(rng, n, bcs, min_len=min_len, max_len=max_len, plant=0.0, n_prob=0.01):
            r = bytearray(r)
            if rng.random() < 0.9:
                bc = bcs[int(rng.integers(0, len(bcs)))].encode()
                mb = synth.mutate(rng, bc, int(rng.choice([0, 0, 0, 1, 1, 2, 3, 4, 5])))
                gap = int(rng.integers(0, 9))
                st = max(len(r) - gap - len(mb), 0)
                r[st:st + len(mb)] = mb
                r = r[:max(len(r), 1)]
            out.append(bytes(r))
        return out
    return synth.random_reads(rng, n, bcs, min_len=min_len, max_len=max_len, start_hi=start_hi, max_edits=5,
                              n_prob=0.02, lower_prob=0.02)


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_seed_var_variants(name):
    n1, (m_lo, m_hi), opts, rk = VARIANTS[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    bcs = synth.random_barcodes(rng, n1, m_lo, m_hi)
    rk = dict(rk)
    if rk.pop("close", False):           # barcodes one or two edits apart: runner-ups inside min_delta
        for k in range(0, n1 - 1, 3):
            bcs[k + 1] = synth.mutate(rng, bcs[k].encode(), int(rng.integers(1, 3))).decode()
        bcs[5] = bcs[2]                  # identical sequences: delta 0
    cfg = _cfg(bcs, **opts)
    reads = _reads(rng, 6000, bcs, **rk) + [b"", b"ACG", bcs[0].encode(), b"N" * 50]
    compare(cfg, reads, label=name)


def _fuzz_case(seed):
    rng = np.random.default_rng(seed)
    n1 = int(rng.choice([8, 30, 96, 200, 384, 600]))
    m_lo = int(rng.choice([8, 12, 16, 20, 24]))
    m_hi = m_lo + int(rng.choice([0, 2, 6, 12]))
    bcs = synth.random_barcodes(rng, n1, m_lo, m_hi)
    if rng.random() < 0.3:
        for _ in range(3):
            bcs[int(rng.integers(0, n1))] = bcs[int(rng.integers(0, n1))]
    kw = dict(max_error_rate=float(rng.choice([0.0, 0.1, 0.2, 0.2, 0.25, 0.3])),
              min_delta=float(rng.choice([0.0, 0.05, 0.1, 0.1, 0.15])))
    shape = int(rng.integers(0, 5))
    a, w = int(rng.integers(1, 8)), int(rng.integers(30, 70))
    start_hi, at_end = None, False
    if shape == 0:       # constrained start near the read start
        kw.update(ref_search_range=R(f"{a}:{a + w}"), barcode_start_range=R(f"{a}:{a + int(rng.integers(0, 12))}"))
        start_hi = a + 10
    elif shape == 1:     # constrained end near the read end
        kw.update(ref_search_range=R(f"end-{w}:end"), barcode_end_range=R(f"end-{int(rng.integers(0, 12))}:end"))
        at_end = True
    elif shape == 2:     # both
        kw.update(ref_search_range=R(f"{a}:{a + w}"), barcode_start_range=R(f"1:{a + int(rng.integers(2, 15))}"),
                  barcode_end_range=R(f"{a + int(rng.integers(5, 30))}:end"))
        start_hi = a + 12
    elif shape == 3:     # sub-range only (default start / end ranges)
        kw.update(ref_search_range=R(f"{a}:{a + w}"))
        start_hi = a + w - m_lo
    cfg = _cfg(bcs, **kw)
    reads = _reads(rng, 1500, bcs, start_hi=start_hi, at_end=at_end, min_len=int(rng.choice([40, 100])),
                   max_len=int(rng.choice([100, 150, 220])))
    reads += [b"", b"A", bcs[0].encode(), b"N" * 60, bcs[-1].encode() * 3]
    return cfg, reads


@pytest.mark.parametrize("seed", range(40))
def test_seed_var_fuzz(seed):
    cfg, reads = _fuzz_case(7000 + seed)
    compare(cfg, reads, label=f"seedvar fuzz {seed}")


@pytest.mark.parametrize("seed", range(12))
def test_seed_var_on_uniform_default_geometry(seed):
    """k_seed_var forced (BDX_DEBUG_PREFER_SEED_VAR) onto the sets k_seed's levels normally take -- uniform length,
    default start / end ranges, 150-column search ranges, with and without min_delta: the exact regime, where a
    candidate's value is its verified distance itself."""
    rng = np.random.default_rng(9000 + seed)
    m = int(rng.choice([12, 16, 20, 24, 28, 32]))
    n_bc = int(rng.choice([24, 96, 200, 384]))
    bcs = synth.random_barcodes(rng, n_bc, m, m)
    kw = dict(max_error_rate=float(rng.choice([0.1, 0.2, 0.25])), min_delta=float(rng.choice([0.0, 0.0, 0.1])))
    if rng.random() < 0.4:
        kw["ref_search_range"] = R(f"{int(rng.integers(1, 20))}:{int(rng.integers(60, 170))}")
    cfg = _cfg(bcs, **kw)
    reads = synth.random_reads(rng, 3000, bcs, min_len=int(rng.choice([60, 150])), max_len=int(rng.choice([150, 175, 200])),
                               max_edits=5, n_prob=0.02, lower_prob=0.02)
    blob, off = bdx.pack_reads(reads)
    want = orc.Oracle(cfg).classify_mt(blob, off)
    with capi.Engine(cfg, max_reads=len(reads), max_bytes=int(off[-1]) + 16, debug=capi.DEBUG_PREFER_SEED_VAR) as eng:
        got = eng.classify_packed(blob, off)
    for f in ("status", "bc1", "bc2", "keep_start", "keep_end"):
        bad = np.nonzero(got[f] != want[f])[0]
        assert bad.size == 0, (f, int(bad[0]), got[bad[0]], want[bad[0]], reads[bad[0]])


def _classify_debug(cfg, blob, off, debug):
    with capi.Engine(cfg, max_reads=len(off) - 1, max_bytes=int(off[-1]) + 16, debug=debug) as eng:
        return eng.classify_packed(blob, off)


@pytest.mark.parametrize("shape", ["dense_in_scan", "constrained_list", "short_reads_edges"])
def test_qgram_filter_is_transparent(shape):
    """The 3-gram filter (seed_var.cu, sv_qgram_pass) only drops hits that cannot verify: results with it, without it
    (BDX_DEBUG_NO_QGRAM_FILTER) and the oracle's are the same -- in its in-scan mode (dense geometry), its list mode
    (constrained start) and with reads barely longer than the barcode, where every window crosses a range edge."""
    rng = np.random.default_rng({"dense_in_scan": 1, "constrained_list": 2, "short_reads_edges": 3}[shape])
    debug = 0
    if shape == "dense_in_scan":
        bcs = synth.random_barcodes(rng, 96, 24, 24)
        cfg = _cfg(bcs, min_delta=0.1)
        reads = synth.random_reads(rng, 8000, bcs, min_len=120, max_len=160, max_edits=6, n_prob=0.02, lower_prob=0.02)
        debug = capi.DEBUG_PREFER_SEED_VAR
    elif shape == "constrained_list":
        bcs = synth.random_barcodes(rng, 384, 16, 28)
        cfg = _cfg(bcs, ref_search_range=R("1:40"), barcode_start_range=R("1:6"), min_delta=0.1)
        reads = _reads(rng, 8000, bcs, start_hi=5)
    else:
        bcs = synth.random_barcodes(rng, 200, 20, 28)
        cfg = _cfg(bcs, max_error_rate=0.25)
        reads = []
        for _ in range(8000):
            bc = bcs[int(rng.integers(0, len(bcs)))].encode()
            core = synth.mutate(rng, bc, int(rng.integers(0, 8)))
            pre = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(0, 4))).astype(np.uint8))
            post = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(0, 4))).astype(np.uint8))
            cut = int(rng.integers(0, 3))
            reads.append((pre + core + post)[cut:])
        debug = capi.DEBUG_PREFER_SEED_VAR
    blob, off = bdx.pack_reads(reads)
    want = orc.Oracle(cfg).classify_mt(blob, off)
    with_filter = _classify_debug(cfg, blob, off, debug)
    without = _classify_debug(cfg, blob, off, debug | capi.DEBUG_NO_QGRAM_FILTER)
    for f in ("status", "bc1", "bc2", "keep_start", "keep_end"):
        for got, label in ((with_filter, "filter on"), (without, "filter off")):
            bad = np.nonzero(got[f] != want[f])[0]
            assert bad.size == 0, (label, f, int(bad[0]), got[bad[0]], want[bad[0]], reads[bad[0]])
