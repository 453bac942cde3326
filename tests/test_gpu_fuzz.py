"""Randomised differential fuzz: random option combinations (costs, thresholds, ranges,
trims, algorithms, dual sets, N wildcards, odd alphabets) -- CUDA path vs oracle, bit-exact."""
import os

import numpy as np
import pytest

import bdx_b200 as bdx
import synth
from gpu_common import compare

pytestmark = pytest.mark.gpu
R = bdx.parse_dynamic_range


# BDX_FUZZ_EXTRA=<n> adds n more seeds per algorithm (soak runs on the GPU box; the default stays short)
_EXTRA = int(os.environ.get("BDX_FUZZ_EXTRA", "0"))


def _range(rng, kind):
    """A random 'start:end' expression (mix of absolute and end-relative parts)."""
    if kind == "any" and rng.random() < 0.4:
        return "1:end"
    a = int(rng.integers(1, 40))
    b = int(rng.integers(a, a + 60))
    forms = [f"{a}:{b}", f"{a}:end", f"1:{b}", f"end-{b}:end", f"end-{b}:end-{max(a - 1, 0)}", f"{a}:end-{a}"]
    return forms[int(rng.integers(0, len(forms)))]


def _random_case(seed):
    rng = np.random.default_rng(seed)
    algo = ["semiglobal", "semiglobal", "semiglobal", "hamming", "exact"][int(rng.integers(0, 5))]
    n1 = int(rng.choice([1, 3, 9, 24, 40, 97]))
    m_lo = int(rng.choice([4, 8, 16, 24, 30, 40]))
    m_hi = m_lo + int(rng.choice([0, 0, 3, 8, 30]))
    n_frac = float(rng.choice([0.0, 0.0, 0.1]))
    alphabet = [b"ACGT", b"ACGT", b"ACGTN", b"ACGTRY"][int(rng.integers(0, 4))]
    b1 = synth.random_barcodes(rng, n1, m_lo, m_hi, alphabet=alphabet, n_frac=n_frac)
    kw = dict(matching_algorithm=algo,
              max_error_rate=float(rng.choice([0.0, 0.1, 0.2, 0.25, 0.34, 0.5])),
              min_delta=float(rng.choice([0.0, 0.0, 0.05, 0.15, 0.3])),
              match=int(rng.choice([0, 0, 0, 1, -1])), mismatch=int(rng.choice([1, 1, 2, 3])),
              indel=int(rng.choice([1, 1, 2, 3])),
              nindel=(None if rng.random() < 0.6 else int(rng.choice([1, 2, 3]))),
              trim_side=[None, 3, 5][int(rng.integers(0, 3))],
              ref_search_range=R(_range(rng, "any")), barcode_start_range=R(_range(rng, "any")),
              barcode_end_range=R(_range(rng, "any")))
    cfg = bdx.DemuxConfig(bc_seqs=b1, bc_lengths_no_N=[sum(c != "N" for c in x) for x in b1],
                          ids=[f"a{i}" for i in range(n1)], **kw)
    b2 = None
    if rng.random() < 0.4:
        n2 = int(rng.choice([1, 5, 33, 70]))
        b2 = synth.random_barcodes(rng, n2, m_lo, m_hi, alphabet=alphabet, n_frac=n_frac)
        cfg.is_dual = True
        cfg.bc_seqs2, cfg.bc_lengths_no_N2 = b2, [sum(c != "N" for c in x) for x in b2]
        cfg.ids2 = [f"b{i}" for i in range(n2)]
        cfg.trim_side2 = [None, 3, 5][int(rng.integers(0, 3))]
        cfg.ref_search_range2 = R(_range(rng, "any"))
        cfg.barcode_start_range2 = R(_range(rng, "any"))
        cfg.barcode_end_range2 = R(_range(rng, "any"))
    want_stats = bool(rng.random() < 0.3)
    reads = synth.random_reads(rng, 500, b1, barcodes2=b2, min_len=int(rng.choice([20, 60, 100])),
                               max_len=int(rng.choice([100, 150, 260])), max_edits=5, lower_prob=0.02,
                               at_end2=bool(rng.random() < 0.5))
    return cfg, reads, want_stats


@pytest.mark.parametrize("seed", list(range(80)) + list(range(3000, 3000 + _EXTRA)))
def test_fuzz_random_config(seed):
    cfg, reads, want_stats = _random_case(1000 + seed)
    compare(cfg, reads, want_stats=want_stats, label=f"seed{seed}: {cfg.matching_algorithm}")


def _fast_case(seed, algo):
    """The regimes the shortcut kernels own: uniform barcode length, ACGT barcodes, unit costs; for :semiglobal
    default start / end ranges (k_prefilter -> k_seed levels -> k_filter; min_delta, trimming and stats included), for
    :hamming anything (k_hamming_scan).  Random set sizes, lengths, thresholds, search ranges, read lengths
    (beyond the kernels' staging capacity too), duplicated barcodes, N / lower-case bases in the reads."""
    rng = np.random.default_rng(seed)
    m = int(rng.choice([8, 12, 16, 20, 24, 24, 28, 32]))
    n_bc = int(rng.choice([8, 24, 96, 96, 200, 384, 700]))
    bcs = synth.random_barcodes(rng, n_bc, m, m)
    if rng.random() < 0.3:                      # identical sequences: the lowest index has to win
        for _ in range(3):
            bcs[int(rng.integers(0, n_bc))] = bcs[int(rng.integers(0, n_bc))]
    kw = dict(matching_algorithm=algo, max_error_rate=float(rng.choice([0.0, 0.05, 0.1, 0.13, 0.2, 0.2, 0.25, 0.3])))
    if rng.random() < 0.5:
        kw["ref_search_range"] = R(_range(rng, "any"))
    kw["trim_side"] = [None, 3, 5][int(rng.integers(0, 3))]       # positions: verbatim hits / winner re-aligned
    kw["min_delta"] = float(rng.choice([0.0, 0.0, 0.05, 0.1, 0.13]))   # semiglobal: decided from the seeds' exact distances
    if algo == "hamming":
        if rng.random() < 0.4:
            kw["barcode_start_range"] = R(_range(rng, "any"))
        if rng.random() < 0.4:
            kw["barcode_end_range"] = R(_range(rng, "any"))
    cfg = bdx.DemuxConfig(bc_seqs=bcs, bc_lengths_no_N=[m] * n_bc, ids=[f"a{i}" for i in range(n_bc)], **kw)
    lo = int(rng.choice([m, 40, 100, 150]))
    hi = int(rng.choice([150, 150, 185, 200, 400]))
    reads = synth.random_reads(rng, 1500, bcs, min_len=min(lo, hi), max_len=hi, max_edits=int(rng.choice([3, 5, 6])),
                               lower_prob=0.02, n_prob=0.03)
    reads += [b"", b"A", bcs[0].encode(), bcs[-1].encode() * 3, b"N" * 60,
              b"ACGT" * 5 + bcs[1].encode() + b"TTG" + bcs[1].encode() + b"CA",          # two verbatim occurrences
              bcs[2].encode() + b"GATTACA" + bcs[2].encode() + b"C" + bcs[2].encode()]   # three of them
    return cfg, reads, bool(rng.random() < 0.3)


@pytest.mark.parametrize("seed", list(range(40)) + list(range(1000, 1000 + _EXTRA)))
@pytest.mark.parametrize("algo", ["semiglobal", "hamming"])
def test_fuzz_shortcut_regimes(seed, algo):
    cfg, reads, want_stats = _fast_case(5000 + seed, algo)
    compare(cfg, reads, want_stats=want_stats, label=f"fast seed{seed}: {algo}")
