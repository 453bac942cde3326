"""4-bit packed read input (bdx_submit_packed4): identical results to the byte input and to the oracle, for every
matching algorithm, with read bytes outside the barcode alphabet (N, lower case, IUPAC letters), dual sets with
different alphabets, ragged read lengths (odd totals, nibble boundaries inside bytes) and several batches in flight."""
import numpy as np
import pytest

import bdx_b200 as bdx
import orc
import synth
from bdx_b200 import capi

pytestmark = pytest.mark.gpu
R = bdx.parse_dynamic_range

CASES = {
    "semiglobal": dict(),
    "semiglobal_trim_delta": dict(trim_side=3, min_delta=0.1),
    "weighted": dict(mismatch=2, indel=3, max_error_rate=0.3, trim_side=5),
    "hamming": dict(matching_algorithm="hamming", trim_side=5),
    "exact": dict(matching_algorithm="exact"),
    "nindel": dict(nindel=1, max_error_rate=0.3, n_frac=0.15),
    "iupac_dual": dict(alphabet=b"ACGTRY", dual=True, trim_side=5, trim_side2=3),
    "ranges": dict(ref_search_range=R("1:40"), barcode_start_range=R("1:6"), min_delta=0.1, varlen=True),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_packed_equals_bytes_and_oracle(name):
    spec = dict(CASES[name])
    rng = np.random.default_rng(sum(map(ord, name)))
    alphabet = spec.pop("alphabet", b"ACGT")
    dual = spec.pop("dual", False)
    varlen = spec.pop("varlen", False)
    b1 = synth.random_barcodes(rng, 96, 16 if varlen else 24, 28 if varlen else 24, alphabet=alphabet,
                               n_frac=spec.pop("n_frac", 0.0))
    cfg = bdx.DemuxConfig(bc_seqs=b1, bc_lengths_no_N=[sum(c != "N" for c in b) for b in b1],
                          ids=[f"a{i}" for i in range(len(b1))], **spec)
    b2 = None
    if dual:
        b2 = synth.random_barcodes(rng, 20, 20, 24, alphabet=b"ACGTN")
        cfg.is_dual, cfg.bc_seqs2, cfg.bc_lengths_no_N2 = True, b2, [len(b) for b in b2]
        cfg.ids2 = [f"b{i}" for i in range(len(b2))]
    reads = synth.random_reads(rng, 5000, b1, barcodes2=b2, min_len=30, max_len=151, n_prob=0.05, lower_prob=0.05,
                               start_hi=5 if varlen else None)
    reads += [b"", b"A", b"NNNNN", b"acgtacgtacgt", b1[0].encode("latin-1")]
    blob, off = bdx.pack_reads(reads)
    off32 = off.astype(np.int32)
    want = orc.Oracle(cfg).classify_mt(blob, off)
    with capi.Engine(cfg, max_reads=len(reads), max_bytes=int(off[-1]) + 16) as eng:
        plain = eng.classify_packed(blob, off)
        packed = eng.config.pack4(blob)
        assert packed.size == (blob.size + 1) // 2
        st = eng.stream
        st.submit_packed4(packed, off32, tag=7)
        tag, got = st.fetch()
        assert tag == 7
        # several batches in flight, odd split points (a batch's nibbles start at its own byte 0)
        cut = [0, 1234, 2501, len(reads)]
        for k in range(3):
            a, b = cut[k], cut[k + 1]
            sub = blob[off[a]:off[b]]
            st.submit_packed4(eng.config.pack4(sub), (off[a:b + 1] - off[a]).astype(np.int32), tag=k)
        parts = [st.fetch()[1] for _ in range(3)]
    for f in ("status", "bc1", "bc2", "keep_start", "keep_end"):
        assert (got[f] == plain[f]).all(), f
        assert (got[f] == want[f]).all(), f
        assert (np.concatenate(parts)[f] == want[f]).all(), f


def test_packed_refused_for_wide_alphabets():
    bcs = ["ABCDEFGHIJKLMNOP", "QRSTUVWXABCDEFGH"]
    cfg = bdx.DemuxConfig(bc_seqs=bcs, bc_lengths_no_N=[16, 16], ids=["x", "y"])
    with capi.Engine(cfg, max_reads=10, max_bytes=1000) as eng:
        with pytest.raises(capi.BdxError) as ei:
            eng.stream.submit_packed4(np.zeros(8, np.uint8), np.array([0, 16], np.int32))
        assert ei.value.code == capi.BDX_ERR_INVALID
