"""DemuxStats for long reads searched near their end: more than 2^20 matched passes whose start lies beyond the
device histograms' 1024 positions.  Every one of them has to come back as an exact record (the overflow list is
drained / grown by the host, never dropped), so that the position / length Dicts equal the reference's
(classification.jl:827-865)."""
import numpy as np
import pytest

import bdx_b200 as bdx
import orc
from bdx_b200 import capi
from bdx_b200.stats import stats_from_counters

pytestmark = pytest.mark.gpu
R = bdx.parse_dynamic_range


def _long_reads(n, L, bcs, seed):
    rng = np.random.default_rng(seed)
    m = rng.integers(0, 4, size=(n, L), dtype=np.uint8)
    blob = np.frombuffer(b"ACGT", dtype=np.uint8)[m]
    which = rng.integers(0, len(bcs), n)
    plant = rng.random(n) < 0.95
    start = L - 40 + rng.integers(0, 8, n)                       # 0-based start column, beyond position 1024
    for b, bc in enumerate(bcs):
        rows = np.nonzero(plant & (which == b))[0]
        code = np.frombuffer(bc.encode(), dtype=np.uint8)
        for k in range(len(code)):
            blob[rows, start[rows] + k] = code[k]
    sub = np.nonzero(plant & (rng.random(n) < 0.3))[0]            # one substitution inside some barcodes
    blob[sub, start[sub] + 5] = ord("A")
    return blob.reshape(-1), np.arange(n + 1, dtype=np.int64) * L


@pytest.mark.parametrize("path", ["host_batches", "device_one_call"])
def test_more_than_2_pow_20_overflow_records(path):
    n, L = 1_200_000, 1100
    bcs = ["ACGGTCATGCAT", "TTGACCGTAAGC", "GGCATTACGGTA", "CATCGATTGCCA"]
    cfg = bdx.DemuxConfig(bc_seqs=bcs, bc_lengths_no_N=[12] * 4, ids=list("wxyz"), ref_search_range=R("end-60:end"),
                          trim_side=3, max_error_rate=0.1, summary=True)
    blob, off = _long_reads(n, L, bcs, 7)
    want = orc.Oracle(cfg, want_stats=True).classify_mt(blob, off)
    if path == "host_batches":          # 12 batches of 100 000 reads: the list is drained every few batches
        with capi.Engine(cfg, max_reads=100_000, max_bytes=100_000 * L, want_stats=True) as eng:
            got = eng.classify_packed(blob, off)
            counters, ovf, lay = eng.stream.stats(), eng.stream.stats_overflow(), eng.config.layout
    else:                               # one device-resident call: the list grows to 2 records per read
        import torch
        config = capi.Config(cfg, want_stats=True)
        st = capi.Stream(config, device=0, max_reads=0, max_bytes=0)
        d_seq = torch.from_numpy(blob).cuda()
        d_off = torch.from_numpy(off.astype(np.int32)).cuda()
        d_res = torch.empty(n * bdx.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        st.classify_device(d_seq.data_ptr(), d_off.data_ptr(), n, d_res.data_ptr())
        st.sync()
        got = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=bdx.RESULT_DTYPE)
        counters, ovf, lay = st.stats(), st.stats_overflow(), config.layout
        st.close()
    for f in ("status", "bc1", "keep_start", "keep_end"):
        assert (got[f] == want[f]).all(), f
    p1 = want["passes"][:, 0]
    m = p1["status"] == 0
    assert int(m.sum()) > (1 << 20)                                   # more matched passes than the old fixed list held
    assert len(ovf) == int(m.sum())                                    # every one of them starts beyond column 1024
    stats = stats_from_counters(counters, lay, cfg, ovf)
    assert stats.total_reads == n and stats.matched_reads == int(m.sum())
    for attr, vals in (("bc1_pos_counts", p1["start"][m]), ("bc1_len_counts", (p1["end"] - p1["start"] + 1)[m])):
        keys, cnt = np.unique(vals, return_counts=True)
        assert getattr(stats, attr) == {int(k): int(c) for k, c in zip(keys, cnt)}, attr
    for b in range(1, 5):
        keys, cnt = np.unique(p1["start"][m & (p1["bc"] == b)], return_counts=True)
        assert stats.bc1_per_bc_pos_counts[b] == {int(k): int(c) for k, c in zip(keys, cnt)}, b
