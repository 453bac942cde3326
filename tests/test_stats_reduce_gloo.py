"""N > 1 path on CPU: shard reads over 2 ranks, sum the flat DemuxStats counter buffers
(layout = bdx_stats_layout from the C library) with a gloo all_reduce, and check the result
equals the single-process DemuxStats (== reference merge_stats, reporting.jl:1-58)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _counters(layout, cfg, ref):
    """Flat counter buffer from oracle results, bin for bin like k_finalize."""
    L = layout
    buf = np.zeros(L.total_len, dtype=np.int64)
    for i in range(len(ref)):
        buf[0] += 1
        st = int(ref["status"][i])
        for p in (0, 1):
            ps = ref["passes"][i, p]
            if ps["status"] != 0:
                continue
            b, s, e = int(ps["bc"]), int(ps["start"]), int(ps["end"])
            from bdx_b200.stats import pass_norms
            norm = pass_norms(cfg, p == 1)[b - 1]
            d = int(round(float(ps["score"]) * norm))
            pos = min(max(s + L.pos_bias, 0), L.pos_bins - 1)
            ln = min(max(e - s + 1, 0), L.len_bins - 1)
            d = min(max(d + L.dist_bias, 0), L.dist_bins - 1)
            for row in (0, b):
                buf[L.pos_off[p] + row * L.pos_bins + pos] += 1
                buf[L.len_off[p] + row * L.len_bins + ln] += 1
                buf[L.dist_off[p] + row * L.dist_bins + d] += 1
        if st == 0:
            buf[1] += 1
            buf[L.sample_off + int(ref["bc1"][i]) * (L.b2 + 1) + int(ref["bc2"][i])] += 1
        elif st == 1:
            buf[2] += 1
        else:
            buf[3] += 1
    return buf


def _make(seed=5):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bdx_b200 as bdx
    import synth
    rng = np.random.default_rng(seed)
    b1 = synth.random_barcodes(rng, 20, 12, 20)
    b2 = synth.random_barcodes(rng, 12, 10, 16)
    cfg = bdx.DemuxConfig(bc_seqs=b1, bc_lengths_no_N=[len(x) for x in b1], ids=[f"a{i}" for i in range(20)],
                          is_dual=True, bc_seqs2=b2, bc_lengths_no_N2=[len(x) for x in b2],
                          ids2=[f"b{i}" for i in range(12)], summary=True, min_delta=0.05, trim_side=5)
    reads = synth.random_reads(rng, 600, b1, barcodes2=b2, min_len=60, max_len=90, start_hi=8)
    return cfg, reads


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bdx_b200 import capi
    import orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, reads = _make()
    shard = reads[rank::world]                       # host dispatcher: round-robin batches
    ref = orc.Oracle(cfg, want_stats=True).classify_reads(shard)
    lay = capi.Config(cfg).layout
    t = torch.from_numpy(_counters(lay, cfg, ref))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put(t.numpy().copy())
    dist.destroy_process_group()


def test_counter_allreduce_equals_merge_stats():
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from bdx_b200 import capi
    from bdx_b200.stats import stats_from_counters, stats_from_passes
    import orc
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    summed = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg, reads = _make()
    ref = orc.Oracle(cfg, want_stats=True).classify_reads(reads)
    lay = capi.Config(cfg).layout
    got = stats_from_counters(summed, lay, cfg)
    passes = [[tuple(ref["passes"][i, p][k] for k in ("status", "bc", "start", "end", "score")) for p in (0, 1)]
              for i in range(len(reads))]
    want = stats_from_passes(ref["status"], ref["bc1"], ref["bc2"], passes, cfg)
    assert got == want
    assert got.total_reads == len(reads) and got.matched_reads > 0
