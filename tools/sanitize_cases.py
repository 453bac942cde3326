#!/usr/bin/env python
"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): every kernel family
of libbdx on a few hundred reads each, checked against the oracle.

    compute-sanitizer --tool memcheck python tools/sanitize_cases.py

Cases: smoke()'s two configs (k_prefilter, k_seed x2, k_filter, k_literal, k_finalize); dual + trim + stats
(k_seed_deep, windowed k_literal, atomics of the DemuxStats counters); the config-3 shape (k_seed_var levels, shared
-memory hit lists, candidate-parallel literal); :hamming (k_hamming_scan) and :exact (k_prefilter<1>); CUDA-graph
replay of repeated full-size chunks; 4-bit packed input (k_unpack4); the device FASTQ demultiplexer."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bdx_b200 as bdx          # noqa: E402
import orc                      # noqa: E402
import synth                    # noqa: E402
from bdx_b200 import capi       # noqa: E402

R = bdx.parse_dynamic_range
N = int(os.environ.get("BDX_SANITIZE_READS", "300"))


def check(name, cfg, reads, want_stats=False):
    blob, off = bdx.pack_reads(reads)
    with capi.Engine(cfg, max_reads=len(reads), max_bytes=int(off[-1]) + 16, want_stats=want_stats) as eng:
        got = eng.classify_packed(blob, off)
        launches = eng.stream.launch_count
    want = orc.Oracle(cfg, want_stats=want_stats).classify(blob, off)
    for f in ("status", "bc1", "bc2", "keep_start", "keep_end"):
        assert (got[f] == want[f]).all(), (name, f)
    print(f"{name}: {len(reads)} reads, {launches} launches, matched {(got['status'] == 0).sum()}", flush=True)


def main():
    rng = np.random.default_rng(1)
    b96 = synth.random_barcodes(rng, 96, 24)
    r150 = synth.random_reads(rng, N, b96, min_len=150)
    mk = lambda bcs, **kw: bdx.DemuxConfig(bc_seqs=bcs, bc_lengths_no_N=[len(b) for b in bcs],  # noqa: E731
                                           ids=[str(i) for i in range(len(bcs))], **kw)
    check("default", mk(b96), r150)
    check("weighted_delta_trim", mk(b96, min_delta=0.1, trim_side=5, mismatch=1, indel=2, max_error_rate=0.25), r150)
    check("trim3_stats", mk(b96, trim_side=3, summary=True), r150, want_stats=True)
    adapter = ["AGATCGGAAGAGCACACGTCTGAACTCCAGTCA"]
    c4 = mk(b96, ref_search_range=R("1:32"), trim_side=5, trim_side2=3)
    c4.is_dual, c4.bc_seqs2, c4.bc_lengths_no_N2, c4.ids2 = True, adapter, [33], ["adapter"]
    check("dual_adapter", c4, synth.random_reads(rng, N, b96, barcodes2=adapter, min_len=150, start_hi=4, at_end2=False))
    b1, b2 = synth.random_barcodes(rng, 384, 16, 28), synth.random_barcodes(rng, 384, 16, 28)
    c3 = mk(b1, ref_search_range=R("1:40"), barcode_start_range=R("1:6"), ref_search_range2=R("end-39:end"),
            barcode_end_range2=R("end-5:end"), min_delta=0.1)
    c3.is_dual, c3.bc_seqs2, c3.bc_lengths_no_N2, c3.ids2 = True, b2, [len(b) for b in b2], [str(i) for i in range(384)]
    check("config3_shape", c3, synth.random_reads(rng, N, b1, barcodes2=b2, min_len=150, start_hi=4))
    check("hamming", mk(b96, matching_algorithm="hamming", trim_side=5), r150)
    check("exact", mk(b96, matching_algorithm="exact"), r150)
    check("nindel_literal_only", mk(synth.random_barcodes(rng, 12, 20, 26, n_frac=0.2), nindel=1, indel=2, max_error_rate=0.4), r150[:100])
    # CUDA-graph replay: the same slot sees full-size chunks again and again
    cfg = mk(b96)
    with capi.Engine(cfg, max_reads=100, max_bytes=100 * 150) as eng:
        for rep in range(10):
            chunk = synth.random_reads(rng, 100, b96, min_len=150)
            blob, off = bdx.pack_reads(chunk)
            got = eng.classify_packed(blob, off)
            want = orc.Oracle(cfg).classify(blob, off)
            assert (got["bc1"] == want["bc1"]).all() and (got["status"] == want["status"]).all(), rep
        print(f"graph_replay: 10 chunks, {eng.stream.launch_count} launches", flush=True)
        # 4-bit packed input on the same stream
        packed = eng.config.pack4(blob)
        eng.stream.submit_packed4(packed, off.astype(np.int32), tag=1)
        _, got = eng.stream.fetch()
        assert (got["bc1"] == want["bc1"]).all()
        print("packed4: ok", flush=True)
    # device FASTQ demultiplexer
    recs = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(r150[:200]))
    cfg = mk(b96, trim_side=5)
    with capi.Engine(cfg, max_reads=4000) as eng:
        out = eng.stream.demux_block(np.frombuffer(recs, dtype=np.uint8), None, final_block=1)
        print(f"demux_block: {out.n_records} records, {out.n_buckets} buckets", flush=True)
    print("sanitize cases ok", flush=True)


if __name__ == "__main__":
    main()
