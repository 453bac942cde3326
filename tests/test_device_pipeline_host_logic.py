"""Host logic of the device FASTQ pipeline (run_pipeline_device: block reading, tail carry-over, per-file final
flags, bucket appends) on the CPU: the stream is replaced by a stand-in that implements bdx_demux_block's contract
with the C++ FASTQ scanner, the ORACLE as classifier and the Python mirror of writer_task.  The outputs must equal
run_pipeline's record-by-record path for every block size.  (The real kernel is tested in test_gpu_demux.py.)"""
import os
import types

import numpy as np
import pytest

import bdx_b200 as bdx
import hostref
import synth
from bdx_b200 import capi
from bdx_b200.demux import run_pipeline, run_pipeline_device


class FakeDemuxStream:
    def __init__(self, cfg):
        self.cfg = cfg
        self.classify = hostref.oracle_classifier(cfg)
        self.do_trim = cfg.trim_side is not None or cfg.trim_side2 is not None

    def _records(self, buf, final):
        recs, consumed = capi.fastq_scan(buf, bool(final), max(buf.size // 4 + 2, 4))
        raw = buf.tobytes()
        out = [tuple(raw[r[f + "_off"]:r[f + "_off"] + r[f + "_len"]] for f in ("header", "seq", "plus", "qual"))
               for r in recs]
        ends = [0]
        pos = 0
        # bytes consumed after each record: re-scan prefix lengths (records are contiguous from 0)
        for k in range(len(recs)):
            _, c = capi.fastq_scan(buf, bool(final), k + 1)
            ends.append(c)
        return out, ends

    def demux_block(self, b1, b2=None, final_block=1, mode=0):
        single = mode == capi.DEMUX_SINGLE
        f1 = (final_block != 0) if single else bool(final_block & 1)
        r1, e1 = self._records(b1, f1)
        n = len(r1)
        if not single:
            r2, e2 = self._records(b2, bool(final_block & 2))
            n = min(n, len(r2))
        res = self.classify([r[1] for r in r1[:n]])
        keys, chunks1, chunks2 = {}, {}, {}
        for i in range(n):
            st, a, b = int(res["status"][i]), int(res["bc1"][i]), int(res["bc2"][i])
            key = 0 if st == 1 else (1 if st == 2 else 2 + (a - 1) * max(len(self.cfg.bc_seqs2 or []), 1) + max(b - 1, 0))
            keys[key] = (st, a, b)
            h, s, p, q = r1[i]
            if self.do_trim and int(res["keep_start"][i]) != -1:
                lo, hi = max(int(res["keep_start"][i]), 1), min(int(res["keep_end"][i]), len(s))
                s, q = (s[lo - 1:hi], q[lo - 1:hi]) if lo <= hi else (b"", b"")
            chunks1.setdefault(key, []).append(b"\n".join((h, s, p, q, b"")))
            if not single:
                chunks2.setdefault(key, []).append(b"\n".join(r2[i] + (b"",)))
        buckets = np.zeros(len(keys), dtype=capi.BUCKET_DTYPE)
        o1, o2 = b"", b""
        for k, key in enumerate(sorted(keys)):
            st, a, b = keys[key]
            c1 = b"".join(chunks1[key]) if mode != capi.DEMUX_MATES else b""
            c2 = b"".join(chunks2[key]) if not single else b""
            buckets[k] = (st, a, b, len(chunks1[key]), len(o1), len(c1), len(o2), len(c2))
            o1 += c1
            o2 += c2
        out = types.SimpleNamespace(n_records=n, n_buckets=len(keys), consumed1=e1[n], consumed2=0 if single else e2[n])
        out._views = (buckets, np.frombuffer(o1, dtype=np.uint8), np.frombuffer(o2, dtype=np.uint8), res)
        return out

    @staticmethod
    def demux_views(out):
        return out._views


def _tree(d):
    return {f: open(os.path.join(d, f), "rb").read() for f in sorted(os.listdir(d))}


@pytest.mark.parametrize("variant", ["single_trim", "crlf_no_final_newline", "paired_mates", "paired_both_unequal"])
def test_block_reader_and_bucket_appends(tmp_path, variant):
    rng = np.random.default_rng(len(variant))
    bcs = synth.random_barcodes(rng, 12, 10)
    cfg = bdx.DemuxConfig(bc_seqs=bcs, bc_lengths_no_N=[10] * 12, ids=[f"s{i}" for i in range(12)],
                          trim_side=5 if variant == "single_trim" else None)
    reads = synth.random_reads(rng, 300, bcs, min_len=15, max_len=50)
    eol = b"\r\n" if variant.startswith("crlf") else b"\n"

    def text(rs, final_nl=True):
        recs = [b"@r%d" % i + eol + s + eol + b"+" + eol + b"I" * len(s) for i, s in enumerate(rs)]
        return eol.join(recs) + (eol if final_nl else b"")

    f1 = str(tmp_path / "a_R1.fastq")
    open(f1, "wb").write(text(reads, final_nl=not variant.startswith("crlf")))
    f2 = None
    if variant.startswith("paired"):
        mates = synth.random_reads(rng, 300 - (37 if variant.endswith("unequal") else 0), bcs, min_len=10, max_len=30)
        f2 = str(tmp_path / "a_R2.fastq")
        open(f2, "wb").write(text(mates))
        cfg.classify_both = "both" in variant
    want = str(tmp_path / "want")
    run_pipeline(cfg, hostref.oracle_classifier(cfg), f1, f2, want, "p1", "p2", chunk_size=64)
    for bb in (1 << 20, 257, 1500, 40):
        got = str(tmp_path / f"got{bb}")
        run_pipeline_device(cfg, FakeDemuxStream(cfg), f1, f2, got, "p1", "p2", block_bytes=bb)
        assert _tree(got) == _tree(want), (variant, bb)
