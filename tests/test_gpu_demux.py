"""Device FASTQ block demultiplexer (bdx_demux_block, SURVEY.md 8f-1 + 8f-3): FASTQ text in, per-file
record runs out.  Checked byte for byte against (a) the reference's golden output files and (b) the
host mirror of reader_task / writer_task fed with the same CUDA classification."""
import os
import zlib

import numpy as np
import pytest

import bdx_b200 as bdx
import hostref
import synth
from bdx_b200 import capi
from bdx_b200.demux import Chunk, Writer, output_filename, run_pipeline, run_pipeline_device

pytestmark = pytest.mark.gpu


def _tree(d):
    return {f: open(os.path.join(d, f), "rb").read() for f in sorted(os.listdir(d))}


@pytest.mark.parametrize("idx", [0, 1, 2])
@pytest.mark.parametrize("block_bytes", [32 << 20, 3001])
def test_golden_files_through_device_io(refdata, tmp_path, idx, block_bytes):
    """Config 1 end to end on the device: file text -> bdx_demux_block -> appended buckets must equal
    the reference's golden output files (single-end, paired routed by read 1, classify_both)."""
    name, files, bc_file, kw, ideal = hostref.demo_cases(refdata)[idx]
    out_dir = str(tmp_path / "out")
    for f1, f2, p1, p2 in files:
        kw2 = dict(kw)
        if f2 is None:
            kw2["classify_both"] = False
        cfg = hostref.build_cfg(bc_file, [f1] if f2 is None else [f1, f2], **kw2)
        with capi.Engine(cfg, max_reads=0) as eng:
            run_pipeline_device(cfg, eng.stream, f1, f2, out_dir, p1, p2, block_bytes=block_bytes)
    assert hostref.check_output_files(out_dir, ideal) == {0: 24, 1: 24, 2: 76}[idx]


def _fastq_text(rng, reads, eol=b"\n", final_newline=True, ragged_qual=False):
    recs = []
    for i, s in enumerate(reads):
        q = bytes(rng.integers(33, 74, len(s)).astype(np.uint8))
        if ragged_qual and i % 7 == 3:
            q = q + b"II"      # longer quality line: cut with the same range (core.jl:160-161)
        recs.append(b"@r%d some comment" % i + eol + s + eol + b"+" + (b"r%d" % i if i % 3 == 0 else b"") + eol + q)
    text = eol.join(recs)
    if final_newline and recs:
        text += eol
    return text


def _host_pipeline(cfg, eng, path1, path2, out_dir, p1, p2):
    run_pipeline(cfg, eng.classify_reads, path1, path2, out_dir, p1, p2, chunk_size=977)


@pytest.mark.parametrize("variant", ["plain", "crlf", "no_final_newline", "trim5", "trim3_dual", "blank_tail",
                                     "paired_mates", "paired_both", "paired_unequal", "dual_many_files"])
def test_device_io_equals_host_pipeline(tmp_path, variant):
    rng = np.random.default_rng(zlib.crc32(variant.encode()))
    bcs = synth.random_barcodes(rng, 40, 12)
    kw = {}
    if variant == "trim5":
        kw = dict(trim_side=5)
    reads = synth.random_reads(rng, 5000, bcs, min_len=20, max_len=90)
    reads[17] = b""
    cfg = bdx.DemuxConfig(bc_seqs=bcs, bc_lengths_no_N=[12] * 40, ids=[f"s{i}" for i in range(40)], **kw)
    if variant == "trim3_dual":
        bcs2 = synth.random_barcodes(rng, 9, 10)
        reads = [r + bcs2[i % 9].encode() + b"ACGT" for i, r in enumerate(reads)]
        cfg = bdx.DemuxConfig(bc_seqs=bcs, bc_lengths_no_N=[12] * 40, ids=[f"s{i}" for i in range(40)],
                              is_dual=True, bc_seqs2=bcs2, bc_lengths_no_N2=[10] * 9, ids2=[f"t{i}" for i in range(9)],
                              trim_side=5, trim_side2=3, min_delta=0.05)
    if variant == "dual_many_files":             # 300 x 300 output files: keys beyond 2^16, three radix passes
        bcs = synth.random_barcodes(rng, 300, 12)
        bcs2 = synth.random_barcodes(rng, 300, 10)
        reads = [r + bcs2[int(rng.integers(0, 300))].encode() + b"AC" for r in synth.random_reads(rng, 900, bcs, min_len=20, max_len=60)]   # < 1024 output files stay open at once
        cfg = bdx.DemuxConfig(bc_seqs=bcs, bc_lengths_no_N=[12] * 300, ids=[f"s{i}" for i in range(300)],
                              is_dual=True, bc_seqs2=bcs2, bc_lengths_no_N2=[10] * 300, ids2=[f"t{i}" for i in range(300)])
    eol = b"\r\n" if variant == "crlf" else b"\n"
    text1 = _fastq_text(rng, reads, eol=eol, final_newline=variant != "no_final_newline",
                        ragged_qual=variant in ("trim5", "trim3_dual"))
    if variant == "blank_tail":
        text1 += b"\n@partial\nACGT\n+\n\n@cut\nAC"   # records with empty lines, then a truncated one (missing lines read as "")
    f1 = str(tmp_path / "in_R1.fastq")
    open(f1, "wb").write(text1)
    f2 = None
    if variant.startswith("paired"):
        n2 = len(reads) - 123 if variant == "paired_unequal" else len(reads)
        mates = synth.random_reads(rng, n2, bcs, min_len=10, max_len=60)
        f2 = str(tmp_path / "in_R2.fastq")
        open(f2, "wb").write(_fastq_text(rng, mates))
        cfg.classify_both = variant != "paired_mates"
    want_dir, got_dir = str(tmp_path / "want"), str(tmp_path / "got")
    with capi.Engine(cfg, max_reads=1000) as eng:
        _host_pipeline(cfg, eng, f1, f2, want_dir, "p1", "p2")
        for k, bb in enumerate((1 << 20, 4099, 70001)):
            d = got_dir + str(k)
            run_pipeline_device(cfg, eng.stream, f1, f2, d, "p1", "p2", block_bytes=bb)
            want, got = _tree(want_dir), _tree(d)
            assert sorted(want) == sorted(got), f"{variant}/{bb}: file sets differ"
            for name in want:
                assert want[name] == got[name], f"{variant}/{bb}: {name} differs"


def test_demux_block_raw_api():
    """Bucket table, consumed bytes and per-record results of one call; empty and tiny inputs."""
    rng = np.random.default_rng(5)
    bcs = synth.random_barcodes(rng, 300, 16)            # > 256 keys: two radix passes
    cfg = bdx.DemuxConfig(bc_seqs=bcs, bc_lengths_no_N=[16] * 300, ids=[f"s{i}" for i in range(300)])
    reads = synth.random_reads(rng, 20000, bcs, min_len=60, max_len=60)
    text = np.frombuffer(_fastq_text(rng, reads) + b"@cut\nACG", dtype=np.uint8)
    with capi.Engine(cfg, max_reads=len(reads), max_bytes=len(reads) * 60) as eng:
        ref = eng.classify_reads(reads)
        out = eng.stream.demux_block(text, None, final_block=0, mode=capi.DEMUX_SINGLE)
        buckets, o1, _, res = eng.stream.demux_views(out)
        assert out.n_records == len(reads) and out.consumed1 == text.size - len(b"@cut\nACG")
        for f in ("status", "bc1", "bc2", "keep_start", "keep_end"):
            assert (res[f] == ref[f]).all()
        assert int(buckets["n_records"].sum()) == len(reads)
        assert int(buckets["length1"].sum()) == out.out1_len == out.consumed1
        keys = [0 if b["status"] == 1 else (1 if b["status"] == 2 else 1 + int(b["bc1"])) for b in buckets]
        assert all(a < b for a, b in zip(keys, keys[1:]))
        for b in buckets[:50]:
            chunk = bytes(o1[b["offset1"]:b["offset1"] + b["length1"]])
            assert chunk.count(b"\n") == 4 * b["n_records"]
        # final: the truncated record is completed
        out = eng.stream.demux_block(text, None, final_block=1, mode=capi.DEMUX_SINGLE)
        assert out.n_records == len(reads) + 1 and out.consumed1 == text.size
        assert out.out1_len == text.size + 3
        # nothing complete / nothing at all
        out = eng.stream.demux_block(np.frombuffer(b"@a\nAC", dtype=np.uint8), None, final_block=0)
        assert out.n_records == 0 and out.consumed1 == 0
        out = eng.stream.demux_block(np.zeros(0, np.uint8), None, final_block=1)
        assert out.n_records == 0 and out.n_buckets == 0
        ms = eng.stream.demux_stage_ms()
        assert len(ms) == 8
