"""bdx_fastq_scan / bdx_fastq_pack (C++ block parser) against the 4x-readline
record semantics of the reference's reader_task (core.jl:43-110), mirrored by fileio.fastq_records."""
import os
import time

import numpy as np
import pytest

import bdx_b200 as bdx
from bdx_b200 import capi

CASES = {
    "plain": b"@r1\nACGT\n+\nIIII\n@r2\nTT\n+\nII\n",
    "crlf": b"@r1\r\nACGT\r\n+\r\nIIII\r\n@r2\r\nTT\r\n+\r\nII\r\n",
    "no_final_newline": b"@r1\nACGT\n+\nIIII\n@r2\nTT\n+\nII",
    "truncated_record": b"@r1\nACGT\n+\nIIII\n@r2\nTT\n",
    "blank_tail": b"@r1\nACGT\n+\nIIII\n\n",
    "empty_seq": b"@r1\n\n+\n\n@r2\nA\n+\nI\n",
    "lone_cr": b"@r1\nAC\rGT\n+\nIIIII\n",
    "empty": b"",
}


def _python_records(tmp_path, data):
    p = tmp_path / "x.fastq"
    p.write_bytes(data)
    return list(bdx.fastq_records(str(p)))


def _c_records(data, block):
    """Feed the file in blocks of `block` bytes, carrying the unconsumed tail like a streaming host."""
    out, tail, pos = [], b"", 0
    while True:
        chunk = data[pos:pos + block]
        pos += len(chunk)
        final = pos >= len(data)
        buf = np.frombuffer(tail + chunk, dtype=np.uint8)
        recs, consumed = capi.fastq_scan(buf, final, 1000)
        raw = buf.tobytes()
        for r in recs:
            out.append(tuple(raw[r[f + "_off"]:r[f + "_off"] + r[f + "_len"]] for f in ("header", "seq", "plus", "qual")))
        tail = raw[consumed:]
        if final:
            assert consumed == len(raw) or len(recs) == 1000
            if consumed == len(raw):
                return out


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("block", [1 << 20, 7, 16])
def test_scan_matches_readline_semantics(tmp_path, name, block):
    data = CASES[name]
    assert _c_records(data, block) == _python_records(tmp_path, data)


def test_pack_layout_and_fixture(refdata):
    path = os.path.join(refdata, "FASTQ_files", "demo1_R1", "demo-001_R1.fastq")
    data = open(path, "rb").read()
    buf = np.frombuffer(data, dtype=np.uint8)
    recs, consumed = capi.fastq_scan(buf, True, 100)
    assert consumed == len(data) and len(recs) == 10
    seq, off = capi.fastq_pack(buf, recs)
    py = list(bdx.fastq_records(path))
    want_blob, want_off = bdx.pack_reads([r[1] for r in py])
    assert (seq == want_blob).all() and (off == want_off).all()
    small = np.zeros(10, dtype=np.uint8)
    with pytest.raises(capi.BdxError) as ei:
        capi.fastq_pack(buf, recs, seq_out=small)
    assert ei.value.code == capi.BDX_ERR_TOO_LARGE


def test_scan_throughput_smoke():
    rec = b"@read/1 some header text\n" + b"ACGT" * 37 + b"AC\n+\n" + b"I" * 150 + b"\n"
    data = np.frombuffer(rec * 200_000, dtype=np.uint8)
    t0 = time.perf_counter()
    recs, consumed = capi.fastq_scan(data, True, 200_000)
    seq, off = capi.fastq_pack(data, recs)
    dt = time.perf_counter() - t0
    assert len(recs) == 200_000 and consumed == data.size and off[-1] == 200_000 * 150
    assert data.size / dt > 2e8, f"scan+pack slower than 0.2 GB/s: {data.size / dt / 1e9:.2f} GB/s"
