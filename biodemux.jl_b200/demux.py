"""Host streaming pipeline around the hot path: ``execute_demultiplexing``.

Mirrors the reference's orchestration (src/core.jl:43-224, 360-631): FASTQ chunks
in, per-barcode output files out, same option names and defaults, same output
naming and append semantics.  The per-chunk classification (the body of
``worker_task``, core.jl:235-269) is done by the CUDA engine through the C ABI
(``capi.Engine``); there is no CPU classification path in this package.
"""
from __future__ import annotations

import gzip
import os
import re
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from .config import DemuxConfig, build_config
from .fileio import fastq_records

# bdx_result (include/bdx.h)
RESULT_DTYPE = np.dtype([("status", "<i4"), ("bc1", "<i4"), ("bc2", "<i4"),
                         ("keep_start", "<i4"), ("keep_end", "<i4")])
# bdx_pass_detail (include/bdx.h)
DETAIL_DTYPE = np.dtype([("status", "<i4"), ("bc", "<i4"), ("dist", "<i4"), ("norm", "<i4"),
                         ("start", "<i4"), ("end", "<i4")])
MATCH, UNKNOWN, AMBIGUOUS = 0, 1, 2


class Chunk:
    """core.jl:5-39 -- up to ``chunk_size`` FASTQ records (and their mates)."""

    __slots__ = ("id", "headers", "seqs", "pluses", "quals", "mate")

    def __init__(self, cid: int):
        self.id = cid
        self.headers: List[bytes] = []
        self.seqs: List[bytes] = []
        self.pluses: List[bytes] = []
        self.quals: List[bytes] = []
        self.mate: Optional["Chunk"] = None

    def __len__(self):
        return len(self.headers)


def iter_chunks(fastq1: str, fastq2: Optional[str], chunk_size: int):
    """reader_task (core.jl:43-110)."""
    it1 = fastq_records(fastq1)
    it2 = fastq_records(fastq2) if fastq2 is not None else None
    cid = 1
    while True:
        c1 = Chunk(cid)
        c2 = Chunk(cid) if it2 is not None else None
        for _ in range(chunk_size):
            try:
                rec1 = next(it1)
                rec2 = next(it2) if it2 is not None else None
            except StopIteration:
                break
            c1.headers.append(rec1[0]); c1.seqs.append(rec1[1]); c1.pluses.append(rec1[2]); c1.quals.append(rec1[3])
            if c2 is not None:
                c2.headers.append(rec2[0]); c2.seqs.append(rec2[1]); c2.pluses.append(rec2[2]); c2.quals.append(rec2[3])
        if len(c1) == 0:
            return
        c1.mate = c2
        yield c1
        cid += 1
        if len(c1) < chunk_size:
            return


def pack_reads(seqs: Sequence[bytes]):
    """Packed batch layout of the C ABI: concatenated bytes + int32 offsets (n+1)."""
    off = np.zeros(len(seqs) + 1, dtype=np.int32)
    if len(seqs):
        off[1:] = np.cumsum([len(s) for s in seqs])
    blob = np.frombuffer(b"".join(seqs), dtype=np.uint8) if len(seqs) else np.zeros(0, np.uint8)
    return blob, off


def output_filename(cfg: DemuxConfig, status: int, bc1: int, bc2: int) -> str:
    """classification.jl:877-900 -- a pure function of (status, bc1, bc2)."""
    suffix = ".fastq.gz" if cfg.gzip_output else ".fastq"
    if status == UNKNOWN:
        return "unknown" + suffix
    if status == AMBIGUOUS:
        return "ambiguous_classification" + suffix
    if cfg.is_dual:
        return f"{cfg.ids[bc1 - 1]}.{cfg.ids2[bc2 - 1]}{suffix}"
    return f"{cfg.ids[bc1 - 1]}{suffix}"


class Writer:
    """writer_task (core.jl:118-224): append-mode handle cache, trimming of read 1,
    routing of mates by the read-1 classification."""

    def __init__(self, output_dir: str, prefix1: str, prefix2: str, cfg: DemuxConfig):
        self.dir, self.p1, self.p2, self.cfg = output_dir, prefix1, prefix2, cfg
        self.handles: Dict[str, object] = {}
        self.do_trim = cfg.trim_side is not None or cfg.trim_side2 is not None

    def _handle(self, filename: str):
        h = self.handles.get(filename)
        if h is None:
            path = os.path.join(self.dir, filename)
            if self.cfg.gzip_output or path.lower().endswith(".gz"):
                h = gzip.open(path, "ab")
            else:
                h = open(path, "ab")
            self.handles[filename] = h
        return h

    def write_chunk(self, chunk: Chunk, results: np.ndarray):
        cfg = self.cfg
        for i in range(len(chunk)):
            st, b1, b2 = int(results["status"][i]), int(results["bc1"][i]), int(results["bc2"][i])
            filename = output_filename(cfg, st, b1, b2)
            h1, s1, p1, q1 = chunk.headers[i], chunk.seqs[i], chunk.pluses[i], chunk.quals[i]
            # trim_ranges[i] === nothing when t_start == -1 (core.jl:250-254)
            if self.do_trim and int(results["keep_start"][i]) != -1:
                lo = max(int(results["keep_start"][i]), 1)
                hi = min(int(results["keep_end"][i]), len(s1))
                if lo <= hi:
                    s1, q1 = s1[lo - 1:hi], q1[lo - 1:hi]
                else:
                    s1, q1 = b"", b""
            mate = chunk.mate
            if cfg.classify_both and mate is not None:
                self._handle(self.p1 + "." + filename).write(b"\n".join((h1, s1, p1, q1, b"")))
                self._handle(self.p2 + "." + filename).write(
                    b"\n".join((mate.headers[i], mate.seqs[i], mate.pluses[i], mate.quals[i], b"")))
            elif mate is not None:
                self._handle(self.p2 + "." + filename).write(
                    b"\n".join((mate.headers[i], mate.seqs[i], mate.pluses[i], mate.quals[i], b"")))
            else:
                self._handle(self.p1 + "." + filename).write(b"\n".join((h1, s1, p1, q1, b"")))

    def close(self):
        for h in self.handles.values():
            h.close()
        self.handles.clear()


def run_pipeline(cfg: DemuxConfig, classify_chunk: Callable[[Sequence[bytes]], np.ndarray],
                 fastq1: str, fastq2: Optional[str], output_dir: str, prefix1: str, prefix2: str,
                 chunk_size: int = 4000):
    """reader -> classify -> writer, in chunk order (core.jl:139-148)."""
    os.makedirs(output_dir, exist_ok=True)
    w = Writer(output_dir, prefix1, prefix2, cfg)
    try:
        for chunk in iter_chunks(fastq1, fastq2, chunk_size):
            w.write_chunk(chunk, classify_chunk(chunk.seqs))
    finally:
        w.close()


class _BlockReader:
    """Blocks of (inflated) FASTQ text with the unconsumed tail carried over to the next block."""

    def __init__(self, path: Optional[str], block_bytes: int):
        from .fileio import smart_open
        self.fh = smart_open(path, "r") if path is not None else None
        self.block_bytes = block_bytes
        self.tail = b""
        self.eof = self.fh is None

    def next_block(self) -> np.ndarray:
        data = b""
        if not self.eof and len(self.tail) < self.block_bytes:
            data = self.fh.read(self.block_bytes)
            if not self.fh.peek(1):
                self.eof = True
        buf = self.tail + data
        self.tail = b""
        return np.frombuffer(buf, dtype=np.uint8)

    def keep_tail(self, buf: np.ndarray, consumed: int):
        self.tail = buf[consumed:].tobytes()

    def close(self):
        if self.fh is not None:
            self.fh.close()


def run_pipeline_device(cfg: DemuxConfig, stream, fastq1: str, fastq2: Optional[str], output_dir: str,
                        prefix1: str, prefix2: str, block_bytes: int = 32 << 20):
    """The host side of the device FASTQ path (``bdx_demux_block``): read blocks of text, hand them to the
    GPU, append every returned bucket to its file.  Replaces reader_task's record splitting
    (core.jl:43-110) and writer_task's per-record trimming and routing (core.jl:118-224); file naming and
    append semantics are the reference's."""
    from .capi import DEMUX_BOTH, DEMUX_MATES, DEMUX_SINGLE

    os.makedirs(output_dir, exist_ok=True)
    paired = fastq2 is not None
    mode = DEMUX_SINGLE if not paired else (DEMUX_BOTH if cfg.classify_both else DEMUX_MATES)
    r1, r2 = _BlockReader(fastq1, block_bytes), _BlockReader(fastq2, block_bytes)
    handles: Dict[str, object] = {}

    def handle(filename: str):
        h = handles.get(filename)
        if h is None:
            path = os.path.join(output_dir, filename)
            h = gzip.open(path, "ab") if (cfg.gzip_output or path.lower().endswith(".gz")) else open(path, "ab")
            handles[filename] = h
        return h

    try:
        while True:
            b1 = r1.next_block()
            b2 = r2.next_block() if paired else None
            if b1.size == 0 or (paired and b2.size == 0):
                break       # `while !eof(io1) && !eof(io2)` (core.jl:48): an exhausted input ends the run
            final = (1 if r1.eof else 0) | (2 if paired and r2.eof else 0)
            out = stream.demux_block(b1, b2, final_block=final, mode=mode)
            buckets, o1, o2, _ = stream.demux_views(out)
            for bk in buckets:
                name = output_filename(cfg, int(bk["status"]), int(bk["bc1"]), int(bk["bc2"]))
                if mode != DEMUX_MATES:
                    handle(prefix1 + "." + name).write(o1[bk["offset1"]:bk["offset1"] + bk["length1"]].tobytes())
                if mode != DEMUX_SINGLE:
                    handle(prefix2 + "." + name).write(o2[bk["offset2"]:bk["offset2"] + bk["length2"]].tobytes())
            r1.keep_tail(b1, out.consumed1)
            if paired:
                r2.keep_tail(b2, out.consumed2)
            if out.n_records == 0:
                if r1.eof and r2.eof:
                    break
                r1.block_bytes *= 2     # a record longer than the block: take more text next time
                r2.block_bytes *= 2
    finally:
        r1.close()
        r2.close()
        for h in handles.values():
            h.close()


def _strip_fastq_ext(path: str) -> str:
    return re.sub(r"\.fastq(\.gz)?$", "", os.path.basename(path))


def execute_demultiplexing(fastq1: str, a: str, b: str, c: Optional[str] = None, *,
                           barcode_file2: Optional[str] = None,
                           output_prefix: str = "", output_prefix1: str = "", output_prefix2: str = "",
                           gzip_output: Optional[bool] = None,
                           max_error_rate: float = 0.2, min_delta: float = 0.0,
                           match: int = 0, mismatch: int = 1, indel: int = 1, nindel: Optional[int] = None,
                           classify_both: bool = False, bc_complement: bool = False, bc_rev: bool = False,
                           ref_search_range: str = "1:end", barcode_start_range: str = "1:end",
                           barcode_end_range: str = "1:end", ref_search_range2: str = "1:end",
                           barcode_start_range2: str = "1:end", barcode_end_range2: str = "1:end",
                           chunk_size: int = 4000, channel_capacity: int = 64,
                           trim_side: Optional[int] = None, trim_side2: Optional[int] = None,
                           summary: bool = False, summary_format: str = "html",
                           matching_algorithm: str = "semiglobal", log: bool = False,
                           device: int = 0, device_io: bool = False, block_bytes: int = 32 << 20):
    """Both reference methods (core.jl:360-392 paired, :500-529 single):

        execute_demultiplexing(FASTQ_file, barcode_file, output_directory; ...)
        execute_demultiplexing(FASTQ_file1, FASTQ_file2, barcode_file, output_directory; ...)

    Returns the DemuxStats-like dict when ``summary`` is set, else None.

    Reporting is out of scope here (SURVEY.md section 2, rows 11-12 stay with the Julia host): the reference
    writes summary.{html,txt,json} / stdout through generate_summary_report (reporting.jl:567-577) and a run
    log to stderr with ``log=true``; this mirror returns the counters instead and says so when either is asked for.
    """
    import warnings
    from .capi import Engine  # needs libbdx + a CUDA device; fails loudly otherwise

    if summary:
        warnings.warn(f"summary=True: the DemuxStats counters are returned; the reference's summary.{summary_format} "
                      "report is written by the Julia host (reporting.jl), not by this mirror", stacklevel=2)
    if log:
        warnings.warn("log=True is ignored: the run log is printed by the Julia host (core.jl:531-543)", stacklevel=2)

    if c is None:
        fastq2, barcode_file, output_directory = None, a, b
        prefix1 = output_prefix or output_prefix1 or _strip_fastq_ext(fastq1)
        prefix2 = ""
        fastqs = [fastq1]
        classify_both = False  # core.jl:559
    else:
        fastq2, barcode_file, output_directory = a, b, c
        prefix1 = output_prefix1 or _strip_fastq_ext(fastq1)
        prefix2 = output_prefix2 or _strip_fastq_ext(fastq2)
        fastqs = [fastq1, fastq2]
    cfg = build_config(barcode_file, barcode_file2, fastqs, gzip_output, bc_complement, bc_rev,
                       classify_both, max_error_rate, min_delta, match, mismatch, indel, nindel,
                       ref_search_range, barcode_start_range, barcode_end_range,
                       ref_search_range2, barcode_start_range2, barcode_end_range2,
                       trim_side, trim_side2, summary, summary_format, matching_algorithm)
    with Engine(cfg, device=device, max_reads=chunk_size) as eng:
        if device_io:
            # FASTQ text in, per-file record runs out, all on the device (bdx_demux_block)
            run_pipeline_device(cfg, eng.stream, fastq1, fastq2, output_directory, prefix1, prefix2, block_bytes)
        else:
            run_pipeline(cfg, eng.classify_reads, fastq1, fastq2, output_directory, prefix1, prefix2, chunk_size)
        if cfg.summary:
            return eng.demux_stats()
    return None
