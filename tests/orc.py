"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


class OrcRange(C.Structure):
    _fields_ = [("start_offset", C.c_int64), ("start_from_end", C.c_int32),
                ("end_offset", C.c_int64), ("end_from_end", C.c_int32)]


class OrcSet(C.Structure):
    _fields_ = [("n_bc", C.c_int32), ("bc_bytes", C.c_void_p), ("bc_off", C.c_void_p),
                ("bc_len_no_n", C.c_void_p), ("ref_search_range", OrcRange),
                ("barcode_start_range", OrcRange), ("barcode_end_range", OrcRange),
                ("trim_side", C.c_int32)]


class OrcConfig(C.Structure):
    _fields_ = [("max_error_rate", C.c_double), ("min_delta", C.c_double),
                ("match", C.c_int64), ("mismatch", C.c_int64), ("indel", C.c_int64), ("nindel", C.c_int64),
                ("has_nindel", C.c_int32), ("algorithm", C.c_int32), ("is_dual", C.c_int32),
                ("want_stats", C.c_int32), ("set1", OrcSet), ("set2", OrcSet)]


class OrcPass(C.Structure):
    _fields_ = [("status", C.c_int32), ("bc", C.c_int32), ("start", C.c_int64), ("end", C.c_int64),
                ("score", C.c_double)]


class OrcResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("bc1", C.c_int32), ("bc2", C.c_int32),
                ("keep_start", C.c_int64), ("keep_end", C.c_int64), ("passes", OrcPass * 2)]


PASS_DTYPE = np.dtype([("status", "<i4"), ("bc", "<i4"), ("start", "<i8"), ("end", "<i8"), ("score", "<f8")])
RESULT_DTYPE = np.dtype([
    ("status", "<i4"), ("bc1", "<i4"), ("bc2", "<i4"), ("_pad", "<i4"),
    ("keep_start", "<i8"), ("keep_end", "<i8"),
    ("passes", PASS_DTYPE, (2,)),
])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        L = C.CDLL(path)
        L.orc_semiglobal.restype = C.c_double
        L.orc_semiglobal.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_double,
                                     C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int64,
                                     C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                     C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_hamming.restype = C.c_double
        L.orc_hamming.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_double,
                                  C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                  C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_exact.restype = C.c_double
        L.orc_exact.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64,
                                C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_parse_dynamic_range.restype = C.c_int
        L.orc_parse_dynamic_range.argtypes = [C.c_char_p, C.POINTER(OrcRange)]
        L.orc_resolve.restype = None
        L.orc_resolve.argtypes = [C.POINTER(OrcRange), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_find_best.restype = C.c_int32
        L.orc_find_best.argtypes = [C.POINTER(OrcConfig), C.POINTER(OrcSet), C.c_char_p, C.c_int64,
                                    C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                    C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_determine.restype = None
        L.orc_determine.argtypes = [C.POINTER(OrcConfig), C.c_char_p, C.c_int64, C.POINTER(OrcResult)]
        L.orc_classify.restype = None
        L.orc_classify.argtypes = [C.POINTER(OrcConfig), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_round2.restype = C.c_double
        L.orc_round2.argtypes = [C.c_double]
        assert C.sizeof(OrcResult) == RESULT_DTYPE.itemsize, (C.sizeof(OrcResult), RESULT_DTYPE.itemsize)
        _LIB = L
    return _LIB


def _b(x) -> bytes:
    return x.encode("latin-1") if isinstance(x, str) else bytes(x)


def semiglobal(q, r, max_error, match=0, mismatch=1, indel=1, nindel=None, rng=None,
               max_start_pos=None, min_end_pos=1, norm=None, traceback=False, trim_side=None):
    """semiglobal_alignment[_N] (classification.jl:447-477). Returns score or (score,s,e)."""
    q, r = _b(q), _b(r)
    if rng is None:
        rng = (1, len(r))
    if max_start_pos is None:
        max_start_pos = len(r)
    if norm is None:
        norm = len(q)
    s, e = C.c_int64(), C.c_int64()
    ts = trim_side or 0
    score = lib().orc_semiglobal(q, len(q), r, len(r), max_error, match, mismatch, indel,
                                 0 if nindel is None else 1, nindel or 0, rng[0], rng[1],
                                 max_start_pos, min_end_pos, norm, int(traceback), ts, C.byref(s), C.byref(e))
    if traceback or ts:
        return score, s.value, e.value
    return score


def hamming(q, r, max_error_rate, rng, max_start_pos, min_end_pos, trim_side=None):
    q, r = _b(q), _b(r)
    s, e = C.c_int64(), C.c_int64()
    score = lib().orc_hamming(q, len(q), r, len(r), max_error_rate, rng[0], rng[1], max_start_pos,
                              min_end_pos, trim_side or 0, C.byref(s), C.byref(e))
    return score, s.value, e.value


def exact(q, r, rng, max_start_pos, min_end_pos, trim_side=None):
    q, r = _b(q), _b(r)
    s, e = C.c_int64(), C.c_int64()
    score = lib().orc_exact(q, len(q), r, len(r), rng[0], rng[1], max_start_pos, min_end_pos,
                            trim_side or 0, C.byref(s), C.byref(e))
    return score, s.value, e.value


def parse_dynamic_range(s: str):
    out = OrcRange()
    if lib().orc_parse_dynamic_range(s.encode(), C.byref(out)) != 0:
        raise ValueError(f"Invalid range format: {s}")
    return out


def resolve(dr: OrcRange, length: int):
    f, l = C.c_int64(), C.c_int64()
    lib().orc_resolve(C.byref(dr), length, C.byref(f), C.byref(l))
    return f.value, l.value


def _to_range(dr) -> OrcRange:
    if isinstance(dr, OrcRange):
        return dr
    if isinstance(dr, str):
        return parse_dynamic_range(dr)
    return OrcRange(dr.start_offset, int(dr.start_from_end), dr.end_offset, int(dr.end_from_end))


class Oracle:
    """Holds an orc_config (and the numpy buffers it points into)."""

    def __init__(self, cfg, want_stats: Optional[bool] = None):
        """cfg: a host DemuxConfig-like object (biodemux.jl_b200.config.DemuxConfig)."""
        self._keep = []
        c = OrcConfig()
        c.max_error_rate = cfg.max_error_rate
        c.min_delta = cfg.min_delta
        c.match, c.mismatch, c.indel = cfg.match, cfg.mismatch, cfg.indel
        c.has_nindel = 0 if cfg.nindel is None else 1
        c.nindel = cfg.nindel or 0
        c.algorithm = cfg.algorithm_code
        c.is_dual = int(cfg.is_dual)
        c.want_stats = int(cfg.summary if want_stats is None else want_stats)
        c.set1 = self._set(cfg.bc_seqs, cfg.bc_lengths_no_N, cfg.ref_search_range,
                           cfg.barcode_start_range, cfg.barcode_end_range, cfg.trim_side)
        c.set2 = self._set(cfg.bc_seqs2, cfg.bc_lengths_no_N2, cfg.ref_search_range2,
                           cfg.barcode_start_range2, cfg.barcode_end_range2, cfg.trim_side2)
        self.c = c

    def _set(self, seqs: Sequence[str], lens: Sequence[int], rs, bs, be, trim) -> OrcSet:
        s = OrcSet()
        bs_ = [_b(x) for x in seqs]
        blob = np.frombuffer(b"".join(bs_) + b"\0", dtype=np.uint8).copy()
        off = np.zeros(len(bs_) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(x) for x in bs_])
        ln = np.asarray(list(lens) if len(lens) else [0], dtype=np.int64)
        self._keep += [blob, off, ln]
        s.n_bc = len(bs_)
        s.bc_bytes = blob.ctypes.data
        s.bc_off = off.ctypes.data
        s.bc_len_no_n = ln.ctypes.data
        s.ref_search_range = _to_range(rs)
        s.barcode_start_range = _to_range(bs)
        s.barcode_end_range = _to_range(be)
        s.trim_side = trim or 0
        return s

    def classify(self, seq_bytes: np.ndarray, offsets: np.ndarray) -> np.ndarray:
        """seq_bytes uint8, offsets int (n+1).  Returns a RESULT_DTYPE array."""
        seq_bytes = np.ascontiguousarray(seq_bytes, dtype=np.uint8)
        off64 = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(off64) - 1
        out = np.zeros(n, dtype=RESULT_DTYPE)
        if seq_bytes.size == 0:
            seq_bytes = np.zeros(1, dtype=np.uint8)
        lib().orc_classify(C.byref(self.c), seq_bytes.ctypes.data, off64.ctypes.data, n, out.ctypes.data)
        return out

    def classify_reads(self, reads: Sequence[bytes]) -> np.ndarray:
        seq, off = pack_reads(reads)
        return self.classify(seq, off)

    def classify_mt(self, seq_bytes: np.ndarray, offsets: np.ndarray, threads: Optional[int] = None,
                    chunk: int = 4000) -> np.ndarray:
        """classify() over chunks of `chunk` reads on `threads` host threads (ctypes releases the GIL):
        the parity checks over >= 1 M reads use this."""
        from concurrent.futures import ThreadPoolExecutor
        off64 = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(off64) - 1
        out = np.zeros(n, dtype=RESULT_DTYPE)
        spans = [(a, min(a + chunk, n)) for a in range(0, n, chunk)]

        def work(span):
            a, b = span
            out[a:b] = self.classify(seq_bytes[off64[a]:off64[b]], off64[a:b + 1] - off64[a])

        with ThreadPoolExecutor(threads or os.cpu_count() or 1) as ex:
            list(ex.map(work, spans))
        return out

    def find_best(self, read, rng, max_start_pos, min_end_pos, need_traceback=False, pass2=False):
        r = _b(read)
        ms, dl = C.c_double(), C.c_double()
        s, e = C.c_int64(), C.c_int64()
        st = self.c.set2 if pass2 else self.c.set1
        bc = lib().orc_find_best(C.byref(self.c), C.byref(st), r, len(r), rng[0], rng[1],
                                 max_start_pos, min_end_pos, int(need_traceback),
                                 C.byref(ms), C.byref(dl), C.byref(s), C.byref(e))
        return bc, ms.value, dl.value, s.value, e.value


def pack_reads(reads: Sequence[bytes]):
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in reads])
    seq = np.frombuffer(b"".join(reads), dtype=np.uint8) if len(reads) else np.zeros(0, np.uint8)
    return seq, off


def round2(x: float) -> float:
    return lib().orc_round2(x)
