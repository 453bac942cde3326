"""DemuxStats (reference src/classification.jl:736-767) rebuilt from the device counters.

The device keeps integer histograms (include/bdx.h, ``bdx_stats_layout``); this module
turns them into the Dict-shaped structure ``merge_stats`` / ``generate_summary_report``
consume (src/reporting.jl:1-58), including the ``round(score, digits=2)`` keys of
classification.jl:835,853.  Summing counter buffers over streams / GPUs first (NCCL
all-reduce) is equivalent to ``merge_stats``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Tuple

import numpy as np


def julia_round2(x: float) -> float:
    """Base.round(x, digits=2): round(x * 100) / 100 with ties-to-even on the scaled value."""
    y = float(np.rint(x * 100.0)) / 100.0
    return y if np.isfinite(y) else x


@dataclass
class DemuxStats:
    total_reads: int = 0
    matched_reads: int = 0
    unmatched_reads: int = 0
    ambiguous_reads: int = 0
    sample_counts: Dict[Tuple[int, int], int] = field(default_factory=dict)
    bc1_pos_counts: Dict[int, int] = field(default_factory=dict)
    bc1_len_counts: Dict[int, int] = field(default_factory=dict)
    bc1_score_counts: Dict[float, int] = field(default_factory=dict)
    bc1_per_bc_score_counts: Dict[int, Dict[float, int]] = field(default_factory=dict)
    bc1_per_bc_pos_counts: Dict[int, Dict[int, int]] = field(default_factory=dict)
    bc1_per_bc_len_counts: Dict[int, Dict[int, int]] = field(default_factory=dict)
    bc2_pos_counts: Dict[int, int] = field(default_factory=dict)
    bc2_len_counts: Dict[int, int] = field(default_factory=dict)
    bc2_score_counts: Dict[float, int] = field(default_factory=dict)
    bc2_per_bc_score_counts: Dict[int, Dict[float, int]] = field(default_factory=dict)
    bc2_per_bc_pos_counts: Dict[int, Dict[int, int]] = field(default_factory=dict)
    bc2_per_bc_len_counts: Dict[int, Dict[int, int]] = field(default_factory=dict)


def _bump(d, k, c=1):
    d[k] = d.get(k, 0) + c


def pass_norms(cfg, pass2: bool):
    """Normalisation length per barcode (classification.jl:460, :476, :567)."""
    seqs = cfg.bc_seqs2 if pass2 else cfg.bc_seqs
    lens = cfg.bc_lengths_no_N2 if pass2 else cfg.bc_lengths_no_N
    if cfg.algorithm_code == 0 and cfg.nindel is not None:
        return [int(x) for x in lens]
    return [len(s.encode("latin-1")) if isinstance(s, str) else len(s) for s in seqs]


def merge_overflow(st: DemuxStats, overflow) -> DemuxStats:
    """Adds the exact (pass, barcode, start, length) records of ``bdx_stats_overflow_fetch`` -- matched passes whose
    position or length lies outside the device histograms -- to the position / length Dicts."""
    for e in overflow if overflow is not None else ():
        pre, b = f"bc{int(e['pass'])}", int(e["bc"])
        _bump(getattr(st, pre + "_pos_counts"), int(e["start"]))
        _bump(getattr(st, pre + "_len_counts"), int(e["length"]))
        _bump(getattr(st, pre + "_per_bc_pos_counts").setdefault(b, {}), int(e["start"]))
        _bump(getattr(st, pre + "_per_bc_len_counts").setdefault(b, {}), int(e["length"]))
    return st


def stats_from_counters(buf: np.ndarray, lay, cfg, overflow=None) -> DemuxStats:
    st = DemuxStats()
    st.total_reads, st.matched_reads, st.unmatched_reads, st.ambiguous_reads = (int(x) for x in buf[:4])
    b1, b2 = lay.b1, lay.b2
    samp = buf[lay.sample_off: lay.sample_off + (b1 + 1) * (b2 + 1)].reshape(b1 + 1, b2 + 1)
    for i, j in zip(*np.nonzero(samp)):
        st.sample_counts[(int(i), int(j))] = int(samp[i, j])
    for p, nb in ((0, b1), (1, b2)):
        if p == 1 and not cfg.is_dual:
            break
        pre = "bc1" if p == 0 else "bc2"
        norms = pass_norms(cfg, p == 1)
        pos = buf[lay.pos_off[p]: lay.pos_off[p] + (nb + 1) * lay.pos_bins].reshape(nb + 1, lay.pos_bins)
        ln = buf[lay.len_off[p]: lay.len_off[p] + (nb + 1) * lay.len_bins].reshape(nb + 1, lay.len_bins)
        ds = buf[lay.dist_off[p]: lay.dist_off[p] + (nb + 1) * lay.dist_bins].reshape(nb + 1, lay.dist_bins)
        g_pos, g_len, g_sc = getattr(st, pre + "_pos_counts"), getattr(st, pre + "_len_counts"), getattr(st, pre + "_score_counts")
        pb_pos, pb_len, pb_sc = (getattr(st, pre + "_per_bc_pos_counts"), getattr(st, pre + "_per_bc_len_counts"),
                                 getattr(st, pre + "_per_bc_score_counts"))
        for k in np.nonzero(pos[0])[0]:
            g_pos[int(k) - lay.pos_bias] = int(pos[0, k])
        for k in np.nonzero(ln[0])[0]:
            g_len[int(k)] = int(ln[0, k])
        for b in range(1, nb + 1):
            for k in np.nonzero(pos[b])[0]:
                _bump(pb_pos.setdefault(b, {}), int(k) - lay.pos_bias, int(pos[b, k]))
            for k in np.nonzero(ln[b])[0]:
                _bump(pb_len.setdefault(b, {}), int(k), int(ln[b, k]))
            for d in np.nonzero(ds[b])[0]:
                dist = int(d) - lay.dist_bias
                key = julia_round2(float(dist) / float(norms[b - 1])) if norms[b - 1] else float("nan")
                _bump(pb_sc.setdefault(b, {}), key, int(ds[b, d]))
                _bump(g_sc, key, int(ds[b, d]))
    return merge_overflow(st, overflow)


def stats_from_passes(status, bc1, bc2, passes, cfg) -> DemuxStats:
    """Same structure from per-read, per-pass records ``(status, bc, start, end, score)``
    (what determine_filename_and_stats sees, classification.jl:940-978)."""
    st = DemuxStats()
    for i in range(len(status)):
        st.total_reads += 1
        for p in (0, 1):
            ps, pb, s, e, score = passes[i][p]
            if ps != 0:
                continue
            pre = "bc1" if p == 0 else "bc2"
            r = julia_round2(float(score))
            _bump(getattr(st, pre + "_pos_counts"), int(s))
            _bump(getattr(st, pre + "_len_counts"), int(e - s + 1))
            _bump(getattr(st, pre + "_score_counts"), r)
            _bump(getattr(st, pre + "_per_bc_score_counts").setdefault(int(pb), {}), r)
            _bump(getattr(st, pre + "_per_bc_pos_counts").setdefault(int(pb), {}), int(s))
            _bump(getattr(st, pre + "_per_bc_len_counts").setdefault(int(pb), {}), int(e - s + 1))
        if status[i] == 0:
            st.matched_reads += 1
            _bump(st.sample_counts, (int(bc1[i]), int(bc2[i])))
        elif status[i] == 1:
            st.unmatched_reads += 1
        else:
            st.ambiguous_reads += 1
    return st


def stats_from_entries(buf: np.ndarray, lay, entries: np.ndarray) -> DemuxStats:
    """DemuxStats from the counter header + the entry list of ``bdx_stats_entries`` (the C report bridge)."""
    st = DemuxStats()
    st.total_reads, st.matched_reads, st.unmatched_reads, st.ambiguous_reads = (int(x) for x in buf[:4])
    samp = buf[lay.sample_off: lay.sample_off + (lay.b1 + 1) * (lay.b2 + 1)].reshape(lay.b1 + 1, lay.b2 + 1)
    for i, j in zip(*np.nonzero(samp)):
        st.sample_counts[(int(i), int(j))] = int(samp[i, j])
    names = {0: "pos", 1: "len", 2: "score"}
    for e in entries:
        pre, kind, b = f"bc{int(e['pass'])}", names[int(e["kind"])], int(e["bc"])
        key = float(e["score"]) if kind == "score" else int(e["key"])
        d = getattr(st, f"{pre}_{kind}_counts") if b == 0 else getattr(st, f"{pre}_per_bc_{kind}_counts").setdefault(b, {})
        _bump(d, key, int(e["count"]))
    return st
