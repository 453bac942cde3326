"""Range mini-language of the reference ("start:end", ``end``, ``end-5``, ``+``).

Mirrors ``DynamicRange`` / ``parse_part`` / ``parse_dynamic_range`` / ``resolve``
(reference src/classification.jl:9-14, 61-100).  Host-side logic only: the
resolved ranges are evaluated again on the device per read length.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class DynamicRange:
    """classification.jl:9-14"""

    start_offset: int
    start_from_end: bool
    end_offset: int
    end_from_end: bool


def _parse_int(s: str) -> int:
    s = s.strip()
    if not s or not s.lstrip("+-").isdigit() or len(s) - len(s.lstrip("+-")) > 1:
        raise ValueError(f"invalid integer in range: {s!r}")
    return int(s)


def parse_part(s: str):
    """classification.jl:61-81 -- one side of ``start:end``."""
    s = s.strip()
    from_end = "end" in s
    if from_end:
        s = s.replace("end", "0")
    if "-" in s:
        p = s.split("-")
        val = _parse_int(p[0]) - _parse_int(p[1])
    elif "+" in s:
        p = s.split("+")
        val = _parse_int(p[0]) + _parse_int(p[1])
    else:
        val = _parse_int(s)
    return val, from_end


def parse_dynamic_range(range_str: str) -> DynamicRange:
    """classification.jl:83-94.  Raises like the reference's ``error(...)``."""
    parts = range_str.split(":")
    if len(parts) != 2:
        raise ValueError(f"Invalid range format: {range_str}. Expected 'start:end'.")
    so, sf = parse_part(parts[0])
    eo, ef = parse_part(parts[1])
    return DynamicRange(so, sf, eo, ef)


def resolve(dr: DynamicRange, length: int):
    """classification.jl:96-100.  Returns (first, last) of the Julia UnitRange,
    including its normalisation ``last = first - 1`` for empty ranges."""
    s = length + dr.start_offset if dr.start_from_end else dr.start_offset
    e = length + dr.end_offset if dr.end_from_end else dr.end_offset
    first = max(1, s)
    last = min(length, e)
    if last < first:
        last = first - 1
    return first, last
