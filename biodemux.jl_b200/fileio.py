"""Barcode-table and FASTQ file helpers of the host pipeline.

Mirrors ``preprocess_bc_file`` / ``smart_open`` (reference src/fileio.jl:7-113).
These are the data formats either side of the hot path; the
hot path itself only ever sees the preprocessed byte strings.
"""
from __future__ import annotations

import csv
import gzip
import io
import re
from typing import Iterator, List, Tuple

_COMPLEMENT = {
    "A": "T", "T": "A", "G": "C", "C": "G",
    "a": "t", "t": "a", "g": "c", "c": "g",
    "N": "N", "n": "n",
}


def preprocess_bc_file(bc_file: str, complement: bool, rev: bool) -> Tuple[List[str], List[int], List[str]]:
    """fileio.jl:7-72 -> ``(sequences, lengths_no_N, ids)``."""
    sequences: List[str] = []
    ids: List[str] = []
    low = bc_file.lower()
    if low.endswith(".fasta") or low.endswith(".fa"):
        current = ""
        with open(bc_file, "r") as fh:
            for line in fh:
                line = line.rstrip("\n").rstrip("\r")
                if line.startswith(">"):
                    if current:
                        sequences.append(current)
                        current = ""
                    ids.append(re.sub(r"\s.*$", "", line[1:].strip()))
                else:
                    current += line.strip()
            if current:
                sequences.append(current)
        annotations = ["B" * len(s) for s in sequences]
    else:
        delim = "," if low.endswith(".csv") else "\t"
        with open(bc_file, "r", newline="") as fh:
            rows = list(csv.reader(fh, delimiter=delim))
        header = [h.strip() for h in rows[0]]
        try:
            i_seq, i_id, i_ann = header.index("Full_seq"), header.index("ID"), header.index("Full_annotation")
        except ValueError as exc:
            raise KeyError(f"barcode table {bc_file} needs columns Full_seq, ID, Full_annotation") from exc
        body = [r for r in rows[1:] if r and any(c != "" for c in r)]
        sequences = [r[i_seq] for r in body]
        ids = [r[i_id] for r in body]
        annotations = [r[i_ann] for r in body]

    for i in range(len(sequences)):
        if len(sequences[i]) != len(annotations[i]):
            raise ValueError(f"Length mismatch between sequence and annotation for ID: {ids[i]}")
        sequences[i] = "".join(c for c, a in zip(sequences[i], annotations[i]) if a == "B")

    sequences = [s.upper().replace("U", "T") for s in sequences]
    if complement:
        sequences = ["".join(_COMPLEMENT.get(c, c) for c in s) for s in sequences]
    if rev:
        sequences = [s[::-1] for s in sequences]
    lengths_no_n = [sum(1 for c in s if c != "N") for s in sequences]
    return sequences, lengths_no_n, ids


def smart_open(path: str, mode: str):
    """fileio.jl:77-95 -- gzip by file extension."""
    if path.lower().endswith(".gz"):
        return gzip.open(path, mode + "b")
    return open(path, mode + "b")


def _readline(fh) -> bytes:
    """Julia ``readline``: strips one trailing ``\\n`` or ``\\r\\n``."""
    line = fh.readline()
    if line.endswith(b"\n"):
        line = line[:-1]
        if line.endswith(b"\r"):
            line = line[:-1]
    return line


def fastq_records(path: str) -> Iterator[Tuple[bytes, bytes, bytes, bytes]]:
    """4x readline per record until EOF (core.jl:96-101)."""
    with smart_open(path, "r") as raw:
        fh = io.BufferedReader(raw) if not isinstance(raw, io.BufferedReader) else raw
        while True:
            if not fh.peek(1):
                return
            yield _readline(fh), _readline(fh), _readline(fh), _readline(fh)
