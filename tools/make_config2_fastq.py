#!/usr/bin/env python
"""Writes config 2's workload as files the reference itself can read: <out>/reads.fastq (N x 150 bp reads of the
bench's distribution) and <out>/barcodes.csv (the bench's 96 barcodes; columns ID, Full_seq, Full_annotation).
For tools/cpu_baseline.jl.  Runs on the CPU (numpy generator):  python tools/make_config2_fastq.py out_dir [N]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def main():
    out = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    os.makedirs(out, exist_ok=True)
    cfg = bench.make_config()
    with open(os.path.join(out, "barcodes.csv"), "w") as fh:
        fh.write("ID,Full_seq,Full_annotation\n")
        for i, s in zip(cfg.ids, cfg.bc_seqs):
            fh.write(f"{i},{s},{'B' * len(s)}\n")
    qual = b"I" * bench.READ_LEN
    with open(os.path.join(out, "reads.fastq"), "wb") as fh:
        done = 0
        while done < n:
            k = min(200_000, n - done)
            blob, off = bench.numpy_reads(cfg, k, (bench.SEED ^ 0x5A5A5A) + done)   # not the barcodes' own stream
            raw = blob.tobytes()
            fh.write(b"".join(b"@r%d\n%s\n+\n%s\n" % (done + i, raw[off[i]:off[i + 1]], qual) for i in range(k)))
            done += k
    print(f"{n} reads -> {out}/reads.fastq, {len(cfg.bc_seqs)} barcodes -> {out}/barcodes.csv")


if __name__ == "__main__":
    main()
