#!/usr/bin/env python
"""Per-source-line summary of an ncu report's source page.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       python tools/ncu_source_lines.py src.csv [kernel-substring] [top-n]
Prints, for each kernel (launch) in the file, the source lines with the most stall samples and executed
instructions -- where a kernel's time goes, in terms of the .cu text."""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    rows = list(csv.reader(open(path)))
    fpath, func, hdr = "", "", None
    agg = defaultdict(lambda: defaultdict(lambda: [0, 0, ""]))     # func -> (file, line) -> [samples, inst, text]
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1]
        elif r[0] == "Function Name":
            func = r[1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0] != "" and r[0].isdigit():
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
            try:                                   # header text with embedded quotes (inline asm) splits oddly
                smp, ins = int(r[i_s] or 0), int(r[i_i] or 0)
            except (ValueError, IndexError):
                continue
            a = agg[func][(fpath.split("/")[-1], int(r[0]))]
            a[0] += smp
            a[1] += ins
            a[2] = r[1].strip()
    for func, lines in agg.items():
        if want not in func:
            continue
        tot_s = sum(v[0] for v in lines.values()) or 1
        tot_i = sum(v[1] for v in lines.values()) or 1
        print(f"== {func[:90]}  samples={tot_s} warp-inst={tot_i}")
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"{100 * v[0] / tot_s:5.1f}% smp {100 * v[1] / tot_i:5.1f}% inst  {f}:{ln}  {v[2][:100]}")


if __name__ == "__main__":
    main()
