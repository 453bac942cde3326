#!/usr/bin/env python
"""Extracts the hot inner loop of a kernel from `cuobjdump -sass` and counts its instructions per class.

usage: python tools/sass_loop.py <libbdx.so> <substring of the mangled kernel name> [max loop length] [mnemonic]
The loop = the backward branch whose body (at most `max` instructions) holds the most LOP3s (or `mnemonic`s); printed with a
per-mnemonic count so that the per-column instruction figures quoted in DESIGN.md can be checked."""
import re
import subprocess
import sys
from collections import Counter


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    max_len = int(sys.argv[3]) if len(sys.argv) > 3 else 260
    rank_by = sys.argv[4] if len(sys.argv) > 4 else "LOP3"
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        if pat not in name:
            continue
        ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", f)
        addr = [int(a, 16) for a, _ in ins]
        idx = {a: k for k, a in enumerate(addr)}
        best = None
        for k, (a, text) in enumerate(ins):
            m = re.search(r"BRA(?:\.\w+)* (?:\w+, )?0x([0-9a-f]+)", text)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt in idx and idx[tgt] < k and k - idx[tgt] <= max_len:
                body = ins[idx[tgt]:k + 1]
                lop = sum(rank_by in t for _, t in body)
                if best is None or lop > best[0]:
                    best = (lop, idx[tgt], k)
        print(f"== {name}")
        if not best:
            print("   (no loop found)")
            continue
        _, a, b = best
        body = ins[a:b + 1]
        cnt = Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in body)
        print(f"   loop 0x{addr[a]:04x}..0x{addr[b]:04x}: {len(body)} instructions: " +
              ", ".join(f"{k} {v}" for k, v in cnt.most_common()))
        for ad, t in body:
            print(f"   /*{ad}*/ {t};")


if __name__ == "__main__":
    main()
