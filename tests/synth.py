"""Seeded synthetic reads / barcode sets for the parity tests (numpy, CPU)."""
from __future__ import annotations

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_barcodes(rng, n, min_len, max_len=None, alphabet=b"ACGT", n_frac=0.0):
    max_len = max_len or min_len
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    out = []
    for _ in range(n):
        m = int(rng.integers(min_len, max_len + 1))
        b = alpha[rng.integers(0, len(alpha), m)].copy()
        if n_frac > 0:
            b[rng.random(m) < n_frac] = ord("N")
        out.append(bytes(b).decode("latin-1"))
    return out


def mutate(rng, s: bytes, k: int) -> bytes:
    s = bytearray(s)
    for _ in range(k):
        kind = int(rng.integers(0, 3))
        if kind == 0 and len(s):
            s[int(rng.integers(0, len(s)))] = int(BASES[rng.integers(0, 4)])
        elif kind == 1:
            s.insert(int(rng.integers(0, len(s) + 1)), int(BASES[rng.integers(0, 4)]))
        elif len(s) > 1:
            del s[int(rng.integers(0, len(s)))]
    return bytes(s)


def random_reads(rng, n, barcodes, barcodes2=None, min_len=150, max_len=None, plant=0.9,
                 start_hi=None, max_edits=5, n_prob=0.01, lower_prob=0.0, at_end2=True):
    """Reads over ACGT with a mutated barcode planted; optionally a second-set barcode near the end."""
    max_len = max_len or min_len
    reads = []
    for _ in range(n):
        L = int(rng.integers(min_len, max_len + 1))
        r = bytearray(BASES[rng.integers(0, 4, L)].tobytes())
        if L and rng.random() < plant:
            bc = barcodes[int(rng.integers(0, len(barcodes)))].encode("latin-1")
            k = int(rng.choice([0, 0, 0, 1, 1, 2, 3, 4, max_edits]))
            mb = mutate(rng, bc.replace(b"N", b"A"), k)
            hi = start_hi if start_hi is not None else max(L - len(mb), 0)
            st = int(rng.integers(0, max(hi, 0) + 1))
            r[st:st + len(mb)] = mb
            r = r[:L]
            if barcodes2:
                bc2 = barcodes2[int(rng.integers(0, len(barcodes2)))].encode("latin-1")
                mb2 = mutate(rng, bc2.replace(b"N", b"A"), int(rng.choice([0, 0, 1, 2, 3])))
                if at_end2:
                    gap = int(rng.integers(0, 6))
                    st2 = max(L - gap - len(mb2), 0)
                else:
                    st2 = int(rng.integers(0, max(L - len(mb2), 0) + 1))
                r[st2:st2 + len(mb2)] = mb2
                r = r[:L]
        if L and rng.random() < n_prob:
            r[int(rng.integers(0, L))] = ord("N")
        if L and lower_prob and rng.random() < lower_prob:
            p = int(rng.integers(0, L))
            r[p] = ord(chr(r[p]).lower())
        reads.append(bytes(r))
    return reads
