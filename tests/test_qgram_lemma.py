"""The 3-gram filter of k_seed_var (csrc/seed_var.cu, sv_qgram_pass) restated on Python integers, and its lemma checked
by construction: whenever a barcode occurs in a read with <= K edits and one of its segments intact on diagonal
delta, the bit-plane test at (delta, K) passes.  A necessary condition must never reject such a hit -- the GPU tests
check that end to end (test_qgram_filter_is_transparent); this one checks the arithmetic itself, edge windows
included, without a GPU."""
import numpy as np

BASES = b"ACGT"


def planes_of_read(read: bytes):
    """(absent, bit0, bit1) as Python ints; plane bit 32 + c is column c, everything outside the read is absent."""
    n = len(read)
    a = (1 << (n + 128)) - 1
    p0 = p1 = 0
    for c, ch in enumerate(read):
        code = BASES.find(bytes([ch]))
        if code < 0:
            continue
        a &= ~(1 << (32 + c))
        p0 |= (code & 1) << (32 + c)
        p1 |= (code >> 1) << (32 + c)
    return a, p0, p1


def qgram_pass(read_planes, bc: bytes, K: int, delta: int) -> bool:
    m = len(bc)
    need = (m - 2) - 3 * K
    if need <= 0:
        return True
    a, p0, p1 = read_planes
    b0 = sum((BASES.index(bytes([ch])) & 1) << i for i, ch in enumerate(bc))
    b1 = sum((BASES.index(bytes([ch])) >> 1) << i for i, ch in enumerate(bc))
    rows = (1 << m) - 1
    biased = delta - K + 32
    assert biased >= 0
    cover = 0
    for k in range(2 * K + 1):
        sh = biased + k
        mis = ((p0 >> sh) ^ b0) | ((p1 >> sh) ^ b1) | (a >> sh)
        mt = ~mis & rows
        cover |= mt & (mt >> 1) & (mt >> 2)
    return bin(cover).count("1") >= need


def edited_occurrence(rng, bc: bytes, K: int):
    """Applies <= K edits to bc that leave one of its K + 1 segments (lengths as in tables.cu: m // (K + 1), the
    first m % (K + 1) one longer) untouched.  Returns (piece, segment offset in bc, its offset in piece)."""
    m = len(bc)
    n_seg, base, extra = K + 1, m // (K + 1), m % (K + 1)
    starts = [0]
    for i in range(n_seg):
        starts.append(starts[-1] + base + (1 if i < extra else 0))
    keep = int(rng.integers(0, n_seg))
    lo, hi = starts[keep], starts[keep + 1]
    out, seg_at = [], None
    edits = int(rng.integers(0, K + 1))
    # choose edit sites outside the kept segment
    sites = [i for i in range(m) if not (lo <= i < hi)]
    chosen = set(rng.choice(sites, size=min(edits, len(sites)), replace=False).tolist()) if sites and edits else set()
    for i, ch in enumerate(bc):
        if i == lo:
            seg_at = len(out)
        if i in chosen:
            kind = int(rng.integers(0, 3))
            if kind == 0:                                  # substitution
                out.append(BASES[(BASES.index(bytes([ch])) + int(rng.integers(1, 4))) % 4])
            elif kind == 1:                                # insertion into the read (before this base)
                out.append(BASES[int(rng.integers(0, 4))])
                out.append(ch)
            # kind == 2: deletion, nothing emitted
        else:
            out.append(ch)
    return bytes(out), lo, seg_at


def test_filter_never_rejects_a_true_occurrence():
    rng = np.random.default_rng(404)
    checked = 0
    for _ in range(60000):
        m = int(rng.integers(8, 33))
        K = int(rng.integers(0, min(6, m // 4) + 1))
        bc = bytes(BASES[i] for i in rng.integers(0, 4, m))
        piece, seg_off, seg_at = edited_occurrence(rng, bc, K)
        pre = bytes(BASES[i] for i in rng.integers(0, 4, int(rng.integers(0, 40))))
        post = bytes(BASES[i] for i in rng.integers(0, 4, int(rng.integers(0, 40))))
        if rng.random() < 0.3:
            pre = pre[:int(rng.integers(0, 3))]             # occurrences at the very start / end of the range
        if rng.random() < 0.3:
            post = post[:int(rng.integers(0, 3))]
        read = bytearray(pre + piece + post)
        for j in range(len(read)):                          # bytes that occur in no barcode, outside the piece
            if rng.random() < 0.02 and not (len(pre) <= j < len(pre) + len(piece)):
                read[j] = ord("N")
        delta = len(pre) + seg_at - seg_off                 # diagonal of the intact segment
        assert qgram_pass(planes_of_read(bytes(read)), bc, K, delta), (bc, bytes(read), K, delta)
        checked += 1
    assert checked == 60000


def test_filter_rejects_most_random_windows():
    """Selectivity, loosely: for 24-nt barcodes at K = 4 a random window passes in well under a third of the cases."""
    rng = np.random.default_rng(405)
    passed = 0
    for _ in range(4000):
        bc = bytes(BASES[i] for i in rng.integers(0, 4, 24))
        read = bytes(BASES[i] for i in rng.integers(0, 4, 60))
        passed += qgram_pass(planes_of_read(read), bc, 4, int(rng.integers(0, 30)))
    assert passed < 4000 // 3, passed
