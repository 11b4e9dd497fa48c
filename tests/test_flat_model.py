"""A lane-by-lane Python model of `visit_flat` (csrc/k4_emit.cuh, FLAT == 2) checked against a plain gather.

There is no GPU in the build container, so the index arithmetic of the short-run emit path — packed run
table, events, vector bitmap, ownership of boundary vectors, the <= 3-source merge and its byte-loop
fallback, the partial first / last vector — is restated here statement by statement and run on random and
adversarial run structures.  The kernel itself is checked byte for byte against the oracle by the
`-m gpu` tests; this file keeps the arithmetic honest between GPU runs."""
from __future__ import annotations

import numpy as np
import pytest


def popc(x: int) -> int:
    return bin(x & 0xFFFFFFFF).count("1")


def model_visit_flat(tile: np.ndarray, runs, o: int):
    """runs: list of (src, length) kept runs in order (length > 0).  o: misalignment (0..15) of the visit's
    first output byte.  Returns the output bytes [o, r_end) as the kernel would write them, and a count of
    how many times each byte was written (must be exactly 1)."""
    bytes_ = sum(l for _, l in runs)
    r_end = o + bytes_
    out = np.zeros(r_end + 32, dtype=np.int32) - 1
    writes = np.zeros(r_end + 32, dtype=np.int32)
    nr = len(runs)
    A = []
    q = o
    for src, ln in runs:
        assert q < 65536 and src < 65536
        A.append(q | (src << 16))
        q += ln
    A.append(r_end)
    nbm = ((r_end - 1) >> 9) + 1
    vl = r_end >> 4
    # entry i = [event word, skip word] of row i; entries past the last row are skip-all (phase I works on
    # pairs of rows and reads one pair ahead)
    BM = [[0, 0xFFFFFFFF if i > (vl >> 5) else ((0xFFFFFFFF << (vl & 31)) & 0xFFFFFFFF) if i == (vl >> 5) else 0]
          for i in range(nbm + 1)]
    EV = []

    def lds_unaligned16(a: int):                      # a may be negative by < 16 (front pad) or run past the tile (back pad)
        return [int(tile[a + j]) if 0 <= a + j < tile.size else 0x100 + ((a + j) & 0xFF) for j in range(16)]

    def merge16(x, y, k):
        assert 0 <= k <= 16
        return x[:k] + y[k:]

    def st128(p0, vec):
        for j in range(16):
            out[p0 + j] = vec[j]
            writes[p0 + j] += 1

    def st8(pos, b):
        out[pos] = b
        writes[pos] += 1

    # ---- events + phase B
    ne = 0
    for r0 in range(0, nr, 32):
        lasts = []
        for lane in range(32):
            r = r0 + lane
            last = False
            x = 0
            if r < nr:
                x = A[r]
                R = x & 0xFFFF
                v = R >> 4
                p0 = v << 4
                x1 = A[r + 1]
                in1 = r + 1 < nr and ((x1 & 0xFFFF) >> 4) == v
                last = not in1
                xm = A[r - 1] if r > 0 else 0
                first = r == 0 or ((xm & 0xFFFF) >> 4) != v
                if first and p0 >= o and p0 + 16 <= r_end and (R > p0 or in1):
                    x2 = A[r + 2] if r + 2 <= nr else 0
                    in2 = in1 and r + 2 < nr and ((x2 & 0xFFFF) >> 4) == v
                    has_prev = R > p0
                    nsrc = (1 if has_prev else 0) + 1 + (1 if in1 else 0) + (1 if in2 else 0)
                    slow = nsrc > 3
                    if not slow and in2:
                        x3 = A[r + 3] if r + 3 <= nr else 0
                        slow = r + 3 < nr and ((x3 & 0xFFFF) >> 4) == v
                    if not slow:
                        sa, sb, sc = (xm, x, x1) if has_prev else (x, x1, x2)
                        X = lds_unaligned16((sa >> 16) + p0 - (sa & 0xFFFF))
                        Y = lds_unaligned16((sb >> 16) + p0 - (sb & 0xFFFF))
                        ov = merge16(X, Y, (sb & 0xFFFF) - p0)
                        if nsrc == 3:
                            Z = lds_unaligned16((sc >> 16) + p0 - (sc & 0xFFFF))
                            ov = merge16(ov, Z, (sc & 0xFFFF) - p0)
                    else:
                        rc = r - 1 if has_prev else r
                        xc = A[rc]
                        rn = A[rc + 1] & 0xFFFF
                        ov = []
                        for j in range(16):
                            while p0 + j >= rn:
                                rc += 1
                                xc = A[rc]
                                rn = A[rc + 1] & 0xFFFF
                            ov.append(int(tile[(xc >> 16) + p0 + j - (xc & 0xFFFF)]))
                    st128(p0, ov)
                if last:
                    BM[v >> 5][0] |= 1 << (v & 31)
                if R & 15:
                    BM[v >> 5][1] |= 1 << (v & 31)
                x = (x >> 16) - R                      # source offset of output byte 0 in run r's frame (kernel: + tile_a)
            lasts.append((last, x))
        bal = sum(1 << i for i, (l, _) in enumerate(lasts) if l)
        for lane, (l, x) in enumerate(lasts):
            if l:
                idx = ne + popc(bal & ((1 << lane) - 1))
                while len(EV) <= idx:
                    EV.append(None)
                EV[idx] = x
        ne += popc(bal)
    assert all(e is not None for e in EV) and len(EV) == ne

    # ---- phase I (same control flow as the kernel: two rows per step, unconditional source windows,
    # predicated stores, entries fetched one pair ahead)
    le = [(2 << lane) - 1 for lane in range(32)]
    evp = -1                                                 # index of EV[events of earlier rows - 1]
    pa = [lane << 4 for lane in range(32)]

    def rows2(mk):
        nonlocal evp
        ev0, sk0, ev1, sk1 = mk
        n0 = popc(ev0)
        for lane in range(32):
            ea0 = EV[evp + popc(ev0 & le[lane])]
            ea1 = EV[evp + n0 + popc(ev1 & le[lane])]
            v0 = lds_unaligned16(ea0 + pa[lane])
            v1 = lds_unaligned16(ea1 + pa[lane] + 512)
            if not (sk0 >> lane) & 1:
                st128(pa[lane], v0)
            if not (sk1 >> lane) & 1:
                st128(pa[lane] + 512, v1)
            pa[lane] += 1024
        evp += n0 + popc(ev1)

    def load4(i):
        assert i + 1 < len(BM)
        return BM[i][0], BM[i][1], BM[i + 1][0], BM[i + 1][1]

    # the kernel looks one pair ahead (events) and two pairs ahead (bitmap entries); what it reads there is
    # never used, but it must stay inside the per-warp region: entries up to nbm + 4
    for i in range(0, nbm, 2):
        rows2(load4(i))

    # ---- phase E
    for lane in range(32):
        pos = -1
        if lane < 16:
            if o:
                pos = lane
        elif (r_end & 15) and not (o and vl == 0):
            pos = (vl << 4) + (lane - 16)
        if pos >= o and pos < r_end:
            if lane < 16:
                rc = 0
                xc = A[0]
                rn = A[1] & 0xFFFF
                while pos >= rn:
                    rc += 1
                    xc = A[rc]
                    rn = A[rc + 1] & 0xFFFF
            else:
                rc = nr - 1
                xc = A[rc]
                while pos < (xc & 0xFFFF):
                    rc -= 1
                    xc = A[rc]
            st8(pos, int(tile[(xc >> 16) + pos - (xc & 0xFFFF)]))
    return out[o:r_end], writes[o:r_end], writes


def reference_gather(tile, runs):
    return np.concatenate([tile[s:s + l] for s, l in runs]).astype(np.int32) if runs else np.zeros(0, np.int32)


def random_runs(rng, tile_bytes, style):
    """Kept runs of a tile: increasing, non-adjacent source intervals (adjacent kept segments are one run)."""
    runs = []
    pos = int(rng.integers(0, 40))
    while pos < tile_bytes - 1:
        if style == "tiny":
            ln = int(rng.integers(1, 9))
            gap = int(rng.integers(1, 6))
        elif style == "mixed":
            ln = int(rng.choice([1, 2, 5, 15, 16, 17, 31, 32, 33, 100, 126, 500, 512, 513, 1200, 3000]))
            gap = int(rng.choice([1, 3, 900, 2000]))
        else:  # gene-like, ~10 % retention: short intergenic runs, now and then a kept gene
            ln = int(rng.exponential(126)) + 1 if rng.random() > 0.1 else int(rng.integers(600, 2500))
            gap = int(rng.integers(200, 1500))
        ln = min(ln, tile_bytes - pos)
        if ln <= 0:
            break
        runs.append((pos, ln))
        pos += ln + gap
    return runs


@pytest.mark.parametrize("style", ["gene", "mixed", "tiny"])
def test_visit_flat_model_writes_every_byte_once(style):
    rng = np.random.default_rng({"gene": 1, "mixed": 2, "tiny": 3}[style])
    tile_bytes = 49152 if style != "tiny" else 4096
    tile = rng.integers(0, 256, tile_bytes, dtype=np.uint8)
    for trial in range(60 if style != "tiny" else 200):
        runs = random_runs(rng, tile_bytes, style)
        if trial % 7 == 0:
            runs = runs[:int(rng.integers(1, 4))]            # very small visits: everything in one or two vectors
        if not runs:
            continue
        o = int(rng.integers(0, 16)) if trial % 3 else 0
        got, w, w_all = model_visit_flat(tile, runs, o)
        exp = reference_gather(tile, runs)
        assert np.array_equal(w, np.ones_like(w)), (style, trial, o, np.flatnonzero(w != 1)[:8])
        assert w_all[:o].sum() == 0 and w_all[o + exp.size:].sum() == 0      # nothing outside the visit's bytes
        assert np.array_equal(got, exp), (style, trial, o, np.flatnonzero(got != exp)[:8])


def test_visit_flat_model_edge_shapes():
    tile = np.arange(4096, dtype=np.int64).astype(np.uint8)
    cases = [
        ([(0, 1)], 0), ([(0, 1)], 15), ([(5, 16)], 0), ([(5, 16)], 1), ([(0, 15), (20, 1)], 0),
        ([(0, 16), (20, 16)], 0), ([(0, 512)], 0), ([(3, 512)], 7), ([(0, 1), (2, 1), (4, 1), (6, 1), (8, 1), (10, 1)], 0),
        ([(0, 1), (2, 1), (4, 1), (6, 1), (8, 1), (10, 1)] * 1, 13), ([(i * 3, 2) for i in range(200)], 5),
        ([(0, 1024), (2000, 3), (2010, 1024)], 0), ([(0, 33), (100, 7), (200, 8), (300, 600)], 9),
        ([(0, 4), (8, 4), (16, 4), (24, 4), (40, 1000)], 0), ([(0, 4), (8, 4), (16, 4), (24, 5), (40, 1000)], 15),
    ]
    for runs, o in cases:
        got, w, _ = model_visit_flat(tile, runs, o)
        assert np.array_equal(w, np.ones_like(w)), (runs[:4], o)
        assert np.array_equal(got, reference_gather(tile, runs)), (runs[:4], o)


# ----------------------------------------------------------------------------------------------
# table A of visit_flat for tiles of at most 128 slots: lane <-> four consecutive slots, one warp scan
# over {kept bytes : 20 bits | run starts : 12 bits}, at most two run starts per lane
# ----------------------------------------------------------------------------------------------
def model_table_a_four_slots(lens, srcs, words, o):
    """lens/srcs: 128 slot entries (entries past the tile's slots hold garbage), words: 4 kept-bit words."""
    nib, st, p, tot, mine = [0] * 32, [0] * 32, [None] * 32, [0] * 32, [0] * 32
    for lane in range(32):
        nib[lane] = (words[lane >> 3] >> (4 * (lane & 7))) & 0xF
    for lane in range(32):
        up = nib[lane - 1] if lane else 0
        st[lane] = nib[lane] & ~((nib[lane] << 1) | (up >> 3)) & 0xFFFFFFFF
        l4 = lens[4 * lane:4 * lane + 4]
        p1 = l4[0] if nib[lane] & 1 else 0
        p2 = p1 + (l4[1] if nib[lane] & 2 else 0)
        p3 = p2 + (l4[2] if nib[lane] & 4 else 0)
        tot[lane] = p3 + (l4[3] if nib[lane] & 8 else 0)
        p[lane] = (0, p1, p2, p3)
        assert st[lane] < 16 and popc(st[lane]) <= 2
        mine[lane] = tot[lane] | (popc(st[lane]) << 20)
    incl, acc = [], 0
    for lane in range(32):
        acc += mine[lane]
        incl.append(acc)
    assert (incl[31] & 0xFFFFF) < (1 << 20)
    nr = incl[31] >> 20
    A = [None] * (nr + 1)
    for lane in range(32):
        if st[lane]:
            excl = incl[lane] - mine[lane]
            q = o + (excl & 0xFFFFF)
            ra = excl >> 20
            j0 = (st[lane] & -st[lane]).bit_length() - 1
            A[ra] = (q + p[lane][j0]) | (srcs[4 * lane + j0] << 16)
            if st[lane] & (st[lane] - 1):
                at3 = bool(st[lane] & 8)
                A[ra + 1] = (q + (p[lane][3] if at3 else p[lane][2])) | (srcs[4 * lane + (3 if at3 else 2)] << 16)
    A[nr] = o + (incl[31] & 0xFFFFF)
    return A


def test_table_a_four_slot_form_equals_the_word_by_word_form():
    rng = np.random.default_rng(11)
    for trial in range(400):
        nslots = int(rng.choice([32, 64, 96, 128]))
        lens = rng.integers(0, 700, 128).tolist()
        srcs = np.concatenate([[0], np.cumsum(lens[:-1])]).tolist()
        dens = float(rng.choice([0.05, 0.3, 0.5, 0.9, 1.0]))
        bits = (rng.random(128) < dens).astype(int)
        bits[nslots:] = 0                                   # words past the tile are 0
        for i in range(nslots, 128):
            lens[i] = int(rng.integers(0, 1 << 20))         # garbage the kernel may read and must ignore
        words = [sum(int(bits[32 * c + i]) << i for i in range(32)) for c in range(4)]
        o = int(rng.integers(0, 16))
        # word-by-word form (the kernel's general path)
        A, q, carry = [], o, 0
        for c in range(4):
            w = words[c]
            starts = w & ~((w << 1) | carry) & 0xFFFFFFFF
            carry = w >> 31
            for lane in range(32):
                if (starts >> lane) & 1:
                    A.append(q | (srcs[32 * c + lane] << 16))
                if (w >> lane) & 1:
                    q += lens[32 * c + lane]
        A.append(q)
        if q >= 65536:
            continue
        assert model_table_a_four_slots(lens, srcs, words, o) == A, trial
