"""The two-bit wire format of the host path (GM2_CFG_WIRE): its host-side decoder alone, on the CPU.

gm2_emit_host may ship kept bases as 2 bits each and expand them on the host (host_expand.cpp).  Here
the wire format is produced by numpy from a known image, exactly as include/gm2.h describes it, and the
decoder must give back the image byte for byte — scalar, AVX2 and AVX-512 VBMI forms (a level the CPU
lacks falls back to the next one down), any destination alignment, one and many threads.  (The GPU side of the format, k_emit_packed, is covered by the -m gpu tests, which
compare both transports with the oracle.)"""
import ctypes

import numpy as np
import pytest

from genome_minimizer_2_b200 import _native

PREFIX = "Minimized_E_coli_K12_MG1655_"
CODE = np.full(256, 255, dtype=np.uint8)
for i, ch in enumerate(b"ACGT"):
    CODE[ch] = i


def _pack_chunk(seqs, tile_lens, ntiles, first_idx):
    """seqs[i]: bytes of sample i's kept bases; tile_lens[i][t]: how many of them came from tile t."""
    S = len(seqs)
    lengths = np.array([len(s) for s in seqs], dtype=np.int64)
    rec_sizes = np.array([1 + len(PREFIX) + len(str(first_idx + i + 1)) + 1 + len(seqs[i]) + 1 for i in range(S)], dtype=np.int64)
    rec_off = np.zeros(S + 1, dtype=np.int64)
    rec_off[1:] = np.cumsum(rec_sizes)
    tile_off = np.zeros((S, ntiles), dtype=np.int32)
    for i in range(S):
        tile_off[i, 1:] = np.cumsum(tile_lens[i])[:-1]
    words = int((rec_off[S] >> 4) + S * (ntiles + 2) + 8)
    rng = np.random.default_rng(99)
    packed = rng.integers(0, 2**32, words + 4, dtype=np.uint64).astype(np.uint32)      # garbage in the gaps: must not matter
    for i in range(S):
        codes = CODE[np.frombuffer(seqs[i], dtype=np.uint8)]
        for t in range(ntiles):
            a = int(tile_off[i, t]); n = int(tile_lens[i][t])
            if n == 0:
                continue
            w0 = int((rec_off[i] >> 4) + i * (ntiles + 2) + (a >> 4) + t)
            nw = (n + 15) // 16
            c = np.zeros(nw * 16, dtype=np.uint64)
            c[:n] = codes[a:a + n]
            vals = (c.reshape(nw, 16) << (2 * np.arange(16, dtype=np.uint64))).sum(axis=1)
            # bits past the piece's last base are unspecified: leave garbage there
            keep_hi = packed[w0 + nw - 1].astype(np.uint64) & ~np.uint64((1 << (2 * (n - 16 * (nw - 1)))) - 1) if n % 16 else np.uint64(0)
            packed[w0:w0 + nw] = vals.astype(np.uint32)
            if n % 16:
                packed[w0 + nw - 1] = np.uint32((int(vals[-1]) | int(keep_hi)) & 0xffffffff)
    image = b"".join(b">" + PREFIX.encode() + str(first_idx + i + 1).encode() + b"\n" + seqs[i] + b"\n" for i in range(S))
    return packed, tile_off, rec_off, lengths, image


def _decode(packed, tile_off, rec_off, lengths, ntiles, first_idx, threads, simd, misalign=0):
    lib = _native.load()
    total = int(rec_off[-1])
    raw = np.full(total + 128 + misalign, 0x2a, dtype=np.uint8)
    base = (-raw.ctypes.data) % 64 + misalign                   # choose the destination's alignment
    out = raw[base:base + total]
    rc = lib.gm2_diag_expand(packed.ctypes.data, tile_off.ctypes.data, rec_off.ctypes.data, lengths.ctypes.data,
                             len(lengths), ntiles, first_idx, PREFIX.encode(), out.ctypes.data, threads, simd)
    assert rc == 0, lib.gm2_last_error(None)
    assert raw[base + total:].tobytes() == b"\x2a" * (raw.size - base - total), "wrote past the image"
    assert raw[:base].tobytes() == b"\x2a" * base, "wrote before the image"
    return out.tobytes()


@pytest.mark.parametrize("simd", [0, 1, 2, 3])
@pytest.mark.parametrize("threads", [1, 3])
def test_decoder_round_trip(simd, threads):
    rng = np.random.default_rng(7 + simd + 10 * threads)
    for trial in range(12):
        ntiles = int(rng.integers(1, 9))
        S = int(rng.integers(1, 12))
        seqs, tile_lens = [], []
        for _ in range(S):
            tl = [int(x) for x in rng.choice([0, 1, 15, 16, 17, 31, 95, 96, 97, 128, 255, 256, 257, 320, 1000, 4097], size=ntiles)]
            if rng.random() < 0.2:
                tl = [0] * ntiles                                               # everything deleted: ">id\n\n"
            tile_lens.append(tl)
            seqs.append(bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=sum(tl))))
        first = int(rng.choice([0, 8, 99, 999_999]))
        packed, tile_off, rec_off, lengths, image = _pack_chunk(seqs, tile_lens, ntiles, first)
        for misalign in (0, 1, 13, 31, 33, 63):
            got = _decode(packed, tile_off, rec_off, lengths, ntiles, first, threads, simd, misalign)
            assert got == image, (trial, misalign)


def test_decoder_large_pieces_many_threads():
    rng = np.random.default_rng(3)
    ntiles, S = 5, 40
    tile_lens = [[int(x) for x in rng.integers(0, 50_000, ntiles)] for _ in range(S)]
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=sum(tl))) for tl in tile_lens]
    packed, tile_off, rec_off, lengths, image = _pack_chunk(seqs, tile_lens, ntiles, 0)
    for simd in (1, 2, 3):
        assert _decode(packed, tile_off, rec_off, lengths, ntiles, 0, 8, simd) == image
    assert _decode(packed, tile_off, rec_off, lengths, ntiles, 0, 1, 0) == image


def test_decoder_rejects_bad_arguments():
    lib = _native.load()
    assert lib.gm2_diag_expand(None, None, None, None, 1, 1, 0, b"x", None, 1, 1) == _native.ERR_INVALID
    assert lib.gm2_diag_expand(None, None, None, None, 0, 0, 0, b"x", None, 1, 1) == _native.ERR_INVALID
    assert lib.gm2_diag_expand(None, None, None, None, 0, 1, 0, b"x", None, 1, 1) == 0     # nothing to do
