#!/usr/bin/env python
"""Host-only probe of the two-bit decoder (csrc/host_expand.cpp) on a chunk shaped like the real ones: no GPU.

    python tools/host_decode_probe.py [--threads N] [--samples S] [--tile-bytes T] [--kept 0.56] [--simd 1]

Builds S records of a K-12-sized genome (pieces of ~kept*T bases per tile, random 2-bit payload: the
decoder's speed does not depend on the letters), decodes them with gm2_diag_expand and prints GB/s of image
written, beside the non-temporal fill rate of the same buffer with the same threads (gm2_diag_host_fill).
The packed words are read from memory that was written once and is far larger than the caches, as the
DMA-written staging buffers are."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from genome_minimizer_2_b200 import _native

ap = argparse.ArgumentParser()
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--samples", type=int, default=256)
ap.add_argument("--tile-bytes", type=int, default=36864)
ap.add_argument("--genome", type=int, default=4_641_652)
ap.add_argument("--kept", type=float, default=0.56)
ap.add_argument("--simd", type=int, default=1)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
lib = _native.load()
threads = a.threads or os.cpu_count()
prefix = b"Minimized_E_coli_K12_MG1655_"
nt = (a.genome + a.tile_bytes - 1) // a.tile_bytes
rng = np.random.default_rng(1)
S = a.samples
tile_len = np.maximum(rng.normal(a.kept, 0.08, (S, nt)) * a.tile_bytes, 0).astype(np.int64)
tile_len = np.minimum(tile_len, a.tile_bytes)
lengths = tile_len.sum(axis=1).astype(np.int64)
tile_off = np.zeros((S, nt), dtype=np.int32)
tile_off[:, 1:] = np.cumsum(tile_len, axis=1)[:, :-1]
hdr = np.array([1 + len(prefix) + len(str(i + 1)) + 1 for i in range(S)], dtype=np.int64)
rec_off = np.zeros(S + 1, dtype=np.int64)
rec_off[1:] = np.cumsum(hdr + lengths + 1)
total = int(rec_off[-1])
words = int((rec_off[S] >> 4) + S * (nt + 2) + 8)
packed = rng.integers(0, 2**32, words + 16, dtype=np.uint64).astype(np.uint32)
raw = np.zeros(total + 128, dtype=np.uint8)
out = raw[(-raw.ctypes.data) % 64:][:total]
out[:] = 1                                                    # fault the pages in
best = 0.0
for _ in range(a.reps):
    t0 = time.perf_counter()
    rc = lib.gm2_diag_expand(packed.ctypes.data, tile_off.ctypes.data, rec_off.ctypes.data, lengths.ctypes.data,
                             S, nt, 0, prefix, out.ctypes.data, threads, a.simd)
    dt = time.perf_counter() - t0
    assert rc == 0
    best = max(best, total / dt / 1e9)
fill = _native.host_fill_gbs(out, threads, 3)
print(f"threads {threads} simd {a.simd} tile {a.tile_bytes} pieces/sample {nt} mean piece {tile_len.mean():.0f} B: "
      f"decode {best:.1f} GB/s ({best/threads:.2f} per thread), NT fill {fill:.1f} GB/s ({fill/threads:.2f} per thread), ratio {best/fill:.2f}")
