#!/usr/bin/env python
"""Store-only probes of the write patterns the emit kernel could use (B200, one GPU).
Prints GB/s for: a flat grid-stride fill, and k_emit's decomposition (one warp per 36 KB chunk of a
2.6 MB record) for several batch sizes / warps per CTA / CTA orders / store widths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genome_minimizer_2_b200 import _native

dev = torch.device("cuda", 0)
ctx = _native.Context(0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
nrec, ntile = 8000, 71
chunk = 36736                      # bytes per (record, tile): 1148 sectors
stride = ntile * chunk             # 2.6 MB records, contiguous
buf = torch.empty(nrec * stride, dtype=torch.uint8, device=dev)

def timed(fn, reps=5):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(st)
    for _ in range(reps): fn()
    b.record(st); torch.cuda.synchronize()
    return buf.numel() * reps / (a.elapsed_time(b) * 1e-3) / 1e9

print("flat fill           %8.0f GB/s" % timed(lambda: ctx.diag_fill(buf.data_ptr(), buf.numel())))
for warps in (8, 4):
    for batch in (8, 32, 128, 512):
        for order in (0, 1):
            for vec32 in (0, 1):
                r = timed(lambda: ctx.diag_fill_streams(buf.data_ptr(), nrec, stride, ntile, chunk, batch, warps, order, vec32))
                print("streams warps=%d batch=%3d order=%d st%d %8.0f GB/s" % (warps, batch, order, 256 if vec32 else 128, r))
# finer chunks: what if a warp only owned 4 KB at a time?
for chunk2, nt2 in ((4096, 640), (16384, 160), (131072, 20)):
    r = timed(lambda: ctx.diag_fill_streams(buf.data_ptr(), nrec, stride, nt2, chunk2, 128, 8, 0, 0))
    print("streams chunk=%6d ntile=%3d            %8.0f GB/s" % (chunk2, nt2, r))
