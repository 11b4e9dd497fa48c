#!/usr/bin/env python
"""Store-only probes, part 2: does line (128 B) misalignment or run fragmentation cost bandwidth?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genome_minimizer_2_b200 import _native
dev = torch.device("cuda", 0)
ctx = _native.Context(0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
nrec, ntile, chunk = 8000, 71, 36736
stride = ntile * chunk
buf = torch.empty(nrec * stride, dtype=torch.uint8, device=dev)
def timed(fn, nbytes, reps=5):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(st)
    for _ in range(reps): fn()
    b.record(st); torch.cuda.synchronize()
    return nbytes * reps / (a.elapsed_time(b) * 1e-3) / 1e9
for mis in (0, 32, 64, 96):
    r = timed(lambda: ctx.diag_fill_streams(buf.data_ptr(), nrec, stride, ntile, chunk, 32, 8, 1, mis << 8), nrec * ntile * (chunk - mis))
    print("misalign %3d B                         %8.0f GB/s" % (mis, r))
for frag in (1216, 1184, 2432, 608, 4864):
    for mis in (0, 32):
        nfr = (chunk - mis) // frag
        for with_b in (0, 2):
            nbytes = nrec * ntile * nfr * ((frag - 32) // 16 * 16 + (32 if with_b else 0))
            r = timed(lambda: ctx.diag_fill_streams(buf.data_ptr(), nrec, stride, ntile, chunk, 32, 8, 1, (frag << 16) | (mis << 8) | with_b), nbytes)
            print("fragments of %4d B, misalign %2d, boundary sectors %s  %8.0f GB/s" % (frag, mis, "written" if with_b else "skipped", r))
