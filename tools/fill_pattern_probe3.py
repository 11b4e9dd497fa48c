#!/usr/bin/env python
"""Store-only probes, part 3: boundary sectors written in-stream (right after their fragment)."""
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
names = {0: "skipped", 2: "afterwards (1 lane each)", 4: "in-stream, 1 lane st256", 8: "in-stream, 2 lanes st128"}
for frag in (1216, 608, 2432):
    for mode in (0, 2, 4, 8):
        nfr = chunk // frag
        nbytes = nrec * ntile * nfr * ((frag - 32) // 16 * 16 + (32 if mode else 0))
        r = timed(lambda: ctx.diag_fill_streams(buf.data_ptr(), nrec, stride, ntile, chunk, 32, 8, 1, (frag << 16) | mode), nbytes)
        print("fragments of %4d B, boundary sectors %-26s %8.0f GB/s" % (frag, names[mode], r))
