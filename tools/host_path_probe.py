#!/usr/bin/env python
"""Where the host path's time goes: gm2_emit_host per range vs whole image, both transports, and the
Python drain loop on top.  Prints one JSON object."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from genome_minimizer_2_b200 import _native, engine, synth

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
g = synth.make_genome(G=4_641_652, F=4_300, seed=1)
st, en = g.starts_ends()
table = engine.GeneTable(g.gene_names(), st, en)
keep = synth.random_keep_bool(table.F, S, 0.5, seed=3)
rows = synth.pack_keep_rows(keep)
eng = engine.MinimizerEngine(seq=g.seq, table=table, device=0)
eng.plan_keep_rows(rows)
ctx = eng.ctx
total = ctx.image_bytes(0, S)
pinned = _native.PinnedBuffer(total)
out = {"samples": S, "image_gb": total / 1e9}
ranges = eng.chunks()
for wire in (1, 2):
    ctx.configure(_native.CFG_WIRE, wire)
    ctx.emit_host(0, S, pinned)                     # warm (allocations)
    t0 = time.perf_counter(); ctx.emit_host(0, S, pinned); dt = time.perf_counter() - t0
    out[f"wire{wire}_whole_gbs"] = total / dt / 1e9
    off = ctx.record_offsets()
    per = []
    t0 = time.perf_counter()
    for a, b in ranges:
        t1 = time.perf_counter()
        ctx.emit_host(a, b, pinned.array[int(off[a]):int(off[b])])
        per.append(time.perf_counter() - t1)
    dt = time.perf_counter() - t0
    out[f"wire{wire}_by_range_gbs"] = total / dt / 1e9
    out[f"wire{wire}_range_ms"] = {"n": len(per), "median": float(np.median(per)) * 1e3, "max": float(np.max(per)) * 1e3, "first": per[0] * 1e3}
    n = [0]
    def sink(a, b, view): n[0] += view.size
    eng.drain(sink)                                  # warm
    t0 = time.perf_counter(); eng.drain(sink); dt = time.perf_counter() - t0
    out[f"wire{wire}_python_drain_gbs"] = total / dt / 1e9
print(json.dumps(out))
