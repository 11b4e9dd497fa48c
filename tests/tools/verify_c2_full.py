#!/usr/bin/env python
"""BASELINE config 2, every record: 10,000 samples x K-12-shaped genome through the ids path
(K1 -> K2/K3 -> K4) into a device-resident 26 GB image; ALL 10,000 record hashes and lengths are
compared with the C oracle (run on the host cores in threads).  Prints one JSON line."""
import json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from genome_minimizer_2_b200 import _native, engine, synth
from oracle import c_oracle

S = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
g = synth.make_genome(seed=1)
starts, ends = g.starts_ends()
table = engine.GeneTable(g.gene_names(), starts, ends)
rng = np.random.default_rng(2)
keep_names = rng.random((S, table.V)) < 0.5
if "" in table.name_to_id:
    keep_names[:, table.name_to_id[""]] = False
ids, off = synth.ids_csr_from_keep(keep_names, n_noise=2000, V=table.V, seed=3)
name_id = np.asarray([table.name_to_id[n] for n in table.names])
rows = synth.pack_keep_rows(keep_names[:, name_id])

ctx = _native.Context(0)
ctx.set_reference(g.seq, starts, ends)
ctx.set_name_map(table.id2gene_off, table.id2gene_idx)
ctx.load_ids_host(ids, off)
ctx.plan(0)
assert np.array_equal(ctx.keep_rows(), rows), "K1 keep rows differ"
lengths = ctx.lengths()
rec_off = ctx.record_offsets()
img = torch.empty(int(rec_off[-1]), dtype=torch.uint8, device="cuda:0")
ctx.emit_dev(0, S, img.data_ptr(), img.numel())
ctx.sync()
got = ctx.diag_range_hashes(img.data_ptr(), img.numel(), rec_off)

t0 = time.perf_counter()
c_oracle.lib()
def work(lo_hi):
    lo, hi = lo_hi
    L, H, _ = c_oracle.batch(g.seq, starts, ends, rows[lo:hi], first_idx=lo)
    return lo, L, H
nthreads = os.cpu_count() or 1
chunks = [(a, min(a + 25, S)) for a in range(0, S, 25)]
exp_len = np.zeros(S, dtype=np.int64); exp_hash = np.zeros(S, dtype=np.uint64)
with ThreadPoolExecutor(nthreads) as ex:
    for lo, L, H in ex.map(work, chunks):
        exp_len[lo:lo + len(L)] = L; exp_hash[lo:lo + len(H)] = H
dt = time.perf_counter() - t0
bad = int((got != exp_hash).sum() + (lengths != exp_len).sum())
print(json.dumps({"workload": f"C2 full verification: {S} samples, K-12 shape, ids path", "records": S,
                  "image_bytes": int(rec_off[-1]), "mismatching_records": bad,
                  "oracle_seconds": round(dt, 1), "oracle_threads": nthreads, "byte_identical": bad == 0}))
sys.exit(0 if bad == 0 else 1)
