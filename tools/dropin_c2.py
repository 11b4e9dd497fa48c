#!/usr/bin/env python
"""BASELINE config 2 through the drop-in entry function, wall clock, everything included:

    process_multiple_genomes_single_file(k12.gb, lists.npy (10,000 lists), model, out.fasta)

Phases are timed separately with the same internals the entry function calls (GenBank read, native
list tokenisation, engine set-up, plan, drain), then the entry function itself is timed as a whole,
once into /dev/null (the pipeline's own ceiling) and once into a real file when the box has the disk
space (23 GB).  Not part of bench.py: it takes about a minute and is disk-bound.

    python tools/dropin_c2.py [--samples 10000] > gpurun_out/dropin_c2.json
"""
import argparse
import contextlib
import io
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from genome_minimizer_2_b200 import engine, genbank, minimizer_2, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=10_000)
ap.add_argument("--workdir", default=None)
args = ap.parse_args()

d = tempfile.mkdtemp(prefix="gm2_c2_", dir=args.workdir)
out = {"workload": f"C2 drop-in: K-12-shaped GenBank + {args.samples} gene-name lists (.npy) -> one FASTA file"}
try:
    g = synth.make_genome(G=4_641_652, F=4_300, seed=1)
    gb, npy = os.path.join(d, "k12.gb"), os.path.join(d, "lists.npy")
    synth.write_genbank(gb, g)
    t0 = time.perf_counter()
    lists = synth.make_gene_lists(g, args.samples, 0.5, seed=1, extra_names=2000)
    synth.save_gene_lists(npy, lists)
    n_names = sum(len(l) for l in lists)
    del lists
    out["input"] = {"genbank_mb": os.path.getsize(gb) / 1e6, "lists_npy_mb": os.path.getsize(npy) / 1e6,
                    "names_in_lists": n_names, "make_inputs_seconds": time.perf_counter() - t0}

    ph = {}
    t0 = time.perf_counter(); record = genbank.read_genbank(gb); ph["read_genbank"] = time.perf_counter() - t0
    t0 = time.perf_counter(); table = engine.GeneTable.from_record(record); ph["gene_table"] = time.perf_counter() - t0
    t0 = time.perf_counter(); ref = engine.ReferenceGenome.from_file(gb); ph["reference_genome_native"] = time.perf_counter() - t0
    assert ref.native and ref.table.names == table.names and np.array_equal(ref.table.starts, table.starts)
    t0 = time.perf_counter(); tok = engine.load_gene_lists(npy, table); ph["load_gene_lists_native"] = time.perf_counter() - t0
    assert isinstance(tok, engine.TokenizedLists)
    t0 = time.perf_counter()
    py_ids, py_off = table.tokenize(np.load(npy, allow_pickle=True).tolist())
    ph["load_gene_lists_numpy_python"] = time.perf_counter() - t0
    assert np.array_equal(py_ids, tok.ids) and np.array_equal(py_off, tok.off)
    del py_ids, py_off
    t0 = time.perf_counter(); eng = engine.MinimizerEngine(record); ph["engine_setup"] = time.perf_counter() - t0
    t0 = time.perf_counter(); lengths = eng.plan_lists(tok); ph["upload_ids_and_plan"] = time.perf_counter() - t0
    nbytes = [0]

    def sink(sa, sb, view):
        nbytes[0] += view.size
    t0 = time.perf_counter(); eng.drain(sink); ph["drain_to_host_no_write"] = time.perf_counter() - t0
    eng.close()
    out["phases_seconds"] = ph
    out["image_gb"] = nbytes[0] / 1e9
    out["kept_gbp"] = float(np.asarray(lengths, dtype=np.int64).sum()) / 1e9

    def whole(path):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            ret = minimizer_2.process_multiple_genomes_single_file(gb, npy, "bench", path)
        return time.perf_counter() - t0, ret

    dt, ret = whole(os.devnull)
    out["entry_function_to_devnull"] = {"seconds": dt, "gbp_per_s": out["kept_gbp"] / dt, "genome_count": ret["genome_count"]}
    free = shutil.disk_usage(d).free
    if free > 2.5 * nbytes[0] + (8 << 30):
        path = os.path.join(d, "out.fasta")
        dt, ret = whole(path)
        out["entry_function_to_file"] = {"seconds": dt, "gbp_per_s": out["kept_gbp"] / dt, "file_gb": os.path.getsize(path) / 1e9,
                                         "filesystem_free_gb_before": free / 1e9}
        os.remove(path)
    else:
        out["entry_function_to_file"] = {"skipped": f"only {free / 1e9:.0f} GB free in {d}"}
finally:
    shutil.rmtree(d, ignore_errors=True)
print(json.dumps(out))
