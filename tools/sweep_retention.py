#!/usr/bin/env python
"""k_emit alone over gene retention x kernel configuration, one process (run under gpurun).

    python tools/sweep_retention.py [--genomes k12,12mbp] [--retentions 0.1,0.2,...] [--modes 1,2] [--occ 0,3,4]

For every point: random keep rows (each gene kept i.i.d. with probability p), plan, three records
checked against the C oracle by device-side hashes, then k_emit timed with CUDA events (3 warm-ups,
8 launches).  Prints one line per point: image GB, ms, GB/s, fraction of the measured HBM peak."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", default="k12,12mbp")
    ap.add_argument("--retentions", default="0.1,0.2,0.3,0.5,0.9")
    ap.add_argument("--modes", default="1,2", help="GM2_CFG_FLAT_MODE values")
    ap.add_argument("--occ", default="0", help="GM2_CFG_EMIT_OCCUPANCY values")
    ap.add_argument("--flat-run-bytes", default="640")
    ap.add_argument("--tile-bytes", default="0")
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--verify", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch
    from genome_minimizer_2_b200 import _native, synth
    from oracle import c_oracle

    peak = 6552.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    lines = []
    for gname in args.genomes.split(","):
        if gname == "k12":
            g, S = synth.make_genome(seed=1), 10_000
        else:
            g = synth.make_genome(12_000_000, 10_000, seed=4, overlap_frac=0.3, nested=200, join_genes=50,
                                  dup_name_frac=0.006, nameless_frac=0.003, name="SYNTH_12M")
            S = 4_000
        starts, ends = g.starts_ends()
        F = len(starts)
        for ret in [float(x) for x in args.retentions.split(",")]:
            rng = np.random.default_rng(int(ret * 1000) + 17)
            rows = synth.pack_keep_rows(rng.random((S, F)) < ret)
            d_rows = torch.from_numpy(rows.view(np.int32)).to(dev)
            pick = sorted({0, S // 2, S - 1})[:args.verify]
            exp = [c_oracle.batch(g.seq, starts, ends, rows[s:s + 1], first_idx=s) for s in pick]
            for tile in [int(x) for x in args.tile_bytes.split(",")]:
                for frb in [int(x) for x in args.flat_run_bytes.split(",")]:
                    for mode in [int(x) for x in args.modes.split(",")]:
                        for occ in [int(x) for x in args.occ.split(",")]:
                            with _native.Context(0) as ctx:
                                if tile:
                                    ctx.configure(_native.CFG_TILE_BYTES, tile)
                                ctx.configure(_native.CFG_FLAT_MODE, mode)
                                ctx.configure(_native.CFG_FLAT_RUN_BYTES, frb)
                                ctx.configure(_native.CFG_EMIT_OCCUPANCY, occ)
                                ctx.set_stream(stream.cuda_stream)
                                ctx.set_reference(g.seq, starts, ends)
                                ctx.load_keep_dev(d_rows.data_ptr(), S)
                                ctx.plan(0)
                                lengths, rec_off = ctx.lengths(), ctx.record_offsets()
                                nbytes = int(rec_off[-1])
                                image = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                                for _ in range(3):
                                    ctx.emit_dev(0, S, image.data_ptr(), nbytes)
                                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                                torch.cuda.synchronize(dev)
                                a.record(stream)
                                for _ in range(args.reps):
                                    ctx.emit_dev(0, S, image.data_ptr(), nbytes)
                                b.record(stream)
                                torch.cuda.synchronize(dev)
                                ms = a.elapsed_time(b) / args.reps
                                ok = True
                                for (L, H, _), s in zip(exp, pick):
                                    got = ctx.diag_range_hashes(image.data_ptr(), nbytes, rec_off[s:s + 2])
                                    ok = ok and int(L[0]) == int(lengths[s]) and int(H[0]) == int(got[0])
                                ctas = ctx.query(_native.Q_LAST_EMIT_CTAS)
                                tile_used = ctx.query(_native.Q_TILE_BYTES)
                                del image
                            alg = nbytes + S * ((F + 7) // 8) + g.G + 16 * F
                            gbs = alg / (ms * 1e-3) / 1e9
                            ln = (f"{gname:>5} ret {ret:.2f} tile {tile_used:6d} frb {frb:7d} mode {mode} occ {occ} ctas {ctas}: "
                                  f"image {nbytes/1e9:6.2f} GB  k_emit {ms:7.3f} ms  {gbs:7.1f} GB/s  frac {gbs/peak:.3f}  "
                                  f"{'ok' if ok else 'MISMATCH'}")
                            print(ln, flush=True)
                            lines.append(ln)
            del d_rows
            torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as fh:
            fh.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
