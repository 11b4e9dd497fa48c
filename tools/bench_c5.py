#!/usr/bin/env python
"""BASELINE config 5 (one GPU, or one rank per GPU under torchrun; samples are per GPU): VAE v0 decoder (latent 64 -> 1024 -> 1024 -> 1024 -> V, random-init
Xavier weights as training/model.py:116-120, eval mode) -> `> 0.5` -> column->gene keep mask ->
minimize, all on the device.  The decoder is plain torch (library GEMMs, outside the graded
kernels); everything after the probabilities is libgm2.  Prints one JSON line."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from genome_minimizer_2_b200 import _native, engine, synth

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=12_500)
ap.add_argument("--columns", type=int, default=55_039)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()

import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(5)
g = synth.make_genome(seed=1)
starts, ends = g.starts_ends()
table = engine.GeneTable(g.gene_names(), starts, ends)
names = [n for n in table.name_to_id if n]
cols = names + [f"group_{i}" for i in range(args.columns - len(names))]
rng = np.random.default_rng(5)
cols = [cols[i] for i in rng.permutation(len(cols))]
essential = names[::15]

def block(i, o):
    lin = torch.nn.Linear(i, o); torch.nn.init.xavier_uniform_(lin.weight); torch.nn.init.zeros_(lin.bias)
    return [lin, torch.nn.BatchNorm1d(o), torch.nn.ReLU()]
last = torch.nn.Linear(1024, args.columns); torch.nn.init.xavier_uniform_(last.weight); torch.nn.init.zeros_(last.bias)
decoder = torch.nn.Sequential(*block(64, 1024), *block(1024, 1024), *block(1024, 1024), last, torch.nn.Sigmoid()).to(dev).eval()

eng = engine.MinimizerEngine(seq=g.seq, table=table, device=local)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st); eng.ctx.set_stream(st.cuda_stream)
space = engine.ColumnSpace(table, cols, essential)
S = args.samples
with torch.no_grad():
    torch.manual_seed(5 + 1000 * rank)          # same weights everywhere, different samples per rank
    z = torch.randn(S, 64, device=dev)
    probs = decoder(z)
    lengths, counts = engine.plan_from_probabilities(eng, space, probs, first_idx=rank * S)
    off = eng.ctx.record_offsets()
    image = torch.empty(int(off[-1]), dtype=torch.uint8, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_dec = t_plan = t_emit = 0.0
    for it in range(args.steps + 2):
        ev[0].record(st)
        probs = decoder(z)
        ev[1].record(st)
        eng.ctx.load_probs_dev(probs.data_ptr(), S, probs.stride(0), 0.5)
        eng.ctx.plan_async(rank * S)
        ev[2].record(st)
        eng.ctx.emit_dev(0, S, image.data_ptr(), image.numel())
        ev[3].record(st)
        torch.cuda.synchronize()
        if it >= 2:
            t_dec += ev[0].elapsed_time(ev[1]); t_plan += ev[1].elapsed_time(ev[2]); t_emit += ev[2].elapsed_time(ev[3])
n = args.steps
kept = int(lengths.sum())
if world > 1:
    t = torch.tensor([t_dec, t_plan, t_emit], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_dec, t_plan, t_emit = (float(x) for x in t.tolist())
    k = torch.tensor([kept], dtype=torch.int64, device=dev)
    dist.all_reduce(k, op=dist.ReduceOp.SUM)
    kept = int(k.item())
if rank == 0:
  print(json.dumps({"workload": f"C5: VAE v0 decode -> threshold -> keep mask -> minimize ({world} GPU(s), max over ranks)", "samples": S * world, "columns": args.columns,
                  "decode_ms": t_dec / n, "keepmask_plan_ms": t_plan / n, "emit_ms": t_emit / n,
                  "gbp_per_s_incl_decode": kept / ((t_dec + t_plan + t_emit) / n * 1e-3) / 1e9,
                  "gbp_per_s_after_decode": kept / ((t_plan + t_emit) / n * 1e-3) / 1e9,
                  "mean_retained_fraction": float(lengths.mean() / g.G), "mean_list_length": float(counts.mean()),
                  "image_gb_per_gpu": image.numel() / 1e9}))
eng.close()
if world > 1:
    dist.destroy_process_group()
