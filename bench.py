#!/usr/bin/env python
"""bench.py — the minimizer hot path on B200: output Gbp/s, roofline fraction, e2e, CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W]                       (ours)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]      (reference CPU arm)
    torchrun --nproc-per-node N ... bench.py --gpus N ...                     (N > 1)

Workload (BASELINE.json configs[1], per GPU): synthetic K-12-shaped genome (G = 4,641,652,
~4.4k `gene` features, seed 1) and S = 10,000 samples, every distinct gene name kept i.i.d. with
p = 0.5 (seed 2) plus 2,000 non-matching name ids per sample.  One STEP = one pass of the hot
path over that batch: K1 keep-mask builder -> K2/K3 plan + record scan -> K4 emit, from
device-resident name-id lists to the device-resident FASTA image (~26 GB).  N > 1: every rank
runs its own 10,000-sample shard of a 10,000*N-sample job (weak scaling, reference replicated,
global record ids), and all-gathers its image size for the host-side concatenation offsets.

value     = kept bases of all ranks / max-over-ranks device time        (Gbp/s)
e2e       = same metric through the C-ABI call gm2_minimize_host with HOST buffers: H2D of the
            id lists and D2H of the whole FASTA image inside the timed region
roofline  = k_emit alone: algorithmic bytes per launch / its CUDA-event time, vs the measured
            HBM copy peak in MEASURED_PEAKS.json
cpu_baseline = the oracle's literal port of the reference's Python loop, one thread, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "minimized-genome output Gbp/s"
UNIT = "Gbp/s"
FALLBACK_HBM_GBS = 6650.0            # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--samples", type=int, default=10_000, help="samples per GPU per step")
    ap.add_argument("--retention", type=float, default=0.5)
    ap.add_argument("--noise-ids", type=int, default=2000)
    ap.add_argument("--genome", choices=["k12", "12mbp"], default="k12")
    ap.add_argument("--tile-bytes", type=int, default=0)
    ap.add_argument("--emit-warps", type=int, default=0)
    ap.add_argument("--emit-batch", type=int, default=-1)
    ap.add_argument("--store-policy", type=int, default=-1)
    ap.add_argument("--packing", type=int, default=0)
    ap.add_argument("--emit-order", type=int, default=-1)
    ap.add_argument("--wire", type=int, default=-1, help="GM2_CFG_WIRE for the e2e leg (0 auto, 1 bytes, 2 two-bit)")
    ap.add_argument("--host-threads", type=int, default=-1, help="GM2_CFG_HOST_THREADS")
    ap.add_argument("--flat-run-bytes", type=int, default=-1, help="GM2_CFG_FLAT_RUN_BYTES (0 never, 1048576 always)")
    ap.add_argument("--run-table", type=int, default=0, help="GM2_CFG_RUN_TABLE (kept-run table entries per warp)")
    ap.add_argument("--emit-occupancy", type=int, default=-1, help="GM2_CFG_EMIT_OCCUPANCY (0 auto, 3, 4)")
    ap.add_argument("--emit-debug", type=int, default=0, help="timing experiments only (wrong output)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dropin", action="store_true")
    ap.add_argument("--cpu-samples", type=int, default=16)
    ap.add_argument("--ref-samples-per-core", type=int, default=2)
    ap.add_argument("--verify", type=int, default=4, help="records checked against the oracle after the run")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------
def make_workload(args, rank: int):
    from genome_minimizer_2_b200 import engine, synth
    if args.genome == "k12":
        g = synth.make_genome(seed=1)
        wname = "C2: K-12-shaped synthetic (4,641,652 bp, 4,400 gene features)"
    else:
        g = synth.make_genome(12_000_000, 10_000, seed=4, overlap_frac=0.3, nested=200, join_genes=50,
                              dup_name_frac=0.006, nameless_frac=0.003, name="SYNTH_12M")
        wname = "C4-shaped: 12 Mbp synthetic, 10,000 gene features, overlapping/antisense"
    starts, ends = g.starts_ends()
    table = engine.GeneTable(g.gene_names(), starts, ends)
    S = args.samples
    rng = np.random.default_rng(2 + 1000 * rank)
    keep_names = rng.random((S, table.V)) < args.retention
    # "" (nameless genes) is a legal name id but real lists never contain it: drop it
    if "" in table.name_to_id:
        keep_names[:, table.name_to_id[""]] = False
    counts = keep_names.sum(1) + args.noise_ids
    off = np.zeros(S + 1, dtype=np.int64)
    off[1:] = np.cumsum(counts)
    ids = np.empty(int(off[-1]), dtype=np.int32)
    noise = rng.integers(table.V, table.V + 50_000, (S, args.noise_ids), dtype=np.int32) if args.noise_ids else None
    for s in range(S):
        k = np.flatnonzero(keep_names[s]).astype(np.int32)
        ids[off[s]:off[s] + k.size] = k
        if noise is not None:
            ids[off[s] + k.size:off[s + 1]] = noise[s]
    return g, table, keep_names, ids, off, wname


def name_lists_for(table, keep_names, rows, noise_ids):
    """The same samples as Python lists of names (what the reference consumes)."""
    names = list(table.name_to_id.keys())
    out = []
    for s in rows:
        l = [names[i] for i in np.flatnonzero(keep_names[s])]
        l += [f"group_{i}" for i in range(noise_ids)]
        out.append(l)
    return out


def oracle_record_for(g):
    """Duck-typed record for the oracle's literal port (str sequence + gene features)."""
    from oracle.genbank_reader import OracleFeature, OracleLocation, OracleRecord
    feats = []
    for gene in g.genes:
        q = {} if gene.name is None else {"gene": [gene.name]}
        feats.append(OracleFeature("gene", OracleLocation(gene.start, gene.end), q))
        feats.append(OracleFeature("CDS", OracleLocation(gene.start, gene.end), dict(q)))
    return OracleRecord(seq=g.seq.tobytes().decode("ascii"), features=feats)


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        with open(self.path) as fh:
            for ln in fh:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arms (oracle's literal port of the reference loop)
# ----------------------------------------------------------------------------------------------
_POOL_STATE = {}


def _pool_init(seq_str, feats):
    from oracle.genbank_reader import OracleRecord
    _POOL_STATE["rec"] = OracleRecord(seq=seq_str, features=feats)


def _pool_run(needed):
    from oracle import minimizer_oracle as mo
    return len(mo.minimize_literal(_POOL_STATE["rec"], needed))


def cpu_literal_single_thread(g, table, keep_names, noise_ids, n):
    from oracle import minimizer_oracle as mo
    rec = oracle_record_for(g)
    lists = name_lists_for(table, keep_names, range(n), noise_ids)
    t0 = time.perf_counter()
    bases = 0
    for needed in lists:
        bases += len(mo.minimize_literal(rec, needed))
    dt = time.perf_counter() - t0
    return bases / dt / 1e9, dt, bases


def cpu_c_port_all_cores(g, table, keep_names, n):
    """For context only: the oracle's optimised C restatement (difference array + one pass) threaded
    over all host cores — a far stronger CPU program than the reference's Python loop."""
    from concurrent.futures import ThreadPoolExecutor
    from genome_minimizer_2_b200 import synth
    from oracle import c_oracle
    starts, ends = g.starts_ends()
    name_id = np.asarray([table.name_to_id[x] for x in table.names])
    n = min(n, keep_names.shape[0])
    rows = synth.pack_keep_rows(keep_names[:n][:, name_id])
    c_oracle.lib()
    cores = os.cpu_count() or 1
    chunks = [(a, min(a + 4, n)) for a in range(0, n, 4)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        bases = sum(int(L.sum()) for L in ex.map(lambda ab: c_oracle.batch(g.seq, starts, ends, rows[ab[0]:ab[1]], first_idx=ab[0])[0], chunks))
    dt = time.perf_counter() - t0
    return {"value": bases / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port (optimised C, not the reference's algorithmic cost)",
            "sample": f"first {n} samples of the workload, oracle/minimizer_c.c threaded over {cores} cores, records hashed, {dt:.2f} s"}


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU algorithm (oracle literal port; the reference is
    pure Python and cannot travel to the GPU box) over all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    g, table, keep_names, ids, off, wname = make_workload(argparse.Namespace(**{**vars(args), "samples": max(
        args.ref_samples_per_core * (os.cpu_count() or 1), 1)}), 0)
    rec = oracle_record_for(g)
    cores = os.cpu_count() or 1
    n = keep_names.shape[0]
    lists = name_lists_for(table, keep_names, range(n), args.noise_ids)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_pool_init, initargs=(rec.seq, rec.features)) as pool:
        for _ in range(args.warmup):
            pool.map(_pool_run, lists[:cores], chunksize=1)
        t0 = time.perf_counter()
        bases = 0
        for _ in range(args.steps):
            bases += sum(pool.map(_pool_run, lists, chunksize=1))
        dt = time.perf_counter() - t0
    value = bases / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wname, "samples_per_step": n, "retention": args.retention,
                   "note": "reference's per-sample Python algorithm (oracle literal port, minimizer_2.py:50-101), "
                           "one process per host core; each step is a bounded sample of the GPU arm's workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} samples per step x {args.steps} steps, fork pool over {cores} cores"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def dropin_c1(g, table, args):
    """BASELINE config 1 through the drop-in entry function itself: GenBank file + 100 gene-name lists
    (.npy, ~50 % of the names + 2,000 non-matching names each) -> one FASTA file on disk.  Wall clock,
    everything included (parse, unpickle, tokenise, GPU, D2H, file write); 3 records re-checked."""
    import hashlib
    import shutil
    from genome_minimizer_2_b200 import minimizer_2, synth
    from oracle import c_oracle, minimizer_oracle as mo
    d = tempfile.mkdtemp(prefix="gm2_c1_")
    try:
        gb, npy, out = os.path.join(d, "k12.gb"), os.path.join(d, "lists.npy"), os.path.join(d, "out.fasta")
        synth.write_genbank(gb, g)
        lists = synth.make_gene_lists(g, 100, 0.5, seed=1, extra_names=2000)
        synth.save_gene_lists(npy, lists)
        import contextlib, io
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            ret = minimizer_2.process_multiple_genomes_single_file(gb, npy, "bench", out)
        dt = time.perf_counter() - t0
        data = open(out, "rb").read()
        body = data.split(b"\n", 3)[3]
        starts, ends = g.starts_ends()
        pos = 0
        ok = True
        for s in range(100):
            keep = mo.keep_vector(table.names, lists[s])
            L = int(mo.kept_mask_numpy(g.G, starts, ends, keep).sum()) if s in (0, 57, 99) else None
            hdr = len(mo.HEADER_PREFIX) + len(str(s + 1)) + 2
            end = body.index(b"\n", pos + hdr)
            if L is not None:
                _, _, img = c_oracle.batch(g.seq, starts, ends, synth.pack_keep_rows(keep[None, :]), first_idx=s, want_image=True)
                ok = ok and body[pos:end + 1] == img.tobytes()
            pos = end + 1
        ok = ok and pos == len(body)
        if not ok:
            raise SystemExit("bench.py: drop-in C1 output differs from the oracle")
        return {"workload": "C1: process_multiple_genomes_single_file, K-12-shaped GenBank, 100 name lists, one FASTA file",
                "seconds": dt, "output_bytes": len(data), "genome_count": ret["genome_count"],
                "byte_identical_to_oracle_records": 3}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def lists_tokenize_bench(g, table, n_lists=1000):
    """SURVEY.md §8 f2 beside the GPU numbers: the `.npy` gene-lists file -> id CSR on the host, natively
    (gm2_tokenize_pickle) vs the reference's loading path restated (np.load(...).tolist() + per-name
    lookup).  Wall clock on one host core; results compared."""
    import shutil
    from genome_minimizer_2_b200 import engine, synth
    d = tempfile.mkdtemp(prefix="gm2_tok_")
    try:
        npy = os.path.join(d, "lists.npy")
        lists = synth.make_gene_lists(g, n_lists, 0.5, seed=2, extra_names=2000)
        synth.save_gene_lists(npy, lists)
        del lists
        t0 = time.perf_counter()
        tok = engine.load_gene_lists(npy, table)
        t_native = time.perf_counter() - t0
        t0 = time.perf_counter()
        ids, off = table.tokenize(np.load(npy, allow_pickle=True).tolist())
        t_python = time.perf_counter() - t0
        if not (isinstance(tok, engine.TokenizedLists) and np.array_equal(tok.ids, ids) and np.array_equal(tok.off, off)):
            raise SystemExit("bench.py: native tokeniser disagrees with the NumPy/Python path")
        return {"workload": f"{n_lists} gene-name lists (.npy, {os.path.getsize(npy) / 1e6:.1f} MB, {ids.size} matching names)",
                "native_seconds": t_native, "numpy_python_seconds": t_python, "cores": 1}
    finally:
        shutil.rmtree(d, ignore_errors=True)


# ----------------------------------------------------------------------------------------------
# ours
# ----------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from genome_minimizer_2_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path to measure")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    g, table, keep_names, ids, off, wname = make_workload(args, rank)
    S = args.samples
    first_idx = rank * S                       # global record ids: rank order == file order

    ctx = _native.Context(local_rank)
    if args.tile_bytes:
        ctx.configure(_native.CFG_TILE_BYTES, args.tile_bytes)
    if args.packing:
        ctx.configure(_native.CFG_PACKING, args.packing)
    if args.emit_warps:
        ctx.configure(_native.CFG_EMIT_WARPS, args.emit_warps)
    if args.emit_batch >= 0:
        ctx.configure(_native.CFG_EMIT_BATCH, args.emit_batch)
    if args.store_policy >= 0:
        ctx.configure(_native.CFG_STORE_POLICY, args.store_policy)
    if args.emit_order >= 0:
        ctx.configure(_native.CFG_ORDER, args.emit_order)
    if args.flat_run_bytes >= 0:
        ctx.configure(_native.CFG_FLAT_RUN_BYTES, args.flat_run_bytes)
    if args.run_table:
        ctx.configure(_native.CFG_RUN_TABLE, args.run_table)
    if args.emit_occupancy >= 0:
        ctx.configure(_native.CFG_EMIT_OCCUPANCY, args.emit_occupancy)
    if args.wire >= 0:
        ctx.configure(_native.CFG_WIRE, args.wire)
    if args.host_threads >= 0:
        ctx.configure(_native.CFG_HOST_THREADS, args.host_threads)
    if args.emit_debug:
        ctx.configure(_native.CFG_DEBUG, args.emit_debug)
        args.verify = 0
    # a real (non-default) torch stream: the kernels are launched on it and the CUDA events below are
    # recorded on it.  (torch's default stream has handle 0, which gm2_set_stream reads as "own stream".)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    starts, ends = g.starts_ends()
    ctx.set_reference(g.seq, starts, ends)
    ctx.set_name_map(table.id2gene_off, table.id2gene_idx)

    # device-resident inputs
    d_ids = torch.from_numpy(ids).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    ctx.load_ids_dev(d_ids.data_ptr(), d_off.data_ptr(), S, ids.size)
    ctx.plan(first_idx)                        # sizing pass (outside the timed region)
    lengths = ctx.lengths()
    rec_off = ctx.record_offsets()
    image_bytes = int(rec_off[-1])
    kept_bases = int(lengths.sum())
    image = torch.empty(image_bytes, dtype=torch.uint8, device=dev)
    size_t = torch.tensor([image_bytes], dtype=torch.int64, device=dev)
    sizes_all = [torch.zeros_like(size_t) for _ in range(world)]

    def step():
        ctx.plan_async(first_idx)
        ctx.emit_dev(0, S, image.data_ptr(), image_bytes)
        if world > 1:
            dist.all_gather(sizes_all, size_t)  # per-rank image sizes -> global file offsets

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # timed region: exactly K steps, CUDA events on the launching stream, clocks sampled meanwhile
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    l0 = ctx.query(_native.Q_LAUNCHES)
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_stop = torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for k in range(args.steps):
        ev[k][0].record(stream)
        ctx.plan_async(first_idx)
        ev[k][1].record(stream)
        ctx.emit_dev(0, S, image.data_ptr(), image_bytes)
        ev[k][2].record(stream)
        if world > 1:
            dist.all_gather(sizes_all, size_t)
    t_stop.record(stream)
    barrier()
    launches = ctx.query(_native.Q_LAUNCHES) - l0
    total_ms = t_start.elapsed_time(t_stop)
    plan_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    emit_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        kb = torch.tensor([kept_bases], dtype=torch.int64, device=dev)
        dist.all_reduce(kb, op=dist.ReduceOp.SUM)
        kept_all = int(kb.item())
        ln = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        launches = int(ln.item())
    else:
        kept_all = kept_bases
    ms_per_step = total_ms / args.steps
    value = kept_all / (ms_per_step * 1e-3) / 1e9

    # correctness spot-check of the timed output against the oracle (device hashes, no big copies)
    verify = {}
    if args.verify > 0:
        from oracle import c_oracle
        pick = sorted(set(np.linspace(0, S - 1, args.verify).astype(int).tolist()))
        name_id = np.asarray([table.name_to_id[n] for n in table.names])
        keep_genes = keep_names[pick][:, name_id]
        if "" in table.name_to_id:
            pass
        from genome_minimizer_2_b200 import synth
        rows = synth.pack_keep_rows(keep_genes)
        ok = True
        for j, s in enumerate(pick):
            L, H, _ = c_oracle.batch(g.seq, starts, ends, rows[j:j + 1], first_idx=first_idx + s)
            got = ctx.diag_range_hashes(image.data_ptr(), image_bytes, rec_off[s:s + 2])
            ok = ok and int(L[0]) == int(lengths[s]) and int(H[0]) == int(got[0])
        verify = {"records_checked": len(pick), "byte_identical_to_oracle": bool(ok)}
        if not ok:
            raise SystemExit("bench.py: timed output differs from the oracle — refusing to report a number")

    # write-only fill of the same buffer: the write roofline of this device, for context
    fill_gbs = None
    if rank == 0:
        nfill = min(image_bytes, 8 << 30) // 16 * 16
        for _ in range(2):
            ctx.diag_fill(image.data_ptr(), nfill, 0x41414141)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        a.record(stream)
        for _ in range(5):
            ctx.diag_fill(image.data_ptr(), nfill, 0x41414141)
        b.record(stream)
        torch.cuda.synchronize(dev)
        fill_gbs = nfill * 5 / (a.elapsed_time(b) * 1e-3) / 1e9

    # end to end through the C-ABI with host buffers
    e2e = None
    if not args.no_e2e:
        del image
        torch.cuda.empty_cache()
        avail = 64 << 30
        try:
            with open("/proc/meminfo") as fh:
                for ln_ in fh:
                    if ln_.startswith("MemAvailable:"):
                        avail = int(ln_.split()[1]) * 1024
        except OSError:
            pass
        budget = int(avail * 0.35 / max(world, 1))
        S_e = S
        while S_e > 1 and int(rec_off[S_e]) > budget:
            S_e //= 2
        e_bytes = int(rec_off[S_e])
        pinned = _native.PinnedBuffer(e_bytes)
        ids_e, off_e = ids[:int(off[S_e])], off[:S_e + 1]
        ctx.set_stream(None)
        # raw pinned D2H rate of this host path with all ranks copying at once: the ceiling of e2e
        nraw = min(e_bytes, 4 << 30)
        dsrc = torch.empty(nraw, dtype=torch.uint8, device=dev)
        hview = torch.from_numpy(pinned.array[:nraw])
        hview.copy_(dsrc, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            hview.copy_(dsrc, non_blocking=True)
        torch.cuda.synchronize(dev)
        raw_gbs = 2 * nraw / (time.perf_counter() - t0) / 1e9
        if world > 1:
            t = torch.tensor([raw_gbs], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            raw_gbs = float(t.item())
        del dsrc, hview
        for _ in range(1):
            ctx.minimize_host(pinned, ids=ids_e, off=off_e, first_idx=first_idx)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            le, ro = ctx.minimize_host(pinned, ids=ids_e, off=off_e, first_idx=first_idx)
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        wire_used = ctx.query(_native.Q_LAST_WIRE)
        d2h_moved = ctx.query(_native.Q_LAST_D2H_BYTES)
        # the other transport beside it (image bytes over PCIe), one warm-up + one timed call
        dt_bytes = None
        if wire_used == 2:
            ctx.configure(_native.CFG_WIRE, 1)
            ctx.minimize_host(pinned, ids=ids_e, off=off_e, first_idx=first_idx)
            barrier()
            t0 = time.perf_counter()
            ctx.minimize_host(pinned, ids=ids_e, off=off_e, first_idx=first_idx)
            barrier()
            dt_bytes = time.perf_counter() - t0
            ctx.configure(_native.CFG_WIRE, args.wire if args.wire >= 0 else 0)
            le, ro = ctx.minimize_host(pinned, ids=ids_e, off=off_e, first_idx=first_idx)   # the checked image is the default path's
        if world > 1:
            t = torch.tensor([dt, dt_bytes or 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0].item())
            dt_bytes = float(t[1].item()) or None
            kb = torch.tensor([int(le.sum())], dtype=torch.int64, device=dev)
            dist.all_reduce(kb, op=dist.ReduceOp.SUM)
            kept_e = int(kb.item())
        else:
            kept_e = int(le.sum())
        # check a record of the host image too
        from oracle import c_oracle
        s_chk = S_e - 1
        name_id = np.asarray([table.name_to_id[n] for n in table.names])
        from genome_minimizer_2_b200 import synth
        Lc, Hc, _ = c_oracle.batch(g.seq, starts, ends, synth.pack_keep_rows(keep_names[s_chk:s_chk + 1][:, name_id]),
                                   first_idx=first_idx + s_chk)
        got = c_oracle.range_hash(pinned.array[int(ro[s_chk]):int(ro[s_chk + 1])])
        if int(Hc[0]) != got:
            raise SystemExit("bench.py: e2e host image differs from the oracle")
        e2e = {"value": kept_e / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(ids_e.nbytes + off_e.nbytes),
               "d2h_bytes_per_step": int(d2h_moved + le.nbytes + ro.nbytes),
               "ms_per_step": dt * 1e3, "samples_per_step": S_e,
               "image_bytes_per_step": e_bytes,
               "delivered_image_gbs_per_gpu": e_bytes / dt / 1e9,
               "raw_pinned_d2h_gbs_per_gpu": raw_gbs,
               "transport": ("two bits per base over PCIe (k_emit_packed), expanded into the caller's buffer by host threads"
                             if wire_used == 2 else "image bytes over PCIe (k_emit)"),
               "image_bytes_transport": (None if dt_bytes is None else
                                         {"value": kept_e / dt_bytes / 1e9, "ms_per_step": dt_bytes * 1e3,
                                          "delivered_image_gbs_per_gpu": e_bytes / dt_bytes / 1e9}),
               "note": "raw = plain pinned cudaMemcpy of image-sized data, all ranks concurrently: the ceiling of the image-bytes transport",
               "api": "gm2_minimize_host (C-ABI): host id lists in, pinned host FASTA image out"}
        pinned.free()

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # roofline of the dominant kernel (k_emit)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = FALLBACK_HBM_GBS, "fallback"
    if os.path.exists(peaks_path):
        try:
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
        except Exception:
            pass
    F = table.F
    alg_bytes = image_bytes + S * ((F + 7) // 8) + g.G + 16 * F
    achieved = alg_bytes / (emit_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"{args.genome}:{S}")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_emit", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": emit_ms, "plan_ms": plan_ms,
                "write_fill_gbs": fill_gbs}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        v, dt, bases = cpu_literal_single_thread(g, table, keep_names, args.noise_ids, args.cpu_samples)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": f"first {args.cpu_samples} samples of the workload, oracle literal port of "
                                  f"minimizer_2.py:50-101 (list scan + position set + per-base loop), {dt:.1f} s",
                        "host_cores_available": os.cpu_count()}

    cpu_c = None
    if not args.no_cpu_baseline and world == 1:
        cpu_c = cpu_c_port_all_cores(g, table, keep_names, 32 * (os.cpu_count() or 1))

    dropin = None
    if world == 1 and not args.no_dropin:
        dropin = dropin_c1(g, table, args)
        dropin["lists_tokenize"] = lists_tokenize_bench(g, table)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wname, "samples_per_gpu": S, "retention": args.retention,
                   "noise_ids_per_sample": args.noise_ids, "input": "device-resident name-id lists (CSR)",
                   "output": f"device-resident FASTA image, {image_bytes/1e9:.2f} GB per GPU per step",
                   "l2": "no flush needed: each step writes an image >> 126 MB L2",
                   "sharding": "samples; reference replicated; all-gather of image sizes only",
                   "tile_bytes": args.tile_bytes or 49152, "kept_bases_per_gpu": kept_bases,
                   "emit_ctas_per_sm": ctx.query(_native.Q_LAST_EMIT_CTAS)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "cpu_port_c": cpu_c, "verify": verify, "dropin_c1": dropin,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
