#!/usr/bin/env python
"""bench.py — the minimizer hot path on B200: output Gbp/s, roofline fraction, e2e, CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W]                       (ours)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]      (reference CPU arm)
    torchrun --nproc-per-node N ... bench.py --gpus N ...                     (N > 1)

Headline workload (BASELINE.json configs[1], per GPU): synthetic K-12-shaped genome (G = 4,641,652,
~4.4k `gene` features, seed 1) and S = 10,000 samples, every distinct gene name kept i.i.d. with
p = 0.5 plus 2,000 non-matching name ids per sample.  One STEP = one pass of the hot path over that
batch: K1 keep-mask builder -> K2/K3 plan + record scan -> K4 emit, from device-resident name-id
lists to the device-resident FASTA image (~26 GB).  N > 1: every rank runs its own 10,000-sample
shard of a 10,000*N-sample job (weak scaling, reference replicated, global record ids) and the
step ends with the real exchange of the sharded product path: an NCCL all-gather of this step's
per-sample lengths (8 bytes per sample), from which every rank derives its file offset.

value     = kept bases of all ranks / max-over-ranks device time        (Gbp/s)
e2e       = same metric through the C-ABI call gm2_minimize_host with HOST buffers: H2D of the
            id lists and D2H of the whole FASTA image inside the timed region
roofline  = k_emit alone: algorithmic bytes per launch / its CUDA-event time, vs the measured
            HBM copy peak in MEASURED_PEAKS.json
cpu_baseline = the reference's per-sample Python algorithm, one thread, bounded sample
sharded   = BASELINE.json configs[2], the FIXED 100,000-sample job cut over the N ranks (strong
            scaling): `device` = every rank streams its shard through a ring of two output buffers
            (plan + emit per 10,000-sample chunk, lengths all-gathered), CUDA events, max over
            ranks; `file` = the product's own sharded entry point (dist.run_single_file_sharded, at
            N = 1 engine.run_single_file) writing one FASTA file into tmpfs, every length and the file
            size checked against the oracle and sampled records (all shard boundaries) hashed;
            `entry` = process_multiple_genomes_single_file itself (GenBank file + .npy lists) under
            the same launch, every record compared with the oracle
retention_sweep = k_emit alone at gene retention 0.1 ... 0.9 on both genome shapes (N = 1)
config5   = BASELINE.json configs[4]: VAE v0 decoder output (torch) -> `> 0.5` -> column -> gene keep mask
            (k_keep_from_probs) -> plan -> emit, 100,000 samples over the ranks (12,500 per GPU), V = 55,039
            columns, random-init weights; per-phase CUDA-event times, three records per rank checked against
            the converter + minimizer oracles
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import shutil
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
REF_TREE = os.path.join(ROOT, "baseline", "_ref")


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
    ap.add_argument("--job-samples", type=int, default=100_000, help="the fixed job of the `sharded` leg (config 3)")
    ap.add_argument("--file-gb-per-rank", type=float, default=12.0, help="tmpfs budget per rank of the sharded file leg")
    ap.add_argument("--tile-bytes", type=int, default=0)
    ap.add_argument("--emit-warps", type=int, default=0)
    ap.add_argument("--emit-batch", type=int, default=-1)
    ap.add_argument("--store-policy", type=int, default=-1)
    ap.add_argument("--packing", type=int, default=0)
    ap.add_argument("--emit-order", type=int, default=-1)
    ap.add_argument("--wire", type=int, default=-1, help="GM2_CFG_WIRE for the e2e leg (0 auto, 1 bytes, 2 two-bit)")
    ap.add_argument("--host-threads", type=int, default=-1, help="GM2_CFG_HOST_THREADS")
    ap.add_argument("--flat-run-bytes", type=int, default=-1, help="GM2_CFG_FLAT_RUN_BYTES (0 never, 1048576 always)")
    ap.add_argument("--flat-mode", type=int, default=0, help="GM2_CFG_FLAT_MODE (1 cursor per lane, 2 bitmap-indexed)")
    ap.add_argument("--run-table", type=int, default=0, help="GM2_CFG_RUN_TABLE (kept-run table entries per warp)")
    ap.add_argument("--emit-occupancy", type=int, default=-1, help="GM2_CFG_EMIT_OCCUPANCY (0 auto, 3, 4)")
    ap.add_argument("--emit-debug", type=int, default=0, help="timing experiments only (wrong output)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-pipeline", action="store_true", help="one context for every step (the plan of step k+1 does not start before the emit of step k has finished)")
    ap.add_argument("--stream-priorities", default="-1,0", help="CUDA priorities of the two contexts' streams (lower = higher)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dropin", action="store_true")
    ap.add_argument("--no-sharded", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--cpu-samples", type=int, default=16)
    ap.add_argument("--ref-samples-per-core", type=int, default=2)
    ap.add_argument("--verify", type=int, default=64, help="records checked against the oracle after the run")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------
def make_genome(kind: str):
    from genome_minimizer_2_b200 import synth
    if kind == "k12":
        return synth.make_genome(seed=1), "K-12-shaped synthetic (4,641,652 bp, 4,400 gene features)"
    g = synth.make_genome(12_000_000, 10_000, seed=4, overlap_frac=0.3, nested=200, join_genes=50,
                          dup_name_frac=0.006, nameless_frac=0.003, name="SYNTH_12M")
    return g, "12 Mbp synthetic, 10,000 gene features, overlapping/antisense"


class Job:
    """A synthetic job of `total` samples, generated block by block from (seed, block index), so that
    any rank can produce any contiguous range and the job is the same however it is cut.  Sample i keeps
    every distinct gene name i.i.d. with probability p_i and carries `noise_ids` ids that name no gene."""
    BLOCK = 500

    def __init__(self, table, total: int, retention, noise_ids: int, seed: int):
        self.table, self.total, self.noise_ids, self.seed = table, int(total), int(noise_ids), int(seed)
        self.retention = retention                                # float, or callable(global index array) -> p
        self.V = table.V
        self.nameless = table.name_to_id.get("")                 # "" is a legal id but real lists never hold it
        self.name_id = np.asarray([table.name_to_id[n] for n in table.names], dtype=np.int64)

    def _p(self, idx: np.ndarray) -> np.ndarray:
        if callable(self.retention):
            return np.asarray(self.retention(idx), dtype=np.float64)
        return np.full(idx.size, float(self.retention))

    def _block(self, b: int):
        lo = b * self.BLOCK
        nb = min(self.BLOCK, self.total - lo)
        rng = np.random.default_rng([self.seed, b])
        keep = rng.random((nb, self.V)) < self._p(np.arange(lo, lo + nb))[:, None]
        if self.nameless is not None:
            keep[:, self.nameless] = False
        noise = (rng.integers(self.V, self.V + 50_000, (nb, self.noise_ids), dtype=np.int32)
                 if self.noise_ids else np.zeros((nb, 0), dtype=np.int32))
        return keep, noise

    def keep_names(self, lo: int, hi: int) -> np.ndarray:
        """bool [hi-lo, V]: which distinct names sample lo..hi-1 keep."""
        parts = []
        for b in range(lo // self.BLOCK, (max(hi, lo + 1) - 1) // self.BLOCK + 1):
            keep, _ = self._block(b)
            a0 = b * self.BLOCK
            parts.append(keep[max(lo - a0, 0):hi - a0])
        return np.concatenate(parts) if parts else np.zeros((0, self.V), dtype=bool)

    def keep_genes(self, samples) -> np.ndarray:
        """bool [len(samples), F] for arbitrary global sample indices (what the oracle consumes)."""
        rows = [self.keep_names(int(s), int(s) + 1)[0] for s in samples]
        return (np.stack(rows) if rows else np.zeros((0, self.V), dtype=bool))[:, self.name_id]

    def csr(self, lo: int, hi: int):
        """(ids int32, off int64[n+1], counts int64[n]) of samples [lo, hi)."""
        ids_parts, counts = [], []
        col = np.arange(self.V, dtype=np.int32)
        for b in range(lo // self.BLOCK, (max(hi, lo + 1) - 1) // self.BLOCK + 1):
            keep, noise = self._block(b)
            a0 = b * self.BLOCK
            sl = slice(max(lo - a0, 0), hi - a0)
            keep, noise = keep[sl], noise[sl]
            full = np.concatenate([np.where(keep, col, np.int32(-1)), noise], axis=1)
            ids_parts.append(full[full >= 0])
            counts.append(keep.sum(axis=1) + self.noise_ids)
        counts = np.concatenate(counts).astype(np.int64) if counts else np.zeros(0, dtype=np.int64)
        off = np.zeros(counts.size + 1, dtype=np.int64)
        off[1:] = np.cumsum(counts)
        ids = np.concatenate(ids_parts).astype(np.int32) if ids_parts else np.zeros(0, dtype=np.int32)
        return ids, off, counts


def name_lists_for(table, keep_names, noise_ids):
    """Samples as Python lists of names (what the reference consumes)."""
    names = list(table.name_to_id.keys())
    out = []
    for row in keep_names:
        l = [names[i] for i in np.flatnonzero(row)]
        l += [f"group_{i}" for i in range(noise_ids)]
        out.append(l)
    return out


def oracle_record_for(g):
    """Duck-typed record for the reference's code / its literal port (str sequence + gene features)."""
    from oracle.genbank_reader import OracleFeature, OracleLocation, OracleRecord
    feats = []
    for gene in g.genes:
        q = {} if gene.name is None else {"gene": [gene.name]}
        feats.append(OracleFeature("gene", OracleLocation(gene.start, gene.end), q))
        feats.append(OracleFeature("CDS", OracleLocation(gene.start, gene.end), dict(q)))
    return OracleRecord(seq=g.seq.tobytes().decode("ascii"), features=feats)


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as fh:
            for ln in fh:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def mem_available() -> int:
    try:
        with open("/proc/meminfo") as fh:
            for ln in fh:
                if ln.startswith("MemAvailable:"):
                    return int(ln.split()[1]) * 1024
    except OSError:
        pass
    return 64 << 30


# ----------------------------------------------------------------------------------------------
# oracle checks (bench.py may use oracle/ as the checker only)
# ----------------------------------------------------------------------------------------------
def oracle_records(g, job: Job, samples, first_idx_of=lambda s: s, threads: int = 0):
    """(lengths, record hashes) of the given global samples from the C oracle, threaded over the host cores."""
    from concurrent.futures import ThreadPoolExecutor
    from genome_minimizer_2_b200 import synth
    from oracle import c_oracle
    c_oracle.lib()
    starts, ends = g.starts_ends()
    samples = [int(s) for s in samples]
    rows = synth.pack_keep_rows(job.keep_genes(samples)) if samples else np.zeros((0, 1), dtype=np.uint32)

    def one(j):
        L, H, _ = c_oracle.batch(g.seq, starts, ends, rows[j:j + 1], first_idx=first_idx_of(samples[j]))
        return int(L[0]), int(H[0])

    with ThreadPoolExecutor(threads or min(os.cpu_count() or 1, 32)) as ex:
        res = list(ex.map(one, range(len(samples))))
    return [r[0] for r in res], [r[1] for r in res]


def oracle_lengths(g, job: Job, lo: int, hi: int) -> np.ndarray:
    """Every length of samples [lo, hi) from the oracle's vectorised union-of-ranges form."""
    from oracle import minimizer_oracle as mo
    starts, ends = g.starts_ends()
    out = np.empty(hi - lo, dtype=np.int64)
    for a in range(lo, hi, 2000):
        b = min(a + 2000, hi)
        out[a - lo:b - lo] = mo.kept_lengths_numpy(g.G, starts, ends, job.keep_names(a, b)[:, job.name_id])
    return out


def record_sizes(lengths: np.ndarray, first_idx: int = 0) -> np.ndarray:
    """'>' + 28-byte id prefix + decimal(idx+1) + '\\n' + bases + '\\n' (minimizer_2.py:476-477)."""
    idx1 = np.arange(first_idx + 1, first_idx + lengths.size + 1)
    return 1 + 28 + np.char.str_len(idx1.astype(str)) + 1 + lengths + 1


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
# CPU arms: the reference's own file when a runtime copy is present, else the oracle's literal port
# ----------------------------------------------------------------------------------------------
_POOL_STATE = {}


def load_reference_module():
    """The reference's UNMODIFIED minimizer_2.py from baseline/_ref (staged by __graft_entry__.build() from
    /root/reference, git-ignored), imported with `Bio` / `matplotlib` stubbed — it imports both at the top
    (minimizer_2.py:10-12) but only duck-types the record on this path.  None when the copy is absent."""
    path = os.path.join(REF_TREE, "src", "genome_minimizer_2", "minimizer", "minimizer_2.py")
    if not os.path.exists(path):
        return None
    import importlib
    import types
    bio, seqio, seqrecord = types.ModuleType("Bio"), types.ModuleType("Bio.SeqIO"), types.ModuleType("Bio.SeqRecord")
    seqrecord.SeqRecord = object
    bio.SeqIO, bio.SeqRecord = seqio, seqrecord
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    for name, mod in (("Bio", bio), ("Bio.SeqIO", seqio), ("Bio.SeqRecord", seqrecord),
                      ("matplotlib", mpl), ("matplotlib.pyplot", plt)):
        sys.modules.setdefault(name, mod)
    if REF_TREE not in sys.path:
        sys.path.insert(0, REF_TREE)
    try:                         # the module imports `..utils.directories` and `src.genome_minimizer_2...`: import it as the package it is
        return importlib.import_module("src.genome_minimizer_2.minimizer.minimizer_2")
    except Exception as e:  # noqa: BLE001 - the port stays available
        print(f"bench.py: reference copy in baseline/_ref not importable ({e!r}); using the port", file=sys.stderr)
        return None


def _pool_init(seq_str, feats, use_reference):
    from oracle.genbank_reader import OracleRecord
    _POOL_STATE["rec"] = OracleRecord(seq=seq_str, features=feats)
    _POOL_STATE["ref"] = load_reference_module() if use_reference else None


def _pool_run(args):
    idx, needed = args
    ref = _POOL_STATE["ref"]
    if ref is not None:
        with contextlib.redirect_stdout(io.StringIO()):
            return len(ref.GenomeMinimiser(record=_POOL_STATE["rec"], needed_genes_list=needed, idx=idx).reduced_genome_str)
    from oracle import minimizer_oracle as mo
    return len(mo.minimize_literal(_POOL_STATE["rec"], needed))


def cpu_single_thread(g, table, job: Job, n: int):
    """The reference's per-sample algorithm on ONE thread (the reference has no parallelism)."""
    rec = oracle_record_for(g)
    lists = name_lists_for(table, job.keep_names(0, n), job.noise_ids)
    _pool_init(rec.seq, rec.features, True)
    kind = "reference" if _POOL_STATE["ref"] is not None else "port"
    t0 = time.perf_counter()
    bases = sum(_pool_run((i, needed)) for i, needed in enumerate(lists))
    dt = time.perf_counter() - t0
    return bases / dt / 1e9, dt, kind


def cpu_c_port_all_cores(g, job: Job, n: int):
    """For context only: the oracle's optimised C restatement (difference array + one pass) threaded
    over all host cores — a far stronger CPU program than the reference's Python loop."""
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    lengths, _ = oracle_records(g, job, range(n), threads=cores)
    dt = time.perf_counter() - t0
    return {"value": sum(lengths) / dt / 1e9, "unit": UNIT, "cores": cores,
            "kind": "port (optimised C, not the reference's algorithmic cost)",
            "sample": f"first {n} samples of the workload, oracle/minimizer_c.c threaded over {cores} cores, records hashed, {dt:.2f} s"}


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path over all host cores, each step a
    bounded sample of the GPU arm's workload.  Rank 0 alone runs it."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    import multiprocessing as mp
    from genome_minimizer_2_b200 import engine
    g, wname = make_genome(args.genome)
    starts, ends = g.starts_ends()
    table = engine.GeneTable(g.gene_names(), starts, ends)
    cores = os.cpu_count() or 1
    n = max(args.ref_samples_per_core * cores, 1)
    job = Job(table, n, args.retention, args.noise_ids, seed=2)
    rec = oracle_record_for(g)
    lists = list(enumerate(name_lists_for(table, job.keep_names(0, n), args.noise_ids)))
    use_ref = load_reference_module() is not None
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_pool_init, initargs=(rec.seq, rec.features, use_ref)) as pool:
        for _ in range(args.warmup):
            pool.map(_pool_run, lists[:cores], chunksize=1)
        t0 = time.perf_counter()
        bases = 0
        for _ in range(args.steps):
            bases += sum(pool.map(_pool_run, lists, chunksize=1))
        dt = time.perf_counter() - t0
    value = bases / dt / 1e9
    kind = "reference" if use_ref else "port"
    what = ("the reference's own GenomeMinimiser (unmodified minimizer_2.py from baseline/_ref, Bio/matplotlib stubbed)"
            if use_ref else "oracle literal port of minimizer_2.py:50-101")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C2: " + wname, "samples_per_step": n, "retention": args.retention,
                   "note": f"{what}, one process per host core; each step is a bounded sample of the GPU arm's workload",
                   "cpu_model": cpu_model()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{n} samples per step x {args.steps} steps, fork pool over {cores} cores"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
# the drop-in entry function itself (config 1), at any N
# ----------------------------------------------------------------------------------------------
def shared_tmpdir(rank: int, world: int, tag: str) -> str:
    """A directory every rank of this job sees (tmpfs when there is one), created by rank 0."""
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    d = os.path.join(base, f"gm2_bench_{os.environ.get('MASTER_PORT', 'solo')}_{os.getppid() if world > 1 else os.getpid()}_{tag}")
    if rank == 0:
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d, exist_ok=True)
    return d


def dropin_entry(g, table, rank, world, barrier, n_lists=100):
    """BASELINE config 1 through the drop-in entry function: GenBank file + gene-name lists (.npy, ~50 % of
    the names + 2,000 non-matching names each) -> one FASTA file.  Wall clock, everything included (parse,
    unpickle, tokenise, GPU, D2H, file write); under torchrun the function shards the samples by itself.
    Every record is compared with the oracle."""
    from genome_minimizer_2_b200 import minimizer_2, synth
    from oracle import c_oracle, minimizer_oracle as mo
    d = shared_tmpdir(rank, world, "c1")
    gb, npy, out = os.path.join(d, "k12.gb"), os.path.join(d, "lists.npy"), os.path.join(d, "out.fasta")
    lists = synth.make_gene_lists(g, n_lists, 0.5, seed=1, extra_names=2000)
    if rank == 0:
        synth.write_genbank(gb, g)
        synth.save_gene_lists(npy, lists)
    barrier()
    try:
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            ret = minimizer_2.process_multiple_genomes_single_file(gb, npy, "bench", out)
        barrier()
        dt = time.perf_counter() - t0
        res = None
        if rank == 0:
            data = open(out, "rb").read()
            body = data.split(b"\n", 3)[3]
            starts, ends = g.starts_ends()
            keep = np.stack([mo.keep_vector(table.names, l) for l in lists])
            _, hashes, _ = c_oracle.batch(g.seq, starts, ends, synth.pack_keep_rows(keep))
            sizes = record_sizes(mo.kept_lengths_numpy(g.G, starts, ends, keep))
            off = np.concatenate([[0], np.cumsum(sizes)])
            ok = int(off[-1]) == len(body) and all(
                c_oracle.range_hash(body[int(off[s]):int(off[s + 1])]) == int(hashes[s]) for s in range(n_lists))
            if not ok:
                raise SystemExit("bench.py: drop-in entry function output differs from the oracle")
            res = {"workload": f"C1: process_multiple_genomes_single_file, K-12-shaped GenBank, {n_lists} name lists, one FASTA file",
                   "ranks": world, "seconds": dt, "output_bytes": len(data), "genome_count": ret["genome_count"],
                   "records_checked": n_lists, "byte_identical": True}
        barrier()
        return res
    finally:
        barrier()
        if rank == 0:
            shutil.rmtree(d, ignore_errors=True)


def lists_tokenize_bench(g, table, n_lists=1000):
    """SURVEY.md §8 f2 beside the GPU numbers: the `.npy` gene-lists file -> id CSR on the host, natively
    (gm2_tokenize_pickle) vs the reference's loading path restated (np.load(...).tolist() + per-name
    lookup).  Wall clock on one host core; results compared."""
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
# k_emit over gene retention, both genome shapes (N = 1)
# ----------------------------------------------------------------------------------------------
def retention_sweep(ctx_k12, g_k12, make_ctx, torch, dev, stream, peak, retentions=(0.1, 0.2, 0.3, 0.5, 0.9)):
    from genome_minimizer_2_b200 import _native, synth
    from oracle import c_oracle
    out = []
    for gname in ("k12", "12mbp"):
        if gname == "k12":
            ctx, g, S, own = ctx_k12, g_k12, 10_000, False
        else:
            g, _ = make_genome("12mbp")
            ctx, S, own = make_ctx(), 4_000, True
            ctx.set_reference(g.seq, *g.starts_ends())
        starts, ends = g.starts_ends()
        F = len(starts)
        try:
            for ret in retentions:
                rng = np.random.default_rng(int(ret * 1000) + 17)
                rows = synth.pack_keep_rows(rng.random((S, F)) < ret)
                d_rows = torch.from_numpy(rows.view(np.int32)).to(dev)
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
                for _ in range(6):
                    ctx.emit_dev(0, S, image.data_ptr(), nbytes)
                b.record(stream)
                torch.cuda.synchronize(dev)
                ms = a.elapsed_time(b) / 6
                ok = True
                for s in (0, S // 2, S - 1):
                    L, H, _ = c_oracle.batch(g.seq, starts, ends, rows[s:s + 1], first_idx=s)
                    got = ctx.diag_range_hashes(image.data_ptr(), nbytes, rec_off[s:s + 2])
                    ok = ok and int(L[0]) == int(lengths[s]) and int(H[0]) == int(got[0])
                if not ok:
                    raise SystemExit(f"bench.py: retention sweep ({gname}, {ret}) differs from the oracle")
                alg = nbytes + S * ((F + 7) // 8) + g.G + 16 * F
                gbs = alg / (ms * 1e-3) / 1e9
                out.append({"genome": gname, "samples": S, "gene_retention": ret, "image_gb": nbytes / 1e9,
                            "k_emit_ms": ms, "gbs": gbs, "frac": gbs / peak,
                            "emit_ctas_per_sm": ctx.query(_native.Q_LAST_EMIT_CTAS), "records_checked": 3})
                del image, d_rows
                torch.cuda.empty_cache()
        finally:
            if own:
                ctx.close()
    return out


# ----------------------------------------------------------------------------------------------
# config 3: the fixed 100,000-sample job over the N ranks
# ----------------------------------------------------------------------------------------------
def sharded_device_leg(pair, g, job: Job, torch, dist, dev, stream, rank, world, barrier, chunk_samples):
    """Every rank streams its count-based shard of the fixed job through a ring of two output buffers, chunk i on
    context i & 1: chunk i+1 is planned under the emit of chunk i (gm2_order_after keeps the emits in order)."""
    ctx = pair[0]
    from genome_minimizer_2_b200 import engine
    lo, hi = engine.shard_range(job.total, rank, world)
    chunks = [(a, min(a + chunk_samples, hi)) for a in range(lo, hi, chunk_samples)]
    staged = []
    for a, b in chunks:
        ids, off, _ = job.csr(a, b)
        staged.append((torch.from_numpy(ids).to(dev), torch.from_numpy(off).to(dev), int(ids.size)))
    sizes, kept, last_plan = [], 0, None
    for (a, b), (d_ids, d_off, n) in zip(chunks, staged):        # sizing pass (also sizes the context's buffers)
        ctx.load_ids_dev(d_ids.data_ptr(), d_off.data_ptr(), b - a, n)
        ctx.plan(a)
        last_plan = (ctx.lengths(), ctx.record_offsets())
        sizes.append(int(last_plan[1][-1]))
        kept += int(last_plan[0].sum())
    cap = max(sizes) if sizes else 16
    ring = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(2)]
    width = (job.total + world - 1) // world
    my_len = torch.zeros(width, dtype=torch.int64, device=dev)
    all_len = torch.zeros(width * world, dtype=torch.int64, device=dev)

    def run():
        pair[1].order_after(pair[0])
        for i, ((a, b), (d_ids, d_off, n)) in enumerate(zip(chunks, staged)):
            c, o = pair[i & 1], pair[(i + 1) & 1]
            c.load_ids_dev(d_ids.data_ptr(), d_off.data_ptr(), b - a, n)
            c.plan_async(a)                                   # under the other context's emit of chunk i-1
            c.lengths_dev(my_len.data_ptr() + 8 * (a - lo))
            c.order_after(o)                                  # emits in file order; ring[i & 1] is this context's own
            c.emit_dev(0, b - a, ring[i & 1].data_ptr(), cap)
        pair[0].order_after(pair[1])                          # stream A (the timing / collective stream) sees everything
        if world > 1:
            dist.all_gather_into_tensor(all_len, my_len)
        else:
            all_len.copy_(my_len)

    run()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    t0.record(stream)
    for _ in range(reps):
        run()
    t1.record(stream)
    barrier()
    ms = t0.elapsed_time(t1) / reps
    t = torch.tensor([ms, float(kept)], dtype=torch.float64, device=dev)
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, kept_all = float(tm[0].item()), int(t[1].item())
    else:
        kept_all = kept
    # checks: the last chunk's image (first and last record) and the gathered lengths against the oracle
    ok = True
    if chunks:
        a, b = chunks[-1]
        lengths, rec_off = last_plan
        img = ring[(len(chunks) - 1) & 1]
        picks = sorted({a, b - 1})
        Ls, Hs = oracle_records(g, job, picks)
        for s, L, H in zip(picks, Ls, Hs):
            got = ctx.diag_range_hashes(img.data_ptr(), cap, rec_off[s - a:s - a + 2])
            ok = ok and L == int(lengths[s - a]) and H == int(got[0])
    gathered = all_len.cpu().numpy().reshape(world, width)
    probe = np.unique(np.linspace(0, job.total - 1, 400).astype(np.int64))
    exp = {int(s): L for s, L in zip(probe, oracle_records(g, job, probe)[0])} if rank == 0 else {}
    for s, L in exp.items():
        r = min(int(s * world // job.total), world - 1)
        while s < engine.shard_range(job.total, r, world)[0]:
            r -= 1
        while s >= engine.shard_range(job.total, r, world)[1]:
            r += 1
        ok = ok and int(gathered[r, s - engine.shard_range(job.total, r, world)[0]]) == L
    flag = torch.tensor([0 if ok else 1], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if int(flag.item()):
        raise SystemExit("bench.py: sharded device leg differs from the oracle")
    del ring, staged
    torch.cuda.empty_cache()
    return {"samples_total": job.total, "samples_per_rank": hi - lo, "chunk_samples": chunk_samples, "chunks_per_rank": len(chunks),
            "seconds": ms * 1e-3, "gbp_per_s": kept_all / (ms * 1e-3) / 1e9, "kept_bases_total": kept_all,
            "collective": "all_gather_into_tensor of the per-sample lengths of every rank's shard (int64, NCCL)",
            "records_checked": 2 * world, "lengths_checked": int(probe.size), "byte_identical": True,
            "timing": "CUDA events on the launching stream around the whole shard (2 passes averaged), max over ranks"}


def sharded_file_leg(g, table, job: Job, rank, world, barrier, gb_per_rank):
    """The product's sharded single-file path on the first n samples of the fixed job, into tmpfs."""
    from genome_minimizer_2_b200 import dist as gdist, engine
    from oracle import c_oracle
    bytes_per_sample = g.G * 0.57 + 40
    budget = min(gb_per_rank * 1e9 * world, 0.25 * mem_available())
    try:
        budget = min(budget, 0.5 * shutil.disk_usage("/dev/shm").free)
    except OSError:
        pass
    n = int(max(min(job.total, budget // bytes_per_sample), min(job.total, 4 * world)))
    ids, off, counts = job.csr(0, n)
    lists = engine.TokenizedLists(ids, off, counts)
    ref = engine.ReferenceGenome(g.seq, table)
    d = shared_tmpdir(rank, world, "c3")
    out = os.path.join(d, "job.fasta")
    barrier()
    try:
        t0 = time.perf_counter()
        if world > 1:
            gdist.run_single_file_sharded(ref, lists, "bench", out, timestamp="<TS>", quiet=True)
        else:
            with contextlib.redirect_stdout(io.StringIO()):
                engine.run_single_file(ref, lists, "bench", out)
        barrier()
        dt = time.perf_counter() - t0
        res = None
        if rank == 0:
            lengths = oracle_lengths(g, job, 0, n)
            sizes = record_sizes(lengths)
            rec_off = np.concatenate([[0], np.cumsum(sizes)])
            with open(out, "rb") as fh:
                pre = b"".join(fh.readline() for _ in range(3))
                file_size = os.fstat(fh.fileno()).st_size
                picks = set(np.linspace(0, n - 1, 32).astype(int).tolist())
                for r in range(world):
                    for cut in engine.shard_range(n, r, world) + engine.shard_range_by_bytes(sizes, r, world):
                        picks.update(s for s in (cut - 1, cut) if 0 <= s < n)
                picks = sorted(picks)
                _, hashes = oracle_records(g, job, picks)
                ok = file_size == len(pre) + int(rec_off[-1]) and pre.startswith(b"# Minimized genomes generated using model: bench\n")
                for s, H in zip(picks, hashes):
                    fh.seek(len(pre) + int(rec_off[s]))
                    ok = ok and c_oracle.range_hash(fh.read(int(sizes[s]))) == H
            if not ok:
                raise SystemExit("bench.py: sharded file leg differs from the oracle")
            res = {"samples": n, "full_job": n == job.total, "file_bytes": int(file_size), "seconds": dt,
                   "gbp_per_s": float(lengths.sum()) / dt / 1e9, "balance": gdist.balance_mode(),
                   "api": ("dist.run_single_file_sharded (NCCL; each rank produces its shard" if world > 1
                           else "engine.run_single_file (one process; records produced") +
                          (" with pwrite)" if os.environ.get("GM2_FILE_SINK", "map") == "write"
                           else " straight into a shared mapping of the file)"),
                   "gb_per_s_file": file_size / dt / 1e9,
                   "target": out.rsplit("/", 2)[0] + "/… (tmpfs)" if out.startswith("/dev/shm") else "temp dir",
                   "lengths_checked": n, "records_checked": len(picks), "byte_identical": True}
        barrier()
        return res
    finally:
        barrier()
        if rank == 0:
            shutil.rmtree(d, ignore_errors=True)


# ----------------------------------------------------------------------------------------------
# BASELINE config 5: decoder output -> threshold -> column -> gene keep mask -> minimize
# ----------------------------------------------------------------------------------------------
def config5_leg(g, table, torch, dist, dev, local_rank, rank, world, barrier, total_samples=100_000, columns=55_039, steps=5):
    """VAE v0 decoder (latent 64 -> 1024 -> 1024 -> 1024 -> V, random-init Xavier weights as training/model.py:116-120,
    eval mode; plain torch, outside the graded kernels) -> `> 0.5` (utils/extras.py:200-201) -> column -> gene keep mask
    with forced essentials (binary_converter.py:49-64, :91-110) -> minimize, everything after the probabilities in
    libgm2 on the device.  100,000 samples over the ranks (12,500 per GPU: one GPU alone takes 12,500 too).  Three
    records per rank are checked against the converter + minimizer oracles."""
    from genome_minimizer_2_b200 import engine, synth
    from oracle import c_oracle, converter_oracle as co, minimizer_oracle as mo
    S = total_samples // max(world, 8)
    torch.manual_seed(5)
    names = [n for n in table.name_to_id if n]
    cols = names + [f"group_{i}" for i in range(columns - len(names))]
    rng = np.random.default_rng(5)
    cols = [cols[i] for i in rng.permutation(len(cols))]
    essential = names[::15]

    def block(i, o):
        lin = torch.nn.Linear(i, o)
        torch.nn.init.xavier_uniform_(lin.weight)
        torch.nn.init.zeros_(lin.bias)
        return [lin, torch.nn.BatchNorm1d(o), torch.nn.ReLU()]

    last = torch.nn.Linear(1024, columns)
    torch.nn.init.xavier_uniform_(last.weight)
    torch.nn.init.zeros_(last.bias)
    decoder = torch.nn.Sequential(*block(64, 1024), *block(1024, 1024), *block(1024, 1024), last, torch.nn.Sigmoid()).to(dev).eval()
    eng = engine.MinimizerEngine(seq=g.seq, table=table, device=local_rank)
    st = torch.cuda.current_stream(dev)
    eng.ctx.set_stream(st.cuda_stream)
    try:
        space = engine.ColumnSpace(table, cols, essential)
        first = rank * S
        with torch.no_grad():
            torch.manual_seed(5 + 1000 * rank)          # same weights everywhere, different samples per rank
            z = torch.randn(S, 64, device=dev)
            probs = decoder(z)
            lengths, counts = engine.plan_from_probabilities(eng, space, probs, first_idx=first)
            off = eng.ctx.record_offsets()
            image = torch.empty(int(off[-1]), dtype=torch.uint8, device=dev)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            t_dec = t_plan = t_emit = 0.0
            for it in range(steps + 2):
                ev[0].record(st)
                probs = decoder(z)
                ev[1].record(st)
                eng.ctx.load_probs_dev(probs.data_ptr(), S, probs.stride(0), 0.5)
                eng.ctx.plan_async(first)
                ev[2].record(st)
                eng.ctx.emit_dev(0, S, image.data_ptr(), image.numel())
                ev[3].record(st)
                torch.cuda.synchronize(dev)
                if it >= 2:
                    t_dec += ev[0].elapsed_time(ev[1]); t_plan += ev[1].elapsed_time(ev[2]); t_emit += ev[2].elapsed_time(ev[3])
            # parity: the reference chain on three samples of this rank
            starts, ends = g.starts_ends()
            pick = sorted({0, S // 2, S - 1})
            host = probs[pick].float().cpu().numpy()
            lists = co.add_essentials(co.masks_to_gene_lists(co.threshold_samples(host), cols), essential)
            ok = True
            for s, needed in zip(pick, lists):
                keep = mo.keep_vector(table.names, needed)
                L, H, _ = c_oracle.batch(g.seq, starts, ends, synth.pack_keep_rows(keep[None, :]), first_idx=first + s)
                got = eng.ctx.diag_range_hashes(image.data_ptr(), image.numel(), off[s:s + 2])
                ok = ok and int(L[0]) == int(lengths[s]) and int(H[0]) == int(got[0]) and len(needed) == int(counts[s])
        kept = int(lengths.sum())
        t = torch.tensor([t_dec, t_plan, t_emit, 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
        k = torch.tensor([kept], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(k, op=dist.ReduceOp.SUM)
        t_dec, t_plan, t_emit, bad = (float(x) for x in t.tolist())
        if bad:
            raise SystemExit("bench.py: config 5 leg differs from the converter + minimizer oracles")
        n = steps
        return {"workload": f"C5: VAE v0 decode -> `> 0.5` -> column -> gene keep mask -> minimize, {S * world} samples over {world} GPU(s), V = {columns} columns",
                "samples_total": S * world, "samples_per_rank": S, "decode_ms": t_dec / n, "keepmask_plan_ms": t_plan / n, "emit_ms": t_emit / n,
                "gbp_per_s_after_decode": int(k.item()) / ((t_plan + t_emit) / n * 1e-3) / 1e9,
                "gbp_per_s_incl_decode": int(k.item()) / ((t_dec + t_plan + t_emit) / n * 1e-3) / 1e9,
                "mean_retained_fraction": float(lengths.mean() / g.G), "mean_list_length": float(counts.mean()),
                "image_gb_per_gpu": image.numel() / 1e9, "records_checked": 3 * world, "byte_identical": True,
                "timing": "CUDA events per phase, max over ranks; the decoder is torch (library GEMMs), outside the graded kernels"}
    finally:
        eng.close()


# ----------------------------------------------------------------------------------------------
# ours
# ----------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from genome_minimizer_2_b200 import _native, engine

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

    g, gdesc = make_genome(args.genome)
    wname = ("C2: " if args.genome == "k12" else "C4-shaped: ") + gdesc
    starts, ends = g.starts_ends()
    table = engine.GeneTable(g.gene_names(), starts, ends)
    S = args.samples
    first_idx = rank * S                       # global record ids: rank order == file order
    job = Job(table, S * world, args.retention, args.noise_ids, seed=2)
    ids, off, _ = job.csr(first_idx, first_idx + S)

    def make_ctx(on_stream=None):
        c = _native.Context(local_rank)
        for key, val, on in ((_native.CFG_TILE_BYTES, args.tile_bytes, args.tile_bytes > 0),
                             (_native.CFG_PACKING, args.packing, args.packing > 0),
                             (_native.CFG_EMIT_WARPS, args.emit_warps, args.emit_warps > 0),
                             (_native.CFG_EMIT_BATCH, args.emit_batch, args.emit_batch >= 0),
                             (_native.CFG_STORE_POLICY, args.store_policy, args.store_policy >= 0),
                             (_native.CFG_ORDER, args.emit_order, args.emit_order >= 0),
                             (_native.CFG_FLAT_RUN_BYTES, args.flat_run_bytes, args.flat_run_bytes >= 0),
                             (_native.CFG_FLAT_MODE, args.flat_mode, args.flat_mode > 0),
                             (_native.CFG_RUN_TABLE, args.run_table, args.run_table > 0),
                             (_native.CFG_EMIT_OCCUPANCY, args.emit_occupancy, args.emit_occupancy >= 0),
                             (_native.CFG_WIRE, args.wire, args.wire >= 0),
                             (_native.CFG_HOST_THREADS, args.host_threads, args.host_threads >= 0),
                             (_native.CFG_DEBUG, args.emit_debug, args.emit_debug != 0)):
            if on:
                c.configure(key, val)
        c.set_stream((on_stream or stream).cuda_stream)
        return c

    if args.emit_debug:
        args.verify = 0
    # real (non-default) torch streams: the kernels are launched on them and the CUDA events below are
    # recorded on them.  (torch's default stream has handle 0, which gm2_set_stream reads as "own stream".)
    prio = [int(x) for x in args.stream_priorities.split(",")]
    stream = torch.cuda.Stream(dev, priority=prio[0])
    stream_b = torch.cuda.Stream(dev, priority=prio[1])
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0 and stream_b.cuda_stream != 0
    # Two contexts on two streams take the steps in turn: a context is one plan slot, so step k runs on context
    # k & 1 and its plan (K1-K3, issue/latency-bound) is issued while the other context's emit of step k-1
    # (write-bound) is still running; gm2_order_after keeps the emits themselves one after the other:
    #   stream A:  plan(k) | .......... emit(k) .......... |                  plan(k+2) | ....
    #   stream B:            plan(k+1) |                     .......... emit(k+1) ..........
    ctx = make_ctx()
    ctx_b = ctx if args.no_pipeline else make_ctx(stream_b)
    pair = (ctx, ctx_b)
    streams = (stream, stream if args.no_pipeline else stream_b)

    # device-resident inputs
    d_ids = torch.from_numpy(ids).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    for c in set(pair):
        c.set_reference(g.seq, starts, ends)
        c.set_name_map(table.id2gene_off, table.id2gene_idx)
        c.load_ids_dev(d_ids.data_ptr(), d_off.data_ptr(), S, ids.size)
        c.plan(first_idx)                      # sizing pass (outside the timed region)
    lengths = ctx.lengths()
    rec_off = ctx.record_offsets()
    image_bytes = int(rec_off[-1])
    kept_bases = int(lengths.sum())
    image = torch.empty(image_bytes, dtype=torch.uint8, device=dev)
    my_len = torch.zeros(S, dtype=torch.int64, device=dev)
    all_len = torch.zeros(S * world, dtype=torch.int64, device=dev)

    def step(k, ev=None):
        """One pass over the batch on context k & 1.  ev: optional [plan issued, emit start, emit end] events."""
        c, o, st = pair[k & 1], pair[(k + 1) & 1], streams[k & 1]
        if ev: ev[0].record(st)
        c.plan_async(first_idx)                # runs under the other context's emit of step k-1
        if world > 1:
            c.lengths_dev(my_len.data_ptr())
            if c is not ctx:
                ctx.order_after(c)             # the collective is issued on stream A: this step's lengths first
        c.order_after(o)                       # emit(k) after emit(k-1)
        if ev: ev[1].record(st)
        c.emit_dev(0, S, image.data_ptr(), image_bytes)
        if ev: ev[2].record(st)
        if world > 1:                          # this step's per-sample lengths -> every rank (global file offsets)
            dist.all_gather_into_tensor(all_len, my_len)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for k in range(2 * ((max(args.warmup, 3) + 1) // 2)):
        step(k)
    barrier()

    # the plan and k_emit alone (nothing else on the GPU): the plan's own cost, and k_emit unperturbed
    solo = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    solo_plan, solo_emit = [], []
    for _ in range(3):
        solo[0].record(stream)
        ctx.plan_async(first_idx)
        solo[1].record(stream)
        ctx.emit_dev(0, S, image.data_ptr(), image_bytes)
        solo[2].record(stream)
        torch.cuda.synchronize(dev)
        solo_plan.append(solo[0].elapsed_time(solo[1]))
        solo_emit.append(solo[1].elapsed_time(solo[2]))
    barrier()

    # timed region: exactly K steps, CUDA events on the launching streams, clocks sampled meanwhile
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    l0 = sum(c.query(_native.Q_LAUNCHES) for c in set(pair))
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_stop = torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    ctx_b.order_after(ctx)                     # stream B starts inside the timed region
    for k in range(args.steps):
        step(k, ev[k])
    ctx.order_after(ctx_b)                     # ... and ends inside it
    t_stop.record(stream)
    barrier()
    launches = sum(c.query(_native.Q_LAUNCHES) for c in set(pair)) - l0
    total_ms = t_start.elapsed_time(t_stop)
    emit_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))   # k_emit inside the timed region (a plan may run beside it)
    plan_ms = float(np.mean(solo_plan))        # the plan alone; inside the timed region it runs under the previous step's emit
    clocks = sampler.stop() if rank == 0 else None
    emit_ctas = ctx.query(_native.Q_LAST_EMIT_CTAS)
    tile_bytes_used = ctx.query(_native.Q_TILE_BYTES)

    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        kb = torch.tensor([kept_bases, launches], dtype=torch.int64, device=dev)
        dist.all_reduce(kb, op=dist.ReduceOp.SUM)
        kept_all, launches = int(kb[0].item()), int(kb[1].item())
        if not np.array_equal(all_len.cpu().numpy()[rank * S:(rank + 1) * S], lengths):
            raise SystemExit("bench.py: the gathered lengths are not this rank's lengths")
    else:
        kept_all = kept_bases
    ms_per_step = total_ms / args.steps
    value = kept_all / (ms_per_step * 1e-3) / 1e9

    # correctness check of the timed output against the oracle (device hashes, no big copies)
    verify = {}
    if args.verify > 0:
        pick = sorted(set(np.linspace(0, S - 1, args.verify).astype(int).tolist()))
        Ls, Hs = oracle_records(g, job, [first_idx + s for s in pick])
        ok = True
        for s, L, H in zip(pick, Ls, Hs):
            got = ctx.diag_range_hashes(image.data_ptr(), image_bytes, rec_off[s:s + 2])
            ok = ok and L == int(lengths[s]) and H == int(got[0])
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
    del image
    torch.cuda.empty_cache()

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = FALLBACK_HBM_GBS, "fallback"
    if os.path.exists(peaks_path):
        try:
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
        except Exception:
            pass

    def side_leg(name, fn):
        """The legs beside the headline step.  A parity failure inside one aborts the run (SystemExit); any other
        failure (a full tmpfs, an allocation on a busy box) is recorded in the leg's place instead of costing the
        whole line — on one GPU only: with several ranks a leg's collectives need every rank, so errors propagate."""
        if world > 1:
            return fn()
        try:
            return fn()
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f"bench.py: leg '{name}' failed: {e!r}\n")
            torch.cuda.synchronize(dev)
            return {"error": repr(e)[:400]}

    sweep = None
    if world == 1 and not args.no_sweep and args.genome == "k12" and not args.emit_debug:
        sweep = side_leg("retention_sweep", lambda: retention_sweep(ctx, g, make_ctx, torch, dev, stream, peak))
        ctx.set_name_map(table.id2gene_off, table.id2gene_idx)

    sharded = None
    if not args.no_sharded and args.genome == "k12" and not args.emit_debug:
        job3 = Job(table, args.job_samples, args.retention, args.noise_ids, seed=3)
        sharded = {"job": f"C3: {args.job_samples} samples, K-12-shaped genome, gene retention {args.retention}, cut over {world} rank(s)",
                   "device": side_leg("sharded.device", lambda: sharded_device_leg(pair, g, job3, torch, dist, dev, stream, rank, world, barrier, S))}
        sharded["file"] = side_leg("sharded.file", lambda: sharded_file_leg(g, table, job3, rank, world, barrier, args.file_gb_per_rank))
    config5 = None
    if not args.no_c5 and args.genome == "k12" and not args.emit_debug:
        config5 = side_leg("config5", lambda: config5_leg(g, table, torch, dist, dev, local_rank, rank, world, barrier))
    dropin = None
    if not args.no_dropin and args.genome == "k12":
        dropin = side_leg("dropin_c1", lambda: dropin_entry(g, table, rank, world, barrier))
        if sharded is not None:
            sharded["entry"] = dropin

    # end to end through the C-ABI with host buffers
    e2e = None
    if not args.no_e2e:
        budget = int(mem_available() * 0.35 / max(world, 1))
        S_e = S
        while S_e > 1 and int(rec_off[S_e]) > budget:
            S_e //= 2
        e_bytes = int(rec_off[S_e])
        pinned = _native.PinnedBuffer(e_bytes)
        # the step's inputs live in pinned host memory too (the H2D copy is part of every timed call)
        n_ids_e = int(off[S_e])
        pin_ids = _native.PinnedBuffer(max(4 * n_ids_e, 16))
        pin_off = _native.PinnedBuffer(8 * (S_e + 1))
        ids_e = pin_ids.array[:4 * n_ids_e].view(np.int32)
        off_e = pin_off.array.view(np.int64)
        ids_e[:] = ids[:n_ids_e]
        off_e[:] = off[:S_e + 1]
        ctx.load_ids_dev(d_ids.data_ptr(), d_off.data_ptr(), S, ids.size)
        ctx.set_stream(None)
        # raw pinned D2H rate of this host path with all ranks copying at once: the ceiling of the image-bytes transport
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
        del dsrc, hview
        # host write ceiling of the two-bit transport: the expansion's threads filling the same buffer with
        # non-temporal stores, all ranks at once
        barrier()
        fill_host = _native.host_fill_gbs(pinned.array[:min(e_bytes, 8 << 30)], threads=max(args.host_threads, 0), reps=3)
        if world > 1:
            t = torch.tensor([raw_gbs, fill_host], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            raw_gbs, fill_host = float(t[0].item()), float(t[1].item())
        ctx.minimize_host(pinned, ids=ids_e, off=off_e, first_idx=first_idx)
        times = []
        for _ in range(args.e2e_steps):
            barrier()
            t0 = time.perf_counter()
            le, ro = ctx.minimize_host(pinned, ids=ids_e, off=off_e, first_idx=first_idx)
            barrier()
            times.append(time.perf_counter() - t0)
        wire_used = ctx.query(_native.Q_LAST_WIRE)
        d2h_moved = ctx.query(_native.Q_LAST_D2H_BYTES)
        # every record of the host image against the oracle (N = 1), a spread of 256 per rank otherwise
        from oracle import c_oracle
        from concurrent.futures import ThreadPoolExecutor
        chk = list(range(S_e)) if world == 1 else sorted(set(np.linspace(0, S_e - 1, 256).astype(int).tolist()))
        Ls, Hs = oracle_records(g, job, [first_idx + s for s in chk], threads=max((os.cpu_count() or 1) // world, 1))
        with ThreadPoolExecutor(max((os.cpu_count() or 1) // world, 1)) as ex:
            got = list(ex.map(lambda s: c_oracle.range_hash(pinned.array[int(ro[s]):int(ro[s + 1])]), chk))
        if any(int(le[s]) != L for s, L in zip(chk, Ls)) or got != Hs:
            raise SystemExit("bench.py: e2e host image differs from the oracle")
        # the other transport beside it (image bytes over PCIe), one warm-up + two timed calls
        dt_bytes = None
        if wire_used == 2:
            ctx.configure(_native.CFG_WIRE, 1)
            ctx.minimize_host(pinned, ids=ids_e, off=off_e, first_idx=first_idx)
            tb = []
            for _ in range(2):
                barrier()
                t0 = time.perf_counter()
                ctx.minimize_host(pinned, ids=ids_e, off=off_e, first_idx=first_idx)
                barrier()
                tb.append(time.perf_counter() - t0)
            dt_bytes = min(tb)
            ctx.configure(_native.CFG_WIRE, args.wire if args.wire >= 0 else 0)
        dt_mean, dt_med, dt_min = float(np.mean(times)), float(np.median(times)), float(np.min(times))
        kept_e = int(le.sum())
        if world > 1:
            t = torch.tensor([dt_mean, dt_med, dt_min, dt_bytes or 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_mean, dt_med, dt_min, dt_bytes = (float(x) for x in t.tolist())
            dt_bytes = dt_bytes or None
            kb = torch.tensor([kept_e], dtype=torch.int64, device=dev)
            dist.all_reduce(kb, op=dist.ReduceOp.SUM)
            kept_e = int(kb.item())
        delivered = e_bytes / dt_mean / 1e9
        e2e = {"value": kept_e / dt_mean / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(ids_e.nbytes + off_e.nbytes),
               "d2h_bytes_per_step": int(d2h_moved + le.nbytes + ro.nbytes),
               "steps": args.e2e_steps, "ms_per_step": dt_mean * 1e3, "ms_median": dt_med * 1e3, "ms_min": dt_min * 1e3,
               "value_median": kept_e / dt_med / 1e9, "value_best": kept_e / dt_min / 1e9,
               "samples_per_step": S_e, "image_bytes_per_step": e_bytes,
               "delivered_image_gbs_per_gpu": delivered,
               "host_write_ceiling_gbs_per_rank": fill_host,
               "e2e_frac_of_host_ceiling": delivered / fill_host if fill_host else None,
               # what the transport moves through the host's memory controllers per delivered byte: the image itself,
               # plus (two-bit form) 0.25 B written by the DMA and 0.25 B read back by the decoder, or (image bytes)
               # 1 B written by the DMA and read back by the copy into the caller's buffer
               "host_memory_traffic_gbs_per_rank": delivered * (1.5 if wire_used == 2 else 3.0),
               "host_traffic_frac_of_host_ceiling": (delivered * (1.5 if wire_used == 2 else 3.0) / fill_host) if fill_host else None,
               "raw_pinned_d2h_gbs_per_gpu": raw_gbs,
               "transport": ("two bits per base over PCIe (k_emit_packed), expanded into the caller's buffer by host threads"
                             if wire_used == 2 else "image bytes over PCIe (k_emit)"),
               "image_bytes_transport": (None if dt_bytes is None else
                                         {"value": kept_e / dt_bytes / 1e9, "ms_per_step": dt_bytes * 1e3,
                                          "delivered_image_gbs_per_gpu": e_bytes / dt_bytes / 1e9,
                                          "frac_of_raw_pinned_d2h": e_bytes / dt_bytes / 1e9 / raw_gbs}),
               "records_checked": len(chk), "byte_identical": True,
               "note": ("host_write_ceiling = the expansion's threads filling the same pinned buffer with non-temporal stores, "
                        "all ranks at once (ceiling of the two-bit transport); raw_pinned_d2h = plain pinned cudaMemcpy of "
                        "image-sized data, all ranks at once (ceiling of the image-bytes transport)"),
               "api": "gm2_minimize_host (C-ABI): pinned host id lists in, pinned host FASTA image out"}
        pinned.free()
        del ids_e, off_e
        pin_ids.free()
        pin_off.free()

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # roofline of the dominant kernel (k_emit)
    F = table.F
    alg_bytes = image_bytes + S * ((F + 7) // 8) + g.G + 16 * F
    achieved = alg_bytes / (emit_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get(f"{args.genome}:{S}")
            if traffic is not None:
                traffic_src = tj.get("source", "profiles/ (ncu --set full capture of k_emit at this config), not this run")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_emit", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": emit_ms,
                "kernel_ms_alone": float(np.mean(solo_emit)), "frac_alone": alg_bytes / (float(np.mean(solo_emit)) * 1e-3) / 1e9 / peak,
                "plan_ms": plan_ms,
                "plan_overlap": ("none (--no-pipeline)" if args.no_pipeline else
                                 "the plan (K1-K3) of step k+1 runs under k_emit of step k: two contexts on two streams take the "
                                 "steps in turn, gm2_order_after keeps the emits in order; plan_ms is the plan alone"),
                "step_frac": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak if world == 1 else None,
                "write_fill_gbs": fill_gbs}

    cpu_baseline = cpu_c = None
    if not args.no_cpu_baseline and world == 1:
        v, dt, kind = cpu_single_thread(g, table, job, args.cpu_samples)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                        "sample": f"first {args.cpu_samples} samples of the workload, "
                                  + ("the reference's own GenomeMinimiser (baseline/_ref)" if kind == "reference"
                                     else "oracle literal port of minimizer_2.py:50-101")
                                  + f" (list scan + position set + per-base loop), {dt:.1f} s",
                        "host_cores_available": os.cpu_count(), "cpu_model": cpu_model()}
        cpu_c = cpu_c_port_all_cores(g, job, min(16 * (os.cpu_count() or 1), S))
    if dropin is not None and world == 1:
        dropin["lists_tokenize"] = lists_tokenize_bench(g, table)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wname, "samples_per_gpu": S, "retention": args.retention,
                   "noise_ids_per_sample": args.noise_ids, "input": "device-resident name-id lists (CSR)",
                   "output": f"device-resident FASTA image, {image_bytes/1e9:.2f} GB per GPU per step",
                   "l2": "no flush needed: each step writes an image >> 126 MB L2",
                   "sharding": "samples; reference replicated; all-gather of the step's per-sample lengths only",
                   "tile_bytes": tile_bytes_used, "kept_bases_per_gpu": kept_bases,
                   "emit_ctas_per_sm": emit_ctas, "cpu_model": cpu_model(), "host_cores": os.cpu_count()},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "cpu_port_c": cpu_c, "verify": verify,
        "sharded": sharded, "config5": config5, "retention_sweep": sweep, "dropin_c1": dropin,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
