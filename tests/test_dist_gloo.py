"""World-size-2 gloo test of the sharded entry point's host logic (shard ranges, length
all-gather, global offsets, pwrite placement).  No GPU here, so the per-rank compute is a test
double backed by the oracle; the product's own engine is exercised by the -m gpu tests."""
from __future__ import annotations

import contextlib
import io
import json
import os

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden


class OracleEngine:
    """Stand-in with MinimizerEngine's plan_lists / drain surface (tests only)."""

    def __init__(self, record):
        from oracle import minimizer_oracle as mo
        self.mo, self.record = mo, record
        self.images, self.first = [], 0

    def plan_lists(self, lists, first_idx=0):
        self.first = first_idx
        seqs = [self.mo.minimize_literal(self.record, l) for l in lists]
        self.images = [self.mo.record_bytes(first_idx + i, s.encode()) for i, s in enumerate(seqs)]
        return np.asarray([len(s) for s in seqs], dtype=np.int64)

    # the surface engine.drain_to_file uses (mapped file: chunks + emit_into; portable form: drain)
    @property
    def S(self):
        return len(self.images)

    @property
    def ctx(self):
        return self

    def record_offsets(self):
        return np.concatenate([[0], np.cumsum([len(x) for x in self.images])]).astype(np.int64)

    def chunks(self, max_bytes=0, s0=0, s1=None):
        s1 = self.S if s1 is None else s1
        return [(i, i + 1) for i in range(s0, s1)]           # one record per chunk: worst case for offsets

    def emit_into(self, s0, s1, out):
        out[:] = np.frombuffer(b"".join(self.images[s0:s1]), dtype=np.uint8)

    def drain(self, sink, max_bytes=0, s0=0, s1=None):
        for a, b in self.chunks(max_bytes, s0, s1):
            sink(a, b, np.frombuffer(b"".join(self.images[a:b]), dtype=np.uint8))


def _worker_multi(rank, world, port, case_name, out_dir, ret_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import tempfile
        from oracle import genbank_reader
        from genome_minimizer_2_b200 import dist as gdist
        case = load_golden(case_name)
        with tempfile.NamedTemporaryFile("w", suffix=".gb", delete=False) as fh:
            fh.write(case["genbank"])
        rec = genbank_reader.read_genbank(fh.name)
        os.unlink(fh.name)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ret = gdist.run_multi_file_sharded(rec, case["lists"], case["model_name"], out_dir,
                                               make_engine=lambda: OracleEngine(rec))
        with open(f"{ret_path}.{rank}", "w") as fh:
            json.dump({"ret": ret, "stdout": buf.getvalue()}, fh)
    finally:
        dist.destroy_process_group()


def _worker(rank, world, port, case_name, out_path, ret_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import tempfile
        from oracle import genbank_reader
        from genome_minimizer_2_b200 import dist as gdist
        case = load_golden(case_name)
        with tempfile.NamedTemporaryFile("w", suffix=".gb", delete=False) as fh:
            fh.write(case["genbank"])
        rec = genbank_reader.read_genbank(fh.name)
        os.unlink(fh.name)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ret = gdist.run_single_file_sharded(rec, case["lists"], case["model_name"], out_path,
                                                make_engine=lambda: OracleEngine(rec), timestamp="<TS>")
        with open(f"{ret_path}.{rank}", "w") as fh:
            json.dump({"ret": ret, "stdout": buf.getvalue()}, fh)
    finally:
        dist.destroy_process_group()


def _run(case_name, tmp_path, world=2):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "sharded.fasta")
    retp = str(tmp_path / "ret")
    mp.spawn(_worker, args=(world, port, case_name, out, retp), nprocs=world, join=True)
    case = load_golden(case_name)
    assert open(out, "rb").read().decode() == case["single_file"]
    r = [json.load(open(f"{retp}.{k}")) for k in range(world)]
    assert all(x["ret"] == case["single_return"] for x in r)
    assert r[0]["stdout"] == case["single_stdout"]
    assert all(x["stdout"] == "" for x in r[1:])


def test_two_ranks_reproduce_the_reference_file_kat(tmp_path):
    _run("kat_appB", tmp_path)


def test_two_ranks_hundred_and_one_samples(tmp_path):
    _run("hundred_and_one", tmp_path)          # ids cross 9->10 and 99->100 digits inside shards


def test_three_ranks_uneven_shards(tmp_path):
    _run("rand_small_1", tmp_path, world=3)    # 17 samples over 3 ranks


def test_two_ranks_multi_file_mode(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "multi")
    retp = str(tmp_path / "ret")
    mp.spawn(_worker_multi, args=(2, port, "hundred_and_one", out, retp), nprocs=2, join=True)
    case = load_golden("hundred_and_one")
    files = {fn: open(os.path.join(out, fn), "rb").read().decode() for fn in sorted(os.listdir(out))}
    assert files == case["multi_files"]
    r = [json.load(open(f"{retp}.{k}")) for k in range(2)]
    assert all(x["ret"] == case["multi_return"] for x in r)
    assert r[0]["stdout"].replace(out, "<OUTDIR>") == case["multi_stdout"] and r[1]["stdout"] == ""


# ----------------------------------------------------------------------------------------------
# the entry functions themselves under a torchrun-style environment (WORLD_SIZE / RANK / MASTER_*):
# they join the job and shard the samples without being told to (minimizer_2._ranks, dist.ensure_process_group)
# ----------------------------------------------------------------------------------------------
def _worker_entry(rank, world, port, case_name, mode, target, ret_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import engine_double
    from genome_minimizer_2_b200 import minimizer_2 as m2
    engine_double.install()                                   # no GPU here
    case = load_golden(case_name)
    work = os.path.join(os.path.dirname(ret_path), f"in{rank}")
    os.makedirs(work, exist_ok=True)
    gb, npy = os.path.join(work, "g.gb"), os.path.join(work, "l.npy")
    with open(gb, "w") as fh:
        fh.write(case["genbank"])
    arr = np.empty(len(case["lists"]), dtype=object)
    for i, l in enumerate(case["lists"]):
        arr[i] = l
    np.save(npy, arr, allow_pickle=True)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        if mode == "single":
            ret = m2.process_multiple_genomes_single_file(gb, npy, case["model_name"], target)
        else:
            ret = m2.process_multiple_genomes_multiple_files(gb, npy, case["model_name"], target)
    assert dist.is_initialized() and dist.get_world_size() == world
    with open(f"{ret_path}.{rank}", "w") as fh:
        json.dump({"ret": ret, "stdout": buf.getvalue()}, fh)


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_entry_function_shards_by_itself_single_file(tmp_path):
    out, retp = str(tmp_path / "o" / "auto.fasta"), str(tmp_path / "ret")
    mp.spawn(_worker_entry, args=(2, _free_port(), "hundred_and_one", "single", out, retp), nprocs=2, join=True)
    case = load_golden("hundred_and_one")
    lines = open(out, "rb").read().decode().split("\n")
    assert lines[2].startswith("# Generated on: ")
    lines[2] = "# Generated on: <TS>"
    assert "\n".join(lines) == case["single_file"]
    r = [json.load(open(f"{retp}.{k}")) for k in range(2)]
    assert all(x["ret"] == case["single_return"] for x in r)
    assert r[0]["stdout"] == case["single_stdout"] and r[1]["stdout"] == ""


def test_entry_function_shards_by_itself_multi_file(tmp_path):
    out, retp = str(tmp_path / "many"), str(tmp_path / "ret")
    mp.spawn(_worker_entry, args=(3, _free_port(), "rand_small_1", "multi", out, retp), nprocs=3, join=True)
    case = load_golden("rand_small_1")
    files = {fn: open(os.path.join(out, fn), "rb").read().decode() for fn in sorted(os.listdir(out))}
    assert files == case["multi_files"]
    r = [json.load(open(f"{retp}.{k}")) for k in range(3)]
    assert all(x["ret"] == case["multi_return"] for x in r)
    assert r[0]["stdout"].replace(out, "<OUTDIR>") == case["multi_stdout"] and all(x["stdout"] == "" for x in r[1:])


def test_gm2_shard_0_keeps_the_whole_job_in_every_rank(monkeypatch):
    from genome_minimizer_2_b200 import dist as gdist, minimizer_2 as m2
    for k in ("RANK", "MASTER_ADDR", "MASTER_PORT"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("WORLD_SIZE", "4")
    assert m2._ranks() == 1 and gdist.launched_ranks() == 1      # a stray WORLD_SIZE is not a torchrun job
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("MASTER_ADDR", "127.0.0.1")
    monkeypatch.setenv("MASTER_PORT", "29500")
    assert m2._ranks() == 4 and gdist.launched_ranks() == 4
    monkeypatch.setenv("GM2_SHARD", "0")
    assert m2._ranks() == 1
    monkeypatch.delenv("GM2_SHARD")
    monkeypatch.setenv("WORLD_SIZE", "1")
    assert m2._ranks() == 1
