"""Two ranks, two GPUs, NCCL: the sharded single-file entry point with the REAL engine must
reproduce the reference's file.  Skipped unless >= 2 CUDA devices are visible (gpurun --gpus 2)."""
from __future__ import annotations

import contextlib
import io
import json
import os
import socket

import pytest

pytestmark = pytest.mark.gpu

from conftest import load_golden  # noqa: E402


def _worker(rank, world, port, case_name, out_path, ret_path):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import tempfile
        from genome_minimizer_2_b200 import dist as gdist, genbank
        case = load_golden(case_name)
        with tempfile.NamedTemporaryFile("w", suffix=".gb", delete=False) as fh:
            fh.write(case["genbank"])
        rec = genbank.read_genbank(fh.name)
        os.unlink(fh.name)
        lists = case["lists"]
        if case_name in ("hundred_and_one", "medium_k12"):
            # the way a driver script gets them: the .npy file, tokenised natively, sliced per rank
            from genome_minimizer_2_b200 import engine, synth
            npy = f"{ret_path}.lists.{rank}.npy"
            synth.save_gene_lists(npy, lists)
            lists = engine.load_gene_lists(npy, engine.GeneTable.from_record(rec))
            assert isinstance(lists, engine.TokenizedLists)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ret = gdist.run_single_file_sharded(rec, lists, case["model_name"], out_path, timestamp="<TS>")
        with open(f"{ret_path}.{rank}", "w") as fh:
            json.dump({"ret": ret, "stdout": buf.getvalue()}, fh)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case_name", ["kat_appB", "hundred_and_one", "rand_small_1", "medium_k12"])
def test_two_gpus_reproduce_the_reference_file(case_name, tmp_path):
    import hashlib
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "sharded.fasta")
    retp = str(tmp_path / "ret")
    mp.spawn(_worker, args=(2, port, case_name, out, retp), nprocs=2, join=True)
    case = load_golden(case_name)
    data = open(out, "rb").read()
    if "single_file" in case:
        assert data.decode() == case["single_file"]
    else:
        assert hashlib.sha256(data).hexdigest() == case["single_file_sha256"]
    r = [json.load(open(f"{retp}.{k}")) for k in range(2)]
    assert all(x["ret"] == case["single_return"] for x in r)
    assert r[0]["stdout"] == case["single_stdout"] and r[1]["stdout"] == ""


def _worker_multi(rank, world, port, case_name, out_dir, ret_path):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import tempfile
        from genome_minimizer_2_b200 import dist as gdist, genbank
        case = load_golden(case_name)
        with tempfile.NamedTemporaryFile("w", suffix=".gb", delete=False) as fh:
            fh.write(case["genbank"])
        rec = genbank.read_genbank(fh.name)
        os.unlink(fh.name)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ret = gdist.run_multi_file_sharded(rec, case["lists"], case["model_name"], out_dir)
        with open(f"{ret_path}.{rank}", "w") as fh:
            json.dump({"ret": ret, "stdout": buf.getvalue()}, fh)
    finally:
        dist.destroy_process_group()


def test_two_gpus_multi_file_mode(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
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
    assert r[0]["stdout"].replace(out, "<OUTDIR>") == case["multi_stdout"]
