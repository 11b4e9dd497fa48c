"""Multi-GPU form of the batch entry points: one process per GPU (torchrun), samples sharded.

Samples are independent (minimizer_2.py:469-477 reads only the shared record and the sample's own
list), so rank r of R takes the contiguous range [r*S//R, (r+1)*S//R): rank order is file order
and record ids stay global (`first_idx`).  The reference genome is replicated.  The only exchange
is an all-gather of the per-sample lengths (8 bytes per sample) from which every rank derives the
global byte offset of its shard; the FASTA bytes never cross NVLink — each rank `pwrite`s its own
shard into the shared output file.  Backend: NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import engine as _engine


def ensure_process_group() -> None:
    """Join the job torchrun started (env:// rendezvous: RANK, WORLD_SIZE, MASTER_ADDR/PORT) unless the caller
    has already initialised a process group.  NCCL when a GPU is visible (device = LOCAL_RANK), gloo otherwise
    (CPU tests).  A group created here is destroyed at interpreter exit."""
    if dist.is_initialized():
        return
    if torch.cuda.is_available():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    import atexit

    def _teardown():
        if dist.is_initialized():
            dist.destroy_process_group()

    atexit.register(_teardown)


def launched_ranks() -> int:
    """How many ranks share this job: the size of an initialised process group, else torchrun's WORLD_SIZE.
    GM2_SHARD=0 turns the automatic sharding of the entry functions off (every rank then does the whole job,
    which is what the reference's CLI would do under torchrun)."""
    if os.environ.get("GM2_SHARD", "1") == "0":
        return 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    try:
        return max(int(os.environ.get("WORLD_SIZE", "1")), 1)
    except ValueError:
        return 1


def _device_for_backend() -> torch.device:
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_gather_lengths(local: np.ndarray) -> List[np.ndarray]:
    """All-gather of variable-length int64 vectors (padded to the longest), in rank order."""
    world = dist.get_world_size()
    dev = _device_for_backend()
    n = torch.tensor([local.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    width = max(max(sizes), 1)
    buf = torch.zeros(width, dtype=torch.int64, device=dev)
    buf[:local.size] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.int64)).to(dev)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return [p[:k].cpu().numpy() for p, k in zip(parts, sizes)]


def header_len(idx: int) -> int:
    """Bytes of '>' + prefix + decimal(idx+1) + '\\n' for the global 0-based sample index."""
    return 1 + len(_engine.SEQ_ID_PREFIX) + len(str(idx + 1)) + 1


def run_single_file_sharded(record, all_lists: Sequence, model_name: str, output_file: str,
                            make_engine: Optional[Callable[[], object]] = None,
                            timestamp: Optional[str] = None) -> dict:
    """`process_multiple_genomes_single_file` (reference :447-495) over all ranks of the default
    process group.  Every rank returns the same dict; rank 0 prints the progress lines."""
    rank, world = dist.get_rank(), dist.get_world_size()
    n = len(all_lists)
    G = len(record.seq)
    lo, hi = _engine.shard_range(n, rank, world)
    eng = make_engine() if make_engine is not None else _engine.MinimizerEngine(record)
    try:
        local_len = eng.plan_lists(all_lists[lo:hi], first_idx=lo)
        gathered = all_gather_lengths(np.asarray(local_len, dtype=np.int64))
        lengths = np.concatenate(gathered) if gathered else np.zeros(0, dtype=np.int64)
        assert lengths.size == n
        rec_sizes = np.asarray([header_len(i) for i in range(n)], dtype=np.int64) + lengths + 1
        pre = (f"# Minimized genomes generated using model: {model_name}\n"
               f"# Total genomes: {n}\n"
               f"# Generated on: {timestamp if timestamp is not None else np.datetime64('now')}\n").encode()
        rec_off = np.zeros(n + 1, dtype=np.int64)
        rec_off[1:] = np.cumsum(rec_sizes)
        if rank == 0:
            with open(output_file, "wb") as fh:
                fh.write(pre)
                fh.truncate(len(pre) + int(rec_off[-1]))
        dist.barrier()
        fd = os.open(output_file, os.O_WRONLY)
        try:
            base = len(pre) + int(rec_off[lo])
            pos = [base]

            def sink(sa: int, sb: int, view: np.ndarray) -> None:
                os.pwrite(fd, view, pos[0])
                pos[0] += view.size

            eng.drain(sink)
            assert pos[0] == len(pre) + int(rec_off[hi])
        finally:
            os.close(fd)
        dist.barrier()
    finally:
        if make_engine is None:
            eng.close()
    if rank == 0:
        for idx in range(n):
            print(f"[{idx+1}/{n}] genes present: {len(all_lists[idx])}")
            if _engine._sampled(idx):
                L = int(lengths[idx])
                print(f"  → {L:,} bp ({_engine._pct(G, L):.1f}% reduction)")
    tot_red, tot_len = 0.0, 0
    for idx in range(n):
        if _engine._sampled(idx):
            tot_red += _engine._pct(G, int(lengths[idx]))
            tot_len += int(lengths[idx])
    return {"genome_count": n, "average_reduction_pct": tot_red / n, "average_length_bp": tot_len / n}


def run_multi_file_sharded(record, all_lists: Sequence, model_name: str, output_dir,
                           filename_template: str = "minimized_{model}_{idx:04d}.fasta",
                           make_engine: Optional[Callable[[], object]] = None) -> dict:
    """`process_multiple_genomes_multiple_files` (reference :499-560) over all ranks: every rank writes
    the files of its own contiguous shard (global idx in the names and record ids); the lengths are
    all-gathered so that every rank returns the reference's dict and rank 0 prints its lines."""
    rank, world = dist.get_rank(), dist.get_world_size()
    n = len(all_lists)
    G = len(record.seq)
    lo, hi = _engine.shard_range(n, rank, world)
    if rank == 0:
        os.makedirs(output_dir, exist_ok=True)
    dist.barrier()
    eng = make_engine() if make_engine is not None else _engine.MinimizerEngine(record)
    try:
        local_len = np.asarray(eng.plan_lists(all_lists[lo:hi], first_idx=lo), dtype=np.int64)
        sizes = np.asarray([header_len(lo + i) for i in range(hi - lo)], dtype=np.int64) + local_len + 1
        rel = np.zeros(hi - lo + 1, dtype=np.int64)
        rel[1:] = np.cumsum(sizes)

        def sink(sa: int, sb: int, view: np.ndarray) -> None:
            base = int(rel[sa])
            for s in range(sa, sb):
                fname = filename_template.format(model=model_name, idx=lo + s)
                with open(os.path.join(output_dir, fname), "wb") as fh:
                    fh.write(view[int(rel[s]) - base:int(rel[s + 1]) - base])

        eng.drain(sink)
        lengths = np.concatenate(all_gather_lengths(local_len)) if n else np.zeros(0, dtype=np.int64)
        dist.barrier()
    finally:
        if make_engine is None:
            eng.close()
    if rank == 0:
        print(f"Writing {n} individual FASTA files to: {output_dir}")
        for idx in range(n):
            print(f"[{idx+1}/{n}] genes present: {len(all_lists[idx])}")
            if _engine._sampled(idx):
                L = int(lengths[idx])
                fname = filename_template.format(model=model_name, idx=idx)
                print(f"  → saved {fname} | {L:,} bp ({_engine._pct(G, L):.1f}% reduction)")
    tot_red, tot_len = 0.0, 0
    for idx in range(n):
        tot_red += _engine._pct(G, int(lengths[idx]))
        tot_len += int(lengths[idx])
    return {"genome_count": n, "average_reduction_pct": tot_red / n, "average_length_bp": tot_len / n}
