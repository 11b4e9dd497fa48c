"""Multi-GPU form of the batch entry points: one process per GPU (torchrun), samples sharded.

Samples are independent (minimizer_2.py:469-477 reads only the shared record and the sample's own
list), so every rank takes a CONTIGUOUS range of samples: rank order is file order and record ids
stay global (`first_idx`).  The reference genome is replicated.  The only exchange is an all-gather
of the per-sample lengths (8 bytes per sample) from which every rank derives the global byte offset
of its shard; the FASTA bytes never cross NVLink — each rank `pwrite`s its own shard into the shared
output file.  Backend: NCCL on GPUs, gloo in the CPU tests.

Shards are cut by OUTPUT BYTES, not by sample count: every rank first plans the
count-based range [r*S//R, (r+1)*S//R) — lengths only, K1-K3, ~0.2 ms per 10,000 samples — the
lengths are all-gathered, and the ranges are then re-cut at equal cumulative record bytes, so a job
whose retention drifts along the file (config 4) still finishes on all GPUs at the same time.
`GM2_SHARD_BALANCE=count` keeps the count-based ranges.

A failure on one rank (bad input, full disk, CUDA error) is agreed on by all ranks before the next
collective, so no rank is left blocked in an all-gather or barrier: the failing rank re-raises its
own exception, the others raise RuntimeError naming the step.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Tuple

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


_RENDEZVOUS_ENV = ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")


def launched_ranks() -> int:
    """How many ranks share this job.  Inside an initialised process group: its size.  Otherwise > 1 only for a
    complete torchrun-style environment (RANK, WORLD_SIZE, MASTER_ADDR and MASTER_PORT all present) — a stray
    WORLD_SIZE from some other launcher does not switch the entry functions to the sharded path, they then do
    the whole job as the reference would.  GM2_SHARD=0 turns the automatic sharding off altogether."""
    if os.environ.get("GM2_SHARD", "1") == "0":
        return 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    if any(not os.environ.get(k) for k in _RENDEZVOUS_ENV):
        return 1
    try:
        return max(int(os.environ["WORLD_SIZE"]), 1)
    except ValueError:
        return 1


def _device_for_backend() -> torch.device:
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def agree(error: Optional[BaseException], step: str) -> None:
    """Collective: every rank reports whether `step` succeeded; if any rank failed, ALL ranks raise here
    (the failing ones their own exception), so nobody enters the next collective alone."""
    flag = torch.tensor([0 if error is None else 1], dtype=torch.int32, device=_device_for_backend())
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if error is not None:
        raise error
    if int(flag.item()):
        raise RuntimeError(f"genome-minimizer sharded run: another rank failed during '{step}'")


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


def record_sizes(lengths: np.ndarray, first_idx: int = 0) -> np.ndarray:
    """Bytes of every FASTA record: header + bases + '\\n' (minimizer_2.py:476-477), vectorised over samples."""
    n = int(lengths.size)
    idx1 = np.arange(first_idx + 1, first_idx + n + 1, dtype=np.int64)
    digits = np.ones(n, dtype=np.int64)
    p = 10
    while n and p <= int(idx1[-1]):
        digits += idx1 >= p
        p *= 10
    return 1 + len(_engine.SEQ_ID_PREFIX) + digits + 1 + np.asarray(lengths, dtype=np.int64) + 1


def balance_mode() -> str:
    return "count" if os.environ.get("GM2_SHARD_BALANCE", "bytes") == "count" else "bytes"


pwrite_all = _engine.pwrite_all


def _planned_lengths(eng, all_lists: Sequence, n: int, rank: int, world: int) -> Tuple[np.ndarray, int, int, bool]:
    """Plan the count-based range, all-gather the lengths, choose the final contiguous range of this rank.
    Returns (all n lengths, lo, hi, replan) — replan says the engine's current plan is for another range."""
    lo, hi = _engine.shard_range(n, rank, world)
    err, local_len = None, np.zeros(0, dtype=np.int64)
    try:
        local_len = np.asarray(eng.plan_lists(all_lists[lo:hi], first_idx=lo), dtype=np.int64)
    except BaseException as e:  # noqa: BLE001 - agreed on below, then re-raised
        err = e
    agree(err, "plan")
    gathered = all_gather_lengths(local_len)
    lengths = np.concatenate(gathered) if gathered else np.zeros(0, dtype=np.int64)
    assert lengths.size == n
    if balance_mode() == "bytes" and world > 1:
        lo2, hi2 = _engine.shard_range_by_bytes(record_sizes(lengths), rank, world)
        return lengths, lo2, hi2, (lo2, hi2) != (lo, hi)
    return lengths, lo, hi, False


def run_single_file_sharded(record, all_lists: Sequence, model_name: str, output_file: str,
                            make_engine: Optional[Callable[[], object]] = None,
                            timestamp: Optional[str] = None, quiet: bool = False) -> dict:
    """`process_multiple_genomes_single_file` (reference :447-495) over all ranks of the default
    process group.  Every rank returns the same dict; rank 0 prints the progress lines.  All ranks must see
    `output_file` on one shared filesystem (rank 0 creates and sizes it, every rank writes its own part)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    n = len(all_lists)
    G = len(record.seq)
    eng = make_engine() if make_engine is not None else _engine.MinimizerEngine(record)
    try:
        lengths, lo, hi, replan = _planned_lengths(eng, all_lists, n, rank, world)
        pre = (f"# Minimized genomes generated using model: {model_name}\n"
               f"# Total genomes: {n}\n"
               f"# Generated on: {timestamp if timestamp is not None else np.datetime64('now')}\n").encode()
        rec_off = np.zeros(n + 1, dtype=np.int64)
        rec_off[1:] = np.cumsum(record_sizes(lengths))
        err = None
        try:
            if rank == 0:
                with open(output_file, "wb") as fh:
                    fh.write(pre)
                    fh.truncate(len(pre) + int(rec_off[-1]))
            if replan:
                eng.plan_lists(all_lists[lo:hi], first_idx=lo)
        except BaseException as e:  # noqa: BLE001
            err = e
        agree(err, "create output file")
        try:
            end = _engine.drain_to_file(eng, output_file, len(pre) + int(rec_off[lo]))      # mapped file, or pwrite until done
            if end != len(pre) + int(rec_off[hi]):
                raise RuntimeError(f"rank {rank} wrote up to byte {end}, expected {len(pre) + int(rec_off[hi])}")
        except BaseException as e:  # noqa: BLE001
            err = e
        agree(err, "write shard")                     # doubles as the closing barrier
    finally:
        if make_engine is None:
            eng.close()
    if rank == 0 and not quiet:
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
                           make_engine: Optional[Callable[[], object]] = None, quiet: bool = False) -> dict:
    """`process_multiple_genomes_multiple_files` (reference :499-560) over all ranks: every rank writes
    the files of its own contiguous shard (global idx in the names and record ids); the lengths are
    all-gathered so that every rank returns the reference's dict and rank 0 prints its lines."""
    rank, world = dist.get_rank(), dist.get_world_size()
    n = len(all_lists)
    G = len(record.seq)
    err = None
    try:
        if rank == 0:
            os.makedirs(output_dir, exist_ok=True)
    except BaseException as e:  # noqa: BLE001
        err = e
    agree(err, "create output directory")
    eng = make_engine() if make_engine is not None else _engine.MinimizerEngine(record)
    try:
        lengths, lo, hi, replan = _planned_lengths(eng, all_lists, n, rank, world)
        try:
            if replan:
                eng.plan_lists(all_lists[lo:hi], first_idx=lo)
            rel = np.zeros(hi - lo + 1, dtype=np.int64)
            rel[1:] = np.cumsum(record_sizes(lengths[lo:hi], first_idx=lo))

            def sink(sa: int, sb: int, view: np.ndarray) -> None:
                names = [filename_template.format(model=model_name, idx=lo + s) for s in range(sa, sb)]
                _engine.write_record_files(output_dir, names, view, rel[sa:sb + 1] - int(rel[sa]))

            eng.drain(sink)
        except BaseException as e:  # noqa: BLE001
            err = e
        agree(err, "write shard")
    finally:
        if make_engine is None:
            eng.close()
    if rank == 0 and not quiet:
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
