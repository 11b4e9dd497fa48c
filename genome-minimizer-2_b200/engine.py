"""Host side above the C-ABI: gene table, name interning, plan/emit orchestration, drains.

What the reference does per sample in Python (minimizer_2.py:20-101) is done here once
per batch: the record is reduced to a gene table + bytes and uploaded (`gm2_set_reference`),
gene names are interned to ids (`gm2_set_name_map`), every sample's list becomes a row of
ids (`gm2_load_ids_host`), and the device produces lengths, record offsets and the FASTA
image (`gm2_plan`, `gm2_emit_*`).  All compute is in libgm2.so; there is no CPU path here.
"""
from __future__ import annotations

import errno
import mmap
import os
import stat
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from .genbank import read_genbank, sequence_bytes

DEFAULT_CHUNK_BYTES = int(os.environ.get("GM2_CHUNK_BYTES", 256 << 20))
FILE_RANGE_BYTES = int(os.environ.get("GM2_FILE_RANGE_BYTES", 1 << 30))    # progress granularity of drain_to_file
SEQ_ID_PREFIX = "Minimized_E_coli_K12_MG1655_"      # the reference's literal (minimizer_2.py:476, :537)


_COORD_MAX = (1 << 62)


class GeneTable:
    """`gene` features of a record in file order (minimizer_2.py:59-61, :78-79) + name interning."""

    def __init__(self, names: Sequence[str], starts: np.ndarray, ends: np.ndarray, features: Optional[list] = None):
        self.names = list(names)
        self.starts = np.asarray(starts, dtype=np.int64)
        self.ends = np.asarray(ends, dtype=np.int64)
        self.features = features            # the record's gene feature objects (for GenomeMinimiser.features)
        # intern: id = order of first appearance; CSR id -> gene indices (1:many for duplicate names)
        self.name_to_id: Dict[str, int] = {}
        buckets: List[List[int]] = []
        for g, nm in enumerate(self.names):
            i = self.name_to_id.get(nm)
            if i is None:
                i = len(buckets)
                self.name_to_id[nm] = i
                buckets.append([])
            buckets[i].append(g)
        self.id2gene_off = np.zeros(len(buckets) + 1, dtype=np.int32)
        if buckets:
            self.id2gene_off[1:] = np.cumsum([len(b) for b in buckets])
        self.id2gene_idx = np.asarray([g for b in buckets for g in b], dtype=np.int32)

    @property
    def F(self) -> int:
        return len(self.names)

    @property
    def V(self) -> int:
        return len(self.name_to_id)

    @classmethod
    def from_record(cls, record) -> "GeneTable":
        names, starts, ends, feats = [], [], [], []
        for feat in record.features:
            if feat.type == "gene":
                names.append(feat.qualifiers.get("gene", [""])[0])
                # AttributeError on a None location, as in the reference (minimizer_2.py:78)
                # (coordinates beyond the genome never meet a base, :94-96; gm2_set_reference clamps to
                # G, so anything past int64 is clamped here first)
                starts.append(min(int(feat.location.start), _COORD_MAX))
                ends.append(min(int(feat.location.end), _COORD_MAX))
                feats.append(feat)
        return cls(names, np.asarray(starts, dtype=np.int64), np.asarray(ends, dtype=np.int64), feats)

    def ids_for(self, needed) -> List[int]:
        """Ids of the table's names that satisfy Python's `name in needed` (minimizer_2.py:62)."""
        if isinstance(needed, np.ndarray):
            needed = needed.tolist()
        if isinstance(needed, (str, bytes)):
            # a bare string: `in` is a substring test — reproduce it rather than "fix" it
            if isinstance(needed, bytes):
                return []
            return [i for nm, i in self.name_to_id.items() if nm in needed]
        get = self.name_to_id.get
        out = []
        for x in needed:
            try:
                i = get(x)
            except TypeError:              # unhashable element: never equal to a str
                continue
            if i is not None:
                out.append(i)
        return out

    def tokenize(self, all_lists: Iterable) -> Tuple[np.ndarray, np.ndarray]:
        """Gene-name lists -> CSR (ids int32, off int64).  Names that match no gene are dropped
        here (they cannot influence the result); duplicates are kept (idempotent on the device)."""
        rows = [self.ids_for(needed) for needed in all_lists]
        off = np.zeros(len(rows) + 1, dtype=np.int64)
        if rows:
            off[1:] = np.cumsum([len(r) for r in rows])
        ids = np.fromiter((i for r in rows for i in r), dtype=np.int32, count=int(off[-1]))
        return ids, off

    def vocabulary(self) -> List[str]:
        """Distinct gene names, position = id (the id space of `tokenize` and gm2_set_name_map)."""
        out = [""] * len(self.name_to_id)
        for nm, i in self.name_to_id.items():
            out[i] = nm
        return out

    def keep_rows_from_bool(self, keep: np.ndarray) -> np.ndarray:
        """Boolean [S, F] -> packed little-endian uint32 rows [S, ceil(F/32)]."""
        keep = np.atleast_2d(np.asarray(keep, dtype=bool))
        S, F = keep.shape
        fw = (F + 31) // 32
        padded = np.zeros((S, fw * 32), dtype=np.uint8)
        padded[:, :F] = keep
        return np.packbits(padded, axis=1, bitorder="little").view("<u4").reshape(S, fw)


class ReferenceGenome:
    """What the batch entry functions need of the GenBank record: `record.seq` as bytes and the gene
    table (minimizer_2.py:455, :35, :59-61, :78-79).  `from_file` reads the file with the native scanner
    (gm2_genbank_parse, csrc/host_genbank.hpp) and, for anything outside that scanner's plain subset,
    with `genbank.read_genbank` — same table either way, and the general reader raises the errors."""

    def __init__(self, seq: np.ndarray, table: GeneTable, native: bool = False):
        self.seq = np.ascontiguousarray(seq, dtype=np.uint8)       # len(ref.seq) == len(record.seq)
        self.table = table
        self.native = native                                       # which reader produced it (tests, timing)

    @classmethod
    def from_record(cls, record) -> "ReferenceGenome":
        return cls(sequence_bytes(record), GeneTable.from_record(record))

    @classmethod
    def from_file(cls, genome_path: str) -> "ReferenceGenome":
        path = os.fspath(genome_path)
        got = _native.scan_genbank(path)
        if got is None:
            return cls.from_record(read_genbank(path))
        seq, names, starts, ends, _ = got
        return cls(seq, GeneTable(names, starts, ends), native=True)


class _Sized:
    """Stands in for one gene list where only its len() is still needed (the progress lines)."""
    __slots__ = ("_n",)

    def __init__(self, n: int):
        self._n = n

    def __len__(self) -> int:
        return self._n


class TokenizedLists:
    """The gene-name lists of a file as an id CSR: what `GeneTable.tokenize` makes of
    `np.load(genes_path, allow_pickle=True).tolist()` (minimizer_2.py:456, :518), produced straight from
    the file's pickle stream by gm2_tokenize_pickle.  Ids are those of a GeneTable built from the same
    record.  Behaves like the list of lists where the entry functions use it: len(), [a:b], len(x[i])."""

    def __init__(self, ids: np.ndarray, off: np.ndarray, counts: np.ndarray):
        self.ids = np.ascontiguousarray(ids, dtype=np.int32)
        self.off = np.ascontiguousarray(off, dtype=np.int64)
        self.counts = np.ascontiguousarray(counts, dtype=np.int64)

    def __len__(self) -> int:
        return int(self.counts.size)

    def __getitem__(self, i):
        if isinstance(i, slice):
            lo, hi, step = i.indices(len(self))
            if step != 1:
                raise ValueError("TokenizedLists supports contiguous slices only")
            hi = max(hi, lo)
            a, b = int(self.off[lo]), int(self.off[hi])
            return TokenizedLists(self.ids[a:b], self.off[lo:hi + 1] - a, self.counts[lo:hi])
        return _Sized(int(self.counts[i]))


def load_gene_lists(genes_path: str, table: GeneTable):
    """`np.load(genes_path, allow_pickle=True).tolist()` (minimizer_2.py:456, :518), tokenised natively
    when the file is an object array of lists of str; any other content goes through NumPy itself and
    comes back as the plain Python object (same errors, same `in` semantics downstream)."""
    tok = _native.tokenize_npy(os.fspath(genes_path), table.vocabulary())
    if tok is not None:
        return TokenizedLists(*tok)
    return np.load(genes_path, allow_pickle=True).tolist()


def default_device() -> int:
    for key in ("GM2_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(key)
        if v is not None and v != "":
            return int(v)
    return 0


class MinimizerEngine:
    """One reference genome resident on one GPU; plans and emits batches of samples."""

    def __init__(self, record=None, *, seq: Optional[np.ndarray] = None, table: Optional[GeneTable] = None,
                 device: Optional[int] = None, config: Optional[Dict[int, int]] = None):
        if isinstance(record, ReferenceGenome):
            seq, table = record.seq, record.table
        elif record is not None:
            seq = sequence_bytes(record)
            table = GeneTable.from_record(record)
        if seq is None or table is None:
            raise ValueError("MinimizerEngine needs a record, or seq + table")
        self.seq = np.ascontiguousarray(seq, dtype=np.uint8)
        self.table = table
        self.G = int(self.seq.size)
        self.device = default_device() if device is None else int(device)
        self.ctx = _native.Context(self.device)            # raises if CUDA is unusable: no fallback
        for k, v in (config or {}).items():
            self.ctx.configure(k, v)
        self.ctx.set_reference(self.seq, table.starts, table.ends)
        self.ctx.set_name_map(table.id2gene_off, table.id2gene_idx)
        self._map_owner: object = table                    # whose id space the device name map holds (see _use_table_map)
        self.first_idx = 0
        self._pinned: List[Optional[_native.PinnedBuffer]] = [None, None]

    def close(self):
        for b in self._pinned:
            if b is not None:
                b.free()
        self._pinned = [None, None]
        self.ctx.close()

    # -- planning ---------------------------------------------------------------------------
    def plan_lists(self, all_lists: Iterable, first_idx: int = 0) -> np.ndarray:
        if isinstance(all_lists, TokenizedLists):
            return self.plan_ids(all_lists.ids, all_lists.off, first_idx)
        ids, off = self.table.tokenize(all_lists)
        return self.plan_ids(ids, off, first_idx)

    def _use_table_map(self) -> None:
        """Ids handed to plan_ids / plan_lists are GeneTable ids.  plan_from_probabilities installs a
        ColumnSpace's map (column ids) and forced bitmaps on the same context; put the table's own map back
        before such ids are read (gm2_set_name_map also drops the forced bitmaps)."""
        if self._map_owner is not self.table:
            self.ctx.set_name_map(self.table.id2gene_off, self.table.id2gene_idx)
            self._map_owner = self.table

    def plan_ids(self, ids: np.ndarray, off: np.ndarray, first_idx: int = 0) -> np.ndarray:
        self._use_table_map()
        self.ctx.load_ids_host(ids, off)
        self.ctx.plan(first_idx)
        self.first_idx = first_idx
        return self.ctx.lengths()

    def plan_keep_rows(self, rows: np.ndarray, first_idx: int = 0) -> np.ndarray:
        self.ctx.load_keep_host(rows)
        self.ctx.plan(first_idx)
        self.first_idx = first_idx
        return self.ctx.lengths()

    @property
    def S(self) -> int:
        return self.ctx.S

    # -- results ------------------------------------------------------------------------------
    def image(self, s0: int = 0, s1: Optional[int] = None) -> np.ndarray:
        """Concatenated FASTA records [s0,s1) as one uint8 array (host)."""
        s1 = self.S if s1 is None else s1
        n = self.ctx.image_bytes(s0, s1)
        out = np.empty(n, dtype=np.uint8)
        self.ctx.emit_host(s0, s1, out)
        return out

    def emit_into(self, s0: int, s1: int, out: np.ndarray) -> None:
        """Records [s0,s1) into caller-owned host memory (pageable, pinned or a file mapping)."""
        self.ctx.emit_host(s0, s1, out)

    def sequence(self, s: int) -> str:
        """Minimized sequence of sample s as str (what GenomeMinimiser.reduced_genome_str holds)."""
        off = self.ctx.record_offsets()
        L = int(self.ctx.lengths()[s])
        img = self.image(s, s + 1)
        hdr = int(off[s + 1] - off[s]) - L - 1
        return img[hdr:hdr + L].tobytes().decode("ascii")

    def minimize_one(self, needed, idx: int = 0) -> Tuple[np.ndarray, str]:
        """One sample: (indices of the REMOVED genes in file order, minimized sequence)."""
        self.plan_lists([needed], first_idx=idx)
        row = self.ctx.keep_rows()[0]
        kept = np.unpackbits(row.view(np.uint8), bitorder="little")[:self.table.F].astype(bool)
        return np.flatnonzero(~kept), self.sequence(0)

    def duplicate_stats(self) -> dict:
        """The reference's `check_sequence_duplicates` (minimizer_2.py:273-303) over the planned samples,
        from device-side sequence hashes: sequences are grouped by (length, 64-bit hash) instead of by
        the strings themselves; `duplicates_detail` maps a representative id to the ids of its group
        (the reference keys it by the sequence string, which never leaves the GPU here)."""
        n = self.S
        lengths = self.ctx.lengths()
        hashes = self.ctx.sequence_hashes(0, n)
        groups: Dict[Tuple[int, int], List[str]] = {}
        for s in range(n):
            groups.setdefault((int(lengths[s]), int(hashes[s])), []).append(f"{SEQ_ID_PREFIX}{self.first_idx + s + 1}")
        dup = {ids[0]: ids for ids in groups.values() if len(ids) > 1}
        return {
            "total_sequences": n,
            "unique_sequences": len(groups),
            "duplicate_groups": len(dup),
            "duplicated_sequences": sum(len(v) for v in dup.values()),
            "unique_only_sequences": sum(1 for v in groups.values() if len(v) == 1),
            "duplicates_detail": dup,
            "compression_ratio": len(groups) / n if n else 0,
        }

    def _pin(self, i: int, nbytes: int) -> _native.PinnedBuffer:
        b = self._pinned[i]
        if b is None or b.nbytes < nbytes:
            if b is not None:
                b.free()
            b = _native.PinnedBuffer(nbytes)
            self._pinned[i] = b
        return b

    def chunks(self, max_bytes: int = DEFAULT_CHUNK_BYTES, s0: int = 0, s1: Optional[int] = None
               ) -> List[Tuple[int, int]]:
        """Split [s0,s1) into contiguous sample ranges of at most max_bytes of image each
        (a single record larger than max_bytes gets a range of its own)."""
        s1 = self.S if s1 is None else s1
        off = self.ctx.record_offsets()
        out = []
        a = s0
        while a < s1:
            b = int(np.searchsorted(off, off[a] + max_bytes, side="right")) - 1
            b = min(max(b, a + 1), s1)
            out.append((a, b))
            a = b
        return out

    def drain(self, sink: Callable[[int, int, np.ndarray], None], max_bytes: int = 0,
              s0: int = 0, s1: Optional[int] = None) -> int:
        """Produce records [s0,s1) chunk by chunk into two pinned buffers and hand each chunk to
        `sink(sa, sb, bytes_view)`; the GPU fills one buffer while the sink consumes the other.
        Returns the total number of bytes delivered."""
        ranges = self.chunks(max_bytes or DEFAULT_CHUNK_BYTES, s0, s1)
        if not ranges:
            return 0
        off = self.ctx.record_offsets()
        need = max(int(off[b] - off[a]) for a, b in ranges)
        bufs = [self._pin(0, need), self._pin(1, need)]
        total = 0
        err: List[BaseException] = []

        def produce(i: int):
            try:
                a, b = ranges[i]
                self.ctx.emit_host(a, b, bufs[i & 1])
            except BaseException as e:  # noqa: BLE001 - re-raised on the caller's thread
                err.append(e)

        produce(0)
        for i, (a, b) in enumerate(ranges):
            if err:
                raise err[0]
            t = None
            if i + 1 < len(ranges):
                t = threading.Thread(target=produce, args=(i + 1,))
                t.start()                                   # ctypes releases the GIL during the call
            n = int(off[b] - off[a])
            sink(a, b, bufs[i & 1].array[:n])
            total += n
            if t is not None:
                t.join()
        if err:
            raise err[0]
        return total


class ContextPair:
    """Two contexts over one reference on one GPU, for device-resident pipelines.

    A context is ONE plan slot (its plan state is overwritten by its next plan), so a job that is processed
    chunk by chunk with the output staying on the GPU alternates two of them: chunk i goes to context i & 1,
    `gm2_order_after` keeps the emits in file order on the device, and the plan of chunk i+1 (K1-K3:
    issue/latency-bound) runs on the other context's stream UNDER the emit of chunk i (write-bound) instead of
    between two emits.  Nothing blocks on the host.  (The host paths — gm2_emit_host, drain — leave the GPU
    idle most of the time anyway and use one context.)"""

    def __init__(self, seq: np.ndarray, table: GeneTable, device: Optional[int] = None,
                 config: Optional[Dict[int, int]] = None, streams: Optional[Sequence[int]] = None):
        dev = default_device() if device is None else int(device)
        self.table = table
        self.ctx = [_native.Context(dev), _native.Context(dev)]
        for i, c in enumerate(self.ctx):
            for k, v in (config or {}).items():
                c.configure(k, v)
            if streams is not None:
                c.set_stream(streams[i])
            c.set_reference(np.ascontiguousarray(seq, dtype=np.uint8), table.starts, table.ends)
            c.set_name_map(table.id2gene_off, table.id2gene_idx)

    def close(self):
        for c in self.ctx:
            c.close()

    def run_chunks(self, chunks: Iterable[Tuple[int, int, int, int, int]], ring: Sequence[Tuple[int, int]],
                   after_emit: Optional[Callable[[int, "_native.Context"], None]] = None) -> int:
        """chunks: (ids_dev_ptr, off_dev_ptr, samples, n_ids, first_idx) of device-resident CSR id lists, in file
        order; ring: two (dev_ptr, capacity) output buffers, chunk i is emitted into ring[i & 1].
        `after_emit(i, ctx)` may issue the chunk's consumer on ctx's stream (hashing, a copy, ...): ring[i & 1] is
        not overwritten before that work has finished.  Returns the number of chunks issued; call sync() on both
        contexts (or order a stream after them) before reading results on the host.  Nothing here waits for a
        plan, so the capacity of a ring buffer cannot be checked against the chunk's image size (gm2_emit_dev
        checks it only when the plan is already on the host): size the ring from a planning pass, as
        bench.py's sharded leg does, or for the worst case (every base of every sample kept)."""
        a, b = self.ctx
        b.order_after(a)
        n = 0
        for i, (ids_ptr, off_ptr, S, n_ids, first_idx) in enumerate(chunks):
            c, o = self.ctx[i & 1], self.ctx[(i + 1) & 1]
            c.load_ids_dev(ids_ptr, off_ptr, S, n_ids)
            c.plan_async(first_idx)
            c.order_after(o)
            c.emit_dev(0, S, ring[i & 1][0], ring[i & 1][1])
            if after_emit is not None:
                after_emit(i, c)
            n += 1
        a.order_after(b)
        return n


def drain_to_file(eng, target, file_pos: int, s0: int = 0, s1: Optional[int] = None,
                  progress: Optional[Callable[[int, int], None]] = None) -> int:
    """Records [s0,s1) into `target` (a path, or an open read-write descriptor that stays open) starting at
    byte `file_pos`; returns the position after the last byte.

    A regular file is grown to its final size and MAPPED, and `gm2_emit_host` produces straight into the
    mapping: the expansion workers (two-bit transport) or the staging copy (image-bytes transport) write
    the page-cache pages themselves, in parallel, with no intermediate buffer and no write() call —
    into tmpfs that is 10x a single-threaded write().  Space is checked before anything is produced
    (a full filesystem under a mapping is a SIGBUS, not an OSError).  Anything that is not a regular
    file (/dev/null, a FIFO: written sequentially), a filesystem without shared mappings, or
    GM2_FILE_SINK=write takes the portable form: pinned ping-pong buffers + (p)write until every byte is down.
    `progress(sa, sb)` is called after each range of about GM2_FILE_RANGE_BYTES, in order."""
    s1 = eng.S if s1 is None else s1
    off = eng.ctx.record_offsets()
    total = int(off[s1] - off[s0])
    end_pos = file_pos + total
    own_fd = not isinstance(target, int)
    fd = os.open(target, os.O_RDWR) if own_fd else target
    path = target if own_fd else f"<fd {fd}>"
    try:
        mm = None
        regular = stat.S_ISREG(os.fstat(fd).st_mode)
        if total > 0 and regular and os.environ.get("GM2_FILE_SINK", "map") != "write":
            try:
                if os.fstat(fd).st_size < end_pos:
                    os.ftruncate(fd, end_pos)
                vfs = os.fstatvfs(fd)
                have = os.fstat(fd).st_blocks * 512
                if vfs.f_bavail * vfs.f_frsize + have < end_pos:
                    raise OSError(errno.ENOSPC, f"{path}: {end_pos:,} bytes needed, "
                                                f"{vfs.f_bavail * vfs.f_frsize:,} free on the filesystem")
                mm = mmap.mmap(fd, end_pos, access=mmap.ACCESS_WRITE)
            except OSError as e:
                if e.errno == errno.ENOSPC:
                    raise
                mm = None                                  # no shared mappings here: portable form below
        if mm is not None:
            view = np.frombuffer(mm, dtype=np.uint8)
            try:
                for a, b in eng.chunks(FILE_RANGE_BYTES, s0, s1):
                    lo = file_pos + int(off[a] - off[s0])
                    eng.emit_into(a, b, view[lo:lo + int(off[b] - off[a])])
                    if progress is not None:
                        progress(a, b)
            finally:
                del view
                try:
                    mm.close()
                except BufferError:                        # an exception in flight still holds a slice
                    pass
            return end_pos
        pos = [file_pos]

        def sink(sa: int, sb: int, chunk: np.ndarray) -> None:
            pos[0] = pwrite_all(fd, chunk, pos[0], positioned=regular)
            if progress is not None:
                progress(sa, sb)

        eng.drain(sink, s0=s0, s1=s1)
        if pos[0] != end_pos:
            raise RuntimeError(f"wrote up to byte {pos[0]}, expected {end_pos}")
        return end_pos
    finally:
        if own_fd:
            os.close(fd)


def pwrite_all(fd: int, view: np.ndarray, pos: int, positioned: bool = True) -> int:
    """os.pwrite (os.write for pipes and devices) until every byte is down: a single write may be short
    (ENOSPC part-way, a signal, or the kernel's 0x7ffff000 per-call cap).  Returns the position after the
    last byte."""
    mv = memoryview(view).cast("B")
    while len(mv):
        k = os.pwrite(fd, mv, pos) if positioned else os.write(fd, mv)
        if k <= 0:
            raise OSError(f"pwrite wrote {k} bytes at offset {pos} ({len(mv)} left)")
        pos += k
        mv = mv[k:]
    return pos


_FILE_WRITERS = max(1, min(8, (os.cpu_count() or 1)))


def write_record_files(output_dir, names: Sequence[str], view: np.ndarray, rel_off: np.ndarray) -> None:
    """One file per record of a drained chunk (`view`, record k at rel_off[k]:rel_off[k+1]), written by a few
    threads at once: file creation and page-cache fills of ~MB files are per-file latency, not bandwidth
    (write() releases the GIL).  Returns when every file is closed; the first error is re-raised."""
    def one(k: int) -> None:
        with open(os.path.join(output_dir, names[k]), "wb") as fh:
            fh.write(view[int(rel_off[k]):int(rel_off[k + 1])])

    if len(names) < 4 or _FILE_WRITERS == 1:
        for k in range(len(names)):
            one(k)
        return
    with ThreadPoolExecutor(max_workers=_FILE_WRITERS) as pool:
        list(pool.map(one, range(len(names))))


def shard_range(S: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous sample range of `rank` (rank order == file order)."""
    return (rank * S) // world, ((rank + 1) * S) // world


def shard_range_by_bytes(rec_sizes: np.ndarray, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous sample range of `rank` when the file is cut at equal cumulative OUTPUT BYTES instead of
    equal sample counts (skewed retention): rank r starts at the first record whose start
    offset is at or beyond r/world of the image.  Ranges tile [0, S) in rank order; every rank's byte
    count is within one record of the mean."""
    sizes = np.asarray(rec_sizes, dtype=np.int64)
    S = int(sizes.size)
    off = np.zeros(S + 1, dtype=np.int64)
    off[1:] = np.cumsum(sizes)
    total = int(off[-1])

    def cut(r: int) -> int:
        if r <= 0:
            return 0
        if r >= world:
            return S
        return min(int(np.searchsorted(off, (total * r + world - 1) // world, side="left")), S)

    return cut(rank), cut(rank + 1)


# ----------------------------------------------------------------------------------------------
# batch entry points behind minimizer_2.process_multiple_genomes_* (reference :447-560)
# ----------------------------------------------------------------------------------------------
def _pct(original_length: int, genome_length: int) -> float:
    return (original_length - genome_length) / original_length * 100.0


def _sampled(idx: int) -> bool:
    """The reference reports (and, in single-file mode, accumulates) only these samples (:482, :550)."""
    return idx <= 9 or (idx + 1) % 100 == 0


def run_single_file(record, all_lists, model_name: str, output_file: str, engine: Optional[MinimizerEngine] = None) -> dict:
    n = len(all_lists)
    G = len(record.seq)
    eng = engine or MinimizerEngine(record)
    try:
        lengths = eng.plan_lists(all_lists, first_idx=0)
        pre = (f"# Minimized genomes generated using model: {model_name}\n"
               f"# Total genomes: {n}\n"
               f"# Generated on: {np.datetime64('now')}\n").encode()
        fd = os.open(output_file, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o666)

        def progress(sa: int, sb: int) -> None:
            for idx in range(sa, sb):                          # same lines, same order as the reference
                print(f"[{idx+1}/{n}] genes present: {len(all_lists[idx])}")
                if _sampled(idx):
                    L = int(lengths[idx])
                    print(f"  → {L:,} bp ({_pct(G, L):.1f}% reduction)")

        try:
            pwrite_all(fd, np.frombuffer(pre, dtype=np.uint8), 0, positioned=False)
            drain_to_file(eng, fd, len(pre), progress=progress)
        finally:
            os.close(fd)
    finally:
        if engine is None:
            eng.close()
    # F10: only the reported samples are summed, the divisor is still n (float64, same order)
    tot_red, tot_len = 0.0, 0
    for idx in range(n):
        if _sampled(idx):
            tot_red += _pct(G, int(lengths[idx]))
            tot_len += int(lengths[idx])
    return {"genome_count": n, "average_reduction_pct": tot_red / n, "average_length_bp": tot_len / n}


def run_multi_file(record, all_lists, model_name: str, output_dir, filename_template: str,
                   engine: Optional[MinimizerEngine] = None) -> dict:
    n = len(all_lists)
    G = len(record.seq)
    print(f"Writing {n} individual FASTA files to: {output_dir}")
    eng = engine or MinimizerEngine(record)
    try:
        lengths = eng.plan_lists(all_lists, first_idx=0)
        rec_off = eng.ctx.record_offsets()

        def sink(sa: int, sb: int, view: np.ndarray) -> None:
            base = int(rec_off[sa])
            names = [filename_template.format(model=model_name, idx=idx) for idx in range(sa, sb)]
            write_record_files(output_dir, names, view, np.asarray(rec_off[sa:sb + 1]) - base)
            for idx, fname in zip(range(sa, sb), names):       # the reference's lines, in its order
                print(f"[{idx+1}/{n}] genes present: {len(all_lists[idx])}")
                if _sampled(idx):
                    L = int(lengths[idx])
                    print(f"  → saved {fname} | {L:,} bp ({_pct(G, L):.1f}% reduction)")

        eng.drain(sink)
    finally:
        if engine is None:
            eng.close()
    tot_red, tot_len = 0.0, 0
    for idx in range(n):
        tot_red += _pct(G, int(lengths[idx]))
        tot_len += int(lengths[idx])
    return {"genome_count": n, "average_reduction_pct": tot_red / n, "average_length_bp": tot_len / n}


# ----------------------------------------------------------------------------------------------
# VAE decoder output -> keep masks on the device (BASELINE config 5)
# ----------------------------------------------------------------------------------------------
class ColumnSpace:
    """The mask columns of the reference's sample -> convert-samples -> minimizer chain, as name ids.

    Reference semantics collapsed here (explore_data/binary_converter.py):
      * duplicate column names: first occurrence kept, the rest dropped, and mask rows must then have
        the de-duplicated length (:29-36, :50-53);
      * a column is present iff its mask value >= 0.5 (:55), on the 0/1 matrix that `> 0.5` produced
        (utils/extras.py:200-201);
      * check_essential_genes adds every essential name that is missing (:91-98).
    """

    def __init__(self, table: GeneTable, col_names: Sequence[str], essential: Iterable[str] = ()):
        seen: Dict[str, int] = {}
        keep_first = []
        for i, nm in enumerate(col_names):
            nm = str(nm)
            if nm not in seen:
                seen[nm] = len(keep_first)
                keep_first.append(nm)
        self.cols: List[str] = keep_first
        self.col_of: Dict[str, int] = seen
        self.V = len(keep_first)
        self.duplicates_dropped = len(col_names) - self.V
        genes_of: Dict[str, List[int]] = {}
        for g, nm in enumerate(table.names):
            genes_of.setdefault(nm, []).append(g)
        rows = [genes_of.get(nm, []) for nm in self.cols]
        self.id2gene_off = np.zeros(self.V + 1, dtype=np.int32)
        if rows:
            self.id2gene_off[1:] = np.cumsum([len(r) for r in rows])
        self.id2gene_idx = np.asarray([g for r in rows for g in r], dtype=np.int32)
        ess = {str(e) for e in essential}
        fw = (table.F + 31) // 32
        fk = np.zeros(fw * 32, dtype=np.uint8)
        for g, nm in enumerate(table.names):
            if nm in ess:
                fk[g] = 1
        self.force_keep = np.packbits(fk, bitorder="little").view("<u4") if fw else np.zeros(0, dtype=np.uint32)
        vw = (self.V + 31) // 32
        fi = np.zeros(vw * 32, dtype=np.uint8)
        for nm in ess:
            c = seen.get(nm)
            if c is not None:
                fi[c] = 1
        self.forced_ids = np.packbits(fi, bitorder="little").view("<u4") if vw else np.zeros(0, dtype=np.uint32)
        self.essentials_not_in_columns = sum(1 for nm in ess if nm not in seen)


def _device_matrix(probs) -> Tuple[int, int, int, int]:
    """(pointer, S, V, row stride in elements) of a float32 CUDA matrix (torch tensor or anything
    exposing data_ptr / shape / stride / dtype the same way)."""
    if len(probs.shape) != 2:
        raise ValueError("probabilities must be a 2-D [samples, columns] matrix")
    if "float32" not in str(probs.dtype):
        raise ValueError("probabilities must be float32")
    if hasattr(probs, "is_cuda") and not probs.is_cuda:
        raise ValueError("probabilities must live on the GPU (there is no CPU path)")
    st = probs.stride()
    if st[1] != 1:
        raise ValueError("probabilities must be row-major with unit column stride")
    return int(probs.data_ptr()), int(probs.shape[0]), int(probs.shape[1]), int(st[0])


def plan_from_probabilities(eng: MinimizerEngine, space: ColumnSpace, probs, threshold: float = 0.5,
                            first_idx: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Device-resident decoder output -> plan.  Returns (lengths, list lengths the reference would print)."""
    ptr, S, V, ld = _device_matrix(probs)
    if V != space.V:
        # the reference raises here too (binary_converter.py:50-53)
        raise ValueError(f"Mask row has length {V}, but dataset has {space.V} gene columns.")
    if eng._map_owner is not space:
        eng._map_owner = None                              # in between, the device map belongs to nobody
        eng.ctx.set_name_map(space.id2gene_off, space.id2gene_idx)
        eng.ctx.set_forced(space.force_keep, space.forced_ids)
        eng._map_owner = space
    eng.ctx.load_probs_dev(ptr, S, ld, threshold)
    eng.ctx.plan(first_idx)
    eng.first_idx = first_idx
    return eng.ctx.lengths(), eng.ctx.counts() + space.essentials_not_in_columns


def run_single_file_from_probabilities(record, probs, col_names: Sequence[str], essential: Iterable[str],
                                       model_name: str, output_file: str, threshold: float = 0.5,
                                       engine: Optional[MinimizerEngine] = None) -> dict:
    """The reference's three-step chain (`--mode sample` -> `--mode convert-samples` -> `--mode minimizer
    --single-file`) for samples that are still on the GPU: same FASTA file, same progress lines, same
    return dict as process_multiple_genomes_single_file on the `_with_essentials.npy` lists."""
    G = len(record.seq)
    eng = engine or MinimizerEngine(record)
    try:
        space = ColumnSpace(eng.table, col_names, essential)
        lengths, counts = plan_from_probabilities(eng, space, probs, threshold)
        n = len(lengths)
        os.makedirs(os.path.dirname(output_file) or ".", exist_ok=True)
        pre = (f"# Minimized genomes generated using model: {model_name}\n"
               f"# Total genomes: {n}\n"
               f"# Generated on: {np.datetime64('now')}\n").encode()
        fd = os.open(output_file, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o666)

        def progress(sa: int, sb: int) -> None:
            for idx in range(sa, sb):
                print(f"[{idx+1}/{n}] genes present: {int(counts[idx])}")
                if _sampled(idx):
                    L = int(lengths[idx])
                    print(f"  → {L:,} bp ({_pct(G, L):.1f}% reduction)")

        try:
            pwrite_all(fd, np.frombuffer(pre, dtype=np.uint8), 0, positioned=False)
            drain_to_file(eng, fd, len(pre), progress=progress)
        finally:
            os.close(fd)
    finally:
        if engine is None:
            eng.close()
    tot_red, tot_len = 0.0, 0
    for idx in range(n):
        if _sampled(idx):
            tot_red += _pct(G, int(lengths[idx]))
            tot_len += int(lengths[idx])
    return {"genome_count": n, "average_reduction_pct": tot_red / n, "average_length_bp": tot_len / n}


def load_masks_npy(masks_npy_path: str) -> np.ndarray:
    """The masks container `--mode convert-samples` reads (explore_data/binary_converter.py:38-46):
    a 2-D float array [N, P], a 1-D object array of rows, or a single 1-D row."""
    masks = np.load(masks_npy_path, allow_pickle=True)
    if masks.ndim == 1:
        if len(masks) and isinstance(masks[0], (list, np.ndarray)):
            masks = np.array([np.asarray(row) for row in masks], dtype=object)
        else:
            masks = masks[None, :]
    return masks


def run_single_file_from_masks(record, masks_npy_path: str, col_names: Sequence[str], essential: Iterable[str],
                               model_name: str, output_file: str, threshold: float = 0.5,
                               engine: Optional[MinimizerEngine] = None) -> dict:
    """`--mode convert-samples` + `--mode minimizer --single-file` from the masks FILE that `--mode sample`
    writes (main.py:434): rows are thresholded exactly as the reference does (`np.asarray(row, float) >=
    threshold`, binary_converter.py:49-55 — in float64, on the host), uploaded as 0/1 float32 and handed
    to the dense keep-mask builder.  Same file, progress lines and return dict as the reference's chain."""
    masks = load_masks_npy(masks_npy_path)
    eng = engine or MinimizerEngine(record)
    try:
        space = ColumnSpace(eng.table, col_names, essential)
        rows = []
        for i, row in enumerate(masks):
            r = np.asarray(row, dtype=float)
            if r.size != space.V:
                raise ValueError(f"Mask row {i} has length {r.size}, but dataset has {space.V} gene columns.")
            rows.append(r >= threshold)
        present = (np.stack(rows, axis=0) if rows else np.zeros((0, space.V), dtype=bool)).astype(np.float32)
        dev = _DeviceMatrix(eng.ctx, present)
        try:
            return run_single_file_from_probabilities(record, dev, col_names, essential, model_name, output_file,
                                                      threshold=0.5, engine=eng)
        finally:
            dev.free()
    finally:
        if engine is None:
            eng.close()


class _DeviceMatrix:
    """A float32 [S, V] matrix uploaded to the engine's GPU with the CUDA runtime through ctypes, exposing
    the torch-like surface `_device_matrix` reads (torch is plumbing for callers, not a dependency here)."""

    def __init__(self, ctx: "_native.Context", host: np.ndarray):
        host = np.ascontiguousarray(host, dtype=np.float32)
        self.shape = host.shape
        self.dtype = "float32"
        self.is_cuda = True
        self._ctx = ctx
        self._buf = ctx.device_alloc(max(host.nbytes, 4))
        ctx.upload(self._buf, host)

    def data_ptr(self) -> int:
        return self._buf

    def stride(self):
        return (self.shape[1], 1)

    def free(self):
        if self._buf:
            self._ctx.device_free(self._buf)
            self._buf = 0
