"""Drop-in facade for the reference module `src/genome_minimizer_2/minimizer/minimizer_2.py`.

The reference's CLI imports two functions from that module (main.py:554-557) and calls them
at main.py:582-587 / :599-604.  This module exposes the same names with the same arguments,
defaults, stdout lines, output files and return values; the per-sample Python loops of the
reference (minimizer_2.py:50-101) run as one batched job in libgm2.so on a B200.

  GenomeMinimiser                            reference :19-270 (constructor does all the work)
  process_multiple_genomes_single_file       reference :447-495
  process_multiple_genomes_multiple_files    reference :499-560

Preserved on purpose: records are never line-wrapped; the single-file
preamble's third line is `np.datetime64('now')`; single-file averages add up only samples
with idx<=9 or (idx+1)%100==0 yet divide by N; an empty `.npy` ends in ZeroDivisionError.
The reporting helpers (reference :273-444: `check_sequence_duplicates`, `print_duplicate_statistics`,
`generate_summary_file`) have no call sites in the reference; they are mirrored in `reporting.py`
and re-exported here so every public name of the module exists.
"""
from __future__ import annotations

import os
import sys
from typing import Optional

import numpy as np

from . import engine as _engine
from .genbank import read_genbank
from .reporting import (  # noqa: F401  (re-exported: reference :273-444)
    check_sequence_duplicates,
    generate_summary_file,
    print_duplicate_statistics,
)

# default output root, the analogue of utils/directories.py:10 in the reference tree
PROJECT_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SEQ_ID_PREFIX = _engine.SEQ_ID_PREFIX
_GENBANK_SUFFIXES = (".gb", ".genbank", ".gbff")


def _default_dir() -> str:
    return os.path.join(PROJECT_ROOT, "minimized_genomes")


# One resident engine per record object: the reference's loop builds a GenomeMinimiser per sample on a
# shared record (minimizer_2.py:469-475); re-uploading the genome for each would dominate.
_ENGINES: "dict[int, tuple]" = {}


def _record_signature(record, table: "_engine.GeneTable", seq: np.ndarray) -> tuple:
    """Content fingerprint of what the engine holds of a record: sequence CRC + the gene table.  A record
    edited in place (a feature moved or renamed, the sequence replaced by one of equal length) changes
    it, so a stale resident engine is never reused; ~10 ms for a K-12-sized record against the
    reference's ~0.8 s of per-construction work."""
    import zlib
    return (int(seq.size), zlib.crc32(seq), len(record.features), hash(tuple(table.names)),
            zlib.crc32(table.starts.tobytes()), zlib.crc32(table.ends.tobytes()))


def _engine_for(record) -> _engine.MinimizerEngine:
    import weakref
    key = id(record)
    seq = _engine.sequence_bytes(record)
    table = _engine.GeneTable.from_record(record)
    sig = _record_signature(record, table, seq)
    hit = _ENGINES.get(key)
    if hit is not None and hit[0]() is record and hit[2] == sig:
        hit[1].table.features = table.features         # the feature objects themselves may have been replaced
        return hit[1]
    if hit is not None:
        _ENGINES.pop(key, None)
        hit[1].close()
    eng = _engine.MinimizerEngine(_engine.ReferenceGenome(seq, table))

    def _drop(_ref, key=key):
        old = _ENGINES.pop(key, None)
        if old is not None:
            try:
                old[1].close()
            except Exception:
                pass

    try:
        _ENGINES[key] = (weakref.ref(record, _drop), eng, sig)
    except TypeError:                       # record type without weakref support: no caching
        _ENGINES.pop(key, None)
    return eng


class GenomeMinimiser:
    """One sample's minimization; attribute-compatible with the reference class.

    Attributes after construction: idx, model_name, record, wildtype_sequence,
    original_genome_length, needed_genes, features (removed gene features, file order),
    positions_to_remove (set, built on first access), reduced_genome_str.
    `engine=` is an extension: reuse a genome already resident on the GPU.
    """

    def __init__(self, record_path: str = None, needed_genes_path: str = None, idx: int = 0,
                 model_name: str = "", record=None, all_needed_gene_lists: list = None,
                 needed_genes_list: list = None, engine: Optional[_engine.MinimizerEngine] = None):
        self.idx, self.model_name = idx, model_name
        self.record = self.load_genome(record_path) if record is None else record
        self.wildtype_sequence = self.record
        self.original_genome_length = len(self.record.seq)
        # same precedence as reference :38-43
        if needed_genes_list is not None:
            self.needed_genes = needed_genes_list
        elif all_needed_gene_lists is not None:
            self.needed_genes = all_needed_gene_lists[idx]
        else:
            self.needed_genes = self.get_needed_genes(needed_genes_path)[idx]

        eng = engine or _engine_for(self.record)
        removed, self.reduced_genome_str = eng.minimize_one(self.needed_genes, idx)
        feats = eng.table.features
        if feats is None:
            # an engine built from a file by the native scanner holds no feature objects: take them from the record
            feats = [f for f in self.record.features if f.type == "gene"]
            if len(feats) != eng.table.F:
                raise ValueError("engine= was built from a different genome than this record")
        self.features = [feats[g] for g in removed]                                 # reference :50-66
        self._spans = [(int(eng.table.starts[g]), int(eng.table.ends[g])) for g in removed]
        self._positions: Optional[set] = None

    @property
    def positions_to_remove(self) -> set:
        """Reference :68-83.  About 2 M Python ints for a K-12 genome, so built lazily."""
        if self._positions is None:
            self._positions = set()
            for a, b in self._spans:
                self._positions.update(range(a, b))
        return self._positions

    # -- loaders used only when record / lists are not passed in (reference :127-210) ----------
    def load_genome(self, file_path: str):
        if not os.path.isfile(file_path):
            raise FileNotFoundError(f"The file {file_path} does not exist.")
        if not file_path.endswith(_GENBANK_SUFFIXES):
            raise ValueError(f"The file {file_path} could not be read.\nEnsure the file holds a GenBank format.")
        return read_genbank(file_path)

    def get_needed_genes(self, file_path: str) -> list:
        if not os.path.isfile(file_path):
            raise FileNotFoundError(f"The file {file_path} does not exist.")
        if not file_path.endswith(".npy"):
            raise ValueError(f"Invalid file format. Expected .npy file, got: {os.path.splitext(file_path)[1]}")
        return np.load(file_path, allow_pickle=True).tolist()

    # -- small helpers (reference :103-125, :254-270) ----------------------------------------------
    def save_minimized_genome(self, file_path: str):
        os.makedirs(_default_dir(), exist_ok=True)
        with open(file_path, "w") as fh:            # header line + sequence, no trailing newline
            fh.write(f">{SEQ_ID_PREFIX}{self.idx+1}\n{self.reduced_genome_str}")

    def plot(self):
        """Reference :212-252 reads `minimised_genomes_sizes`, an attribute nothing ever sets, so calling it
        ends in AttributeError there as here.  Plotting itself is outside this package."""
        sizes = self.minimised_genomes_sizes
        raise NotImplementedError(f"plotting {len(sizes)} genome sizes is not part of the minimizer hot path")

    def get_reduction_stats(self) -> dict:
        n0, n1 = self.original_genome_length, len(self.reduced_genome_str)
        return {"original_length": n0, "reduced_length": n1,
                "reduction_percentage": (n0 - n1) / n0 * 100,
                "genes_removed": len(self.features),
                "positions_removed": len(self.positions_to_remove)}


def _ranks() -> int:
    """> 1 when the samples should be sharded over the ranks of a job (dist.py) instead of every rank repeating
    the whole job: inside an initialised process group, or in a process started by torchrun (RANK, WORLD_SIZE,
    MASTER_ADDR and MASTER_PORT all set).  Anything less — a stray WORLD_SIZE from another launcher — leaves the
    entry functions doing what the reference does.  GM2_SHARD=0 turns sharding off."""
    started_by_torchrun = os.environ.get("WORLD_SIZE", "1") not in ("", "1")
    if not started_by_torchrun and "torch.distributed" not in sys.modules:
        return 1                 # nothing can have initialised a process group: torch is not imported for this
    from . import dist as _dist
    return _dist.launched_ranks()


def process_multiple_genomes_single_file(genome_path: str, genes_path: str, model_name: str, output_file: str = None):
    """All samples into ONE FASTA file; returns {"genome_count", "average_reduction_pct", "average_length_bp"}.
    Under torchrun (one process per GPU) the samples are sharded over the ranks: every rank writes its own
    contiguous part of the same file, rank 0 prints the progress lines, all ranks return the same dict."""
    if not output_file:
        output_file = os.path.join(_default_dir(), f"minimized_genomes_{model_name}.fasta")
    os.makedirs(os.path.dirname(output_file), exist_ok=True)
    record = _engine.ReferenceGenome.from_file(genome_path)            # reference :455
    all_lists = _engine.load_gene_lists(genes_path, record.table)       # reference :456
    if _ranks() > 1:
        from . import dist as _dist
        _dist.ensure_process_group()
        return _dist.run_single_file_sharded(record, all_lists, model_name, output_file)
    return _engine.run_single_file(record, all_lists, model_name, output_file)


def process_multiple_genomes_multiple_files(genome_path: str, genes_path: str, model_name: str,
                                            output_dir: str = None,
                                            filename_template: str = "minimized_{model}_{idx:04d}.fasta"):
    """Each sample into its own FASTA file under output_dir; same return dict (all samples averaged).
    Under torchrun every rank writes the files of its own contiguous shard of the samples."""
    if output_dir is None:
        output_dir = _default_dir()
    os.makedirs(output_dir, exist_ok=True)
    record = _engine.ReferenceGenome.from_file(genome_path)            # reference :515
    all_lists = _engine.load_gene_lists(genes_path, record.table)       # reference :518
    if _ranks() > 1:
        from . import dist as _dist
        _dist.ensure_process_group()
        return _dist.run_multi_file_sharded(record, all_lists, model_name, output_dir, filename_template)
    return _engine.run_multi_file(record, all_lists, model_name, output_dir, filename_template)
