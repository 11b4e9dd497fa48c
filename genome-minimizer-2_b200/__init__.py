"""genome_minimizer_2_b200 — B200 (sm_100a) build of genome-minimizer-2's `--mode minimizer`
hot path: gene-name lists -> GenBank `gene` intervals -> reduced genomes as FASTA.

Layout
  csrc/gm2.cu     hand-written CUDA kernels + the C-ABI declared in include/gm2.h
  _native.py      ctypes binding of libgm2.so (fails loudly when the library or a GPU is missing)
  engine.py       host side above the C-ABI: gene table, name interning, streaming drain, sharding
  genbank.py      restated `Bio.SeqIO.read(path, "genbank")` (Biopython is a reference dependency)
  minimizer_2.py  drop-in mirror of the reference module
                  src/genome_minimizer_2/minimizer/minimizer_2.py (same names, arguments, prints, returns)
  reporting.py    the module's reporting helpers (duplicate statistics, summary file), host only
  synth.py        synthetic K-12-shaped GenBank + gene-list generator (the reference ships no data)

Nothing here imports `oracle/`: the oracle is test infrastructure.
"""
from .minimizer_2 import (  # noqa: F401
    GenomeMinimiser,
    process_multiple_genomes_single_file,
    process_multiple_genomes_multiple_files,
    check_sequence_duplicates,
    print_duplicate_statistics,
    generate_summary_file,
)

__all__ = [
    "GenomeMinimiser",
    "process_multiple_genomes_single_file",
    "process_multiple_genomes_multiple_files",
    "check_sequence_duplicates",
    "print_duplicate_statistics",
    "generate_summary_file",
]
