"""oracle/ — CPU restatement of genome-minimizer-2's `--mode minimizer` hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `genome-minimizer-2_b200/` (the product) may
import, call, link or execute anything in this directory; the only permitted users are
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs, and there only as the checker or the CPU baseline.

Parity status
  * Everything downstream of a parsed record (minimizer_2.py:50-101, :447-560 of the
    reference) is PINNED: `tests/golden/make_golden.py` imported the reference's own,
    unmodified `minimizer_2.py` in the build container and froze its outputs as
    fixtures under `tests/golden/`; `tests/test_oracle.py` checks every function here
    against them.
  * GenBank parsing (Biopython 1.85 `SeqIO.read(path, "genbank")`, poetry.lock:4-5 of the
    reference; Biopython is not installed and not vendored) is "parity unpinned": the
    reference has no tests or fixtures for it.  `oracle/genbank_reader.py` restates the
    documented behaviour and is cross-checked against the product's
    separately written reader, nothing more.
"""
