#!/usr/bin/env python
"""Stage a runtime copy of the reference for `bench.py --impl reference` (build container only).

    python baseline/stage_reference.py          # needs /root/reference (read-only mount)

The contract's `pip install --no-index --target baseline/_ref /root/reference` cannot work offline: the
reference builds with poetry-core (pyproject.toml:43-45), which is not in the wheelhouse, and pins
python < 3.11.  The reference is pure Python, so an install would only copy its files; this script copies
the UNMODIFIED tree (`main.py`, `src/`) into baseline/_ref/ instead.  That directory is git-ignored (it is
not this repository's source and never enters its history) but travels to the GPU box with `gpurun`, where
/root/reference does not exist.  bench.py imports `src/genome_minimizer_2/minimizer/minimizer_2.py` from it
with `Bio` / `matplotlib` stubbed (neither is installed; the minimizer path only duck-types the record)."""
import os
import shutil
import sys

REFERENCE = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def stage() -> bool:
    if not os.path.isdir(REFERENCE):
        return os.path.isdir(DEST)
    shutil.rmtree(DEST, ignore_errors=True)
    os.makedirs(DEST)
    shutil.copy2(os.path.join(REFERENCE, "main.py"), os.path.join(DEST, "main.py"))
    shutil.copytree(os.path.join(REFERENCE, "src"), os.path.join(DEST, "src"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for name in ("pyproject.toml", "README.md"):
        shutil.copy2(os.path.join(REFERENCE, name), os.path.join(DEST, name))
    return True


if __name__ == "__main__":
    ok = stage()
    print("baseline/_ref staged" if ok else "no /root/reference and no staged copy")
    sys.exit(0 if ok else 1)
