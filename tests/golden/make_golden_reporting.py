#!/usr/bin/env python
"""Mint the golden fixture for the reporting helpers of the reference module (build container only).

    python tests/golden/make_golden_reporting.py        # needs /root/reference (read-only mount)

Runs the reference's UNMODIFIED `check_sequence_duplicates`, `print_duplicate_statistics` and
`generate_summary_file` (minimizer_2.py:273-444), imported exactly as make_golden.py imports the
module, and freezes what they return / print / write into `tests/golden/reporting/reporting.json`.  The only
thing patched is the module global `PROJECT_ROOT` (the reference writes the summary under
`PROJECT_ROOT/minimized_genomes`, and /root/reference is read-only); the "Generated on:" line holds a
wall-clock timestamp and is replaced by "<TS>".
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from make_golden import import_reference  # noqa: E402


def duplicate_cases():
    long_a = "ACGT" * 20                       # > 50 characters: printed truncated with "..."
    exactly_50 = "G" * 50
    cases = {
        "empty": {},
        "all_unique": {"s1": "ACGT", "s2": "ACGA", "s3": ""},
        "one_pair": {"s1": "ACGT", "s2": "TTTT", "s3": "ACGT"},
        "long_and_many_ids": {**{f"id{i}": long_a for i in range(7)}, "x": exactly_50, "y": exactly_50, "z": "A"},
        "empty_strings_duplicate": {"a": "", "b": "", "c": "C"},
    }
    # twelve duplicate groups of different sizes: only the ten largest are printed, ties keep dict order
    many = {}
    for g in range(12):
        for k in range(2 + (g * 5) % 4):
            many[f"g{g}_{k}"] = "AC" * (g + 1)
    many["lonely"] = "T"
    cases["twelve_groups"] = many
    return cases


def summary_cases():
    return {
        "typical": dict(output_file="/some/dir/minimized_genomes_v0.fasta", model_name="v0",
                        genome_path="/data/wild_type_sequence.gb", genes_path="/x/y/lists.npy",
                        original_length=4641652,
                        minimised_sizes=[2.6082, 2.5911, 2.7003, 2.45, 2.6082, 2.8123, 2.3999, 2.6, 2.61, 2.59, 2.9]),
        "no_sizes": dict(output_file="out.fasta", model_name="", genome_path="g.gb", genes_path="l.npy",
                         original_length=1000, minimised_sizes=[]),
        "zero_original_length": dict(output_file="a/b/c.fasta", model_name="m", genome_path="g.gbff",
                                     genes_path="l.npy", original_length=0, minimised_sizes=[0.5, 0.25]),
        "all_equal_sizes": dict(output_file="same.fasta", model_name="same", genome_path="g.gb",
                                genes_path="l.npy", original_length=2000000, minimised_sizes=[1.5, 1.5, 1.5]),
        "single_size": dict(output_file="one.fasta", model_name="one", genome_path="g.gb",
                            genes_path="l.npy", original_length=40, minimised_sizes=[0.00002]),
        "name_without_fasta_suffix": dict(output_file="/tmp/genomes.fa", model_name="fa", genome_path="g.gb",
                                          genes_path="l.npy", original_length=100, minimised_sizes=[0.00005, 0.00006]),
    }


def main():
    ref = import_reference()
    out = {"duplicates": {}, "summaries": {}}
    for name, d in duplicate_cases().items():
        stats = ref.check_sequence_duplicates(d)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ref.print_duplicate_statistics(stats)
        out["duplicates"][name] = {
            "sequences": [[k, v] for k, v in d.items()],          # pairs: json sort_keys must not reorder the input
            "stats": {k: v for k, v in stats.items() if k != "duplicates_detail"},
            "duplicates_detail": [[seq, ids] for seq, ids in stats["duplicates_detail"].items()],   # insertion order
            "printed": buf.getvalue(),
        }
    dup_stats = ref.check_sequence_duplicates(duplicate_cases()["twelve_groups"])
    for name, kw in summary_cases().items():
        with tempfile.TemporaryDirectory() as d:
            ref.PROJECT_ROOT = d                    # module global: where the reference puts the summary
            ref.generate_summary_file(duplicate_stats=dup_stats, **kw)
            sub = os.path.join(d, "minimized_genomes")
            files = {}
            for fn in sorted(os.listdir(sub)):
                with open(os.path.join(sub, fn)) as fh:
                    lines = fh.read().split("\n")
                lines = ["Generated on: <TS>" if ln.startswith("Generated on: ") else ln for ln in lines]
                files[fn] = "\n".join(lines)
        out["summaries"][name] = {"args": kw, "files": files}
    path = os.path.join(HERE, "reporting", "reporting.json")      # a sub-directory: tests/conftest.py globs golden/*.json for minimizer cases
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print(f"wrote {path}")


if __name__ == "__main__":
    main()
