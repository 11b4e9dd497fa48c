"""The reporting helpers of the drop-in module (`check_sequence_duplicates`,
`print_duplicate_statistics`, `generate_summary_file`, `GenomeMinimiser.plot`) against what the
reference's own functions returned / printed / wrote (tests/golden/reporting/reporting.json, minted by
tests/golden/make_golden_reporting.py from minimizer_2.py:273-444).  CPU only."""
from __future__ import annotations

import json
import os

import pytest

from genome_minimizer_2_b200 import minimizer_2 as m2, reporting

with open(os.path.join(os.path.dirname(__file__), "golden", "reporting", "reporting.json")) as _fh:
    GOLD = json.load(_fh)


@pytest.mark.parametrize("name", sorted(GOLD["duplicates"]))
def test_duplicate_statistics_match_the_reference(name, capsys):
    case = GOLD["duplicates"][name]
    stats = m2.check_sequence_duplicates(dict(case["sequences"]))
    detail = stats.pop("duplicates_detail")
    assert stats == case["stats"]
    assert [[seq, ids] for seq, ids in detail.items()] == case["duplicates_detail"]      # same order too
    stats["duplicates_detail"] = detail
    m2.print_duplicate_statistics(stats)
    assert capsys.readouterr().out == case["printed"]


@pytest.mark.parametrize("name", sorted(GOLD["summaries"]))
def test_summary_file_matches_the_reference(name, tmp_path, monkeypatch):
    case = GOLD["summaries"][name]
    dup = reporting.check_sequence_duplicates(dict(GOLD["duplicates"]["twelve_groups"]["sequences"]))
    monkeypatch.setattr(m2, "PROJECT_ROOT", str(tmp_path))            # the reference's module global, same role
    m2.generate_summary_file(duplicate_stats=dup, **case["args"])
    out_dir = tmp_path / "minimized_genomes"
    assert sorted(os.listdir(out_dir)) == sorted(case["files"])
    for fn, expected in case["files"].items():
        lines = (out_dir / fn).read_text().split("\n")
        stamp = [ln for ln in lines if ln.startswith("Generated on: ")]
        assert len(stamp) == 1 and len(stamp[0]) == len("Generated on: 2026-01-01T00:00:00")
        lines = ["Generated on: <TS>" if ln.startswith("Generated on: ") else ln for ln in lines]
        assert "\n".join(lines) == expected


def test_summary_failures_are_logged_not_raised(tmp_path, caplog):
    import numpy as np
    dup = reporting.check_sequence_duplicates({})
    # a NumPy array has no truth value: the reference's `if minimised_sizes` raises inside its try block
    reporting.generate_summary_file("x.fasta", "m", "g.gb", "l.npy", 10, np.array([1.0, 2.0]), dup,
                                    project_root=str(tmp_path))
    assert not os.path.exists(tmp_path / "minimized_genomes" / "x_summary.txt")
    assert any("Failed to generate summary file" in r.getMessage() for r in caplog.records)


def test_plot_behaves_like_the_reference_without_sizes(capsys):
    gm = object.__new__(m2.GenomeMinimiser)            # no GPU here: only the method is under test
    gm.model_name = "m"
    with pytest.raises(AttributeError):                # the reference never sets minimised_genomes_sizes
        gm.plot()
    gm.minimised_genomes_sizes = [2.5] * 99             # plotting itself is out of scope
    with pytest.raises(NotImplementedError):
        gm.plot()


def test_every_public_name_of_the_reference_module_exists():
    for name in ("GenomeMinimiser", "process_multiple_genomes_single_file", "process_multiple_genomes_multiple_files",
                 "check_sequence_duplicates", "print_duplicate_statistics", "generate_summary_file", "PROJECT_ROOT"):
        assert hasattr(m2, name), name
    for meth in ("save_minimized_genome", "load_genome", "get_needed_genes", "plot", "get_reduction_stats"):
        assert callable(getattr(m2.GenomeMinimiser, meth)), meth
