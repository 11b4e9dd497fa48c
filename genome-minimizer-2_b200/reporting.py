"""Reporting helpers of the reference module (`minimizer_2.py:273-444`): duplicate-sequence statistics,
their printed form, and the summary text file.  None of them is reachable from the reference's CLI;
they are here so the drop-in module exposes every public name of the one it replaces.

Same inputs, return values, printed text and file contents as the reference (golden fixture
`tests/golden/reporting/reporting.json`, minted from the reference's own functions).  Host-only code: on the
GPU path the statistics come from `MinimizerEngine.duplicate_stats()` (device-side sequence hashes,
`gm2_sequence_hashes`), whose result has the same keys and can be passed to the two printers below.
"""
from __future__ import annotations

import logging
import os
from typing import Dict, List

import numpy as np

RULE = "=" * 80
SUBRULE = "-" * 40
SHOWN_GROUPS, SHOWN_CHARS, SHOWN_IDS = 10, 50, 5


def check_sequence_duplicates(sequences_dict: dict) -> dict:
    """Group ids by identical sequence string (reference :273-303).  Groups keep first-seen order."""
    by_sequence: Dict[str, List[str]] = {}
    for seq_id, sequence in sequences_dict.items():
        by_sequence.setdefault(sequence, []).append(seq_id)
    repeated = {seq: ids for seq, ids in by_sequence.items() if len(ids) > 1}
    total = len(sequences_dict)
    return {
        "total_sequences": total,
        "unique_sequences": len(by_sequence),
        "duplicate_groups": len(repeated),
        "duplicated_sequences": sum(map(len, repeated.values())),
        "unique_only_sequences": len(by_sequence) - len(repeated),
        "duplicates_detail": repeated,
        "compression_ratio": len(by_sequence) / total if total else 0,
    }


def _clip(items: list, limit: int, sep: str) -> str:
    return sep.join(items[:limit]) + ("..." if len(items) > limit else "")


def print_duplicate_statistics(duplicate_stats: dict):
    """Print the statistics block (reference :306-343): overview, then the ten largest groups."""
    st = duplicate_stats
    out = ["\n" + RULE, "SEQUENCE DUPLICATION ANALYSIS", RULE, " Overview:"]
    out += [f"- {label}: {st[key]:,}" for label, key in (
        ("Total sequences generated", "total_sequences"), ("Unique sequences", "unique_sequences"),
        ("Duplicate groups", "duplicate_groups"), ("Sequences with duplicates", "duplicated_sequences"),
        ("Truly unique sequences", "unique_only_sequences"))]
    out.append(f"- Percentage of unique sequences: {st['compression_ratio']:.2%}")
    if st["duplicate_groups"] > 0:
        out.append("\n Duplicate Details:")
        # stable sort: equally large groups stay in first-seen order
        ranked = sorted(st["duplicates_detail"].items(), key=lambda kv: -len(kv[1]))
        for rank, (sequence, ids) in enumerate(ranked[:SHOWN_GROUPS], start=1):
            shown = sequence[:SHOWN_CHARS] + ("..." if len(sequence) > SHOWN_CHARS else "")
            out += [f"Group {rank}: {len(ids)} identical sequences", f"- Sequence: {shown}",
                    f"- IDs: {_clip(list(ids), SHOWN_IDS, ', ')}", ""]
        if len(ranked) > SHOWN_GROUPS:
            out.append(f"  ... and {len(ranked) - SHOWN_GROUPS} more duplicate groups")
    else:
        out.append("\n✓ No duplicate sequences found!")
    out.append(RULE)
    print("\n".join(out))


def _section(title: str, lines: List[str]) -> List[str]:
    return [title, SUBRULE, *lines]


def summary_text(output_file: str, summary_file: str, model_name: str, genome_path: str, genes_path: str,
                 original_length: int, minimised_sizes: list, duplicate_stats: dict, timestamp) -> str:
    """The summary file's content (reference :374-441); sizes are in Mbp."""
    sizes = minimised_sizes
    have = bool(sizes)                                   # a NumPy array here raises, as in the reference
    mean, median, lo, hi, std = ((float(np.mean(sizes)), float(np.median(sizes)), float(np.min(sizes)),
                                  float(np.max(sizes)), float(np.std(sizes))) if have else (0, 0, 0, 0, 0))
    body = [RULE, "GENOME MINIMIZATION SUMMARY REPORT", RULE, ""]
    body += _section("GENERATION INFORMATION", [
        f"Model Name: {model_name}", f"Generated on: {timestamp}",
        f"Output FASTA file: {os.path.basename(output_file)}",
        f"Summary file: {os.path.basename(summary_file)}", ""])
    body += _section("INPUT FILES", [
        f"Genome template: {os.path.basename(genome_path)}",
        f"Gene lists file: {os.path.basename(genes_path)}",
        f"Original genome length: {original_length:,} bp", ""])
    body += _section("PROCESSING STATISTICS", [f"Successfully processed: {len(sizes):,}", ""])
    body += _section("MINIMIZED GENOME SIZE STATISTICS", [
        *(f"{label} size: {v:.3f} Mbp ({v * 1e6:,.0f} bp)" for label, v in
          (("Mean", mean), ("Median", median), ("Minimum", lo), ("Maximum", hi))),
        f"Standard deviation: {std:.3f} Mbp", f"Size range: {hi - lo:.3f} Mbp", ""])
    if original_length > 0:
        def red(v):
            return (original_length - v * 1e6) / original_length * 100
        body += _section("GENOME REDUCTION STATISTICS", [
            f"Mean reduction: {red(mean):.2f}%",
            f"Minimum reduction: {red(hi):.2f}% (largest genome)",
            f"Maximum reduction: {red(lo):.2f}% (smallest genome)", ""])
    ds = duplicate_stats
    body += _section("SEQUENCE DUPLICATION ANALYSIS", [
        f"Total sequences: {ds['total_sequences']:,}", f"Unique sequences: {ds['unique_sequences']:,}",
        f"Duplicate groups: {ds['duplicate_groups']:,}",
        f"Sequences with duplicates: {ds['duplicated_sequences']:,}",
        f"Uniqueness ratio: {ds['compression_ratio']:.2%}"])
    if have:
        edges = np.linspace(lo, hi, 6)
        counts, _ = np.histogram(sizes, bins=edges)
        body += _section("\nSIZE DISTRIBUTION SUMMARY", [
            f"{edges[i]:.2f} - {edges[i + 1]:.2f} Mbp: {counts[i]:,} genomes ({counts[i] / len(sizes) * 100:.1f}%)"
            for i in range(len(counts))])
    return "\n".join(body) + "\n"


def generate_summary_file(output_file: str, model_name: str, genome_path: str, genes_path: str,
                          original_length: int, minimised_sizes: list, duplicate_stats: dict,
                          project_root: str = None):
    """Write `<project_root>/minimized_genomes/<fasta name with .fasta -> _summary.txt>` (reference
    :346-444).  Like the reference, any failure is logged and swallowed.  `project_root` is an
    extension (default: the drop-in module's PROJECT_ROOT)."""
    try:
        if project_root is None:
            from .minimizer_2 import PROJECT_ROOT as project_root
        out_dir = os.path.join(project_root, "minimized_genomes")
        os.makedirs(out_dir, exist_ok=True)
        summary_file = os.path.join(out_dir, os.path.basename(output_file).replace(".fasta", "_summary.txt"))
        logging.info(f"Generating summary file: {os.path.basename(summary_file)}")
        text = summary_text(output_file, summary_file, model_name, genome_path, genes_path,
                            original_length, minimised_sizes, duplicate_stats, np.datetime64("now"))
        with open(summary_file, "w") as fh:
            fh.write(text)
        logging.info(f"✓ Summary file saved: {summary_file}")
    except Exception as exc:                                # reference :443-444
        logging.error(f"✗ Failed to generate summary file: {exc}")
