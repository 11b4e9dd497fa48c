"""Host logic of the `GenomeMinimiser` facade (reference minimizer_2.py:19-270) that needs no GPU: argument
precedence, loaders and their errors, attributes, `save_minimized_genome`, `get_reduction_stats` — against
the golden `class_sample` blocks minted from the reference class.  The engine is the oracle-backed test
double (tests/engine_double.py); the same checks run through the CUDA engine in tests/test_gpu_parity.py."""
from __future__ import annotations

import hashlib

import numpy as np
import pytest

import engine_double
from genome_minimizer_2_b200 import engine, genbank, minimizer_2 as m2


@pytest.fixture(autouse=True)
def _double(monkeypatch):
    monkeypatch.setattr(engine, "MinimizerEngine", engine_double.OracleEngine)
    monkeypatch.setattr(m2, "_ENGINES", {})                     # the per-record engine cache starts empty


def test_class_attributes_match_the_reference(golden, golden_paths):
    cs = golden.get("class_sample")
    if cs is None:
        pytest.skip("fixture without a class sample")
    gb, npy = golden_paths
    rec = genbank.read_genbank(gb)
    k = cs["idx"]
    forms = [
        m2.GenomeMinimiser(record=rec, needed_genes_list=golden["lists"][k], idx=k, model_name="m"),
        m2.GenomeMinimiser(record=rec, all_needed_gene_lists=golden["lists"], idx=k),
        m2.GenomeMinimiser(record_path=gb, needed_genes_path=npy, idx=k),
        # needed_genes_list wins over the other two sources (reference :38-43)
        m2.GenomeMinimiser(record=rec, needed_genes_list=golden["lists"][k], all_needed_gene_lists=[["nomatch"]] * 99, idx=k),
        # containers other than list behave as Python's `in` does on them
        m2.GenomeMinimiser(record=rec, needed_genes_list=tuple(golden["lists"][k]), idx=k),
        m2.GenomeMinimiser(record=rec, needed_genes_list=np.array(golden["lists"][k], dtype=object), idx=k),
    ]
    for gm in forms:
        assert gm.idx == k and gm.original_genome_length == len(rec.seq)
        assert hashlib.sha256(gm.reduced_genome_str.encode()).hexdigest() == cs["reduced_genome_str_sha256"]
        assert [[int(f.location.start), int(f.location.end)] for f in gm.features] == cs["removed_gene_spans"]
        assert all(f.type == "gene" for f in gm.features)
        assert len(gm.positions_to_remove) == cs["positions_removed"]
        assert gm.get_reduction_stats() == cs["stats"]
    assert forms[0].model_name == "m" and forms[0].wildtype_sequence is rec and forms[0].record is rec


def test_save_minimized_genome_writes_header_and_sequence_without_final_newline(golden, golden_paths, tmp_path, monkeypatch):
    if not golden["lists"]:
        pytest.skip("no samples")
    gb, _ = golden_paths
    monkeypatch.setattr(m2, "PROJECT_ROOT", str(tmp_path / "root"))
    gm = m2.GenomeMinimiser(record=genbank.read_genbank(gb), needed_genes_list=golden["lists"][0], idx=4)
    out = tmp_path / "one.fasta"
    gm.save_minimized_genome(str(out))
    assert out.read_text() == f">Minimized_E_coli_K12_MG1655_5\n{gm.reduced_genome_str}"     # reference :120-121
    assert (tmp_path / "root" / "minimized_genomes").is_dir()                                 # reference :116-117


def test_loader_errors_are_the_reference_errors(tmp_path):
    gm = object.__new__(m2.GenomeMinimiser)
    gm.idx = 0
    with pytest.raises(FileNotFoundError, match="does not exist"):
        gm.load_genome(str(tmp_path / "absent.gb"))
    wrong = tmp_path / "genome.txt"
    wrong.write_text("LOCUS\n//\n")
    with pytest.raises(ValueError, match="Ensure the file holds a GenBank format"):
        gm.load_genome(str(wrong))
    with pytest.raises(FileNotFoundError, match="does not exist"):
        gm.get_needed_genes(str(tmp_path / "absent.npy"))
    with pytest.raises(ValueError, match=r"Expected \.npy file, got: \.txt"):
        gm.get_needed_genes(str(wrong))
    for suffix in (".gb", ".genbank", ".gbff"):                          # reference :142
        p = tmp_path / f"g{suffix}"
        p.write_text("LOCUS       A 4 bp DNA linear\nORIGIN\n        1 acgt\n//\n")
        assert gm.load_genome(str(p)).seq == "ACGT"
    two = tmp_path / "two.gb"
    two.write_text("LOCUS       A 4 bp\nORIGIN\n        1 acgt\n//\n" * 2)
    with pytest.raises(ValueError, match="More than one record found in handle"):
        gm.load_genome(str(two))


def test_one_engine_per_record_object(golden_paths, golden):
    if len(golden["lists"]) < 2:
        pytest.skip("needs two samples")
    gb, _ = golden_paths
    rec = genbank.read_genbank(gb)
    a = m2.GenomeMinimiser(record=rec, needed_genes_list=golden["lists"][0], idx=0)
    assert len(m2._ENGINES) == 1
    b = m2.GenomeMinimiser(record=rec, needed_genes_list=golden["lists"][1], idx=1)
    assert len(m2._ENGINES) == 1                                         # the genome is not uploaded again
    assert a.reduced_genome_str == golden.get("sequences", [a.reduced_genome_str])[0]
    if "sequences" in golden:
        assert b.reduced_genome_str == golden["sequences"][1]
    del a, b, rec
    import gc
    gc.collect()
    assert len(m2._ENGINES) == 0                                         # dropped with the record


# ---- the two entry functions over every fixture (host logic only: loaders, progress lines, F10 averages, files)
def test_single_file_entry_host_logic(golden, golden_paths, tmp_path, capsys):
    gb, npy = golden_paths
    out = tmp_path / "nested" / "dir" / "out.fasta"                      # parents are created (reference :453)
    ret = m2.process_multiple_genomes_single_file(gb, npy, golden["model_name"], str(out))
    stdout = capsys.readouterr().out
    data = out.read_bytes().decode()
    lines = data.split("\n")
    assert lines[2].startswith("# Generated on: ") and len(lines[2]) == len("# Generated on: 2026-01-01T00:00:00")
    lines[2] = "# Generated on: <TS>"
    data = "\n".join(lines)
    if "single_file" in golden:
        assert data == golden["single_file"]
    else:
        assert hashlib.sha256(data.encode()).hexdigest() == golden["single_file_sha256"]
    assert stdout == golden["single_stdout"]
    assert ret == golden["single_return"]


def test_multi_file_entry_host_logic(golden, golden_paths, tmp_path, capsys):
    import os
    from pathlib import Path
    gb, npy = golden_paths
    out_dir = tmp_path / "multi"
    ret = m2.process_multiple_genomes_multiple_files(gb, npy, golden["model_name"], Path(out_dir))   # main.py passes a Path
    stdout = capsys.readouterr().out
    files = {fn: (out_dir / fn).read_bytes().decode() for fn in sorted(os.listdir(out_dir))}
    if "multi_files" in golden:
        assert files == golden["multi_files"]
    else:
        assert {k: hashlib.sha256(v.encode()).hexdigest() for k, v in files.items()} == golden["multi_files_sha256"]
    assert stdout.replace(str(out_dir), "<OUTDIR>") == golden["multi_stdout"]
    assert ret == golden["multi_return"]
