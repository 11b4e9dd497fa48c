"""SURVEY.md §8 f1 (BASELINE config 5): decoder output -> gene lists -> minimizer.
CPU part: the converter oracle and the host-side column mapping against fixtures minted from the
reference's own binary_converter.py.  GPU part (-m gpu): the fused device path reproduces the file
the reference's three-step chain writes."""
from __future__ import annotations

import contextlib
import io
import re

import numpy as np
import pytest

from conftest import load_golden
from oracle import converter_oracle as co, genbank_reader, minimizer_oracle as mo
from genome_minimizer_2_b200 import engine, genbank

CASES = ["converter_0", "converter_1"]


@pytest.mark.parametrize("name", CASES)
def test_converter_oracle_matches_reference(name):
    c = load_golden(name)
    decoded = np.asarray(c["decoded"], dtype=np.float32)
    binary = co.threshold_samples(decoded)
    lists = co.masks_to_gene_lists(binary, c["columns"])
    assert lists == c["lists_before_essentials"]
    assert co.add_essentials(lists, c["essential"]) == c["lists"]
    assert binary[0].sum() == 0 and binary[2].all()                     # 0.5 is not present (strict >)


@pytest.mark.parametrize("name", CASES)
def test_column_space_equals_list_membership(name, tmp_path):
    c = load_golden(name)
    p = tmp_path / "g.gb"
    p.write_text(c["genbank"])
    table = engine.GeneTable.from_record(genbank.read_genbank(str(p)))
    space = engine.ColumnSpace(table, c["columns"], c["essential"])
    decoded = np.asarray(c["decoded"], dtype=np.float32)
    assert space.V == decoded.shape[1]
    fk = np.unpackbits(space.force_keep.view(np.uint8), bitorder="little")[:table.F].astype(bool)
    fi = np.unpackbits(space.forced_ids.view(np.uint8), bitorder="little")[:space.V].astype(bool)
    for s, needed in enumerate(c["lists"]):
        present = decoded[s] > 0.5
        keep = fk.copy()
        for col in np.flatnonzero(present):
            keep[space.id2gene_idx[space.id2gene_off[col]:space.id2gene_off[col + 1]]] = True
        assert keep.tolist() == mo.keep_vector(table.names, needed).tolist()
        count = int(present.sum()) + int((fi & ~present).sum()) + space.essentials_not_in_columns
        assert count == len(needed)                                      # what "[i/N] genes present:" prints


def test_wrong_row_length_raises_like_the_reference():
    with pytest.raises(ValueError, match="gene columns"):
        co.masks_to_gene_lists(np.zeros((2, 3)), ["a", "b"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_chain_matches_reference_file(name, tmp_path):
    import torch
    c = load_golden(name)
    gb = tmp_path / "g.gb"
    gb.write_text(c["genbank"])
    rec = genbank.read_genbank(str(gb))
    decoded = torch.tensor(c["decoded"], dtype=torch.float32, device="cuda:0")
    out = tmp_path / "o" / "chain.fasta"
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ret = engine.run_single_file_from_probabilities(rec, decoded, c["columns"], c["essential"],
                                                        c["model_name"], str(out))
    lines = out.read_bytes().decode().split("\n")
    assert re.fullmatch(r"# Generated on: \d{4}-\d\d-\d\dT\d\d:\d\d:\d\d", lines[2])
    lines[2] = "# Generated on: <TS>"
    assert "\n".join(lines) == c["single_file"]
    assert buf.getvalue() == c["single_stdout"]
    assert ret == c["single_return"]
    # a strided (padded) matrix and an unaligned base pointer give the same answer
    eng = engine.MinimizerEngine(rec)
    try:
        space = engine.ColumnSpace(eng.table, c["columns"], c["essential"])
        wide = torch.zeros(decoded.shape[0], decoded.shape[1] + 5, dtype=torch.float32, device="cuda:0")
        wide[:, 1:1 + decoded.shape[1]] = decoded
        l1, n1 = engine.plan_from_probabilities(eng, space, decoded)
        l2, n2 = engine.plan_from_probabilities(eng, space, wide[:, 1:1 + decoded.shape[1]])
        assert np.array_equal(l1, l2) and np.array_equal(n1, n2)
        assert n1.tolist() == [len(x) for x in c["lists"]]
        with pytest.raises(ValueError, match="gene columns"):
            engine.plan_from_probabilities(eng, space, wide)
    finally:
        eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("container", ["2d-float64", "object-rows"])
def test_masks_file_chain_matches_reference_file(name, container, tmp_path):
    """From the masks FILE `--mode sample` writes (main.py:434) to the FASTA the reference's
    convert-samples + minimizer chain produces."""
    c = load_golden(name)
    gb = tmp_path / "g.gb"
    gb.write_text(c["genbank"])
    rec = genbank.read_genbank(str(gb))
    binary = (np.asarray(c["decoded"], dtype=np.float32) > 0.5).astype(float)       # utils/extras.py:199-201
    masks = tmp_path / "binary_samples.npy"
    if container == "2d-float64":
        np.save(masks, binary)
    else:
        arr = np.empty(len(binary), dtype=object)
        for i, row in enumerate(binary):
            arr[i] = row.tolist()
        np.save(masks, arr, allow_pickle=True)
    out = tmp_path / "chain.fasta"
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ret = engine.run_single_file_from_masks(rec, str(masks), c["columns"], c["essential"], c["model_name"], str(out))
    lines = out.read_bytes().decode().split("\n")
    lines[2] = "# Generated on: <TS>"
    assert "\n".join(lines) == c["single_file"]
    assert buf.getvalue() == c["single_stdout"]
    assert ret == c["single_return"]
    # a row of the wrong length is refused with the reference's message
    bad = tmp_path / "bad.npy"
    np.save(bad, binary[:, :-1])
    with pytest.raises(ValueError, match="gene columns"):
        engine.run_single_file_from_masks(rec, str(bad), c["columns"], c["essential"], "m", str(tmp_path / "x.fasta"))
