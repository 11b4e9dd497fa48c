"""The product's GenBank reader against the oracle's independently written one, plus the
`SeqIO.read` contract.  CPU only.  Parity with Biopython itself is unpinned."""
from __future__ import annotations

import pytest

from oracle import genbank_reader as og
from genome_minimizer_2_b200 import genbank, synth


def _same(a, b):
    assert a.seq == b.seq
    assert len(a.features) == len(b.features)
    for fa, fb in zip(a.features, b.features):
        assert fa.type == fb.type
        assert (fa.location.start, fa.location.end) == (fb.location.start, fb.location.end)
        assert fa.qualifiers == fb.qualifiers


def test_readers_agree_on_golden_files(golden, tmp_path):
    p = tmp_path / "g.gb"
    p.write_text(golden["genbank"])
    _same(genbank.read_genbank(str(p)), og.read_genbank(str(p)))


@pytest.mark.parametrize("seed,kw", [(1, {}), (2, dict(join_genes=20, origin_wrap=True, iupac_runs=4)),
                                     (3, dict(overlap_frac=0.7, nested=30, nameless_frac=0.2, dup_name_frac=0.2))])
def test_readers_agree_on_synthetic(seed, kw, tmp_path):
    g = synth.make_genome(50_000, 120, seed, **kw)
    p = tmp_path / "s.gb"
    synth.write_genbank(str(p), g, seed=seed)
    a = genbank.read_genbank(str(p))
    _same(a, og.read_genbank(str(p)))
    # and both recover exactly what the generator laid out
    assert a.seq == g.seq.tobytes().decode()
    genes = [f for f in a.features if f.type == "gene"]
    assert [f.qualifiers.get("gene", [""])[0] for f in genes] == g.gene_names()
    starts, ends = g.starts_ends()
    assert [f.location.start for f in genes] == starts.tolist()
    assert [f.location.end for f in genes] == ends.tolist()
    assert any(f.type == "CDS" for f in a.features)


@pytest.mark.parametrize("text,exp", [
    ("190..255", (189, 255)),
    ("complement(5683..6459)", (5682, 6459)),
    ("<1..>9", (0, 9)),
    ("42", (41, 42)),
    ("33^34", (33, 33)),
    ("join(10..20,30..40)", (9, 40)),
    ("order(10..20,30..40)", (9, 40)),
    ("complement(join(100..200,\n 5..9))", (4, 200)),
    ("join(4641600..4641652,1..50)", (0, 4641652)),
    ("join(complement(30..40),complement(10..20))", (9, 40)),
    ("join(<5..9, 12..>20)", (4, 20)),
])
def test_location_grammar(text, exp):
    for parse in (genbank.parse_location, og.parse_location):
        loc = parse(text)
        assert (loc.start, loc.end) == exp


@pytest.mark.parametrize("text", ["", "join(", "foo", "10..", "J00194.1:100..202", "12.15", "join(1..2))"])
def test_unsupported_locations_raise(text):
    for parse in (genbank.parse_location, og.parse_location):
        with pytest.raises(ValueError):
            parse(text)


ONE = """LOCUS       A 8 bp DNA linear
FEATURES             Location/Qualifiers
     gene            1..4
                     /gene="x"
ORIGIN
        1 acgtnnry
//
"""


def test_exactly_one_record_rule(tmp_path):
    p = tmp_path / "none.gb"
    p.write_text("just text\n")
    for read in (genbank.read_genbank, og.read_genbank):
        with pytest.raises(ValueError, match="No records found in handle"):
            read(str(p))
    p2 = tmp_path / "two.gb"
    p2.write_text(ONE + ONE)
    for read in (genbank.read_genbank, og.read_genbank):
        with pytest.raises(ValueError, match="More than one record found in handle"):
            read(str(p2))
    p1 = tmp_path / "one.gb"
    p1.write_text(ONE)
    r = genbank.read_genbank(str(p1))
    assert r.seq == "ACGTNNRY" and len(r.features) == 1          # upper-cased, IUPAC passes through


def test_qualifier_forms(tmp_path):
    text = '''LOCUS       Q 10 bp DNA linear
FEATURES             Location/Qualifiers
     gene            1..4
                     /pseudo
                     /gene="first"
                     /gene="second"
                     /note="spans
                     two lines with ""quotes"""
                     /codon_start=1
     gene            5..6
                     /locus_tag="no gene qualifier"
ORIGIN
        1 acgtacgtac
//
'''
    p = tmp_path / "q.gb"
    p.write_text(text)
    for read in (genbank.read_genbank, og.read_genbank):
        r = read(str(p))
        q = r.features[0].qualifiers
        assert q["gene"] == ["first", "second"]
        assert q["pseudo"] == [""]
        assert q["note"] == ['spans two lines with "quotes"']
        assert q["codon_start"] == ["1"]
        assert r.features[1].qualifiers.get("gene", [""])[0] == ""
