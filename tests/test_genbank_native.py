"""The native GenBank scanner (gm2_genbank_parse, csrc/host_genbank.hpp) against the
two Python readers (the product's `genbank.py` and the oracle's independently written one): same
sequence bytes and same gene table wherever the scanner accepts a file, and a decline — never a
different answer — wherever the general reader raises.  Host only, no GPU."""
from __future__ import annotations

import random

import numpy as np
import pytest

from oracle import genbank_reader as og
from genome_minimizer_2_b200 import _native, engine, genbank, synth


def _python_table(path, read=genbank.read_genbank):
    rec = read(path)
    t = engine.GeneTable.from_record(rec)
    return genbank.sequence_bytes(rec), t.names, t.starts.tolist(), t.ends.tolist(), len(rec.features)


def _native_table(path):
    got = _native.scan_genbank(path)
    if got is None:
        return None
    seq, names, starts, ends, nf = got
    return seq, names, starts.tolist(), ends.tolist(), nf


def _assert_same(a, b):
    assert np.array_equal(a[0], b[0])
    assert a[1:] == b[1:]


def test_scanner_equals_both_readers_on_golden_files(golden, tmp_path):
    p = tmp_path / "g.gb"
    p.write_text(golden["genbank"])
    nat = _native_table(str(p))
    assert nat is not None
    _assert_same(nat, _python_table(str(p)))
    _assert_same(nat, _python_table(str(p), og.read_genbank))


@pytest.mark.parametrize("seed,kw", [(1, {}), (2, dict(join_genes=20, origin_wrap=True, iupac_runs=4)),
                                     (3, dict(overlap_frac=0.7, nested=30, nameless_frac=0.2, dup_name_frac=0.2))])
def test_scanner_equals_readers_on_synthetic(seed, kw, tmp_path):
    g = synth.make_genome(60_000, 150, seed, **kw)
    p = tmp_path / "s.gb"
    synth.write_genbank(str(p), g, seed=seed)
    nat = _native_table(str(p))
    assert nat is not None
    _assert_same(nat, _python_table(str(p)))
    starts, ends = g.starts_ends()
    assert nat[1] == g.gene_names() and nat[2] == starts.tolist() and nat[3] == ends.tolist()
    assert np.array_equal(nat[0], g.seq)


ODD = '''LOCUS       ODD 60 bp DNA linear
DEFINITION  odd but legal shapes.
FEATURES             Location/Qualifiers
     source          1..60
                     /note="/gene=""not a qualifier"" inside a value
                     /gene="still inside the note"
                     end of note"
     gene            1..4
                     /pseudo
                     /gene="first"
                     /gene="second"
     gene            5..6
                     /gene
                     /gene="bare came first"
     gene            complement(join(7..9,
                     12..14))
                     /note="spans
                     two lines with ""quotes"""
                     /gene="after ""quoted"" note"
     gene            15
                     /gene=unquoted
     gene            16^17
                     /gene ="key with a blank is another key"
     gene            <18..>19
                       /gene="indented past column 22: a stray line"
                     /locus_tag="only a tag"
     gene            order(20..21,30..31)
                     /gene="closing quote then blank" 
                     /gene="swallowed by the value above"
     gene\t           1..2
     gene_long_key   22..23
                     /gene="key is not gene"
     geneXYZ         24..25
     gene            26..27
/gene="qualifier in column 1 belongs to nobody: the table ended above"
ORIGIN
        1 acgtacgtac gtacgtacgt acgtacgtnn
       31 ryACGTacgt\tacgt acgtacgtac
//
'''


def test_scanner_equals_readers_on_odd_shapes(tmp_path):
    p = tmp_path / "odd.gb"
    p.write_text(ODD)
    nat = _native_table(str(p))
    assert nat is not None
    _assert_same(nat, _python_table(str(p)))
    assert nat[1][:5] == ["first", "", 'after "quoted" note', "unquoted", ""]
    assert (nat[2][2], nat[3][2]) == (6, 14) and (nat[2][4], nat[3][4]) == (16, 16)


ONE = """LOCUS       A 8 bp DNA linear
FEATURES             Location/Qualifiers
     gene            1..4
                     /gene="x"
ORIGIN
        1 acgtnnry
//
"""


@pytest.mark.parametrize("text", [
    "just text\n",                                          # no record: the reader raises ValueError
    ONE + ONE,                                              # two records: the reader raises ValueError
    ONE.replace("\n", "\r\n"),                              # carriage returns: left to universal-newline decoding
    ONE.replace('"x"', '"é"'),                         # non-ASCII
    ONE.replace("1..4", "J00194.1:1..4"),                   # remote location
    ONE.replace("1..4", "1.4"),                             # within-position location
    ONE.replace("1..4", "join(1..2"),                       # malformed
    ONE.replace("1..4", "1..12345678901234567890"),         # beyond the scanner's integer range
])
def test_scanner_declines_what_it_does_not_handle(text, tmp_path):
    p = tmp_path / "d.gb"
    p.write_bytes(text.encode("utf-8"))
    assert _native.scan_genbank(str(p)) is None
    ref = engine.ReferenceGenome                             # and the fall-back is the general reader
    try:
        want = _python_table(str(p))
    except ValueError:
        with pytest.raises(ValueError):
            ref.from_file(str(p))
    else:
        got = ref.from_file(str(p))
        assert not got.native
        assert np.array_equal(got.seq, want[0]) and got.table.names == want[1]


def test_reference_genome_from_file_prefers_the_scanner(golden, tmp_path):
    p = tmp_path / "g.gb"
    p.write_text(golden["genbank"])
    ref = engine.ReferenceGenome.from_file(str(p))
    want = engine.ReferenceGenome.from_record(genbank.read_genbank(str(p)))
    assert ref.native and not want.native
    assert np.array_equal(ref.seq, want.seq)
    assert ref.table.names == want.table.names
    assert np.array_equal(ref.table.starts, want.table.starts) and np.array_equal(ref.table.ends, want.table.ends)
    assert np.array_equal(ref.table.id2gene_off, want.table.id2gene_off)
    assert np.array_equal(ref.table.id2gene_idx, want.table.id2gene_idx)


def test_empty_and_featureless_files(tmp_path):
    p = tmp_path / "e.gb"
    p.write_bytes(b"")
    assert _native.scan_genbank(str(p)) is None
    p.write_text("LOCUS       E 0 bp\n//\n")
    nat = _native_table(str(p))
    _assert_same(nat, _python_table(str(p)))
    assert nat[0].size == 0 and nat[1] == []
    p.write_text("LOCUS       E 4 bp\nFEATURES             Location/Qualifiers\nORIGIN\n        1 acgt\n")   # no terminator
    _assert_same(_native_table(str(p)), _python_table(str(p)))


def test_fuzzed_files_never_disagree(tmp_path):
    """Random edits of a small file: whenever the scanner answers, the general reader gives the same
    answer; whenever the general reader raises, the scanner declines."""
    g = synth.make_genome(900, 14, 5, nested=2, join_genes=2, nameless_frac=0.15, dup_name_frac=0.1)
    base = synth.genbank_text(g, seed=5)
    alphabet = ' "/=\n()0123456789.,^<>:genjoicmplrdt\t_ACGTacgtn'
    rng = random.Random(1234)
    p = tmp_path / "f.gb"
    answered = declined = 0
    for it in range(1500):
        t = list(base)
        for _ in range(rng.randint(1, 4)):
            kind = rng.random()
            i = rng.randrange(len(t))
            if kind < 0.35:
                t[i] = rng.choice(alphabet)
            elif kind < 0.6:
                t.insert(i, rng.choice(alphabet))
            elif kind < 0.8:
                del t[i]
            else:                                            # duplicate or drop a whole line
                text = "".join(t)
                lines = text.split("\n")
                k = rng.randrange(len(lines))
                if rng.random() < 0.5:
                    lines.insert(k, lines[k])
                else:
                    del lines[k]
                t = list("\n".join(lines))
        text = "".join(t)
        p.write_text(text)
        nat = _native_table(str(p))
        try:
            want = _python_table(str(p))
        except (ValueError, AttributeError):
            assert nat is None, f"iteration {it}: the reader raises but the scanner answered"
            declined += 1
            continue
        if nat is None:
            declined += 1
            continue
        answered += 1
        assert np.array_equal(nat[0], want[0]) and nat[1:] == want[1:], f"iteration {it}"
    assert answered > 800 and declined > 20
