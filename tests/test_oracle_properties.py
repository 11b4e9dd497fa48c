"""Property tests (hypothesis) over the three CPU restatements of the hot path — the literal port of
minimizer_2.py:50-101, the NumPy form and the C form — on interval soups the reference's rule must
survive: zero-length, nested, identical, touching, out-of-range and whole-genome spans, duplicate and
empty gene names, empty lists, non-ACGT bases (proposed tiers).  CPU only; the GPU parity
tests compare the CUDA path with these same functions."""
from __future__ import annotations

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import c_oracle, minimizer_oracle as mo
from genome_minimizer_2_b200 import genbank, synth

NAMES = ["aaa", "bbb", "ccc", "ddd", "", "AAA", "thrL", "group_1"]


@st.composite
def cases(draw):
    G = draw(st.integers(0, 120))
    seq = bytes(draw(st.lists(st.sampled_from(b"ACGTNRYKM"), min_size=G, max_size=G)))
    F = draw(st.integers(0, 12))
    genes = []
    for _ in range(F):
        kind = draw(st.integers(0, 5))
        a = draw(st.integers(0, max(G, 1)))
        if kind == 0:
            b = a                                             # zero length
        elif kind == 1:
            a, b = 0, G                                       # origin-wrapping join: the whole genome
        elif kind == 2:
            b = G + draw(st.integers(0, 30))                  # end beyond the sequence
        else:
            b = a + draw(st.integers(0, 40))
        genes.append((draw(st.sampled_from(NAMES)), a, b))
    lists = draw(st.lists(st.lists(st.sampled_from(NAMES + ["nomatch"]), max_size=6), min_size=1, max_size=4))
    return seq, genes, lists


def _record(seq: bytes, genes):
    feats = [genbank.Feature("source", genbank.Location(0, len(seq)), {})]
    for name, a, b in genes:
        quals = {"gene": [name, "synonym"]} if name else {"locus_tag": ["t"]}
        feats.append(genbank.Feature("gene", genbank.Location(a, b), quals))
        feats.append(genbank.Feature("CDS", genbank.Location(a, b), {"gene": [name]}))      # distractor
    return genbank.GenomeRecord(seq.decode(), feats)


@settings(max_examples=200, deadline=None)
@given(cases())
def test_three_restatements_agree(case):
    seq, genes, lists = case
    rec = _record(seq, genes)
    names, starts, ends = mo.gene_table(rec)
    assert names == [g[0] for g in genes]
    arr = np.frombuffer(seq, dtype=np.uint8)
    keep = np.stack([mo.keep_vector(names, needed) for needed in lists]) if genes else np.zeros((len(lists), 0), bool)
    lengths, _, image = c_oracle.batch(arr, starts, ends, synth.pack_keep_rows(keep), want_image=True)
    pos = 0
    for s, needed in enumerate(lists):
        lit = mo.minimize_literal(rec, needed).encode()
        assert mo.minimize_numpy(arr, starts, ends, keep[s]).tobytes() == lit
        r = mo.record_bytes(s, lit)
        assert image[pos:pos + len(r)].tobytes() == r and int(lengths[s]) == len(lit)
        pos += len(r)
    assert pos == image.size
    # the many-samples length form (used by bench.py to check every length of a 100,000-sample job)
    assert np.array_equal(mo.kept_lengths_numpy(len(seq), starts, ends, keep), lengths)


def test_vectorised_lengths_agree_with_the_mask_form_on_a_gene_shaped_genome():
    g = synth.make_genome(150_000, 140, seed=11, nested=6, overlap_frac=0.4, join_genes=3, origin_wrap=True)
    starts, ends = g.starts_ends()
    rng = np.random.default_rng(5)
    keep = rng.random((37, len(starts))) < rng.random((37, 1))
    exp = [int(mo.kept_mask_numpy(g.G, starts, ends, k).sum()) for k in keep]
    assert mo.kept_lengths_numpy(g.G, starts, ends, keep, chunk=8).tolist() == exp


@settings(max_examples=150, deadline=None)
@given(cases())
def test_rule_properties(case):
    """Consequences of `deleted(p) <=> some non-kept gene covers p`."""
    seq, genes, lists = case
    rec = _record(seq, genes)
    names = [g[0] for g in genes]
    every = sorted(set(names))
    assert mo.minimize_literal(rec, every) == seq.decode()                 # all names kept: nothing deleted
    none_kept = mo.minimize_literal(rec, [])
    covered = set()
    for _, a, b in genes:
        covered.update(range(a, min(b, len(seq))))
    assert none_kept == "".join(chr(c) for i, c in enumerate(seq) if i not in covered)      # intergenic bases only
    for needed in lists:
        out = mo.minimize_literal(rec, needed)
        assert len(none_kept) <= len(out) <= len(seq)
        more = mo.minimize_literal(rec, list(needed) + every[:2])
        assert len(more) >= len(out)                                       # keeping more never deletes more
        assert mo.minimize_literal(rec, list(needed) + list(needed) + ["nomatch"]) == out   # duplicates, unknown names
        # the output is a subsequence of the genome in ascending order
        it = iter(seq.decode())
        assert all(ch in it for ch in out)
