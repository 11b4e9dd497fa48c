#!/usr/bin/env python
"""Mint golden fixtures by running the REFERENCE's own code (build container only).

    python tests/golden/make_golden.py            # needs /root/reference (read-only mount)

The reference has no tests, fixtures or golden files of its own, so the oracle
is pinned against outputs of the reference's unmodified
`src/genome_minimizer_2/minimizer/minimizer_2.py`, imported from /root/reference.  That module
imports `Bio` and `matplotlib` at the top (minimizer_2.py:10-12); neither is installed and the
hot path only duck-types the record, so both are stubbed in `sys.modules` and
`Bio.SeqIO.read` is served by `oracle/genbank_reader.py` (GenBank parsing parity is therefore
NOT pinned by these fixtures — only everything downstream of the parsed record is).

/root/reference does not exist on the GPU box: tests never import it, they read the JSON
written here.  Fixture format (one JSON per case):
  genbank        GenBank text fed to the reference
  lists          the gene-name lists (saved as an object .npy exactly like binary_converter.py:71)
  model_name
  single_file    bytes the reference wrote, line 3 (timestamp) replaced by "# Generated on: <TS>"
  single_stdout / single_return
  multi_files    {file name: content}, multi_stdout ("<OUTDIR>" for the temp dir), multi_return
  sequences      the per-sample minimized sequences (or sha256 + length for the large cases)
"""
from __future__ import annotations

import contextlib
import gzip
import hashlib
import io
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import genbank_reader  # noqa: E402
import genome_minimizer_2_b200.synth as synth  # noqa: E402


def import_reference():
    """Import the reference's minimizer_2 with Bio / matplotlib stubbed."""
    if not os.path.isdir(REFERENCE):
        raise SystemExit("make_golden.py needs the reference mounted at /root/reference")
    bio = types.ModuleType("Bio")
    seqio = types.ModuleType("Bio.SeqIO")
    seqrecord = types.ModuleType("Bio.SeqRecord")

    def read(path, fmt):
        assert fmt == "genbank"
        return genbank_reader.read_genbank(path)

    seqio.read = read
    seqrecord.SeqRecord = object
    bio.SeqIO = seqio
    bio.SeqRecord = seqrecord
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.update({"Bio": bio, "Bio.SeqIO": seqio, "Bio.SeqRecord": seqrecord,
                        "matplotlib": mpl, "matplotlib.pyplot": plt})
    sys.path.insert(0, REFERENCE)
    import importlib
    return importlib.import_module("src.genome_minimizer_2.minimizer.minimizer_2")


def save_lists(path, lists):
    arr = np.empty(len(lists), dtype=object)
    for i, l in enumerate(lists):
        arr[i] = l
    np.save(path, arr, allow_pickle=True)


def run_reference(ref, genbank_text, lists, model_name, big=False):
    out = {"genbank": genbank_text, "lists": lists, "model_name": model_name}
    with tempfile.TemporaryDirectory() as d:
        gb = os.path.join(d, "genome.gb")
        npy = os.path.join(d, "genes.npy")
        with open(gb, "w") as fh:
            fh.write(genbank_text)
        save_lists(npy, lists)
        # single file
        fasta = os.path.join(d, "single", "out.fasta")
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ret = ref.process_multiple_genomes_single_file(gb, npy, model_name, fasta)
        with open(fasta, "rb") as fh:
            data = fh.read().decode()
        lines = data.split("\n")
        assert lines[2].startswith("# Generated on: ")
        lines[2] = "# Generated on: <TS>"
        data = "\n".join(lines)
        # sequences may be empty strings -> recover them by record structure
        recs = data.split("\n")[3:]
        seqs = [recs[i + 1] for i in range(0, len(recs) - 1, 2)]
        out["single_stdout"] = buf.getvalue()
        out["single_return"] = ret
        # multi file
        mdir = os.path.join(d, "multi")
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ret2 = ref.process_multiple_genomes_multiple_files(gb, npy, model_name, mdir)
        files = {}
        for fn in sorted(os.listdir(mdir)):
            with open(os.path.join(mdir, fn), "rb") as fh:
                files[fn] = fh.read().decode()
        out["multi_stdout"] = buf.getvalue().replace(mdir, "<OUTDIR>")
        out["multi_return"] = ret2
        if big:
            out["single_file_sha256"] = hashlib.sha256(data.encode()).hexdigest()
            out["multi_files_sha256"] = {k: hashlib.sha256(v.encode()).hexdigest() for k, v in files.items()}
            out["sequence_sha256"] = [hashlib.sha256(s.encode()).hexdigest() for s in seqs]
            out["sequence_lengths"] = [len(s) for s in seqs]
        else:
            out["single_file"] = data
            out["multi_files"] = files
            out["sequences"] = seqs
        # GenomeMinimiser attributes for one sample (index 2 when there is one)
        if lists:
            k = min(2, len(lists) - 1)
            rec = genbank_reader.read_genbank(gb)
            gm = ref.GenomeMinimiser(record=rec, needed_genes_list=lists[k], idx=k, model_name=model_name)
            out["class_sample"] = {
                "idx": k,
                "reduced_genome_str_sha256": hashlib.sha256(gm.reduced_genome_str.encode()).hexdigest(),
                "removed_gene_spans": [[int(f.location.start), int(f.location.end)] for f in gm.features],
                "positions_removed": len(gm.positions_to_remove),
                "stats": gm.get_reduction_stats(),
            }
    return out


KAT_GENBANK = """LOCUS       KAT40                     40 bp    DNA     linear   BCT 01-JAN-2000
DEFINITION  Known-answer test record (SURVEY.md App. B).
ACCESSION   KAT40
VERSION     KAT40.1
KEYWORDS    .
SOURCE      synthetic construct
  ORGANISM  synthetic construct
            other sequences.
FEATURES             Location/Qualifiers
     source          1..40
                     /organism="synthetic construct"
     gene            3..6
                     /gene="aaa"
     CDS             3..6
                     /gene="aaa"
                     /product="a"
     gene            complement(5..10)
                     /gene="bbb"
     gene            13..14
                     /locus_tag="only_a_tag"
     CDS             16..18
                     /gene="ccc"
     gene            17..19
                     /gene="ddd"
     gene            20..21
                     /gene="aaa"
     gene            26..30
                     /gene="eee"
                     /gene="syn"
     gene            33^34
                     /gene="fff"
     gene            37..40
                     /gene="ggg"
ORIGIN
        1 acgttgcaag cttaggccat nnacgtrykm acgtttgaca
//
"""
KAT_LISTS = [["aaa", "bbb", "ddd", "eee", "fff", "ggg", ""], ["aaa", "bbb", "ddd", "eee", "fff", "ggg"],
             ["aaa"], ["bbb", "syn", "group_1"], [], ["AAA", "ggg"]]

KAT2_GENBANK = """LOCUS       KAT20                     20 bp    DNA     linear   BCT 01-JAN-2000
FEATURES             Location/Qualifiers
     gene            3..6
                     /gene="aaa"
     gene            5..10
                     /gene="bbb"
ORIGIN
        1 acgttgcaag cttaggccat
//
"""
KAT2_LISTS = [["aaa", "bbb"], ["aaa"]] + [[] for _ in range(10)]

EDGE_GENBANK = """LOCUS       EDGE60                    60 bp    DNA     circular BCT 01-JAN-2000
FEATURES             Location/Qualifiers
     source          1..60
     gene            join(55..60,1..4)
                     /gene="wrap"
     gene            <8..>12
                     /gene="fuzzy"
     gene            complement(join(15..18,25..28))
                     /gene="cjoin"
     gene            order(30..31,
                     36..37)
                     /gene="ordr"
     gene            40
                     /gene="single"
     gene            45..50
                     /gene=unquoted
     gene            45..50
                     /gene="twin"
     gene            52..200
                     /gene="beyond"
     misc_feature    1..60
                     /note="never deletes anything"
ORIGIN
        1 acgtacgtac gtacgtacgt acgtacgtac gtacgtacgt acgtacgtac gtacgtacgt
//
"""
EDGE_LISTS = [["wrap", "fuzzy", "cjoin", "ordr", "single", "unquoted", "twin", "beyond"],
              ["fuzzy", "cjoin", "ordr", "single", "unquoted", "twin", "beyond"],
              ["wrap"], ["wrap", "fuzzy"], ["wrap", "cjoin"], ["wrap", "ordr"], ["wrap", "single", "beyond"],
              ["wrap", "unquoted"], ["wrap", "twin"], ["wrap", "unquoted", "twin", "fuzzy", "cjoin", "ordr", "single"]]


def converter_cases(ref):
    """The reference's own producer chain: thresholded decoder output -> masks_to_gene_lists ->
    check_essential_genes (explore_data/binary_converter.py) -> minimizer."""
    import importlib
    import pandas as pd
    bc = importlib.import_module("src.genome_minimizer_2.explore_data.binary_converter")
    out = {}
    for k, (G, F, seed, extra_cols, dup) in enumerate([(3000, 40, 41, 25, False), (5000, 70, 42, 60, True)]):
        g = synth.make_genome(G, F, seed, nested=3, dup_name_frac=0.08, nameless_frac=0.05, name=f"CONV{k}")
        text = synth.genbank_text(g, seed=seed)
        rng = np.random.default_rng(seed)
        gene_names = list(dict.fromkeys(n for n in g.gene_names() if n))
        cols = gene_names[: int(len(gene_names) * 0.8)] + [f"group_{i}" for i in range(extra_cols)]
        order = rng.permutation(len(cols))
        cols = [cols[i] for i in order]
        if dup:
            cols = cols + [cols[3], cols[7]]                     # duplicate column names: first occurrence wins
        P = len(dict.fromkeys(cols))
        N = 9
        decoded = rng.random((N, P)).astype(np.float32)
        decoded[0, :] = 0.5                                      # exactly the threshold: not present (strict >)
        decoded[1, :] = 0.0
        decoded[2, :] = 1.0
        decoded[3, ::2] = np.float32(0.50000006)                 # just above
        binary = (decoded > 0.5).astype(float)                   # utils/extras.py:199-201
        essential = set(gene_names[::7][:6]) | {gene_names[-1], "not_a_column_1", "not_a_column_2"}
        with tempfile.TemporaryDirectory() as d:
            masks_path = os.path.join(d, "binary.npy")
            np.save(masks_path, binary)
            ids_path = os.path.join(d, "ids.npy")
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                bc.masks_to_gene_lists(masks_path, pd.Index(cols), ids_path)
                id_lists = np.load(ids_path, allow_pickle=True)
                filled = bc.check_essential_genes(essential, id_lists, ids_path)
            lists = np.load(ids_path, allow_pickle=True).tolist()
            lists_ess = np.load(filled, allow_pickle=True).tolist()
        case = run_reference(ref, text, lists_ess, f"conv{k}")
        case.update({"columns": cols, "decoded": decoded.tolist(), "essential": sorted(essential),
                     "lists_before_essentials": lists})
        out[f"converter_{k}"] = case
    return out


def main():
    ref = import_reference()
    cases = {}
    cases["kat_appB"] = run_reference(ref, KAT_GENBANK, KAT_LISTS, "katmodel")
    cases["kat_stats_quirk"] = run_reference(ref, KAT2_GENBANK, KAT2_LISTS, "quirk")
    cases["edge_locations"] = run_reference(ref, EDGE_GENBANK, EDGE_LISTS, "edge")

    # randomized small genomes: overlaps, nesting, duplicates, nameless, join(), IUPAC, origin wrap
    for k, (G, F, seed, kw) in enumerate([
        (1500, 24, 11, dict(nested=3, join_genes=2, iupac_runs=2)),
        (3000, 40, 12, dict(nested=4, overlap_frac=0.4, dup_name_frac=0.1, nameless_frac=0.1)),
        (2000, 30, 13, dict(nested=2, origin_wrap=True, join_genes=3)),
        (997, 60, 14, dict(nested=6, overlap_frac=0.6, genic_frac=0.97, dup_name_frac=0.05)),
        (4096, 12, 15, dict(nested=0, overlap_frac=0.0, genic_frac=0.5)),
    ]):
        g = synth.make_genome(G, F, seed, name=f"RAND{k}", **kw)
        text = synth.genbank_text(g, seed=seed)
        lists = []
        for p in (0.0, 0.1, 0.5, 0.9, 1.0):
            lists += synth.make_gene_lists(g, 3, p, seed=seed * 100 + int(p * 10), extra_names=5,
                                           sort_lists=(p == 0.9))
        lists.append([""])                     # only the nameless genes are kept
        lists.append(g.gene_names())           # everything incl. "" kept
        cases[f"rand_small_{k}"] = run_reference(ref, text, lists, f"rand{k}")

    # 101 samples on a tiny genome: exercises the (idx+1)%100==0 reporting rule
    g = synth.make_genome(600, 10, 21, nested=1, name="HUNDRED")
    lists = synth.make_gene_lists(g, 101, 0.5, seed=2100, extra_names=2)
    cases["hundred_and_one"] = run_reference(ref, synth.genbank_text(g, seed=21), lists, "hundred")

    # medium, K-12-shaped at 1/20 scale: store hashes only (files would be MBs)
    g = synth.make_genome(232_082, 220, 31, name="K12_SMALL")
    lists = synth.make_gene_lists(g, 12, 0.5, seed=3100, extra_names=100)
    big = run_reference(ref, synth.genbank_text(g, seed=31), lists, "k12small", big=True)
    gb_text = big.pop("genbank")
    with gzip.open(os.path.join(HERE, "medium_k12.gb.gz"), "wt") as fh:
        fh.write(gb_text)
    big["genbank_file"] = "medium_k12.gb.gz"
    cases["medium_k12"] = big

    cases.update(converter_cases(ref))

    for name, case in cases.items():
        with open(os.path.join(HERE, name + ".json"), "w") as fh:
            json.dump(case, fh, indent=1, sort_keys=True)
        print(f"wrote {name}.json ({len(case['lists'])} samples)")


if __name__ == "__main__":
    main()
