"""the native gene-list tokeniser (gm2_tokenize_pickle, host only) against the
reference's own loading path, `np.load(genes_path, allow_pickle=True).tolist()` followed by
`name in needed_genes` (minimizer_2.py:456, :62), restated by `GeneTable.tokenize`.

No GPU is needed: the entry point is host code inside libgm2.so."""
import io
import pickle

import numpy as np
import pytest

from conftest import load_golden
from genome_minimizer_2_b200 import _native, engine, genbank, synth


def _table(names):
    n = len(names)
    return engine.GeneTable(names, np.arange(n, dtype=np.int64) * 10, np.arange(n, dtype=np.int64) * 10 + 5)


def _save_with_protocol(path, arr, protocol):
    """np.save's container (magic, header, one pickle) with a chosen pickle protocol: files written by
    older NumPy releases use protocol 2/3 (BINPUT memo), current ones 4 (MEMOIZE, frames)."""
    from numpy.lib import format as npf
    with open(path, "wb") as fh:
        npf.write_array_header_1_0(fh, npf.header_data_from_array_1_0(arr))
        pickle.dump(arr, fh, protocol=protocol)


def _check(path, table):
    """Native result == Python result on the same file; returns the Python lists."""
    lists = np.load(path, allow_pickle=True).tolist()
    want_ids, want_off = table.tokenize(lists)
    got = _native.tokenize_npy(str(path), table.vocabulary())
    assert got is not None, "file should be inside the native subset"
    ids, off, counts = got
    assert np.array_equal(off, want_off)
    assert np.array_equal(ids, want_ids)
    assert counts.tolist() == [len(x) for x in lists]
    return lists


@pytest.mark.parametrize("case", ["kat_appB", "kat_stats_quirk", "rand_small_0", "rand_small_3", "hundred_and_one"])
def test_golden_lists(tmp_path, case):
    c = load_golden(case)
    gb = tmp_path / "g.gb"
    gb.write_text(c["genbank"])
    table = engine.GeneTable.from_record(genbank.read_genbank(str(gb)))
    p = tmp_path / "lists.npy"
    synth.save_gene_lists(str(p), c["lists"])
    _check(p, table)
    loaded = engine.load_gene_lists(str(p), table)
    assert isinstance(loaded, engine.TokenizedLists) and len(loaded) == len(c["lists"])


def test_fuzz_against_python_path(tmp_path):
    rng = np.random.default_rng(11)
    pool = [f"g{i}" for i in range(700)] + ["", "thrL", "dnaK", "gène", "β-lac", "名前", "x" * 300, "a b", "A", "a"]
    for trial in range(25):
        names = [str(x) for x in rng.choice(pool, size=int(rng.integers(1, 400)), replace=True)]   # duplicate gene names happen
        table = _table(names)
        extra = [f"group_{i}" for i in range(int(rng.integers(0, 500)))]                  # names outside the genome
        cols = np.array(sorted(set(names)) + extra, dtype=object)
        S = int(rng.integers(0, 40))
        lists = []
        for _ in range(S):
            k = int(rng.integers(0, cols.size + 1))
            l = cols[rng.integers(0, cols.size, k)].tolist()                              # with repeats, any order
            if rng.random() < 0.2:
                l = []
            lists.append(l)
        p = tmp_path / f"f{trial}.npy"
        if trial % 5 == 4 and S > 0:
            # every list the same length: np.array(lists, dtype=object) is 2-D and np.save keeps that shape
            k = int(rng.integers(1, 30))
            lists = [cols[rng.integers(0, cols.size, k)].tolist() for _ in range(S)]
            arr = np.array(lists, dtype=object)
            assert arr.ndim == 2
            np.save(p, arr, allow_pickle=True)
        elif trial % 5 == 3:
            arr = np.empty(S, dtype=object)
            for i, l in enumerate(lists):
                arr[i] = tuple(l) if i % 2 else l                                         # tuples behave like lists for `in`
            _save_with_protocol(p, arr, protocol=int(rng.choice([2, 3, 5])))
        else:
            synth.save_gene_lists(str(p), lists)
        _check(p, table)


def test_many_samples_share_memoised_names(tmp_path):
    """The shape of a real file: a few thousand distinct str objects, millions of memo references
    (BINGET for the first 256, LONG_BINGET after), APPENDS in batches of 1000."""
    names = [f"gene{i}" for i in range(3000)]
    table = _table(names)
    cols = np.array(names + [f"group_{i}" for i in range(1500)], dtype=object)
    rng = np.random.default_rng(3)
    lists = [cols[np.flatnonzero(rng.random(cols.size) < 0.5)].tolist() for _ in range(60)]
    p = tmp_path / "big.npy"
    synth.save_gene_lists(str(p), lists)
    _check(p, table)


def test_sliced_tokens_match_sliced_lists(tmp_path):
    names = [f"g{i}" for i in range(50)]
    table = _table(names)
    rng = np.random.default_rng(5)
    lists = [list(rng.choice(names + ["zz"], size=int(rng.integers(0, 30)))) for _ in range(17)]
    lists = [[str(x) for x in l] for l in lists]
    p = tmp_path / "l.npy"
    synth.save_gene_lists(str(p), lists)
    tok = engine.load_gene_lists(str(p), table)
    assert isinstance(tok, engine.TokenizedLists)
    for lo, hi in ((0, 17), (3, 9), (9, 9), (16, 17), (0, 0)):
        part = tok[lo:hi]
        ids, off = table.tokenize(lists[lo:hi])
        assert len(part) == hi - lo
        assert np.array_equal(part.ids, ids) and np.array_equal(part.off, off)
        assert [len(part[i]) for i in range(hi - lo)] == [len(l) for l in lists[lo:hi]]
    assert len(tok[4]) == len(lists[4])


@pytest.mark.parametrize("content", ["non_str_item", "np_str_item", "bare_string", "nested_array", "unicode_array",
                                     "bool_array", "set_element", "three_d"])
def test_outside_the_subset_falls_back_to_numpy(tmp_path, content):
    """Anything for which `name in needed_genes` is not plain str-vs-str equality must go through
    NumPy/Python unchanged: the native tokeniser declines, load_gene_lists returns what np.load gives."""
    names = ["a", "b", "c"]
    table = _table(names)
    p = tmp_path / "x.npy"
    if content == "non_str_item":
        arr = np.empty(2, dtype=object); arr[0] = ["a", 3, None]; arr[1] = ["b"]
    elif content == "np_str_item":
        arr = np.empty(1, dtype=object); arr[0] = [np.str_("a"), "b"]          # np.str_('a') == 'a': Python must decide
    elif content == "bare_string":
        arr = np.empty(2, dtype=object); arr[0] = "abc"; arr[1] = ["a"]         # substring semantics (engine.ids_for)
    elif content == "nested_array":
        arr = np.empty(1, dtype=object); arr[0] = np.array(["a", "b"], dtype=object)
    elif content == "unicode_array":
        arr = np.array([["a", "b"], ["c", "zz"]])                                # dtype <U2: no pickle at all
    elif content == "bool_array":
        arr = np.zeros((2, 3), dtype=bool)
    elif content == "set_element":
        arr = np.empty(1, dtype=object); arr[0] = {"a", "b"}
    else:
        arr = np.empty((2, 2, 2), dtype=object); arr[...] = "a"
    np.save(p, arr, allow_pickle=True)
    assert _native.tokenize_npy(str(p), table.vocabulary()) is None
    got = engine.load_gene_lists(str(p), table)
    want = np.load(p, allow_pickle=True).tolist()
    assert not isinstance(got, engine.TokenizedLists)
    assert repr(got) == repr(want)


def test_corrupt_and_foreign_files(tmp_path):
    table = _table(["a", "b"])
    good = tmp_path / "good.npy"
    synth.save_gene_lists(str(good), [["a", "b"], ["b"]])
    raw = good.read_bytes()
    for name, data in (("truncated.npy", raw[:-7]), ("garbage.npy", raw[:128] + b"\xff" * 40), ("text.npy", b"not an npy file")):
        p = tmp_path / name
        p.write_bytes(data)
        assert _native.tokenize_npy(str(p), table.vocabulary()) is None      # NumPy's loader reports the error
        with pytest.raises(Exception):
            engine.load_gene_lists(str(p), table)
    with pytest.raises(FileNotFoundError):
        engine.load_gene_lists(str(tmp_path / "missing.npy"), table)


def test_shape_mismatch_is_rejected(tmp_path):
    """A header that promises more elements than the pickle holds is not silently accepted."""
    table = _table(["a"])
    arr = np.empty(3, dtype=object)
    for i in range(3):
        arr[i] = ["a"]
    bio = io.BytesIO()
    np.save(bio, arr, allow_pickle=True)
    raw = bio.getvalue().replace(b"(3,)", b"(4,)")
    p = tmp_path / "bad.npy"
    p.write_bytes(raw)
    assert _native.tokenize_npy(str(p), table.vocabulary()) is None


def test_empty_file_and_empty_vocabulary(tmp_path):
    p = tmp_path / "empty.npy"
    synth.save_gene_lists(str(p), [])
    tok = engine.load_gene_lists(str(p), _table(["a"]))
    assert isinstance(tok, engine.TokenizedLists) and len(tok) == 0 and tok.off.tolist() == [0]
    p2 = tmp_path / "l.npy"
    synth.save_gene_lists(str(p2), [["a", "b"], []])
    tok = engine.load_gene_lists(str(p2), engine.GeneTable([], np.zeros(0, np.int64), np.zeros(0, np.int64)))
    assert isinstance(tok, engine.TokenizedLists) and tok.ids.size == 0 and tok.counts.tolist() == [2, 0]
