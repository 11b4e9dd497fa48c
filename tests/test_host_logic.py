"""Host-side logic above the C-ABI that needs no GPU: gene table, interning, packing,
sharding, synthetic-data determinism.  CPU only."""
from __future__ import annotations

import os

import numpy as np
import pytest

from oracle import minimizer_oracle as mo
from genome_minimizer_2_b200 import engine, genbank, synth


def test_gene_table_from_record_follows_reference_rules(golden, tmp_path):
    p = tmp_path / "g.gb"
    p.write_text(golden["genbank"])
    rec = genbank.read_genbank(str(p))
    t = engine.GeneTable.from_record(rec)
    names, starts, ends = mo.gene_table(rec)
    assert t.names == names and t.starts.tolist() == starts.tolist() and t.ends.tolist() == ends.tolist()
    # CSR maps every id back to exactly the genes carrying that name
    for nm, i in t.name_to_id.items():
        genes = t.id2gene_idx[t.id2gene_off[i]:t.id2gene_off[i + 1]].tolist()
        assert genes == [g for g, n in enumerate(names) if n == nm]


def test_tokenize_equals_python_membership(golden, tmp_path):
    p = tmp_path / "g.gb"
    p.write_text(golden["genbank"])
    t = engine.GeneTable.from_record(genbank.read_genbank(str(p)))
    ids, off = t.tokenize(golden["lists"])
    for s, needed in enumerate(golden["lists"]):
        row = set(ids[off[s]:off[s + 1]].tolist())
        keep = np.zeros(t.F, bool)
        for i in row:
            keep[t.id2gene_idx[t.id2gene_off[i]:t.id2gene_off[i + 1]]] = True
        assert keep.tolist() == [n in needed for n in t.names]


def test_tokenize_odd_containers():
    t = engine.GeneTable(["abc", "", "abc", "x"], np.arange(4), np.arange(4) + 1)
    assert t.V == 3
    # numpy rows, tuples, non-string and unhashable members, a bare string (substring semantics)
    ids, off = t.tokenize([np.array(["abc", "zzz"]), ("x", 5, None), [["abc"]], "xabcx", [""], []])
    rows = [ids[off[i]:off[i + 1]].tolist() for i in range(6)]
    assert rows[0] == [t.name_to_id["abc"]]
    assert rows[1] == [t.name_to_id["x"]]
    assert rows[2] == []
    assert sorted(rows[3]) == sorted(t.name_to_id[n] for n in ("abc", "", "x"))   # "" in "xabcx" is True
    assert rows[4] == [t.name_to_id[""]]
    assert rows[5] == []


def test_keep_row_packing_is_little_endian_words():
    t = engine.GeneTable(["g%d" % i for i in range(70)], np.zeros(70), np.ones(70))
    keep = np.zeros((2, 70), bool)
    keep[0, [0, 31, 32, 69]] = True
    rows = t.keep_rows_from_bool(keep)
    assert rows.shape == (2, 3) and rows.dtype == np.uint32
    assert rows[0].tolist() == [1 | (1 << 31), 1, 1 << 5] and rows[1].tolist() == [0, 0, 0]
    assert np.array_equal(rows, synth.pack_keep_rows(keep))


@pytest.mark.parametrize("S,world", [(0, 1), (1, 4), (10, 3), (100_000, 8), (7, 8)])
def test_shard_ranges_partition_in_rank_order(S, world):
    r = [engine.shard_range(S, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == S
    for a, b in zip(r, r[1:]):
        assert a[1] == b[0]
    sizes = [hi - lo for lo, hi in r]
    assert max(sizes) - min(sizes) <= 1


def test_synthetic_genome_is_deterministic_and_k12_shaped():
    a = synth.make_genome(seed=1)
    b = synth.make_genome(seed=1)
    assert a.G == 4_641_652 and len(a.genes) == 4_400
    assert np.array_equal(a.seq, b.seq) and a.gene_names() == b.gene_names()
    st, en = a.starts_ends()
    assert (st >= 0).all() and (en <= a.G).all() and (st < en).all()
    cov = np.zeros(a.G + 1, int)
    np.add.at(cov, st, 1)
    np.add.at(cov, en, -1)
    assert 0.80 < (np.cumsum(cov)[:-1] > 0).mean() < 0.93
    assert set(np.unique(a.seq).tolist()) == set(b"ACGT")
    names = a.gene_names()
    assert "" in names and len(set(names)) < len(names)            # nameless + duplicate names exist


def test_gene_list_file_round_trips_like_the_reference_reads_it(tmp_path):
    g = synth.make_genome(5_000, 30, 3)
    lists = synth.make_gene_lists(g, 5, 0.5, seed=1, extra_names=4)
    p = tmp_path / "l.npy"
    synth.save_gene_lists(str(p), lists)
    assert np.load(p, allow_pickle=True).tolist() == lists
    # equal-length lists become a 2-D object array; .tolist() still yields list[list[str]]
    same = [["a", "b"], ["c", "d"]]
    np.save(p, np.array(same, dtype=object), allow_pickle=True)
    assert np.load(p, allow_pickle=True).tolist() == same


# ----------------------------------------------------------------------------------------------
# engine.drain_to_file: mapped file vs the portable (p)write form, through the engine double
# ----------------------------------------------------------------------------------------------
def _planned_double(name="rand_small_1"):
    import tempfile
    from conftest import load_golden
    from oracle import genbank_reader
    from engine_double import OracleEngine
    case = load_golden(name)
    with tempfile.NamedTemporaryFile("w", suffix=".gb", delete=False) as fh:
        fh.write(case["genbank"])
    rec = genbank_reader.read_genbank(fh.name)
    os.unlink(fh.name)
    eng = OracleEngine(rec)
    eng.plan_lists(case["lists"])
    return eng, b"".join(eng.images)


@pytest.mark.parametrize("sink", ["map", "write"])
def test_drain_to_file_forms_write_the_same_bytes(sink, tmp_path, monkeypatch):
    from genome_minimizer_2_b200 import engine
    monkeypatch.setenv("GM2_FILE_SINK", sink)
    eng, image = _planned_double()
    pre = b"# preamble\n"
    path = tmp_path / "out.fasta"
    path.write_bytes(pre)
    seen = []
    end = engine.drain_to_file(eng, str(path), len(pre), progress=lambda a, b: seen.append((a, b)))
    assert end == len(pre) + len(image)
    assert path.read_bytes() == pre + image
    assert seen == eng.chunks() and seen[0][0] == 0 and seen[-1][1] == eng.S          # every range once, in order
    # a sub-range at its own offset into a larger, pre-sized file (what a rank of the sharded form does)
    off = eng.record_offsets()
    big = tmp_path / "big.fasta"
    with open(big, "wb") as fh:
        fh.truncate(len(pre) + len(image))
    a, b = 1, eng.S - 1
    engine.drain_to_file(eng, str(big), len(pre) + int(off[a]), s0=a, s1=b)
    got = big.read_bytes()
    assert got[len(pre) + int(off[a]):len(pre) + int(off[b])] == image[int(off[a]):int(off[b])]
    assert got[:len(pre) + int(off[a])].count(0) == len(pre) + int(off[a])            # nothing outside the range
    assert got[len(pre) + int(off[b]):].count(0) == len(got) - len(pre) - int(off[b])


def test_drain_to_file_takes_the_sequential_form_for_a_pipe(tmp_path):
    from genome_minimizer_2_b200 import engine
    eng, image = _planned_double()
    r, w = os.pipe()
    import threading
    got = []
    t = threading.Thread(target=lambda: got.append(os.fdopen(r, "rb").read()))
    t.start()
    try:
        assert engine.drain_to_file(eng, w, 0) == len(image)
    finally:
        os.close(w)
    t.join()
    assert got[0] == image


def test_drain_to_file_refuses_a_full_filesystem_before_producing(tmp_path, monkeypatch):
    import errno
    from genome_minimizer_2_b200 import engine
    eng, image = _planned_double()
    real = os.fstatvfs

    class Full:
        def __init__(self, v):
            self.f_bavail, self.f_frsize = 0, v.f_frsize

    monkeypatch.setattr(os, "fstatvfs", lambda fd: Full(real(fd)))
    path = tmp_path / "o.fasta"
    path.write_bytes(b"")
    called = []
    monkeypatch.setattr(type(eng), "emit_into", lambda self, a, b, out: called.append((a, b)))
    with pytest.raises(OSError) as e:
        engine.drain_to_file(eng, str(path), 0)
    assert e.value.errno == errno.ENOSPC and not called
