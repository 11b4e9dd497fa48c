"""The oracle against the golden fixtures minted from the reference's own code
(tests/golden/make_golden.py).  CPU only."""
from __future__ import annotations

import hashlib

import numpy as np
import pytest

from oracle import c_oracle, genbank_reader, minimizer_oracle as mo


def _record(golden, tmp_path):
    p = tmp_path / "g.gb"
    p.write_text(golden["genbank"])
    return genbank_reader.read_genbank(str(p))


def _expected_sequences(golden):
    if "sequences" in golden:
        return golden["sequences"], None
    return None, (golden["sequence_sha256"], golden["sequence_lengths"])


def _check(golden, seqs):
    exp, hashed = _expected_sequences(golden)
    if exp is not None:
        assert seqs == exp
    else:
        assert [len(s) for s in seqs] == hashed[1]
        assert [hashlib.sha256(s.encode()).hexdigest() for s in seqs] == hashed[0]


def test_literal_port_matches_reference(golden, tmp_path):
    if golden["name"] == "medium_k12":
        pytest.skip("literal per-base loop is checked on the small fixtures; numpy/C cover this one")
    rec = _record(golden, tmp_path)
    _check(golden, [mo.minimize_literal(rec, needed) for needed in golden["lists"]])


def test_numpy_port_matches_reference(golden, tmp_path):
    rec = _record(golden, tmp_path)
    names, starts, ends = mo.gene_table(rec)
    seq = np.frombuffer(rec.seq.encode(), dtype=np.uint8)
    seqs = [mo.minimize_numpy(seq, starts, ends, mo.keep_vector(names, needed)).tobytes().decode()
            for needed in golden["lists"]]
    _check(golden, seqs)


def _keep_rows(names, lists):
    keep = np.stack([mo.keep_vector(names, needed) for needed in lists]) if lists else np.zeros((0, len(names)), bool)
    F = len(names)
    fw = (F + 31) // 32
    padded = np.zeros((len(lists), fw * 32), dtype=np.uint8)
    padded[:, :F] = keep
    return np.packbits(padded, axis=1, bitorder="little").view("<u4").reshape(len(lists), fw)


def test_c_port_matches_reference(golden, tmp_path):
    rec = _record(golden, tmp_path)
    names, starts, ends = mo.gene_table(rec)
    seq = np.frombuffer(rec.seq.encode(), dtype=np.uint8)
    rows = _keep_rows(names, golden["lists"])
    lengths, hashes, image = c_oracle.batch(seq, starts, ends, rows, want_image=True)
    # split the image back into sequences
    seqs, pos = [], 0
    img = image.tobytes()
    for i, L in enumerate(lengths):
        hdr = len(mo.HEADER_PREFIX) + len(str(i + 1)) + 2
        rec_bytes = img[pos:pos + hdr + int(L) + 1]
        assert rec_bytes == mo.record_bytes(i, rec_bytes[hdr:hdr + int(L)])
        assert int(hashes[i]) == mo.range_hash(rec_bytes) == c_oracle.range_hash(rec_bytes)
        seqs.append(rec_bytes[hdr:hdr + int(L)].decode())
        pos += len(rec_bytes)
    assert pos == len(img)
    _check(golden, seqs)


def test_entry_function_contract(golden, tmp_path):
    """run_single_file / run_multi_file reproduce file bytes, stdout and return values."""
    rec = _record(golden, tmp_path)
    fast = golden["name"] == "medium_k12"
    if fast:
        names, starts, ends = mo.gene_table(rec)
        seq = np.frombuffer(rec.seq.encode(), dtype=np.uint8)
        minimize = lambda r, needed: mo.minimize_numpy(seq, starts, ends, mo.keep_vector(names, needed)).tobytes().decode()
    else:
        minimize = mo.minimize_literal
    data, log, ret = mo.run_single_file(rec, golden["lists"], golden["model_name"], "<TS>", minimize=minimize)
    assert log == golden["single_stdout"]
    assert ret == golden["single_return"]
    files, mlog, mret = mo.run_multi_file(rec, golden["lists"], golden["model_name"], "<OUTDIR>", minimize=minimize)
    assert mlog == golden["multi_stdout"]
    assert mret == golden["multi_return"]
    if fast:
        assert hashlib.sha256(data).hexdigest() == golden["single_file_sha256"]
        assert {k: hashlib.sha256(v).hexdigest() for k, v in files.items()} == golden["multi_files_sha256"]
    else:
        assert data.decode() == golden["single_file"]
        assert {k: v.decode() for k, v in files.items()} == golden["multi_files"]


def test_class_attributes(golden, tmp_path):
    rec = _record(golden, tmp_path)
    cs = golden["class_sample"]
    needed = golden["lists"][cs["idx"]]
    removed = mo.removed_features(rec, needed)
    assert [[f.location.start, f.location.end] for f in removed] == cs["removed_gene_spans"]
    assert len(mo.positions_to_remove(removed)) == cs["positions_removed"]


def test_stats_quirk_value():
    """12 samples -> 29.2 %, not the true mean."""
    lengths = [20, 14] + [12] * 10
    s = mo.single_file_stats(lengths, 20)
    assert f"{s['average_reduction_pct']:.1f}" == "29.2"
    assert f"{s['average_length_bp']:,.1f}" == "10.8"
    m = mo.multi_file_stats(lengths, 20)
    assert m["average_reduction_pct"] > s["average_reduction_pct"]


def test_range_hash_agrees_between_numpy_and_c():
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 9, 63, 64, 65, 1000, 4097):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert mo.range_hash(b) == c_oracle.range_hash(b)
    assert mo.range_hash(b"ACGTACGTA") != mo.range_hash(b"ACGTACGTC")


def test_random_numpy_vs_c_vs_literal():
    """Property check on random interval soups incl. degenerate intervals."""
    rng = np.random.default_rng(5)
    for trial in range(30):
        G = int(rng.integers(0, 300))
        F = int(rng.integers(0, 40))
        seq = rng.integers(65, 91, G, dtype=np.uint8)
        starts = rng.integers(-5, G + 10, F).astype(np.int64)
        ends = starts + rng.integers(-3, 60, F)
        keep = rng.random((4, F)) < rng.random()
        fw = (F + 31) // 32
        padded = np.zeros((4, fw * 32), dtype=np.uint8)
        padded[:, :F] = keep
        rows = np.packbits(padded, axis=1, bitorder="little").view("<u4").reshape(4, fw)
        lengths, hashes, image = c_oracle.batch(seq, starts, ends, rows, want_image=True)
        pos = 0
        for s in range(4):
            exp = mo.minimize_numpy(seq, starts, ends, keep[s]).tobytes()
            # literal definition straight from the reference's wording
            gone = set()
            for g in range(F):
                if not keep[s, g]:
                    gone.update(range(int(starts[g]), int(ends[g])))
            lit = bytes(b for i, b in enumerate(seq.tobytes()) if i not in gone)
            assert exp == lit
            r = mo.record_bytes(s, exp)
            assert image[pos:pos + len(r)].tobytes() == r
            assert int(lengths[s]) == len(exp)
            pos += len(r)
