"""Parity of the CUDA path against the oracle and the golden fixtures.  Needs a B200: -m gpu.

Everything here goes through the C-ABI (ctypes binding `_native.Context`) or the drop-in
entry functions that sit on top of it.  Bar: bit-exact (byte/integer work).
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c_oracle, genbank_reader, minimizer_oracle as mo  # noqa: E402  (checker only)

import genome_minimizer_2_b200 as gm2  # noqa: E402
from genome_minimizer_2_b200 import _native, engine, genbank, synth  # noqa: E402


# ----------------------------------------------------------------------------------------------
# golden fixtures through the drop-in entry functions
# ----------------------------------------------------------------------------------------------
def _mask_ts(data: bytes) -> str:
    lines = data.decode().split("\n")
    assert re.fullmatch(r"# Generated on: \d{4}-\d\d-\d\dT\d\d:\d\d:\d\d", lines[2]), lines[2]
    lines[2] = "# Generated on: <TS>"
    return "\n".join(lines)


def test_single_file_entry_matches_reference(golden, golden_paths, tmp_path):
    gb, npy = golden_paths
    out = tmp_path / "deep" / "dir" / "out.fasta"
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ret = gm2.process_multiple_genomes_single_file(gb, npy, golden["model_name"], str(out))
    data = _mask_ts(out.read_bytes())
    if "single_file" in golden:
        assert data == golden["single_file"]
    else:
        assert hashlib.sha256(data.encode()).hexdigest() == golden["single_file_sha256"]
    assert buf.getvalue() == golden["single_stdout"]
    assert ret == golden["single_return"]


def test_multi_file_entry_matches_reference(golden, golden_paths, tmp_path):
    gb, npy = golden_paths
    outdir = tmp_path / "multi"
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ret = gm2.process_multiple_genomes_multiple_files(gb, npy, golden["model_name"], outdir)  # a Path, as main.py:603 passes
    files = {fn: (outdir / fn).read_bytes() for fn in sorted(os.listdir(outdir))}
    if "multi_files" in golden:
        assert {k: v.decode() for k, v in files.items()} == golden["multi_files"]
    else:
        assert {k: hashlib.sha256(v).hexdigest() for k, v in files.items()} == golden["multi_files_sha256"]
    assert buf.getvalue().replace(str(outdir), "<OUTDIR>") == golden["multi_stdout"]
    assert ret == golden["multi_return"]


def test_class_facade_matches_reference(golden, golden_paths):
    gb, npy = golden_paths
    cs = golden["class_sample"]
    rec = genbank.read_genbank(gb)
    m = gm2.GenomeMinimiser(record=rec, needed_genes_list=golden["lists"][cs["idx"]], idx=cs["idx"], model_name="x")
    assert hashlib.sha256(m.reduced_genome_str.encode()).hexdigest() == cs["reduced_genome_str_sha256"]
    assert [[f.location.start, f.location.end] for f in m.features] == cs["removed_gene_spans"]
    assert m.get_reduction_stats() == cs["stats"]
    # the path-based constructor (record_path / needed_genes_path) gives the same answer
    m2 = gm2.GenomeMinimiser(record_path=gb, needed_genes_path=npy, idx=cs["idx"])
    assert m2.reduced_genome_str == m.reduced_genome_str


def test_empty_gene_list_file_divides_by_zero(tmp_path):
    """Reference behaviour for 0 lists: preamble written, then ZeroDivisionError (minimizer_2.py:488)."""
    g = synth.make_genome(500, 6, 3, nested=0)
    gb = tmp_path / "g.gb"
    synth.write_genbank(str(gb), g)
    npy = tmp_path / "e.npy"
    np.save(npy, np.empty(0, dtype=object), allow_pickle=True)
    with pytest.raises(ZeroDivisionError):
        gm2.process_multiple_genomes_single_file(str(gb), str(npy), "m", str(tmp_path / "o.fasta"))
    assert (tmp_path / "o.fasta").read_text().startswith("# Minimized genomes generated using model: m\n# Total genomes: 0\n")


# ----------------------------------------------------------------------------------------------
# C-ABI level parity against the C oracle on seeded random inputs
# ----------------------------------------------------------------------------------------------
def _oracle_image(seq, starts, ends, rows, first_idx=0):
    lengths, hashes, image = c_oracle.batch(seq, starts, ends, rows, first_idx=first_idx, want_image=True)
    return lengths, hashes, image


def _gpu_image(ctx, S):
    """The image through gm2_emit_host, by BOTH transports when the reference allows it: image bytes
    over PCIe (k_emit) and two bits per base + host expansion (k_emit_packed, GM2_CFG_WIRE)."""
    n = ctx.image_bytes(0, S)
    outs = []
    for wire in (1, 2):
        ctx.configure(_native.CFG_WIRE, wire)
        out = np.full(n, 0x2a, dtype=np.uint8)
        try:
            ctx.emit_host(0, S, out)
        except _native.Gm2Error as e:
            if wire == 2 and e.code == _native.ERR_STATE:        # reference with letters other than ACGT
                continue
            raise
        assert ctx.query(_native.Q_LAST_WIRE) == (wire if n else 1)
        outs.append(out)
    ctx.configure(_native.CFG_WIRE, 0)
    if len(outs) == 2:
        assert np.array_equal(outs[0], outs[1]), "the two transports disagree"
    return outs[0]


@pytest.mark.parametrize("G,F,seed,kw", [
    (0, 0, 1, {}),
    (1, 0, 2, {}),
    (17, 3, 3, dict(nested=1)),
    (4096, 0, 4, {}),
    (5000, 64, 5, dict(nested=6, overlap_frac=0.5, join_genes=4)),
    (65536, 100, 6, dict(nested=4)),            # exactly one tile
    (65537, 100, 7, dict(nested=4)),            # one base into the second tile
    (200_000, 33, 8, dict(nested=2, iupac_runs=5, origin_wrap=True)),
    (300_001, 900, 9, dict(nested=30, overlap_frac=0.4, genic_frac=0.95)),
])
def test_keep_rows_parity_random(G, F, seed, kw):
    g = synth.make_genome(G, F, seed, **kw)
    starts, ends = g.starts_ends()
    Fg = len(g.genes)
    S = 37
    p = np.linspace(0.0, 1.0, S)
    keep = synth.random_keep_bool(Fg, S, p, seed=seed)
    rows = synth.pack_keep_rows(keep) if Fg else np.zeros((S, 0), dtype=np.uint32)
    exp_len, exp_hash, exp_img = _oracle_image(g.seq, starts, ends, rows, first_idx=9_999_990)
    with _native.Context(0) as ctx:
        ctx.set_reference(g.seq, starts, ends)
        ctx.load_keep_host(rows.reshape(S, -1)) if Fg else ctx.load_keep_host(np.zeros((S, 0), dtype=np.uint32).reshape(S, 0))
        ctx.plan(9_999_990)                      # ids cross 7 -> 8 digits inside this batch
        assert np.array_equal(ctx.lengths(), exp_len)
        off = ctx.record_offsets()
        assert off[0] == 0 and off[-1] == exp_img.size
        img = _gpu_image(ctx, S)
        assert np.array_equal(img, exp_img)


def test_interval_soup_parity():
    """Arbitrary (unsorted, nested, duplicated, empty, out-of-range) intervals."""
    rng = np.random.default_rng(77)
    for trial in range(12):
        G = int(rng.integers(1, 150_000))
        F = int(rng.integers(1, 300))
        seq = rng.integers(65, 91, G, dtype=np.uint8)
        starts = rng.integers(-50, G + 50, F).astype(np.int64)
        ends = starts + rng.integers(-20, max(G // 20, 30), F)
        if trial % 3 == 0:
            starts[0], ends[0] = 0, G                      # a gene spanning everything
        S = 19
        rows = synth.pack_keep_rows(rng.random((S, F)) < rng.random())
        exp_len, _, exp_img = _oracle_image(seq, starts, ends, rows)
        with _native.Context(0) as ctx:
            ctx.configure(_native.CFG_TILE_BYTES, int(rng.choice([4096, 8192, 65536, 131072])))
            ctx.configure(_native.CFG_EMIT_WARPS, int(rng.choice([1, 2, 4, 8])))
            ctx.set_reference(seq, starts, ends)
            ctx.load_keep_host(rows)
            ctx.plan(0)
            assert np.array_equal(ctx.lengths(), exp_len)
            assert np.array_equal(_gpu_image(ctx, S), exp_img)


def test_ids_mode_matches_keep_mode_and_oracle():
    """K1: name-id lists (duplicates, unknown ids, 1:many names) -> keep rows."""
    g = synth.make_genome(120_000, 300, 21, nested=10, dup_name_frac=0.2, nameless_frac=0.05)
    starts, ends = g.starts_ends()
    table = engine.GeneTable(g.gene_names(), starts, ends)
    rng = np.random.default_rng(3)
    S = 64
    keep_names = rng.random((S, table.V)) < 0.5
    ids, off = synth.ids_csr_from_keep(keep_names, n_noise=40, V=table.V, seed=4)
    ids = np.concatenate([ids, ids[:0]])
    # expected keep rows: gene kept iff its name id is in the row
    name_id = np.asarray([table.name_to_id[n] for n in table.names])
    keep = keep_names[:, name_id]
    rows = synth.pack_keep_rows(keep)
    exp_len, _, exp_img = _oracle_image(g.seq, starts, ends, rows)
    with _native.Context(0) as ctx:
        ctx.set_reference(g.seq, starts, ends)
        ctx.set_name_map(table.id2gene_off, table.id2gene_idx)
        # negative and huge ids are legal noise
        ids2 = ids.copy()
        ctx.load_ids_host(np.concatenate([ids2, np.asarray([-1, 2**31 - 1], dtype=np.int32)]),
                          np.concatenate([off[:-1], [off[-1] + 2]]).astype(np.int64))
        ctx.plan(0)
        assert np.array_equal(ctx.keep_rows(), rows)
        assert np.array_equal(ctx.lengths(), exp_len)
        assert np.array_equal(_gpu_image(ctx, S), exp_img)


def test_chunked_and_ranged_emit_equal_whole_image():
    g = synth.make_genome(150_000, 200, 31)
    starts, ends = g.starts_ends()
    S = 50
    rows = synth.pack_keep_rows(synth.random_keep_bool(len(g.genes), S, 0.5, seed=1))
    _, _, exp_img = _oracle_image(g.seq, starts, ends, rows)
    with _native.Context(0) as ctx:
        ctx.set_reference(g.seq, starts, ends)
        ctx.load_keep_host(rows)
        ctx.plan(0)
        off = ctx.record_offsets()
        whole = np.empty(off[-1], dtype=np.uint8)
        ctx.emit_host(0, S, whole, chunk_bytes=200_000)       # forces many small staged chunks
        assert np.array_equal(whole, exp_img)
        for a, b in [(0, 1), (7, 19), (49, 50), (20, 20)]:
            part = np.empty(off[b] - off[a], dtype=np.uint8)
            ctx.emit_host(a, b, part)
            assert np.array_equal(part, exp_img[off[a]:off[b]])
        small = np.empty(10, dtype=np.uint8)
        with pytest.raises(_native.Gm2Error) as ei:
            ctx.emit_host(0, S, small)
        assert ei.value.code == _native.ERR_CAPACITY


def test_engine_drain_delivers_every_byte_once():
    g = synth.make_genome(100_000, 120, 41)
    rec_lists = synth.make_gene_lists(g, 23, 0.5, seed=5, extra_names=10)
    starts, ends = g.starts_ends()
    table = engine.GeneTable(g.gene_names(), starts, ends)
    eng = engine.MinimizerEngine(seq=g.seq, table=table)
    try:
        lengths = eng.plan_lists(rec_lists)
        got = []
        total = eng.drain(lambda a, b, v: got.append((a, b, v.tobytes())), max_bytes=300_000)
        img = b"".join(x[2] for x in got)
        assert total == len(img)
        assert [x[0] for x in got][0] == 0 and got[-1][1] == 23
        keep = np.stack([mo.keep_vector(table.names, l) for l in rec_lists])
        exp_len, _, exp_img = _oracle_image(g.seq, starts, ends, synth.pack_keep_rows(keep))
        assert np.array_equal(lengths, exp_len)
        assert img == exp_img.tobytes()
    finally:
        eng.close()


def test_device_image_hashes_match_oracle_k12_shape():
    """K-12-shaped genome, device-resident image, compared record by record through the
    device-side range hash (no 100+ MB host copies) and byte-for-byte on a subset."""
    import torch
    g = synth.make_genome(seed=1)
    starts, ends = g.starts_ends()
    S = 48
    rows = synth.pack_keep_rows(synth.random_keep_bool(len(g.genes), S, 0.5, seed=2))
    exp_len, exp_hash, _ = c_oracle.batch(g.seq, starts, ends, rows)
    with _native.Context(0) as ctx:
        ctx.set_reference(g.seq, starts, ends)
        ctx.load_keep_host(rows)
        ctx.plan(0)
        assert np.array_equal(ctx.lengths(), exp_len)
        off = ctx.record_offsets()
        img = torch.empty(int(off[-1]), dtype=torch.uint8, device="cuda:0")
        ctx.emit_dev(0, S, img.data_ptr(), img.numel())
        ctx.sync()
        got = ctx.diag_range_hashes(img.data_ptr(), img.numel(), off)
        assert np.array_equal(got, exp_hash)
        # byte-exact on three records
        for s in (0, 17, S - 1):
            _, _, one = c_oracle.batch(g.seq, starts, ends, rows[s:s + 1], first_idx=s, want_image=True)
            assert np.array_equal(img[int(off[s]):int(off[s + 1])].cpu().numpy(), one)
        # size-independent properties: every record ends in '\n', starts with '>', no byte outside ACGT/header
        host = img[:int(off[1])].cpu().numpy()
        assert host[0] == ord(">") and host[-1] == 10
        # all-kept and none-kept rows bracket every length
        allk = synth.pack_keep_rows(np.ones((1, len(g.genes)), bool))
        none = synth.pack_keep_rows(np.zeros((1, len(g.genes)), bool))
        ctx.load_keep_host(np.concatenate([allk, none]))
        ctx.plan(0)
        L = ctx.lengths()
        assert L[0] == g.G and L[1] <= exp_len.min()


def test_header_prefix_and_store_policy_variants():
    g = synth.make_genome(70_000, 80, 51)
    starts, ends = g.starts_ends()
    S = 9
    rows = synth.pack_keep_rows(synth.random_keep_bool(len(g.genes), S, 0.4, seed=3))
    _, _, exp_img = c_oracle.batch(g.seq, starts, ends, rows, prefix="Other_prefix|", want_image=True)
    for policy in (0, 1):
        with _native.Context(0) as ctx:
            ctx.configure(_native.CFG_STORE_POLICY, policy)
            ctx.set_header_prefix("Other_prefix|")
            ctx.set_reference(g.seq, starts, ends)
            ctx.load_keep_host(rows)
            ctx.plan(0)
            assert np.array_equal(_gpu_image(ctx, S), exp_img)


def test_call_order_errors_are_reported():
    with _native.Context(0) as ctx:
        with pytest.raises(_native.Gm2Error) as ei:
            ctx.plan(0)
        assert ei.value.code == _native.ERR_STATE
        with pytest.raises(_native.Gm2Error):
            ctx.configure(_native.CFG_TILE_BYTES, 1000)


def test_dense_tiny_genes_exercise_table_flush_and_global_slot_path():
    """Thousands of tiny genes inside one tile: more slots than the shared slot table holds
    (global fallback), run tables flushed many times, runs shorter than one 16-byte vector."""
    rng = np.random.default_rng(123)
    G = 70_000
    F = 6_000
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, G)]        # ACGT: both transports are exercised
    starts = np.sort(rng.integers(0, G - 12, F)).astype(np.int64)
    ends = starts + rng.integers(1, 12, F)
    S = 21
    rows = synth.pack_keep_rows(rng.random((S, F)) < np.linspace(0.05, 0.95, S)[:, None])
    exp_len, _, exp_img = _oracle_image(seq, starts, ends, rows)
    for rt_cap, warps, flat, mode in ((32, 8, 0, 2), (64, 4, 640, 1), (1024, 2, 1 << 20, 1), (32, 8, 1 << 20, 1), (1024, 1, 0, 1),
                                      (64, 4, 640, 2), (1024, 2, 1 << 20, 2), (32, 8, 1 << 20, 2)):
        with _native.Context(0) as ctx:
            ctx.configure(_native.CFG_RUN_TABLE, rt_cap)
            ctx.configure(_native.CFG_EMIT_WARPS, warps)
            ctx.configure(_native.CFG_FLAT_RUN_BYTES, flat)
            ctx.configure(_native.CFG_FLAT_MODE, mode)
            ctx.set_reference(seq, starts, ends)
            assert ctx.query(_native.Q_NUM_SLOTS) > 4096
            ctx.load_keep_host(rows)
            ctx.plan(0)
            assert np.array_equal(ctx.lengths(), exp_len)
            assert np.array_equal(_gpu_image(ctx, S), exp_img)


@pytest.mark.parametrize("mode", [1, 2])
def test_short_run_visits_match_the_oracle(mode):
    """Low gene retention on gene-shaped genomes: most kept runs are intergenic gaps of a few bytes to a few
    hundred bytes, the case the flat emit forms exist for (GM2_CFG_FLAT_MODE).  Gaps are exponential with a
    small mean so that vectors holding two, three and more run starts, runs shorter than one vector, visits
    that start / end inside a vector and visits with no run at all occur; retention ramps from 0 to 0.6 over
    the samples; default tile and a small one; both launch forms; always-flat and the default threshold."""
    rng = np.random.default_rng(77 + mode)
    for G, F, gap_mean, tile in ((400_000, 900, 40, 0), (300_000, 260, 120, 0), (90_000, 700, 6, 8192), (120_000, 40, 300, 16384)):
        lens = np.maximum(rng.lognormal(np.log(G / F * 0.8), 0.5, F).astype(np.int64), 1)
        gaps = rng.exponential(gap_mean, F).astype(np.int64)
        gaps[rng.random(F) < 0.15] = 0                       # abutting genes
        starts = np.cumsum(gaps + np.concatenate([[0], lens[:-1]]))
        ends = np.minimum(starts + lens, G)
        starts = np.minimum(starts, G)
        seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, G)]
        S = 40
        keep = rng.random((S, F)) < np.linspace(0.0, 0.6, S)[:, None]
        rows = synth.pack_keep_rows(keep)
        exp_len, _, exp_img = _oracle_image(seq, starts, ends, rows, first_idx=7)
        for occ, frb, pack in ((4, 640, 1), (3, 1 << 20, 1), (0, 200, 1)) + (((3, 1 << 20, 2), (4, 640, 2)) if mode == 2 else ()):
            with _native.Context(0) as ctx:
                if tile:
                    ctx.configure(_native.CFG_TILE_BYTES, tile)
                ctx.configure(_native.CFG_PACKING, pack)          # 2: the whole-visit form from the two-bit staged tile
                ctx.configure(_native.CFG_FLAT_MODE, mode)
                ctx.configure(_native.CFG_EMIT_OCCUPANCY, occ)
                ctx.configure(_native.CFG_FLAT_RUN_BYTES, frb)
                ctx.set_reference(seq, starts, ends)
                ctx.load_keep_host(rows)
                ctx.plan(7)
                assert np.array_equal(ctx.lengths(), exp_len)
                assert np.array_equal(_gpu_image(ctx, S), exp_img), (G, F, tile, occ, frb, pack)


def test_many_genes_cover_one_segment():
    """Segments covered by far more than two genes (overflow cover lists)."""
    rng = np.random.default_rng(9)
    G = 20_000
    F = 40
    seq = rng.integers(65, 91, G, dtype=np.uint8)
    starts = rng.integers(0, 2_000, F).astype(np.int64)
    ends = G - rng.integers(0, 2_000, F).astype(np.int64)       # all 40 genes overlap in the middle
    S = 33
    keep = rng.random((S, F)) < 0.9
    keep[0] = True
    keep[1] = False
    keep[2] = True
    keep[2, 17] = False
    rows = synth.pack_keep_rows(keep)
    exp_len, _, exp_img = _oracle_image(seq, starts, ends, rows)
    with _native.Context(0) as ctx:
        ctx.set_reference(seq, starts, ends)
        ctx.load_keep_host(rows)
        ctx.plan(0)
        assert np.array_equal(ctx.lengths(), exp_len)
        assert exp_len[0] == G
        assert np.array_equal(_gpu_image(ctx, S), exp_img)


def test_full_size_c2_properties():
    """BASELINE config 2 at full size (10,000 samples x K-12 shape, 26 GB image, device-resident):
    every length against an independent numpy computation, structural bytes of every record,
    64 records hashed on the device against the C oracle, idempotence of a second emit."""
    import torch
    g = synth.make_genome(seed=1)
    starts, ends = g.starts_ends()
    F = len(g.genes)
    S = 10_000
    keep = synth.random_keep_bool(F, S, 0.5, seed=2)
    rows = synth.pack_keep_rows(keep)
    # independent lengths: elementary segments, kept iff every covering gene is kept
    bp = np.unique(np.concatenate([[0, g.G], starts, ends]))
    seg_len = np.diff(bp)
    removed_cover = np.zeros((S, len(seg_len)), dtype=bool)
    for gi in range(F):
        a, b = np.searchsorted(bp, starts[gi]), np.searchsorted(bp, ends[gi])
        if b > a:
            removed_cover[:, a:b] |= ~keep[:, gi:gi + 1]
    exp_len = (~removed_cover).dot(seg_len.astype(np.int64))
    with _native.Context(0) as ctx:
        ctx.set_reference(g.seq, starts, ends)
        ctx.load_keep_host(rows)
        ctx.plan(0)
        assert np.array_equal(ctx.lengths(), exp_len)
        off = ctx.record_offsets()
        hdr = np.asarray([1 + len("Minimized_E_coli_K12_MG1655_") + len(str(i + 1)) + 1 for i in range(S)])
        assert np.array_equal(np.diff(off), hdr + exp_len + 1)
        img = torch.empty(int(off[-1]), dtype=torch.uint8, device="cuda:0")
        ctx.emit_dev(0, S, img.data_ptr(), img.numel())
        ctx.sync()
        offs = torch.from_numpy(off).to("cuda:0")
        assert bool((img[offs[:-1]] == ord(">")).all())                       # every record starts with '>'
        assert bool((img[offs[1:] - 1] == 10).all())                          # ... and ends with '\n'
        assert bool((img[offs[:-1] + torch.from_numpy(hdr).to("cuda:0") - 1] == 10).all())   # header newline
        # exactly two newlines per record, no byte outside "ACGT" and the header alphabet elsewhere
        step = 1 << 30
        newlines = sum(int(torch.count_nonzero(img[a:a + step] == 10)) for a in range(0, img.numel(), step))
        assert newlines == 2 * S
        pick = np.unique(np.linspace(0, S - 1, 64).astype(int))
        got = ctx.diag_range_hashes(img.data_ptr(), img.numel(), off)
        for s in pick:
            _, h, _ = c_oracle.batch(g.seq, starts, ends, rows[s:s + 1], first_idx=int(s))
            assert int(h[0]) == int(got[s]), s
        # idempotence: a second emit into a fresh buffer gives the same hashes for every record
        img.fill_(0)
        torch.cuda.synchronize()            # fill_ ran on torch's stream, the context has its own
        ctx.emit_dev(0, S, img.data_ptr(), img.numel())
        ctx.sync()
        assert np.array_equal(ctx.diag_range_hashes(img.data_ptr(), img.numel(), off), got)


def test_class_facade_reuses_one_engine_per_record(golden_paths, golden):
    """The reference builds one GenomeMinimiser per sample on a shared record; ours must not
    re-upload the genome each time, and must still give every sample's exact sequence."""
    import gc
    from genome_minimizer_2_b200 import minimizer_2 as m2
    gb, _ = golden_paths
    rec = genbank.read_genbank(gb)
    ref = genbank_reader.read_genbank(gb)
    seqs = [m2.GenomeMinimiser(record=rec, needed_genes_list=l, idx=i).reduced_genome_str
            for i, l in enumerate(golden["lists"][:6])]
    assert id(rec) in m2._ENGINES and len([k for k in m2._ENGINES if k == id(rec)]) == 1
    if "sequences" in golden:
        assert seqs == golden["sequences"][:6]
    else:
        assert [len(s) for s in seqs] == golden["sequence_lengths"][:6]
    key = id(rec)
    del rec
    gc.collect()
    assert key not in m2._ENGINES                      # engine released with the record


def test_duplicate_stats_match_the_reference_definition():
    """Device-side sequence hashes reproduce check_sequence_duplicates' numbers (minimizer_2.py:273-303)."""
    g = synth.make_genome(60_000, 40, 61, nested=2)
    starts, ends = g.starts_ends()
    table = engine.GeneTable(g.gene_names(), starts, ends)
    base = synth.make_gene_lists(g, 7, 0.5, seed=9, extra_names=3)
    lists = base + [base[0], base[3], list(reversed(base[0])), base[0] + ["unknown_name"], [], []]
    eng = engine.MinimizerEngine(seq=g.seq, table=table)
    try:
        eng.plan_lists(lists)
        st = eng.duplicate_stats()
        # reference definition on the actual strings (oracle)
        seqs = {}
        for i, l in enumerate(lists):
            keep = mo.keep_vector(table.names, l)
            seqs[f"id{i}"] = mo.minimize_numpy(g.seq, starts, ends, keep).tobytes()
        from collections import defaultdict
        grp = defaultdict(list)
        for k, v in seqs.items():
            grp[v].append(k)
        dups = {k: v for k, v in grp.items() if len(v) > 1}
        assert st["total_sequences"] == len(lists)
        assert st["unique_sequences"] == len(grp)
        assert st["duplicate_groups"] == len(dups)
        assert st["duplicated_sequences"] == sum(len(v) for v in dups.values())
        assert st["unique_only_sequences"] == sum(1 for v in grp.values() if len(v) == 1)
        assert st["compression_ratio"] == len(grp) / len(lists)
        # hashes are those of the bare sequences
        h = eng.ctx.sequence_hashes()
        assert [int(x) for x in h] == [mo.range_hash(v) for v in seqs.values()]
    finally:
        eng.close()


def test_huge_record_ids_render_all_digits():
    """Headers with 12-19 digit ids (a shard far into a huge job); ids cross a power of ten in the batch."""
    g = synth.make_genome(9_000, 12, 71, nested=1)
    starts, ends = g.starts_ends()
    S = 25
    rows = synth.pack_keep_rows(synth.random_keep_bool(len(g.genes), S, 0.5, seed=7))
    for first in (999_999_999_990, 9_223_372_036_854_775_000):
        exp_len, _, exp_img = _oracle_image(g.seq, starts, ends, rows, first_idx=first)
        with _native.Context(0) as ctx:
            ctx.set_reference(g.seq, starts, ends)
            ctx.load_keep_host(rows)
            ctx.plan(first)
            assert np.array_equal(_gpu_image(ctx, S), exp_img)


def test_two_bit_packing_is_byte_identical_and_rejects_iupac():
    """GM2_CFG_PACKING=2 (2 bits per base in shared memory) must give the same image; a reference
    with any non-ACGT letter is refused for that packing."""
    rng = np.random.default_rng(5)
    for G, F, seed, kw in [(200_000, 180, 81, dict(nested=5)), (65_537, 900, 82, dict(overlap_frac=0.5, nested=20)),
                           (49_152 * 2, 40, 83, {}), (33, 2, 84, dict(nested=0))]:
        g = synth.make_genome(G, F, seed, **kw)
        starts, ends = g.starts_ends()
        S = 29
        rows = synth.pack_keep_rows(synth.random_keep_bool(len(g.genes), S, np.linspace(0.05, 0.95, S), seed=seed))
        exp_len, _, exp_img = _oracle_image(g.seq, starts, ends, rows)
        for policy in (0, 1):
            with _native.Context(0) as ctx:
                ctx.configure(_native.CFG_PACKING, 2)
                ctx.configure(_native.CFG_STORE_POLICY, policy)
                ctx.set_reference(g.seq, starts, ends)
                assert ctx.query(_native.Q_PACKING) == 2
                ctx.load_keep_host(rows)
                ctx.plan(0)
                assert np.array_equal(ctx.lengths(), exp_len)
                assert np.array_equal(_gpu_image(ctx, S), exp_img)
    g = synth.make_genome(5_000, 10, 85, iupac_runs=3)
    with _native.Context(0) as ctx:
        ctx.configure(_native.CFG_PACKING, 2)
        with pytest.raises(_native.Gm2Error) as ei:
            ctx.set_reference(g.seq, *g.starts_ends())
        assert ei.value.code == _native.ERR_INVALID and "ACGT" in str(ei.value)


def test_fuzz_kernel_configurations():
    """Random genomes x random kernel configurations (tile size, warps, run-table size, packing,
    store policy, CTA order, batch) against the C oracle, byte for byte."""
    rng = np.random.default_rng(int(os.environ.get("GM2_FUZZ_SEED", "20261018")))
    for trial in range(int(os.environ.get("GM2_FUZZ_TRIALS", "48"))):          # soak runs: GM2_FUZZ_TRIALS=1000
        G = int(rng.choice([1, 31, 32, 33, 4095, 4096, 4097, int(rng.integers(5_000, 140_000))]))
        F = int(rng.integers(0, 400)) if G > 100 else int(rng.integers(0, 4))
        seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, G)]
        style = trial % 4
        if style == 0:      # gene-like
            starts = np.sort(rng.integers(0, max(G - 1, 1), F)).astype(np.int64)
            ends = starts + rng.integers(1, max(G // max(F, 1) * 2, 2), F)
        elif style == 1:    # tiny runs everywhere
            starts = np.sort(rng.integers(0, max(G - 1, 1), F)).astype(np.int64)
            ends = starts + rng.integers(0, 40, F)
        elif style == 2:    # heavy overlap / nesting
            starts = rng.integers(0, max(G // 2, 1), F).astype(np.int64)
            ends = starts + rng.integers(0, max(G // 2, 2), F)
        else:               # degenerate and out-of-range intervals mixed in
            starts = rng.integers(-100, G + 100, F).astype(np.int64)
            ends = starts + rng.integers(-50, max(G // 10, 5), F)
        S = int(rng.integers(1, 40))
        p = rng.random((S, 1)) if trial % 3 else np.full((S, 1), rng.choice([0.02, 0.98]))
        rows = synth.pack_keep_rows(rng.random((S, F)) < p) if F else np.zeros((S, 0), dtype=np.uint32)
        first = int(rng.choice([0, 9, 99, 12345, 99_999_990]))
        exp_len, _, exp_img = _oracle_image(seq, starts, ends, rows, first_idx=first)
        cfg = {
            _native.CFG_TILE_BYTES: int(rng.choice([0, 4096, 8192, 16384, 32768, 49152, 65536, 131072])),
            _native.CFG_EMIT_WARPS: int(rng.choice([1, 2, 4, 8])),
            _native.CFG_RUN_TABLE: int(rng.choice([32, 34, 64, 128])),
            _native.CFG_PACKING: int(rng.choice([1, 2])),
            _native.CFG_STORE_POLICY: int(rng.choice([0, 1])),
            _native.CFG_ORDER: int(rng.choice([0, 1])),
            _native.CFG_EMIT_BATCH: int(rng.choice([0, 1, 3, 16])),
            _native.CFG_FLAT_RUN_BYTES: int(rng.choice([0, 64, 640, 1 << 20])),
            _native.CFG_EMIT_OCCUPANCY: int(rng.choice([0, 3, 4])),
            _native.CFG_FLAT_MODE: int(rng.choice([0, 1, 2])),
        }
        with _native.Context(0) as ctx:
            for k, v in cfg.items():
                ctx.configure(k, v)
            ctx.set_reference(seq, starts, ends)
            ctx.load_keep_host(rows.reshape(S, -1) if F else np.zeros((S, 0), dtype=np.uint32))
            ctx.plan(first)
            assert np.array_equal(ctx.lengths(), exp_len), (trial, cfg)
            assert np.array_equal(_gpu_image(ctx, S), exp_img), (trial, G, F, S, cfg)


def test_emit_occupancy_follows_the_kept_fraction_and_never_changes_the_bytes():
    """GM2_CFG_EMIT_OCCUPANCY: auto picks the 72-register / 3-CTA build, and the 4-CTA one only for small tiles
    when the plan kept little of the genome; forcing either gives the same image.  GM2_CFG_TILE_BYTES = 0 (the
    default) sizes the tile from the gene density."""
    g = synth.make_genome(300_000, 280, seed=61)
    starts, ends = g.starts_ends()
    rng = np.random.default_rng(61)
    S = 24
    for tile, p_keep, want in ((0, 0.05, 3), (24576, 0.05, 4), (24576, 0.95, 3)):
        rows = synth.pack_keep_rows(rng.random((S, len(starts))) < p_keep)
        exp_len, _, exp_img = _oracle_image(g.seq, starts, ends, rows, first_idx=0)
        images = {}
        for occ in (0, 3, 4):
            with _native.Context(0) as ctx:
                ctx.configure(_native.CFG_EMIT_OCCUPANCY, occ)
                ctx.configure(_native.CFG_TILE_BYTES, tile)
                ctx.set_reference(g.seq, starts, ends)
                # 280 genes on 300 kbp: 34 genes' worth of bases is 36 KB
                assert ctx.query(_native.Q_TILE_BYTES) == (tile or 36864)
                ctx.load_keep_host(rows)
                ctx.plan(0)
                assert np.array_equal(ctx.lengths(), exp_len)
                images[occ] = _gpu_image(ctx, S)
                got = ctx.query(_native.Q_LAST_EMIT_CTAS)
                assert got == (want if occ == 0 else occ), (p_keep, occ, got)
            assert np.array_equal(images[occ], exp_img), (p_keep, occ)
    with _native.Context(0) as ctx:
        with pytest.raises(_native.Gm2Error):
            ctx.configure(_native.CFG_EMIT_OCCUPANCY, 5)


def test_wire_format_selection_and_threads():
    """GM2_CFG_WIRE / GM2_CFG_HOST_THREADS: what gets chosen, what is refused, and that the number of
    expansion threads does not change the bytes."""
    rng = np.random.default_rng(77)
    G, F, S = 150_000, 120, 37
    starts = np.sort(rng.integers(0, G - 10, F)).astype(np.int64)
    ends = starts + rng.integers(1, 2_500, F)
    rows = synth.pack_keep_rows(rng.random((S, F)) < 0.5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, G)]
    iupac = acgt.copy()
    iupac[1000] = ord("N")
    for seq, is_acgt in ((acgt, True), (iupac, False)):
        _, _, exp_img = _oracle_image(seq, starts, ends, rows, first_idx=5)
        with _native.Context(0) as ctx:
            ctx.configure(_native.CFG_TILE_BYTES, 16384)
            ctx.set_reference(seq, starts, ends)
            ctx.load_keep_host(rows)
            ctx.plan(5)
            n = ctx.image_bytes(0, S)
            for wire, threads, want_wire in ((0, 8, 2 if is_acgt else 1), (0, 4, 1), (0, 1, 1), (1, 8, 1), (2, 1, 2), (2, 2, 2), (2, 5, 2), (2, 64, 2)):
                ctx.configure(_native.CFG_WIRE, wire)
                ctx.configure(_native.CFG_HOST_THREADS, threads)
                out = np.zeros(n, dtype=np.uint8)
                if wire == 2 and not is_acgt:
                    with pytest.raises(_native.Gm2Error) as ei:
                        ctx.emit_host(0, S, out)
                    assert ei.value.code == _native.ERR_STATE
                    continue
                for chunk in (0, 1, 200_000):                       # default, one record per chunk, a few per chunk
                    out[:] = 0
                    ctx.emit_host(0, S, out, chunk_bytes=chunk)
                    assert ctx.query(_native.Q_LAST_WIRE) == want_wire
                    assert np.array_equal(out, exp_img), (is_acgt, wire, threads, chunk)
                    d2h = ctx.query(_native.Q_LAST_D2H_BYTES)
                    assert d2h == n if want_wire == 1 else 0 < d2h < n // 2
            # a sub-range, two-bit
            ctx.configure(_native.CFG_WIRE, 0)
            ctx.configure(_native.CFG_HOST_THREADS, 3)
            off = ctx.record_offsets()
            part = np.zeros(int(off[30] - off[11]), dtype=np.uint8)
            ctx.emit_host(11, 30, part)
            assert np.array_equal(part, exp_img[int(off[11]):int(off[30])])


def test_integration_md_ctypes_stub_runs_as_written(tmp_path):
    """The ctypes stub printed in INTEGRATION.md is executed verbatim against a golden case."""
    from conftest import ROOT, load_golden
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = text.split("ctypes stub", 1)[1].split("```python", 1)[1].split("```", 1)[0]
    case = load_golden("kat_appB")
    gb = tmp_path / "g.gb"
    gb.write_text(case["genbank"])
    cwd = os.getcwd()
    os.chdir(ROOT)                                   # the stub loads the library by its in-tree relative path
    try:
        env = {"record": genbank.read_genbank(str(gb)), "all_lists": case["lists"]}
        exec(compile(block, "INTEGRATION.md", "exec"), env)
    finally:
        os.chdir(cwd)
    expected = "".join(case["single_file"].split("\n", 3)[3])      # records without the 3-line preamble
    assert env["image"].tobytes().decode() == expected
