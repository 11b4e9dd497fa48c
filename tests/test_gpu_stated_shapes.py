"""-m gpu parity at the shapes BASELINE.json states for configs 4 and 5 ("next round" 3).

config 4: 12 Mbp reference, 10,000 gene features with overlapping / nested / antisense / join genes,
          gene retention swept 0.1 ... 0.9 (minimizer_2.py:50-101 on that genome);
config 5: decoder output of the v0 model's width (55,039 columns, utils/extras.py:192-203,
          training/model.py:79-107) -> threshold -> column -> gene keep mask (binary_converter.py:49-64,
          :91-110) on the K-12-shaped reference.
Everything goes through the C-ABI; the checker is the oracle (never the product).  Bar: bit-exact.
"""
from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c_oracle, converter_oracle as co, minimizer_oracle as mo  # noqa: E402  (checker only)

from genome_minimizer_2_b200 import _native, engine, synth  # noqa: E402


def _config4_genome():
    return synth.make_genome(12_000_000, 10_000, seed=4, overlap_frac=0.3, nested=200, join_genes=50,
                             dup_name_frac=0.006, nameless_frac=0.003, name="SYNTH_12M")


def test_config4_genome_retention_sweep_both_transports():
    """32 samples on the 12 Mbp / 10k-feature overlapping genome, gene retention 0.1 ... 0.9: every length and
    every record's hash against the C oracle (device-resident image), three records byte-for-byte, and the
    host path by both transports (image bytes / two bits per base) on a sub-range."""
    import torch
    g = _config4_genome()
    starts, ends = g.starts_ends()
    F = len(g.genes)
    S = 32
    p = np.linspace(0.1, 0.9, S)
    keep = synth.random_keep_bool(F, S, p, seed=41)
    rows = synth.pack_keep_rows(keep)
    exp_len, exp_hash, _ = c_oracle.batch(g.seq, starts, ends, rows, first_idx=999_990)
    with _native.Context(0) as ctx:
        ctx.set_reference(g.seq, starts, ends)
        ctx.load_keep_host(rows)
        ctx.plan(999_990)                                  # record ids cross 6 -> 7 digits inside the batch
        assert np.array_equal(ctx.lengths(), exp_len)
        off = ctx.record_offsets()
        img = torch.empty(int(off[-1]), dtype=torch.uint8, device="cuda:0")
        for occupancy in (3, 4):                           # both launch forms of k_emit write the same bytes
            ctx.configure(_native.CFG_EMIT_OCCUPANCY, occupancy)
            img.fill_(0)
            torch.cuda.synchronize()
            ctx.emit_dev(0, S, img.data_ptr(), img.numel())
            ctx.sync()
            got = ctx.diag_range_hashes(img.data_ptr(), img.numel(), off)
            assert np.array_equal(got, exp_hash), occupancy
        ctx.configure(_native.CFG_EMIT_OCCUPANCY, 0)
        for s in (0, 15, S - 1):
            _, _, one = c_oracle.batch(g.seq, starts, ends, rows[s:s + 1], first_idx=999_990 + s, want_image=True)
            assert np.array_equal(img[int(off[s]):int(off[s + 1])].cpu().numpy(), one), s
        # host path, both transports, low-retention end and high-retention end
        for a, b in ((0, 3), (S - 2, S)):
            _, _, exp_img = c_oracle.batch(g.seq, starts, ends, rows[a:b], first_idx=999_990 + a, want_image=True)
            n = ctx.image_bytes(a, b)
            assert n == exp_img.size
            for wire in (1, 2):
                ctx.configure(_native.CFG_WIRE, wire)
                out = np.full(n, 0x2a, dtype=np.uint8)
                ctx.emit_host(a, b, out)
                assert ctx.query(_native.Q_LAST_WIRE) == wire
                assert np.array_equal(out, exp_img), (a, b, wire)
            ctx.configure(_native.CFG_WIRE, 0)


def _config5_columns(table, V, rng):
    """V column labels shaped like the presence/absence table's: ~4.4k of them are names of genes of the
    reference (every distinct name once), a few of those are repeated further down (the reference keeps the
    first and drops the rest, binary_converter.py:29-36), the rest match nothing."""
    gene_names = sorted({n for n in table.names if n})
    cols = [f"group_{i}" for i in range(V + 40)]
    slots = rng.choice(V, size=len(gene_names), replace=False)
    for nm, j in zip(gene_names, slots):
        cols[j] = nm
    dup_at = rng.choice(np.setdiff1d(np.arange(V), slots), size=40, replace=False)
    for k, j in enumerate(sorted(dup_at)):
        cols[j] = gene_names[(7 * k) % len(gene_names)]           # a second column with a gene's name
    return cols, gene_names


def test_config5_width_55039_keep_rows_counts_and_records():
    """k_keep_from_probs at the v0 decoder's width: V = 55,039 de-duplicated columns, ~4.4k of them mapped to
    genes, duplicate column labels, forced essentials (present and absent from the columns), and values at
    exactly 0.5, nextafter(0.5, 1) and nextafter(0.5, 0) (strict > 0.5, utils/extras.py:200): keep rows, the
    list lengths the reference prints, lengths and record hashes against the oracles."""
    import torch
    g = synth.make_genome(seed=1)
    starts, ends = g.starts_ends()
    table = engine.GeneTable(g.gene_names(), starts, ends)
    rng = np.random.default_rng(55)
    raw_cols, gene_names = _config5_columns(table, 55_039, rng)
    V = len(co.dedup_columns(raw_cols))
    assert V == 55_039 and len(raw_cols) == V + 40
    essential = [gene_names[i] for i in range(0, len(gene_names), 97)] + ["not_a_column_1", "not_a_column_2"]
    S = 24
    decoded = rng.random((S, V), dtype=np.float32)
    # column-wise retention ramp so that rows differ a lot, plus the three edge values sprinkled everywhere
    decoded = (decoded * np.linspace(0.6, 1.4, S, dtype=np.float32)[:, None]).astype(np.float32)
    edge = rng.integers(0, 3, size=(S, V))
    pick = rng.random((S, V)) < 0.05
    half = np.float32(0.5)
    vals = np.asarray([half, np.nextafter(half, np.float32(1)), np.nextafter(half, np.float32(0))], dtype=np.float32)
    decoded[pick] = vals[edge[pick]]
    decoded[0, :] = 0.5                                             # nothing present (strict >)
    decoded[1, :] = vals[1]                                         # everything present
    # the reference chain: threshold (extras.py:199-201) -> names (binary_converter.py:49-64) -> essentials (:91-110)
    binary = co.threshold_samples(decoded)
    lists = co.add_essentials(co.masks_to_gene_lists(binary, raw_cols), essential)
    keep = np.stack([mo.keep_vector(table.names, l) for l in lists])
    rows = synth.pack_keep_rows(keep)
    exp_len, exp_hash, _ = c_oracle.batch(g.seq, starts, ends, rows)
    eng = engine.MinimizerEngine(seq=g.seq, table=table, device=0)
    try:
        space = engine.ColumnSpace(eng.table, raw_cols, essential)
        assert space.V == V and space.duplicates_dropped == len(raw_cols) - V
        dev = torch.from_numpy(decoded).to("cuda:0")
        lengths, counts = engine.plan_from_probabilities(eng, space, dev)
        assert counts.tolist() == [len(l) for l in lists]            # "[i/N] genes present:" of the reference
        assert np.array_equal(eng.ctx.keep_rows(), rows)
        assert np.array_equal(lengths, exp_len)
        off = eng.ctx.record_offsets()
        img = torch.empty(int(off[-1]), dtype=torch.uint8, device="cuda:0")
        eng.ctx.emit_dev(0, S, img.data_ptr(), img.numel())
        eng.ctx.sync()
        assert np.array_equal(eng.ctx.diag_range_hashes(img.data_ptr(), img.numel(), off), exp_hash)
        # a padded (strided) matrix at the same width
        wide = torch.zeros(S, V + 9, dtype=torch.float32, device="cuda:0")
        wide[:, 3:3 + V] = dev
        l2, n2 = engine.plan_from_probabilities(eng, space, wide[:, 3:3 + V])
        assert np.array_equal(l2, exp_len) and np.array_equal(n2, counts)
        # name lists planned on the SAME engine afterwards use the gene table's own id space again
        l3 = eng.plan_lists(lists[2:6])
        assert np.array_equal(l3, exp_len[2:6])
        assert np.array_equal(eng.ctx.keep_rows(), rows[2:6])
        # ... and the column space can come back
        l4, _ = engine.plan_from_probabilities(eng, space, dev[5:9])
        assert np.array_equal(l4, exp_len[5:9])
    finally:
        eng.close()


def test_context_pair_pipeline_is_byte_identical_and_ordered():
    """engine.ContextPair / gm2_order_after: six chunks of device-resident id lists alternate between two contexts
    on two streams (plan of chunk i+1 under the emit of chunk i); each chunk's image is hashed on the device right
    after its emit and must match the oracle, although the ring holds only two chunks."""
    import torch
    g = synth.make_genome(400_000, 380, seed=71, nested=6, overlap_frac=0.3)
    starts, ends = g.starts_ends()
    table = engine.GeneTable(g.gene_names(), starts, ends)
    rng = np.random.default_rng(71)
    name_ids = np.asarray([table.name_to_id[n] for n in table.names], dtype=np.int64)
    nchunks, Sc = 6, 40
    chunks, expect, keep_dev = [], [], []
    for i in range(nchunks):
        keep = rng.random((Sc, table.V)) < (0.15 + 0.14 * i)
        ids = [np.flatnonzero(row).astype(np.int32) for row in keep]
        off = np.zeros(Sc + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(x) for x in ids])
        flat = np.concatenate(ids) if off[-1] else np.zeros(1, dtype=np.int32)
        d_ids, d_off = torch.from_numpy(flat).to("cuda:0"), torch.from_numpy(off).to("cuda:0")
        keep_dev.append((d_ids, d_off))
        chunks.append((d_ids.data_ptr(), d_off.data_ptr(), Sc, int(off[-1]), i * Sc))
        gene_keep = keep[:, name_ids]                                        # genes sharing a name are kept together
        expect.append(c_oracle.batch(g.seq, starts, ends, synth.pack_keep_rows(gene_keep), first_idx=i * Sc))
    cap = max(int((e[0] + 64).sum()) for e in expect)
    ring = [torch.empty(cap, dtype=torch.uint8, device="cuda:0") for _ in range(2)]
    pair = engine.ContextPair(g.seq, table, device=0)
    try:
        got = [None] * nchunks
        offs = [np.concatenate([[0], np.cumsum(1 + len(engine.SEQ_ID_PREFIX) + np.char.str_len((np.arange(i * Sc, (i + 1) * Sc) + 1).astype(str)) + 1 + expect[i][0] + 1)]).astype(np.int64)
                for i in range(nchunks)]

        def after_emit(i, ctx):
            # enqueued on the chunk's own stream, right behind its emit; returns host data => that stream is synchronised,
            # while the OTHER context's plan / emit of the next chunk keeps running
            got[i] = ctx.diag_range_hashes(ring[i & 1].data_ptr(), cap, offs[i])

        assert pair.run_chunks(chunks, [(r.data_ptr(), cap) for r in ring], after_emit) == nchunks
        for c in pair.ctx:
            c.sync()
        for i in range(nchunks):
            assert np.array_equal(got[i], expect[i][1]), i
        # the same schedule without host round trips in between: only the last two chunks remain in the ring
        pair.run_chunks(chunks, [(r.data_ptr(), cap) for r in ring])
        for c in pair.ctx:
            c.sync()
        for i in (nchunks - 2, nchunks - 1):
            assert np.array_equal(pair.ctx[i & 1].diag_range_hashes(ring[i & 1].data_ptr(), cap, offs[i]), expect[i][1]), i
            assert np.array_equal(pair.ctx[i & 1].lengths(), expect[i][0])
    finally:
        pair.close()
    with _native.Context(0) as a:
        assert a.order_after(a) is None                                      # ordering a context after itself is a no-op
