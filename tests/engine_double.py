"""Test double for `engine.MinimizerEngine` where there is no GPU: the plan_lists / drain /
ctx.record_offsets / close surface the batch entry points use, computed by the oracle.  Used by the
CPU-only tests of the host wiring (CLI drop-in, torchrun-aware entry functions); the product's own
engine is exercised by the `-m gpu` tests."""
from __future__ import annotations

import numpy as np

from oracle import minimizer_oracle as mo            # checker only (tests)
from genome_minimizer_2_b200 import engine


class OracleEngine:
    def __init__(self, record):
        ref = record if isinstance(record, engine.ReferenceGenome) else engine.ReferenceGenome.from_record(record)
        self.ref, self.images, self.ctx = ref, [], self
        self.table = ref.table

    def plan_lists(self, all_lists, first_idx=0):
        t = self.ref.table
        if isinstance(all_lists, engine.TokenizedLists):
            keeps = []
            for i in range(len(all_lists)):
                keep = np.zeros(t.F, dtype=bool)
                for v in all_lists.ids[all_lists.off[i]:all_lists.off[i + 1]]:
                    keep[t.id2gene_idx[t.id2gene_off[v]:t.id2gene_off[v + 1]]] = True
                keeps.append(keep)
        else:
            keeps = [mo.keep_vector(t.names, needed) for needed in all_lists]
        seqs = [mo.minimize_numpy(self.ref.seq, t.starts, t.ends, k).tobytes() for k in keeps]
        self.images = [mo.record_bytes(first_idx + i, s) for i, s in enumerate(seqs)]
        return np.asarray([len(s) for s in seqs], dtype=np.int64)

    def minimize_one(self, needed, idx=0):
        """(indices of the removed genes in file order, minimized sequence) — MinimizerEngine.minimize_one."""
        t = self.ref.table
        keep = mo.keep_vector(t.names, needed)
        seq = mo.minimize_numpy(self.ref.seq, t.starts, t.ends, keep).tobytes().decode("ascii")
        return np.flatnonzero(~keep), seq

    def record_offsets(self):
        return np.concatenate([[0], np.cumsum([len(x) for x in self.images])]).astype(np.int64)

    @property
    def S(self):
        return len(self.images)

    def chunks(self, max_bytes=0, s0=0, s1=None):
        s1 = self.S if s1 is None else s1
        return [(a, min(a + 3, s1)) for a in range(s0, s1, 3)]          # three records per chunk

    def drain(self, sink, max_bytes=0, s0=0, s1=None):
        for a, b in self.chunks(max_bytes, s0, s1):
            sink(a, b, np.frombuffer(b"".join(self.images[a:b]), dtype=np.uint8))

    def emit_into(self, s0, s1, out):
        data = np.frombuffer(b"".join(self.images[s0:s1]), dtype=np.uint8)
        assert out.size == data.size
        out[:] = data

    def close(self):
        pass


def install():
    engine.MinimizerEngine = OracleEngine
