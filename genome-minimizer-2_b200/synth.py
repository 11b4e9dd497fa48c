"""Synthetic inputs for the minimizer path — the reference ships no data.

  make_genome(...)        K-12-shaped (or 12 Mbp-shaped) random genome + gene table
  write_genbank(...)      GenBank flat file: LOCUS / FEATURES (gene + distractor CDS, tRNA,
                          misc_feature) / ORIGIN in 6 blocks of 10 lower-case bases per line
  make_gene_lists(...)    per-sample name lists as `binary_converter.py:64-71` writes them
                          (np.array(lists, dtype=object)), plus non-matching pangenome names
  random_keep / ids_csr   compact sample forms for the large configs (no Python strings)

Everything is seeded and deterministic.  Only `gene` features matter to the path
(minimizer_2.py:60); the CDS/tRNA/misc_feature entries exist so that a reader that looks
at the wrong feature type is caught.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

K12_G = 4_641_652
K12_F = 4_400


@dataclass
class SynthGene:
    name: Optional[str]          # None -> gene feature without /gene (only /locus_tag)
    start: int                   # 0-based span start
    end: int                     # span end (exclusive)
    strand: int = 1
    parts: Optional[List[Tuple[int, int]]] = None   # join() parts, 0-based half-open, file order
    locus_tag: str = ""
    extra_names: List[str] = field(default_factory=list)   # further /gene values (synonyms)


@dataclass
class SynthGenome:
    seq: np.ndarray              # uint8, UPPER-case ASCII
    genes: List[SynthGene]       # file order
    name: str = "SYNTH"

    @property
    def G(self) -> int:
        return int(self.seq.size)

    def gene_names(self) -> List[str]:
        return [g.name if g.name is not None else "" for g in self.genes]

    def starts_ends(self) -> Tuple[np.ndarray, np.ndarray]:
        return (np.asarray([g.start for g in self.genes], dtype=np.int64),
                np.asarray([g.end for g in self.genes], dtype=np.int64))


_LETTERS = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz", dtype=np.uint8)


def _unique_names(rng: np.random.Generator, n: int) -> List[str]:
    out, seen = [], set()
    while len(out) < n:
        base = bytes(_LETTERS[rng.integers(0, 26, 3)]).decode()
        suf = "" if rng.random() < 0.25 else chr(ord("A") + int(rng.integers(0, 26)))
        nm = base + suf
        if nm not in seen:
            seen.add(nm)
            out.append(nm)
    return out


def make_genome(G: int = K12_G, F: int = K12_F, seed: int = 1, *, overlap_frac: float = 0.15,
                nested: int = 8, dup_name_frac: float = 0.01, nameless_frac: float = 0.005,
                join_genes: int = 0, iupac_runs: int = 0, origin_wrap: bool = False,
                genic_frac: float = 0.88, name: str = "SYNTH_K12") -> SynthGenome:
    """Random genome of G bases with about F genes laid out left to right.

    Gene lengths are log-normal (mean ~ genic_frac*G/F, clipped to [60, 7000] at K-12 scale),
    ~half on the complement strand, `overlap_frac` of neighbours overlap by 1-30 bp, `nested`
    small genes sit inside larger ones, ~1 % duplicate names, ~0.5 % nameless genes."""
    rng = np.random.default_rng(seed)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, G)].copy()
    for _ in range(iupac_runs):
        if G < 50:
            break
        a = int(rng.integers(0, G - 40))
        n = int(rng.integers(1, 40))
        seq[a:a + n] = np.frombuffer(b"NRYKMSWBDHV", dtype=np.uint8)[rng.integers(0, 11, n)]

    n_main = max(F - nested, 0)
    genes: List[SynthGene] = []
    if n_main > 0 and G > 0:
        mean_len = max(genic_frac * G / n_main, 3.0)
        lens = rng.lognormal(mean=np.log(mean_len) - 0.18, sigma=0.6, size=n_main)
        lens = np.clip(lens, min(60, max(mean_len * 0.06, 1)), max(mean_len * 7.4, 2)).astype(np.int64)
        lens = np.maximum(lens, 1)
        is_ov = rng.random(n_main) < overlap_frac
        is_ov[0] = False
        ov = np.where(is_ov, rng.integers(1, 31, n_main), 0)
        ov = np.minimum(ov, np.maximum(lens - 1, 0))
        ov[1:] = np.minimum(ov[1:], np.maximum(lens[:-1] - 1, 0))
        span = int(lens.sum() - ov.sum())
        free = G - span
        if free < n_main + 1:                       # too dense: shrink genes proportionally
            scale = (G * 0.9) / max(span, 1)
            lens = np.maximum((lens * scale).astype(np.int64), 1)
            ov = np.minimum(ov, np.maximum(lens - 1, 0))
            ov[1:] = np.minimum(ov[1:], np.maximum(lens[:-1] - 1, 0))
            span = int(lens.sum() - ov.sum())
            free = G - span
        w = rng.exponential(1.0, n_main + 1)
        w[1:-1][is_ov[1:]] = 0.0                   # an overlapping neighbour has no gap
        gaps = np.floor(w / w.sum() * max(free, 0)).astype(np.int64)
        pos = int(gaps[0])
        for i in range(n_main):
            if i > 0:
                pos += int(gaps[i]) - int(ov[i])
            a, b = pos, min(pos + int(lens[i]), G)
            genes.append(SynthGene(name="", start=a, end=b, strand=1 if rng.random() < 0.5 else -1))
            pos = b
    # nested genes inside random larger genes
    for _ in range(nested if genes else 0):
        host = genes[int(rng.integers(0, len(genes)))]
        hl = host.end - host.start
        if hl < 6:
            continue
        n = int(rng.integers(2, max(hl // 2, 3)))
        a = host.start + int(rng.integers(1, hl - n))
        genes.append(SynthGene(name="", start=a, end=a + n, strand=-host.strand))
    # join() genes: turn some genes into two-part joins with an internal gap (span semantics, F4)
    for k in rng.permutation(len(genes))[:join_genes]:
        g = genes[int(k)]
        L = g.end - g.start
        if L >= 9:
            c1 = g.start + L // 3
            c2 = g.start + 2 * L // 3
            g.parts = [(g.start, c1), (c2, g.end)]
    if origin_wrap and G >= 20:
        genes.append(SynthGene(name="", start=0, end=G, strand=1, parts=[(G - 7, G), (0, 5)]))
    genes.sort(key=lambda g: (g.start, g.end))
    names = _unique_names(rng, len(genes))
    for i, g in enumerate(genes):
        g.name = names[i]
        g.locus_tag = f"b{i+1:04d}"
    n = len(genes)
    for _ in range(int(round(dup_name_frac * n))):
        i, j = rng.integers(0, n, 2)
        genes[int(i)].name = genes[int(j)].name
    for i in rng.permutation(n)[:int(round(nameless_frac * n))]:
        genes[int(i)].name = None
    for k in range(n // 2, n):                    # one gene carries a second /gene value (a synonym)
        if genes[k].name is not None:
            genes[k].extra_names = ["syn" + genes[k].locus_tag]
            break
    return SynthGenome(seq=seq, genes=genes, name=name)


def _loc_text(g: SynthGene) -> str:
    if g.parts:
        body = "join(" + ",".join(f"{a+1}..{b}" for a, b in g.parts) + ")"
    else:
        body = f"{g.start+1}..{g.end}"
    return f"complement({body})" if g.strand < 0 else body


def _feature_lines(key: str, loc: str, quals: Sequence[str]) -> List[str]:
    """One feature: key in columns 6-20, location from column 22 (wrapped at commas), then one
    or more already-formatted qualifier lines (each may contain '\n' for a wrapped value)."""
    chunks, cur = [], ""
    for piece in loc.replace(",", ",\0").split("\0"):
        if len(cur) + len(piece) > 58 and cur:
            chunks.append(cur)
            cur = ""
        cur += piece
    chunks.append(cur)
    lines = [f"     {key:<16}" + chunks[0]] + [" " * 21 + c for c in chunks[1:]]
    for q in quals:
        lines += [" " * 21 + part for part in q.split("\n")]
    return lines


def genbank_text(genome: SynthGenome, *, lowercase: bool = True, distractors: bool = True, seed: int = 0) -> str:
    rng = np.random.default_rng(seed + 7919)
    G = genome.G
    out = [f"LOCUS       {genome.name:<16} {G:>11} bp    DNA     circular BCT 01-JAN-2000",
           "DEFINITION  Synthetic genome for the genome-minimizer-2 B200 build.",
           f"ACCESSION   {genome.name}",
           f"VERSION     {genome.name}.1",
           "KEYWORDS    .",
           "SOURCE      synthetic construct",
           "  ORGANISM  synthetic construct",
           "            other sequences; artificial sequences.",
           "FEATURES             Location/Qualifiers"]
    out += _feature_lines("source", f"1..{max(G, 1)}", ['/organism="synthetic construct"', '/mol_type="genomic DNA"'])
    for g in genome.genes:
        loc = _loc_text(g)
        q: List[str] = []
        if g.name is not None:
            q.append(f'/gene="{g.name}"')
        q += [f'/gene="{extra}"' for extra in g.extra_names]
        q.append(f'/locus_tag="{g.locus_tag}"')
        out += _feature_lines("gene", loc, q)
        if distractors:
            r = rng.random()
            if r < 0.93:
                cds = q + ["/codon_start=1",
                           '/note="contains ""quoted"" text and a description long enough that it\nis wrapped over two lines"',
                           f'/product="hypothetical protein {g.locus_tag}"',
                           '/translation="MKRISTTITTTITITTGNGAGMSLNRWQ\nAARTLLPVIA"']
                if r < 0.02:
                    cds.insert(len(q), "/pseudo")
                out += _feature_lines("CDS", loc, cds)
            elif r < 0.97:
                out += _feature_lines("tRNA", loc, q + ['/product="tRNA-Xxx"'])
            else:
                out += _feature_lines("misc_feature", loc, ['/note="no gene qualifier here"'])
    out.append("ORIGIN")
    s = genome.seq.tobytes().decode("ascii")
    if lowercase:
        s = s.lower()
    for i in range(0, G, 60):
        row = s[i:i + 60]
        out.append(f"{i+1:>9} " + " ".join(row[j:j + 10] for j in range(0, len(row), 10)))
    out.append("//")
    return "\n".join(out) + "\n"


def write_genbank(path: str, genome: SynthGenome, **kw) -> None:
    with open(path, "w") as fh:
        fh.write(genbank_text(genome, **kw))


def make_gene_lists(genome: SynthGenome, S: int, p: float = 0.5, seed: int = 1, extra_names: int = 2000,
                    sort_lists: bool = False) -> List[List[str]]:
    """Per sample: each distinct gene name kept i.i.d. with probability p, plus `extra_names`
    non-matching pangenome-style names (`group_1234`, `abcD_2`), in column order or sorted
    (binary_converter.py:64 vs :110)."""
    rng = np.random.default_rng(seed)
    uniq = list(dict.fromkeys(n for n in genome.gene_names() if n != ""))
    pool = [f"group_{i}" for i in range(4 * extra_names + 8)] + [f"{n}_2" for n in uniq[:extra_names]]
    lists = []
    for _ in range(S):
        keep = rng.random(len(uniq)) < p
        names = [n for n, k in zip(uniq, keep) if k]
        if extra_names:
            ex = rng.choice(len(pool), size=min(extra_names, len(pool)), replace=False)
            names += [pool[int(i)] for i in ex]
            order = rng.permutation(len(names))
            names = [names[int(i)] for i in order]
        if sort_lists:
            names = sorted(names)
        lists.append(names)
    return lists


def save_gene_lists(path: str, lists: List[List[str]]) -> None:
    """Exactly the container `binary_converter.py:71` writes."""
    arr = np.empty(len(lists), dtype=object)
    for i, l in enumerate(lists):
        arr[i] = l
    np.save(path, arr, allow_pickle=True)


def random_keep_bool(F: int, S: int, p, seed: int = 2) -> np.ndarray:
    """Boolean [S, F]; p is a scalar or a per-sample array of retention probabilities."""
    rng = np.random.default_rng(seed)
    p = np.broadcast_to(np.asarray(p, dtype=np.float64).reshape(-1, 1), (S, 1)) if np.ndim(p) else p
    return rng.random((S, F)) < p


def pack_keep_rows(keep: np.ndarray) -> np.ndarray:
    keep = np.atleast_2d(np.asarray(keep, dtype=bool))
    S, F = keep.shape
    fw = (F + 31) // 32
    padded = np.zeros((S, fw * 32), dtype=np.uint8)
    padded[:, :F] = keep
    return np.packbits(padded, axis=1, bitorder="little").view("<u4").reshape(S, fw)


def ids_csr_from_keep(keep_names: np.ndarray, n_noise: int = 0, V: int = 0, seed: int = 3
                      ) -> Tuple[np.ndarray, np.ndarray]:
    """Boolean [S, V_names] over DISTINCT names -> CSR of name ids, each row shuffled and padded
    with `n_noise` ids >= V (names that match no gene: legal, ignored by the device)."""
    rng = np.random.default_rng(seed)
    S, Vn = keep_names.shape
    rows = []
    for s in range(S):
        r = np.flatnonzero(keep_names[s]).astype(np.int32)
        if n_noise:
            r = np.concatenate([r, rng.integers(V, V + 50_000, n_noise).astype(np.int32)])
            rng.shuffle(r)
        rows.append(r)
    off = np.zeros(S + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in rows])
    ids = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int32)
    return ids.astype(np.int32), off
