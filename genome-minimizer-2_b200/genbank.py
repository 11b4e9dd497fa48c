"""GenBank flat-file reader: the host-side stand-in for `Bio.SeqIO.read(path, "genbank")`.

The reference loads its genome with Biopython 1.85 (minimizer_2.py:145, :455, :515;
poetry.lock:4-5) and then reads only `record.seq`, `feature.type`,
`feature.qualifiers.get("gene", [""])[0]`, `int(feature.location.start)` and
`int(feature.location.end)` (minimizer_2.py:35, :59-61, :78-79, :94).  Biopython is not a
dependency of this build, so this module restates the documented behaviour needed by that
path: exactly one LOCUS..// record, feature keys in columns 6-20,
locations from column 22 possibly continued over lines, `/key=value` qualifiers with quoted
multi-line values, span semantics for join()/order()/complement(), between-base `N^M` as a
zero-length location, fuzzy `<`/`>` ignored, ORIGIN lines taken from column 11 with blanks
removed, sequence upper-cased.  Parity with Biopython itself is unpinned (the reference has
no tests; Biopython is not installable offline) — see DESIGN.md.

A record read from a real Biopython install can be used instead: everything downstream
(`engine.GeneTable.from_record`) only duck-types the attributes listed above.
"""
from __future__ import annotations

import re
from typing import Dict, List, Tuple

import numpy as np

__all__ = ["Location", "Feature", "GenomeRecord", "parse_location", "read_genbank"]


class Location:
    __slots__ = ("start", "end", "strand")

    def __init__(self, start: int, end: int, strand: int = 1):
        self.start = start
        self.end = end
        self.strand = strand

    def __repr__(self):
        return f"Location([{self.start}:{self.end}]({'+' if self.strand >= 0 else '-'}))"


class Feature:
    __slots__ = ("type", "location", "qualifiers")

    def __init__(self, type: str, location: Location, qualifiers: Dict[str, List[str]]):
        self.type = type
        self.location = location
        self.qualifiers = qualifiers

    def __repr__(self):
        return f"Feature({self.type!r}, {self.location!r})"


class GenomeRecord:
    """Duck-type of the slice of `Bio.SeqRecord.SeqRecord` the minimizer uses."""

    def __init__(self, seq: str, features: List[Feature], name: str = "", id: str = ""):
        self.seq = seq              # upper-case str: len(), iteration and str() behave like Bio.Seq here
        self.features = features
        self.name = name
        self.id = id or name

    def __len__(self):
        return len(self.seq)


_SIMPLE = re.compile(r"(?:[A-Za-z_][\w.]*:)?([<>]?)(\d+)(?:(\.\.|\^|\.)([<>]?)(\d+))?")
_SHAPE = re.compile(r"^(?:complement\(|join\(|order\(|L|,|\))+$")


def parse_location(text: str) -> Location:
    """Location string -> 0-based half-open span (min start .. max end over all parts)."""
    t = "".join(text.split())
    if not t:
        raise ValueError("empty feature location")
    spans: List[Tuple[int, int]] = []

    def repl(m: "re.Match[str]") -> str:
        if m.group(0).find(":") >= 0:
            raise ValueError(f"remote feature location not supported: {text!r}")
        n = int(m.group(2))
        op = m.group(3)
        if op is None:
            spans.append((n - 1, n))
        elif op == "..":
            spans.append((n - 1, int(m.group(5))))
        elif op == "^":
            spans.append((n, n))
        else:
            raise ValueError(f"within-position location 'N.M' not supported: {text!r}")
        return "L"

    shape = _SIMPLE.sub(repl, t)
    if not _SHAPE.match(shape) or shape.count("(") != shape.count(")") or not spans:
        raise ValueError(f"cannot parse feature location {text!r}")
    strand = -1 if t.startswith("complement(") else 1
    return Location(min(a for a, _ in spans), max(b for _, b in spans), strand)


def _unquote(key: str, value: str) -> str:
    if value[:1] == '"':
        value = value[1:]
    if value[-1:] == '"':
        value = value[:-1]
    value = value.replace('""', '"')
    if key == "translation":
        value = re.sub(r"\s+", "", value)
    return value


def _parse_features(lines: List[str]) -> List[Feature]:
    # group the block into one chunk of lines per feature
    chunks: List[List[str]] = []
    for ln in lines:
        if not ln.strip():
            continue
        if ln.startswith("     ") and len(ln) > 5 and ln[5] != " ":
            chunks.append([ln])
        elif chunks:
            chunks[-1].append(ln)
    feats: List[Feature] = []
    for ch in chunks:
        key = ch[0][5:21].strip()
        loc_parts = [ch[0][21:].strip()]
        i = 1
        while i < len(ch) and not ch[i][21:].lstrip().startswith("/") and not ch[i].lstrip().startswith("/"):
            loc_parts.append(ch[i].strip())
            i += 1
        quals: Dict[str, List[str]] = {}
        while i < len(ch):
            body = ch[i][21:] if ch[i][:21].strip() == "" else ch[i].strip()
            i += 1
            if not body.startswith("/"):
                continue                                   # stray continuation: ignore
            if "=" not in body:
                quals.setdefault(body[1:].strip(), [""])    # bare /key (only if not seen yet)
                continue
            qk, qv = body[1:].split("=", 1)
            if qv.startswith('"'):
                pieces = [qv]
                while (pieces[-1] == '"' and len(pieces) == 1 or not pieces[-1].endswith('"')) and i < len(ch):
                    nxt = ch[i][21:] if ch[i][:21].strip() == "" else ch[i].strip()
                    pieces.append(nxt.strip())
                    i += 1
                qv = " ".join(pieces)
            quals.setdefault(qk, []).append(_unquote(qk, qv))
        feats.append(Feature(key, parse_location("".join(loc_parts)), quals))
    return feats


def _split_records(text: str) -> List[str]:
    recs: List[str] = []
    pos = 0
    while True:
        m = re.compile(r"^LOCUS", re.M).search(text, pos)
        if not m:
            break
        end = re.compile(r"^//", re.M).search(text, m.start())
        stop = len(text) if not end else end.end()
        recs.append(text[m.start():stop])
        pos = stop
    return recs


def read_genbank(path: str) -> GenomeRecord:
    """One-record GenBank file -> GenomeRecord.  Raises ValueError with Biopython's messages
    for zero or several records (`SeqIO.read` contract)."""
    with open(path, "r") as fh:
        text = fh.read()
    recs = _split_records(text)
    if not recs:
        raise ValueError("No records found in handle")
    if len(recs) > 1:
        raise ValueError("More than one record found in handle")
    rec = recs[0]
    lines = rec.split("\n")
    name = ""
    first = lines[0].split()
    if len(first) > 1:
        name = first[1]
    # locate blocks by their column-0 keywords
    feat_lo = feat_hi = org_lo = None
    for i, ln in enumerate(lines):
        if feat_lo is None and ln.startswith("FEATURES"):
            feat_lo = i + 1
        elif feat_lo is not None and feat_hi is None and ln[:1] not in (" ", "") and not ln.startswith("FEATURES"):
            feat_hi = i
        if ln.startswith("ORIGIN"):
            org_lo = i + 1
            if feat_lo is not None and feat_hi is None:
                feat_hi = i
            break
    feats: List[Feature] = []
    if feat_lo is not None:
        feats = _parse_features(lines[feat_lo:feat_hi if feat_hi is not None else len(lines)])
    seq = ""
    if org_lo is not None:
        body = []
        for ln in lines[org_lo:]:
            if ln.startswith("//"):
                break
            body.append(ln[10:])
        seq = "".join(body).replace(" ", "").replace("\r", "").upper()
    return GenomeRecord(seq, feats, name=name)


def sequence_bytes(record) -> np.ndarray:
    """`record.seq` as uint8 (exactly the characters the reference would iterate over)."""
    s = record.seq
    if isinstance(s, (bytes, bytearray)):
        return np.frombuffer(bytes(s), dtype=np.uint8)
    return np.frombuffer(str(s).encode("ascii"), dtype=np.uint8)
