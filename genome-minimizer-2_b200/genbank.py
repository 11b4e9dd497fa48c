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
    """One feature-table entry.  `type` is parsed eagerly; `location` and `qualifiers` are parsed
    from the entry's raw text on first access (the minimizer touches them for `gene` features only,
    minimizer_2.py:59-61, :78-79 — a K-12 file has as many CDS entries that are never looked at)."""
    __slots__ = ("type", "_raw", "_loc", "_quals", "__weakref__")

    def __init__(self, type: str, location: "Location | None" = None, qualifiers: "Dict[str, List[str]] | None" = None,
                 raw: "str | None" = None):
        self.type = type
        self._raw = raw
        self._loc = location
        self._quals = qualifiers

    def _parse(self) -> None:
        loc, quals = _parse_feature_body(self._raw)
        self._loc, self._quals, self._raw = loc, quals, None

    @property
    def location(self) -> Location:
        if self._raw is not None:
            self._parse()
        return self._loc

    @property
    def qualifiers(self) -> Dict[str, List[str]]:
        if self._raw is not None:
            self._parse()
        return self._quals

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


_FEATURE_START = re.compile(r"^ {5}(?=\S)", re.M)
_SIMPLE_LOC = re.compile(r"^(complement\()?(\d+)\.\.(\d+)\)?$")


def _parse_feature_body(raw: str) -> Tuple[Location, Dict[str, List[str]]]:
    """Location and qualifiers of one feature entry (its lines, key line first)."""
    ch = [ln for ln in raw.split("\n") if ln.strip()]
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
    text = "".join(loc_parts)
    m = _SIMPLE_LOC.match(text)
    if m and (m.group(1) is None) == (not text.endswith(")")):
        loc = Location(int(m.group(2)) - 1, int(m.group(3)), -1 if m.group(1) else 1)
    else:
        loc = parse_location(text)
    return loc, quals


def _parse_features(block: str) -> List[Feature]:
    """Feature table text (between the FEATURES line and the next column-0 keyword) -> features.
    One regex pass finds the entries (5 blanks, then the key in columns 6-20); bodies stay raw."""
    starts = [m.start() for m in _FEATURE_START.finditer(block)]
    starts.append(len(block))
    feats: List[Feature] = []
    for a, b in zip(starts, starts[1:]):
        raw = block[a:b]
        feats.append(Feature(raw[5:21].split(None, 1)[0] if raw[5:21].strip() else "", raw=raw))
    return feats


def _split_records(text: str) -> List[str]:
    recs: List[str] = []
    pos = 0 if text.startswith("LOCUS") else text.find("\nLOCUS")
    while pos >= 0:
        if text[pos] == "\n":
            pos += 1
        end = 0 if text.startswith("//", pos) else text.find("\n//", pos)
        stop = len(text) if end < 0 else (text.find("\n", end + 1) + 1 or len(text))
        recs.append(text[pos:stop])
        pos = text.find("\nLOCUS", stop - 1)
    return recs


_COL0_KEYWORD = re.compile(r"^\S", re.M)


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
    first = rec[:rec.find("\n")].split() if "\n" in rec else rec.split()
    name = first[1] if len(first) > 1 else ""
    feats: List[Feature] = []
    f0 = 0 if rec.startswith("FEATURES") else rec.find("\nFEATURES")
    if f0 >= 0:
        body0 = rec.find("\n", f0 + 1) + 1                     # first line after the FEATURES header
        m = _COL0_KEYWORD.search(rec, body0) if body0 > 0 else None
        feats = _parse_features(rec[body0:m.start() if m else len(rec)]) if body0 > 0 else []
    seq = ""
    o0 = rec.find("\nORIGIN")
    if o0 >= 0:
        body0 = rec.find("\n", o0 + 1) + 1
        end = rec.find("\n//", body0 - 1)
        block = rec[body0:end if end >= 0 else len(rec)] if body0 > 0 else ""
        seq = "".join([ln[10:] for ln in block.split("\n")]).replace(" ", "").replace("\r", "").upper()
    return GenomeRecord(seq, feats, name=name)


def sequence_bytes(record) -> np.ndarray:
    """`record.seq` as uint8 (exactly the characters the reference would iterate over)."""
    s = record.seq
    if isinstance(s, (bytes, bytearray)):
        return np.frombuffer(bytes(s), dtype=np.uint8)
    return np.frombuffer(str(s).encode("ascii"), dtype=np.uint8)
