"""Restated `Bio.SeqIO.read(path, "genbank")` — ORACLE side (test infrastructure).

The reference calls Biopython 1.85 at minimizer_2.py:145, :455, :515 and then touches only
`record.seq` (len, iteration), `feature.type`, `feature.qualifiers.get("gene", [""])[0]`
and `int(feature.location.start/.end)` (minimizer_2.py:35, :59-61, :78-79, :94).
Biopython is a third-party dependency (poetry.lock:4-5, >=1.85,<2.0) that is neither
installed here nor vendored in the reference, and the reference has no tests for it:
PARITY UNPINNED at this boundary.  This file follows SURVEY.md App. A.

Written as a line-oriented state machine on purpose — the product's reader
(genome-minimizer-2_b200/genbank.py) is a separately written regex/slice parser, and
tests cross-check the two.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple


@dataclass
class OracleLocation:
    start: int
    end: int
    strand: int = 1


@dataclass
class OracleFeature:
    type: str
    location: OracleLocation
    qualifiers: Dict[str, List[str]] = field(default_factory=dict)


@dataclass
class OracleRecord:
    seq: str
    features: List[OracleFeature]
    name: str = ""


class _LocParser:
    """Recursive-descent parser for the GenBank location grammar -> (start, end, strand)."""

    def __init__(self, text: str):
        self.t = "".join(text.split())
        self.i = 0

    def _peek(self, s: str) -> bool:
        return self.t.startswith(s, self.i)

    def _expect(self, s: str) -> None:
        if not self._peek(s):
            raise ValueError(f"cannot parse location {self.t!r} at offset {self.i}")
        self.i += len(s)

    def parse(self) -> Tuple[int, int, int]:
        out = self._loc()
        if self.i != len(self.t):
            raise ValueError(f"trailing text in location {self.t!r}")
        return out

    def _loc(self) -> Tuple[int, int, int]:
        if self._peek("complement("):
            self.i += len("complement(")
            a, b, _ = self._loc()
            self._expect(")")
            return a, b, -1
        for kw in ("join(", "order("):
            if self._peek(kw):
                self.i += len(kw)
                parts = [self._loc()]
                while self._peek(","):
                    self.i += 1
                    parts.append(self._loc())
                self._expect(")")
                # CompoundLocation.start/.end = min/max over parts (span incl. gaps)
                return min(p[0] for p in parts), max(p[1] for p in parts), parts[0][2]
        return self._simple()

    def _number(self) -> int:
        j = self.i
        while j < len(self.t) and self.t[j].isdigit():
            j += 1
        if j == self.i:
            raise ValueError(f"cannot parse location {self.t!r} at offset {self.i}")
        v = int(self.t[self.i:j])
        self.i = j
        return v

    def _simple(self) -> Tuple[int, int, int]:
        # remote reference "ACC.1:" prefix is not supported (rare; Biopython keeps a ref)
        if ":" in self.t[self.i:].split(",")[0].split(")")[0]:
            raise ValueError(f"remote location not supported: {self.t!r}")
        if self._peek("<") or self._peek(">"):
            self.i += 1
        n = self._number()
        if self._peek(".."):
            self.i += 2
            if self._peek("<") or self._peek(">"):
                self.i += 1
            m = self._number()
            return n - 1, m, 1
        if self._peek("^"):
            self.i += 1
            self._number()
            return n, n, 1          # between-bases: zero length
        if self._peek("."):
            raise ValueError(f"within-position 'N.M' not supported: {self.t!r}")
        return n - 1, n, 1


def parse_location(text: str) -> OracleLocation:
    a, b, s = _LocParser(text).parse()
    return OracleLocation(a, b, s)


def _clean_value(key: str, raw: str) -> str:
    v = raw
    if v.startswith('"'):
        v = v[1:]
    if v.endswith('"'):
        v = v[:-1]
    v = v.replace('""', '"')
    if key == "translation":
        v = "".join(v.split())
    return v


def _parse_one_record(lines: List[str]) -> OracleRecord:
    name = ""
    features: List[OracleFeature] = []
    seq_parts: List[str] = []
    state = "header"
    cur: Optional[OracleFeature] = None
    loc_buf: List[str] = []
    q_key: Optional[str] = None
    q_buf: List[str] = []
    q_open = False          # inside a quoted value that has not met its closing quote yet
    in_loc = False

    def finish_qualifier():
        nonlocal q_key, q_buf, q_open
        q_open = False
        if cur is not None and q_key is not None:
            if not q_buf:                       # bare key
                cur.qualifiers.setdefault(q_key, [""])
            else:
                cur.qualifiers.setdefault(q_key, []).append(_clean_value(q_key, " ".join(q_buf)))
        q_key, q_buf = None, []

    def finish_feature():
        nonlocal cur, loc_buf, in_loc
        finish_qualifier()
        if cur is not None:
            if in_loc:
                cur.location = parse_location("".join(loc_buf))
            features.append(cur)
        cur, loc_buf, in_loc = None, [], False

    for line in lines:
        line = line.rstrip("\r\n")
        if state == "header":
            if line.startswith("LOCUS"):
                parts = line.split()
                name = parts[1] if len(parts) > 1 else ""
            elif line.startswith("FEATURES"):
                state = "features"
            elif line.startswith("ORIGIN"):
                state = "origin"
        elif state == "features":
            if line[:1] not in (" ", ""):
                finish_feature()
                state = "origin" if line.startswith("ORIGIN") else "header"
                continue
            if not line.strip():
                continue
            if line[:5] == "     " and len(line) > 5 and line[5] != " ":
                finish_feature()
                cur = OracleFeature(type=line[5:21].strip(), location=OracleLocation(0, 0), qualifiers={})
                loc_buf = [line[21:].strip()]
                in_loc = True
            else:
                body = line[21:] if line[:21].strip() == "" else line.strip()
                if in_loc and not body.startswith("/"):
                    loc_buf.append(body.strip())
                    continue
                if in_loc:
                    cur.location = parse_location("".join(loc_buf))
                    in_loc = False
                if body.startswith("/") and not q_open:
                    finish_qualifier()
                    if "=" in body:
                        k, v = body[1:].split("=", 1)
                        q_key, q_buf = k, [v]
                        q_open = v.startswith('"') and (v == '"' or not v.endswith('"'))
                    else:
                        q_key, q_buf = body[1:], []
                else:
                    q_buf.append(body.strip())
                    if q_open and body.rstrip().endswith('"'):
                        q_open = False
        elif state == "origin":
            if line.startswith("//"):
                break
            seq_parts.append(line[10:].replace(" ", ""))
    finish_feature()
    return OracleRecord(seq="".join(seq_parts).upper(), features=features, name=name)


def read_genbank(path: str) -> OracleRecord:
    """Exactly-one-record rule of `SeqIO.read`."""
    with open(path, "r") as fh:
        all_lines = fh.readlines()
    records: List[List[str]] = []
    cur: Optional[List[str]] = None
    for ln in all_lines:
        if ln.startswith("LOCUS"):
            cur = [ln]
        elif cur is not None:
            cur.append(ln)
            if ln.startswith("//"):
                records.append(cur)
                cur = None
    if cur is not None:          # unterminated trailing record still counts
        records.append(cur)
    if not records:
        raise ValueError("No records found in handle")
    if len(records) > 1:
        raise ValueError("More than one record found in handle")
    return _parse_one_record(records[0])
