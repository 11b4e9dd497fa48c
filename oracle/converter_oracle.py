"""CPU restatement of the producer chain in front of the minimizer — TEST INFRASTRUCTURE
(see oracle/__init__.py).  SURVEY.md §8 f1 / BASELINE config 5:

  threshold_samples     utils/extras.py:199-201          binary = (decoded > 0.5).astype(float)
  masks_to_gene_lists   explore_data/binary_converter.py:19-76   de-dup columns keeping the first,
                        rows must have the de-duplicated length, present = (row >= 0.5), names = cols[row]
  add_essentials        explore_data/binary_converter.py:78-121  sorted(set(names) | essential_set)

Pinned by tests/golden/converter_*.json, minted by running the reference's own binary_converter.py.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import numpy as np


def threshold_samples(decoded: np.ndarray, threshold: float = 0.5) -> np.ndarray:
    return (decoded > threshold).astype(float)


def dedup_columns(cols: Sequence[str]) -> List[str]:
    seen, out = set(), []
    for c in cols:
        if c not in seen:
            seen.add(c)
            out.append(c)
    return out


def masks_to_gene_lists(masks: np.ndarray, cols: Sequence[str], threshold: float = 0.5) -> List[List[str]]:
    cols = dedup_columns([str(c) for c in cols])
    P = len(cols)
    col_arr = np.asarray(cols, dtype=object)
    out = []
    for i, row in enumerate(np.asarray(masks)):
        r = np.asarray(row, dtype=float)
        if r.size != P:
            raise ValueError(f"Mask row {i} has length {r.size}, but dataset has {P} gene columns.")
        out.append(col_arr[r >= threshold].tolist())
    return out


def add_essentials(id_lists: Iterable[Sequence[str]], essential_set: Iterable[str]) -> List[List[str]]:
    ess = set(essential_set)
    return [sorted(set(lst) | ess) for lst in id_lists]
