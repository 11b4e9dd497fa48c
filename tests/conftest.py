"""Shared test plumbing.  `-m "not gpu"` runs here (no GPU); `-m gpu` runs on a B200."""
from __future__ import annotations

import glob
import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.json")))


def load_golden(name: str) -> dict:
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as fh:
        case = json.load(fh)
    if "genbank" not in case:
        with gzip.open(os.path.join(GOLDEN_DIR, case["genbank_file"]), "rt") as fh:
            case["genbank"] = fh.read()
    case["name"] = name
    return case


@pytest.fixture(params=golden_names())
def golden(request):
    return load_golden(request.param)


@pytest.fixture
def golden_paths(golden, tmp_path):
    """The fixture's GenBank text and gene lists written to disk the way the CLI receives them."""
    import numpy as np
    gb = tmp_path / "genome.gb"
    gb.write_text(golden["genbank"])
    arr = np.empty(len(golden["lists"]), dtype=object)
    for i, l in enumerate(golden["lists"]):
        arr[i] = l
    npy = tmp_path / "genes.npy"
    np.save(npy, arr, allow_pickle=True)
    return str(gb), str(npy)
