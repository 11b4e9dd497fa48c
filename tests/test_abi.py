"""The C-ABI library: loads, exports every symbol include/gm2.h declares, binding covers them.
No compute calls (no GPU here).  CPU only."""
from __future__ import annotations

import ctypes
import os
import re

import pytest

from genome_minimizer_2_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gm2.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gm2_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    lib = build.build_native()
    assert os.path.exists(lib)
    L = _native.load(build_if_missing=False)
    assert L.gm2_abi_version() == _native.ABI_VERSION


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 25
    raw = ctypes.CDLL(build.LIB)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/gm2.h but not exported by libgm2.so"
    assert sorted(_native.SIGNATURES) == names


def test_header_constants_match_binding():
    text = open(os.path.join(ROOT, "include", "gm2.h")).read()
    consts = dict(re.findall(r"#define\s+(GM2_[A-Z0-9_]+)\s+(-?\d+)", text))
    assert int(consts["GM2_ABI_VERSION"]) == _native.ABI_VERSION
    for k, v in (("GM2_ERR_INVALID", _native.ERR_INVALID), ("GM2_ERR_CUDA", _native.ERR_CUDA),
                 ("GM2_ERR_STATE", _native.ERR_STATE), ("GM2_ERR_CAPACITY", _native.ERR_CAPACITY),
                 ("GM2_ERR_NOMEM", _native.ERR_NOMEM), ("GM2_ERR_UNSUPPORTED", _native.ERR_UNSUPPORTED), ("GM2_CFG_TILE_BYTES", _native.CFG_TILE_BYTES),
                 ("GM2_CFG_EMIT_WARPS", _native.CFG_EMIT_WARPS), ("GM2_CFG_EMIT_BATCH", _native.CFG_EMIT_BATCH),
                 ("GM2_CFG_PACKING", _native.CFG_PACKING), ("GM2_CFG_STORE_POLICY", _native.CFG_STORE_POLICY),
                 ("GM2_CFG_FLAT_RUN_BYTES", _native.CFG_FLAT_RUN_BYTES), ("GM2_CFG_WIRE", _native.CFG_WIRE),
                 ("GM2_CFG_HOST_THREADS", _native.CFG_HOST_THREADS), ("GM2_Q_LAST_WIRE", _native.Q_LAST_WIRE),
                 ("GM2_Q_LAST_D2H_BYTES", _native.Q_LAST_D2H_BYTES), ("GM2_CFG_ORDER", _native.CFG_ORDER),
                 ("GM2_Q_LAUNCHES", _native.Q_LAUNCHES), ("GM2_Q_KEEP_WORDS", _native.Q_KEEP_WORDS)):
        assert int(consts[k]) == v, k


def test_no_cpu_fallback_without_a_gpu():
    """Without a CUDA device the product must fail loudly, not compute on the CPU."""
    L = _native.load()
    if L.gm2_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(_native.Gm2Error) as ei:
        _native.Context(0)
    assert ei.value.code == _native.ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "genome-minimizer-2_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "liboracle" not in src, fn
