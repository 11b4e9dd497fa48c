"""ctypes binding of oracle/minimizer_c.c — TEST INFRASTRUCTURE (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_gm2.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "minimizer_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/liboracle_gm2.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        P = ctypes.c_void_p
        L.oracle_minimize.restype = ctypes.c_int64
        L.oracle_minimize.argtypes = [P, ctypes.c_int64, P, P, ctypes.c_int32, P, P, P]
        L.oracle_range_hash.restype = ctypes.c_uint64
        L.oracle_range_hash.argtypes = [P, ctypes.c_int64]
        L.oracle_batch.restype = ctypes.c_int64
        L.oracle_batch.argtypes = [P, ctypes.c_int64, P, P, ctypes.c_int32, P, ctypes.c_int64, ctypes.c_int64,
                                   ctypes.c_char_p, P, P, P]
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def batch(seq: np.ndarray, starts: np.ndarray, ends: np.ndarray, keep_rows: np.ndarray, first_idx: int = 0,
          prefix: str = "Minimized_E_coli_K12_MG1655_", want_image: bool = False
          ) -> Tuple[np.ndarray, np.ndarray, Optional[np.ndarray]]:
    """(lengths int64[S], record hashes uint64[S], image uint8[...] or None)."""
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    ends = np.ascontiguousarray(ends, dtype=np.int64)
    keep_rows = np.ascontiguousarray(keep_rows, dtype=np.uint32)
    F = len(starts)
    FW = (F + 31) // 32
    S = keep_rows.size // FW if FW else keep_rows.shape[0]
    lengths = np.zeros(S, dtype=np.int64)
    hashes = np.zeros(S, dtype=np.uint64)
    image = None
    if want_image:
        image = np.empty(S * (len(seq) + len(prefix) + 32), dtype=np.uint8)
    total = lib().oracle_batch(_p(seq), len(seq), _p(starts), _p(ends), F, _p(keep_rows), S, first_idx,
                               prefix.encode(), _p(lengths), _p(hashes), _p(image))
    if total < 0:
        raise MemoryError("oracle_batch")
    if image is not None:
        image = image[:total]
    return lengths, hashes, image


def range_hash(data) -> int:
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
    return int(lib().oracle_range_hash(_p(a), a.size))
