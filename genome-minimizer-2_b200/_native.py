"""ctypes binding of libgm2.so (include/gm2.h).

There is no CPU fallback: loading fails loudly when the library cannot be found/built,
and `Context()` fails loudly when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

from . import build as _build

# ---- constants mirrored from include/gm2.h -------------------------------------------------
ABI_VERSION = 1
OK = 0
ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_CAPACITY, ERR_NOMEM, ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6
CFG_TILE_BYTES, CFG_EMIT_WARPS, CFG_EMIT_BATCH, CFG_PACKING, CFG_STORE_POLICY, CFG_RUN_TABLE, CFG_DEBUG, CFG_ORDER = 1, 2, 3, 4, 5, 6, 7, 8
CFG_FLAT_RUN_BYTES, CFG_WIRE, CFG_HOST_THREADS, CFG_EMIT_OCCUPANCY, CFG_FLAT_MODE = 9, 10, 11, 12, 13
Q_SM_COUNT, Q_LAUNCHES, Q_NUM_SEGMENTS, Q_NUM_TILES, Q_PACKING, Q_NUM_SLOTS, Q_KEEP_WORDS, Q_LAST_WIRE, Q_LAST_D2H_BYTES, Q_LAST_EMIT_CTAS, Q_LAST_FLAT_MODE, Q_TILE_BYTES = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12

_c = ctypes
_P = _c.c_void_p
_I64 = _c.c_int64

# name -> (restype, argtypes); every symbol include/gm2.h declares
SIGNATURES = {
    "gm2_abi_version": (_c.c_int, []),
    "gm2_device_count": (_c.c_int, []),
    "gm2_create": (_c.c_int, [_c.c_int, _c.POINTER(_P)]),
    "gm2_destroy": (_c.c_int, [_P]),
    "gm2_last_error": (_c.c_char_p, [_P]),
    "gm2_configure": (_c.c_int, [_P, _c.c_int, _I64]),
    "gm2_query": (_c.c_int, [_P, _c.c_int, _c.POINTER(_I64)]),
    "gm2_set_stream": (_c.c_int, [_P, _P]),
    "gm2_sync": (_c.c_int, [_P]),
    "gm2_order_after": (_c.c_int, [_P, _P]),
    "gm2_set_reference": (_c.c_int, [_P, _P, _I64, _P, _P, _c.c_int32]),
    "gm2_set_name_map": (_c.c_int, [_P, _P, _P, _c.c_int32]),
    "gm2_set_header_prefix": (_c.c_int, [_P, _c.c_char_p]),
    "gm2_load_ids_host": (_c.c_int, [_P, _P, _P, _I64]),
    "gm2_load_ids_dev": (_c.c_int, [_P, _P, _P, _I64, _I64]),
    "gm2_set_forced": (_c.c_int, [_P, _P, _P]),
    "gm2_load_probs_dev": (_c.c_int, [_P, _P, _I64, _I64, _c.c_float]),
    "gm2_get_counts": (_c.c_int, [_P, _P]),
    "gm2_load_keep_host": (_c.c_int, [_P, _P, _I64]),
    "gm2_load_keep_dev": (_c.c_int, [_P, _P, _I64]),
    "gm2_plan": (_c.c_int, [_P, _I64]),
    "gm2_plan_async": (_c.c_int, [_P, _I64]),
    "gm2_get_lengths": (_c.c_int, [_P, _P]),
    "gm2_get_lengths_dev": (_c.c_int, [_P, _P]),
    "gm2_get_record_offsets": (_c.c_int, [_P, _P]),
    "gm2_get_keep_rows": (_c.c_int, [_P, _P]),
    "gm2_image_bytes": (_c.c_int, [_P, _I64, _I64, _c.POINTER(_I64)]),
    "gm2_emit_dev": (_c.c_int, [_P, _I64, _I64, _P, _I64]),
    "gm2_emit_host": (_c.c_int, [_P, _I64, _I64, _P, _I64, _I64]),
    "gm2_sequence_hashes": (_c.c_int, [_P, _I64, _I64, _P]),
    "gm2_minimize_host": (_c.c_int, [_P, _P, _P, _P, _I64, _I64, _P, _P, _P, _I64, _I64]),
    "gm2_device_alloc": (_c.c_int, [_P, _c.POINTER(_P), _I64]),
    "gm2_device_free": (_c.c_int, [_P, _P]),
    "gm2_upload": (_c.c_int, [_P, _P, _P, _I64]),
    "gm2_host_alloc": (_c.c_int, [_c.POINTER(_P), _I64]),
    "gm2_host_free": (_c.c_int, [_P]),
    "gm2_diag_fill": (_c.c_int, [_P, _P, _I64, _c.c_uint32]),
    "gm2_diag_fill_streams": (_c.c_int, [_P, _P, _I64, _I64, _c.c_int32, _I64, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_int32]),
    "gm2_diag_range_hashes": (_c.c_int, [_P, _P, _I64, _P, _I64, _P]),
    "gm2_diag_host_fill": (_c.c_int, [_P, _I64, _c.c_int32, _c.c_int32, _c.POINTER(_c.c_double)]),
    "gm2_diag_expand": (_c.c_int, [_P, _P, _P, _P, _I64, _c.c_int32, _I64, _c.c_char_p, _P, _c.c_int32, _c.c_int32]),
    "gm2_tokenize_pickle": (_c.c_int, [_P, _I64, _I64, _I64, _P, _P, _c.c_int32, _P, _I64, _P, _P]),
    "gm2_genbank_parse": (_c.c_int, [_P, _I64, _c.POINTER(_P)]),
    "gm2_genbank_sizes": (_c.c_int, [_P, _c.POINTER(_I64), _c.POINTER(_c.c_int32), _c.POINTER(_I64), _c.POINTER(_I64)]),
    "gm2_genbank_copy": (_c.c_int, [_P, _P, _P, _P, _P, _P]),
    "gm2_genbank_free": (_c.c_int, [_P]),
}

_lib = None


class Gm2Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libgm2 error {code}: {message}")
        self.code = code


def library_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load libgm2.so, building it with nvcc if it is absent.  Raises if that fails —
    the product has no other code path."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise RuntimeError(f"{path} is missing; run `python -m genome_minimizer_2_b200.build` "
                               "(nvcc, sm_100a).  There is no CPU fallback.")
        _build.build_native()
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    got = lib.gm2_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"libgm2.so ABI version {got} != binding {ABI_VERSION}; rebuild the library")
    _lib = lib
    return lib


def _ptr(a) -> Optional[int]:
    """Host numpy array -> void*, int -> itself (device pointer), None -> NULL."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return int(a)
    return a.ctypes.data


def tokenize_npy(path: str, names) -> Optional[tuple]:
    """Gene-name lists file -> (ids int32, off int64[S+1], counts int64[S]) through gm2_tokenize_pickle,
    or None when the file is not in the subset the native tokeniser handles (the caller then uses
    `np.load(path, allow_pickle=True).tolist()`, minimizer_2.py:456).  `names[v]` is the vocabulary
    (id v).  Host only; raises what `np.load` would raise for a missing / unreadable file."""
    from numpy.lib import format as npf
    with open(path, "rb") as fh:
        try:
            version = npf.read_magic(fh)
            if version == (1, 0):
                shape, _fortran, dtype = npf.read_array_header_1_0(fh)
            elif version == (2, 0):
                shape, _fortran, dtype = npf.read_array_header_2_0(fh)
            else:
                return None
        except ValueError:
            return None                                   # not a .npy file (np.load has more formats)
        if not dtype.hasobject or dtype != np.dtype(object) or len(shape) not in (1, 2):
            return None
        body = np.fromfile(fh, dtype=np.uint8)
    S = int(shape[0])
    L = int(shape[1]) if len(shape) == 2 else 0
    if body.size == 0 or (len(shape) == 2 and L == 0):
        return None
    enc = [str(n).encode("utf-8", "surrogatepass") for n in names]
    blob = np.frombuffer(b"".join(enc), dtype=np.uint8) if enc else np.zeros(0, dtype=np.uint8)
    name_off = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        name_off[1:] = np.cumsum([len(e) for e in enc])
    cap = body.size // 2 + 16
    ids = np.empty(cap, dtype=np.int32)
    off = np.empty(S + 1, dtype=np.int64)
    counts = np.empty(S, dtype=np.int64)
    if blob.size == 0:
        blob = np.zeros(1, dtype=np.uint8)
    rc = load().gm2_tokenize_pickle(_ptr(body), body.size, S, L, _ptr(blob), _ptr(name_off), len(enc),
                                    _ptr(ids), cap, _ptr(off), _ptr(counts))
    if rc != OK:
        return None               # unsupported or corrupt: NumPy's loader decides (and raises its own errors)
    return ids[:int(off[-1])].copy(), off, counts


def scan_genbank(path: str) -> Optional[tuple]:
    """GenBank file -> (seq uint8[G], names list[str] (F), starts int64[F], ends int64[F], n_features)
    through gm2_genbank_parse, or None when the file is outside the subset the native scanner handles
    (the caller then uses `genbank.read_genbank`, which also owns every error message).  Host only;
    raises what `open` raises for a missing / unreadable file."""
    text = np.fromfile(path, dtype=np.uint8)
    lib = load()
    h = _P()
    rc = lib.gm2_genbank_parse(_ptr(text) if text.size else None, int(text.size), ctypes.byref(h))
    if rc != OK:
        return None
    try:
        G, F, nb, nf = _I64(), _c.c_int32(), _I64(), _I64()
        rc = lib.gm2_genbank_sizes(h, ctypes.byref(G), ctypes.byref(F), ctypes.byref(nb), ctypes.byref(nf))
        if rc != OK:
            raise Gm2Error(rc, (lib.gm2_last_error(None) or b"").decode())
        seq = np.empty(G.value, dtype=np.uint8)
        starts = np.empty(F.value, dtype=np.int64)
        ends = np.empty(F.value, dtype=np.int64)
        name_off = np.empty(F.value + 1, dtype=np.int64)
        blob = np.empty(max(nb.value, 1), dtype=np.uint8)
        rc = lib.gm2_genbank_copy(h, _ptr(seq), _ptr(starts), _ptr(ends), _ptr(name_off), _ptr(blob))
        if rc != OK:
            raise Gm2Error(rc, (lib.gm2_last_error(None) or b"").decode())
    finally:
        lib.gm2_genbank_free(h)
    raw = blob.tobytes()
    names = [raw[int(name_off[g]):int(name_off[g + 1])].decode("ascii") for g in range(F.value)]
    return seq, names, starts, ends, int(nf.value)


class PinnedBuffer:
    """Page-locked host buffer (gm2_host_alloc) exposed as a numpy uint8 array."""

    def __init__(self, nbytes: int):
        lib = load()
        p = _P()
        rc = lib.gm2_host_alloc(ctypes.byref(p), int(nbytes))
        if rc != OK:
            raise Gm2Error(rc, (lib.gm2_last_error(None) or b"").decode())
        self._p = p
        self.nbytes = int(nbytes)
        buf = (ctypes.c_uint8 * max(self.nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=np.uint8, count=self.nbytes)

    @property
    def ptr(self) -> int:
        return self._p.value

    def free(self):
        if self._p is not None:
            self.array = None
            load().gm2_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def host_fill_gbs(buf, threads: int = 0, reps: int = 3) -> float:
    """Non-temporal fill of a host buffer (numpy uint8 array or PinnedBuffer) by `threads` threads: GB/s."""
    arr = buf.array if isinstance(buf, PinnedBuffer) else buf
    out = _c.c_double(0.0)
    rc = load().gm2_diag_host_fill(_ptr(arr), int(arr.size), int(threads), int(reps), ctypes.byref(out))
    if rc != OK:
        raise Gm2Error(rc, (load().gm2_last_error(None) or b"").decode())
    return float(out.value)


class Context:
    """One gm2 context == one GPU.  Thin, typed wrapper; raises Gm2Error on any failure."""

    def __init__(self, device: int = 0):
        self._lib = load()
        h = _P()
        rc = self._lib.gm2_create(int(device), ctypes.byref(h))
        if rc != OK:
            raise Gm2Error(rc, (self._lib.gm2_last_error(None) or b"").decode())
        self._h = h
        self.device = int(device)
        self.G = 0
        self.F = 0
        self.S = 0

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, rc: int):
        if rc != OK:
            raise Gm2Error(rc, (self._lib.gm2_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._lib.gm2_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def configure(self, key: int, value: int):
        self._ck(self._lib.gm2_configure(self._h, key, int(value)))

    def query(self, key: int) -> int:
        v = _I64(0)
        self._ck(self._lib.gm2_query(self._h, key, ctypes.byref(v)))
        return int(v.value)

    def set_stream(self, cuda_stream: Optional[int]):
        self._ck(self._lib.gm2_set_stream(self._h, cuda_stream))

    def order_after(self, other: "Context"):
        """Work issued on this context from now on starts after everything issued on `other` so far."""
        self._ck(self._lib.gm2_order_after(self._h, other._h))

    def sync(self):
        self._ck(self._lib.gm2_sync(self._h))

    # -- static inputs ----------------------------------------------------------------------
    def set_reference(self, seq: np.ndarray, gene_start: np.ndarray, gene_end: np.ndarray):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        gs = np.ascontiguousarray(gene_start, dtype=np.int64)
        ge = np.ascontiguousarray(gene_end, dtype=np.int64)
        if gs.shape != ge.shape or gs.ndim != 1:
            raise ValueError("gene_start / gene_end must be 1-D and equally long")
        self._ck(self._lib.gm2_set_reference(self._h, _ptr(seq), seq.size, _ptr(gs), _ptr(ge), gs.size))
        self.G, self.F, self.S = int(seq.size), int(gs.size), 0

    def set_name_map(self, id2gene_off: np.ndarray, id2gene_idx: np.ndarray):
        off = np.ascontiguousarray(id2gene_off, dtype=np.int32)
        idx = np.ascontiguousarray(id2gene_idx, dtype=np.int32)
        self._ck(self._lib.gm2_set_name_map(self._h, _ptr(off), _ptr(idx), off.size - 1))

    def set_header_prefix(self, prefix: str):
        self._ck(self._lib.gm2_set_header_prefix(self._h, prefix.encode("ascii")))

    @property
    def keep_words(self) -> int:
        return (self.F + 31) // 32

    # -- samples ------------------------------------------------------------------------------
    def load_ids_host(self, ids: np.ndarray, off: np.ndarray):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        off = np.ascontiguousarray(off, dtype=np.int64)
        self._ck(self._lib.gm2_load_ids_host(self._h, _ptr(ids), _ptr(off), off.size - 1))
        self.S = int(off.size - 1)

    def load_ids_dev(self, ids_ptr: int, off_ptr: int, S: int, n_ids: int):
        self._ck(self._lib.gm2_load_ids_dev(self._h, int(ids_ptr), int(off_ptr), int(S), int(n_ids)))
        self.S = int(S)

    def load_keep_host(self, rows: np.ndarray):
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        fw = self.keep_words
        S = rows.shape[0] if rows.ndim == 2 else (rows.size // fw if fw else 0)
        if rows.size != S * fw:
            raise ValueError(f"keep rows must hold S x {fw} uint32 words")
        self._ck(self._lib.gm2_load_keep_host(self._h, _ptr(rows), S))
        self.S = int(S)

    def load_keep_dev(self, rows_ptr: int, S: int):
        self._ck(self._lib.gm2_load_keep_dev(self._h, int(rows_ptr), int(S)))
        self.S = int(S)

    def set_forced(self, force_keep_genes: Optional[np.ndarray], forced_ids: Optional[np.ndarray]):
        fk = None if force_keep_genes is None else np.ascontiguousarray(force_keep_genes, dtype=np.uint32)
        fi = None if forced_ids is None else np.ascontiguousarray(forced_ids, dtype=np.uint32)
        self._ck(self._lib.gm2_set_forced(self._h, _ptr(fk), _ptr(fi)))

    def load_probs_dev(self, probs_ptr: int, S: int, ld: int, threshold: float = 0.5):
        self._ck(self._lib.gm2_load_probs_dev(self._h, int(probs_ptr), int(S), int(ld), float(threshold)))
        self.S = int(S)

    def counts(self) -> np.ndarray:
        out = np.empty(self.S, dtype=np.int64)
        self._ck(self._lib.gm2_get_counts(self._h, _ptr(out)))
        return out

    # -- plan / results -------------------------------------------------------------------------
    def plan(self, first_idx: int = 0):
        self._ck(self._lib.gm2_plan(self._h, int(first_idx)))

    def plan_async(self, first_idx: int = 0):
        self._ck(self._lib.gm2_plan_async(self._h, int(first_idx)))

    def lengths(self) -> np.ndarray:
        out = np.empty(self.S, dtype=np.int64)
        self._ck(self._lib.gm2_get_lengths(self._h, _ptr(out)))
        return out

    def lengths_dev(self, dev_ptr: int):
        """Lengths of the current plan copied device-to-device into `dev_ptr` (S int64), asynchronously."""
        self._ck(self._lib.gm2_get_lengths_dev(self._h, int(dev_ptr)))

    def record_offsets(self) -> np.ndarray:
        out = np.empty(self.S + 1, dtype=np.int64)
        self._ck(self._lib.gm2_get_record_offsets(self._h, _ptr(out)))
        return out

    def keep_rows(self) -> np.ndarray:
        out = np.empty((self.S, self.keep_words), dtype=np.uint32)
        self._ck(self._lib.gm2_get_keep_rows(self._h, _ptr(out)))
        return out

    def image_bytes(self, s0: int, s1: int) -> int:
        v = _I64(0)
        self._ck(self._lib.gm2_image_bytes(self._h, int(s0), int(s1), ctypes.byref(v)))
        return int(v.value)

    # -- emit -----------------------------------------------------------------------------------
    def emit_dev(self, s0: int, s1: int, dev_ptr: int, cap: int):
        self._ck(self._lib.gm2_emit_dev(self._h, int(s0), int(s1), int(dev_ptr), int(cap)))

    def emit_host(self, s0: int, s1: int, out, chunk_bytes: int = 0) -> int:
        """Records [s0,s1) into `out` (numpy uint8 array or PinnedBuffer); returns bytes written."""
        arr = out.array if isinstance(out, PinnedBuffer) else out
        if arr.dtype != np.uint8 or not arr.flags["C_CONTIGUOUS"]:
            raise ValueError("output must be a contiguous uint8 array")
        n = self.image_bytes(s0, s1)
        self._ck(self._lib.gm2_emit_host(self._h, int(s0), int(s1), _ptr(arr), arr.size, int(chunk_bytes)))
        return n

    def minimize_host(self, out, *, ids=None, off=None, keep_rows=None, first_idx: int = 0, chunk_bytes: int = 0):
        """One call: host samples in -> (lengths, rec_off) + image bytes in `out`."""
        arr = out.array if isinstance(out, PinnedBuffer) else out
        if keep_rows is not None:
            keep_rows = np.ascontiguousarray(keep_rows, dtype=np.uint32)
            fw = self.keep_words
            S = keep_rows.shape[0] if keep_rows.ndim == 2 else (keep_rows.size // fw if fw else 0)
        else:
            ids = np.ascontiguousarray(ids, dtype=np.int32)
            off = np.ascontiguousarray(off, dtype=np.int64)
            S = off.size - 1
        lengths = np.empty(S, dtype=np.int64)
        rec_off = np.empty(S + 1, dtype=np.int64)
        self._ck(self._lib.gm2_minimize_host(self._h, _ptr(ids), _ptr(off), _ptr(keep_rows), S, int(first_idx),
                                              _ptr(lengths), _ptr(rec_off), _ptr(arr), arr.size, int(chunk_bytes)))
        self.S = int(S)
        return lengths, rec_off

    def sequence_hashes(self, s0: int = 0, s1: Optional[int] = None) -> np.ndarray:
        s1 = self.S if s1 is None else s1
        out = np.zeros(s1 - s0, dtype=np.uint64)
        self._ck(self._lib.gm2_sequence_hashes(self._h, int(s0), int(s1), _ptr(out)))
        return out

    def device_alloc(self, nbytes: int) -> int:
        p = _P()
        self._ck(self._lib.gm2_device_alloc(self._h, ctypes.byref(p), int(nbytes)))
        return int(p.value)

    def device_free(self, ptr: int):
        self._ck(self._lib.gm2_device_free(self._h, int(ptr)))

    def upload(self, dev_ptr: int, host: np.ndarray):
        host = np.ascontiguousarray(host)
        self._ck(self._lib.gm2_upload(self._h, int(dev_ptr), _ptr(host), host.nbytes))

    # -- diagnostics ------------------------------------------------------------------------------
    def diag_fill(self, dev_ptr: int, nbytes: int, pattern: int = 0x41414141):
        self._ck(self._lib.gm2_diag_fill(self._h, int(dev_ptr), int(nbytes), pattern))

    def diag_fill_streams(self, dev_ptr, nrec, stride, ntile, chunk, batch, warps, order=0, vec32=0):
        self._ck(self._lib.gm2_diag_fill_streams(self._h, int(dev_ptr), nrec, stride, ntile, chunk, batch, warps, order, vec32))

    def diag_range_hashes(self, dev_ptr: int, dev_bytes: int, off: np.ndarray) -> np.ndarray:
        off = np.ascontiguousarray(off, dtype=np.int64)
        out = np.zeros(off.size - 1, dtype=np.uint64)
        self._ck(self._lib.gm2_diag_range_hashes(self._h, int(dev_ptr), int(dev_bytes), _ptr(off), off.size - 1, _ptr(out)))
        return out
