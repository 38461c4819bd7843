"""Python front-end of tools/gen_csv.c (deterministic synthetic CSV; bench/test tooling)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcsvgen.so")
_lib = None

GiB = 1 << 30


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gen_csv.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcsvgen.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.gen_max_row_bytes.argtypes = [C.c_uint32, C.c_int]
        L.gen_max_row_bytes.restype = C.c_size_t
        L.gen_unquoted.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, C.c_size_t,
                                   C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
        L.gen_unquoted.restype = C.c_size_t
        L.gen_quoted.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t,
                                 C.POINTER(C.c_uint64)]
        L.gen_quoted.restype = C.c_size_t
        L.gen_queries.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.gen_queries.restype = None
        _lib = L
    return _lib


def _alloc(n: int, out: np.ndarray | None):
    if out is not None:
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.size >= n
        return out
    return np.empty(n, dtype=np.uint8)


def unquoted(target: int, seed: int = 42, nfields: int = 16, modulus: int = 10 ** 6, first_row: int = 0,
             with_header: bool = True, out: np.ndarray | None = None):
    """cfg2 (nfields=16, modulus=1e6) / cfg5 (nfields=256, modulus=1e15). Returns (bytes view, rows)."""
    slack = lib().gen_max_row_bytes(nfields, 0) + 8 * nfields + 64
    buf = _alloc(target + slack, out)
    rows = C.c_uint64()
    n = lib().gen_unquoted(seed, nfields, modulus, first_row, int(with_header), target, buf.ctypes.data,
                           buf.size, C.byref(rows))
    return buf[:n], rows.value


def quoted(target: int, seed: int = 43, first_row: int = 0, with_header: bool = True,
           out: np.ndarray | None = None):
    """cfg3 / cfg4 grammar. Returns (bytes view, rows)."""
    slack = lib().gen_max_row_bytes(16, 1) + 256
    buf = _alloc(target + slack, out)
    rows = C.c_uint64()
    n = lib().gen_quoted(seed, first_row, int(with_header), target, buf.ctypes.data, buf.size, C.byref(rows))
    return buf[:n], rows.value


def queries(nq: int, record_cnt: int, field_cnt: int, seed: int = 46):
    rec = np.empty(nq, dtype=np.uint32)
    fld = np.empty(nq, dtype=np.uint32)
    lib().gen_queries(seed, nq, record_cnt, field_cnt, rec.ctypes.data, fld.ctypes.data)
    return rec, fld
